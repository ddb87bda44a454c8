// collective_kernels.cuh -- SURVEY.md section 8(f) rank 4: exit detection, k-nearest-neighbour herding and leader-follower
// steering (reference core/evacuation.py:137-174, core/sensory_region.py:9-16, core/geom2D.py:38-59,
// core/steering/collective_motion.py:16-289; logic nodes simulation/logic.py:168-256).
//
// State: the States fields these nodes read or write live in arrays indexed by the ORIGINAL agent index ("id", the row of
// the host array), so the per-step re-sorts of the planes never move them: is_leader, is_follower (bytes), index_leader,
// familiar_exit (int64).  States.target stays per slot (Soa::target) because Navigation reads it.
// All arithmetic mirrors the reference's operation order (no FMA contraction: the library is built with -fmad=false).
#pragma once
#include "kernels.cuh"

constexpr long long NO_TARGET = -1;          // simulation/agents.py:28
constexpr long long NO_LEADER = -1;          // simulation/agents.py:29
constexpr int KNN_MAX = 32;                  // capacity of the per-agent neighbour table (size_nearest_other <= KNN_MAX)

// geom2D.py:38-59 -- segments (x0, x1) and (y0, y1)
__device__ __forceinline__ bool line_intersect(double x0x, double x0y, double x1x, double x1y, double y0x, double y0y, double y1x, double y1y) {
    const double ux = x1x - x0x, uy = x1y - x0y, vx = y1x - y0x, vy = y1y - y0y, bx = y0x - x0x, by = y0y - x0y;
    const double d = ux * vy - uy * vx;
    if (d == 0.0) return false;
    const double t0 = bx * vy - by * vx, t1 = bx * uy - by * ux;
    const double q0 = t0 / d, q1 = t1 / d;
    return 0.0 <= q0 && q0 <= 1.0 && 0.0 <= q1 && q1 <= 1.0;
}

// sensory_region.py:9-16 -- obstacle records: {p0x, p0y, p1x, p1y, ...} (SEG doubles each)
__device__ __forceinline__ bool is_obstacle_between_points(double p0x, double p0y, double p1x, double p1y, const double *__restrict__ obs, int n_obs) {
    for (int w = 0; w < n_obs; ++w) {
        const double2 a = __ldg(reinterpret_cast<const double2 *>(obs + (size_t)w * SEG));
        const double2 b = __ldg(reinterpret_cast<const double2 *>(obs + (size_t)w * SEG + 2));
        if (line_intersect(p0x, p0y, p1x, p1y, a.x, a.y, b.x, b.y)) return true;
    }
    return false;
}

__device__ __forceinline__ void normalize2(double x, double y, double &ox, double &oy) {   // vector2D.py:152-163
    const double l = hypot(x, y);
    if (l != 0.0) { ox = x / l; oy = y / l; } else { ox = x; oy = y; }
}

// collective_motion.py:25-58, first return value ("agent 1 is not heading towards agent 2's back": follow it)
__device__ __forceinline__ bool is_heading_away(double x1x, double x1y, double x2x, double x2y, double v1x, double v1y, double v2x, double v2y, double cos_phi) {
    if (hypot(v1x, v1y) == 0.0 || hypot(v2x, v2y) == 0.0) return false;
    double ex, ey, bx, by;
    normalize2(x2x - x1x, x2y - x1y, ex, ey);
    normalize2(v2x, v2y, bx, by);
    // (False, False) / (True, False) / (False, True) / (True, True) for the four combinations of cos_phi < c_i < 1 and
    // cos_phi < c_j < 1: the first value only depends on c_j
    const double c_j = -(ex * bx + ey * by);
    return !(cos_phi < c_j && c_j < 1.0);
}

// target = target_in[id] for every slot (cdb_set_states) / the inverse (cdb_get_states)
__global__ void k_target_scatter(Soa s, int n, const long long *__restrict__ target_by_id) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n && s.id[t] >= 0) s.target[t] = target_by_id[s.id[t]];
}
__global__ void k_slot_map(Soa s, int n, int *__restrict__ slot_of_id, long long *__restrict__ target_by_id) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n && s.id[t] >= 0) { slot_of_id[s.id[t]] = t; target_by_id[s.id[t]] = s.target[t]; }
}

// ---- ExitDetection (logic.py:237-256, evacuation.py:137-174) -----------------------------------------------------------------
__global__ void k_exit_detection(Soa s, int n, const double *__restrict__ doors, int n_doors, const double *__restrict__ obs, int n_obs,
                                 double detection_range, long long *__restrict__ detected_by_id, uint8_t *__restrict__ has_by_id,
                                 uint8_t *__restrict__ is_follower, int apply) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n || s.id[t] < 0) return;
    const double px = s(PX, t), py = s(PY, t);
    double distance = detection_range;
    long long detected = -1;
    for (int c = 0; c < n_doors; ++c) {
        const double cx = __ldg(doors + 2 * c), cy = __ldg(doors + 2 * c + 1);
        if (is_obstacle_between_points(px, py, cx, cy, obs, n_obs)) continue;
        const double d = hypot(cx - px, cy - py);
        if (d < distance) { distance = d; detected = c; }
    }
    const int id = s.id[t];
    detected_by_id[id] = detected;
    has_by_id[id] = detected >= 0;
    if (apply && detected >= 0 && is_follower[id]) {   // logic.py:253-255
        s.target[t] = detected;
        is_follower[id] = 0;
    }
}

// ---- find_nearest_neighbors + herding_interaction (collective_motion.py:69-154) -----------------------------------------------
// One thread per agent in cell order over the block list built with cell_size = sight.  The candidates of the 3x3 cells
// are visited in ascending sorted-slot order, which is the order in which the reference's pair iteration presents them to
// this agent, and the table is maintained with the reference's replace-the-first-maximum rule, so the rows (and therefore
// the summation order of herding_interaction) come out identical.  `rec` doubles per neighbour record, {px, py, vx, vy} first.
__global__ void k_herding(Soa s, int n_host, const int *n_dev, const double *__restrict__ nbr, int rec, const Grid *grid,
                          const int *__restrict__ cell_sorted, const int *__restrict__ cell_start, const int *__restrict__ cell_count,
                          const int *__restrict__ order, const double *__restrict__ obs, int n_obs, double sight, int k,
                          const uint8_t *__restrict__ is_follower, int all_agents, double weight_position, double cos_phi,
                          long long *__restrict__ knn_by_id, double *__restrict__ dir_by_id, uint8_t *__restrict__ has_by_id) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= eff_n(n_host, n_dev)) return;
    const int slot = order[t], id = s.id[slot];
    const bool herding = is_follower[id];
    if (!herding && !all_agents) {
        if (dir_by_id) { dir_by_id[2 * id] = 0.0; dir_by_id[2 * id + 1] = 0.0; has_by_id[id] = 0; }
        return;
    }
    double dist[KNN_MAX];
    int nb[KNN_MAX];
    for (int q = 0; q < k; ++q) { dist[q] = sight; nb[q] = -1; }
    double dmax = sight, dmax2 = sight * sight * (1.0 + 1e-9);
    const double2 mp = __ldg(reinterpret_cast<const double2 *>(nbr + (size_t)t * rec));
    const int ny = (int)grid->ny, nxg = (int)grid->nx;
    const int c = cell_sorted[t];
    const int cx = c / ny, cy = c - cx * ny;
    const int ylo = cy > 0 ? cy - 1 : 0, yhi = cy + 1 < ny ? cy + 1 : ny - 1;
    for (int dx = -1; dx <= 1; ++dx) {
        const int x2 = cx + dx;
        if (x2 < 0 || x2 >= nxg) continue;
        const int b = cell_start[x2 * ny + ylo], e = cell_start[x2 * ny + yhi] + cell_count[x2 * ny + yhi];
        for (int u = b; u < e; ++u) {
            if (u == t) continue;
            const double2 op = __ldg(reinterpret_cast<const double2 *>(nbr + (size_t)u * rec));
            const double rx = mp.x - op.x, ry = mp.y - op.y;
            if (rx * rx + ry * ry > dmax2) continue;       // conservative prefilter, the exact test follows
            const double l = hypot(rx, ry);
            if (!(l < dmax)) continue;
            // line of sight in the reference's pair orientation (i = the agent that comes first in cell order)
            const bool me_first = t < u;
            if (is_obstacle_between_points(me_first ? mp.x : op.x, me_first ? mp.y : op.y, me_first ? op.x : mp.x, me_first ? op.y : mp.y, obs, n_obs))
                continue;
            int arg = 0;                                   // set_neighbor (:61-66): np.argmax = first maximum
            for (int q = 1; q < k; ++q) if (dist[q] > dist[arg]) arg = q;
            nb[arg] = u; dist[arg] = l;
            dmax = dist[0];
            for (int q = 1; q < k; ++q) dmax = fmax(dmax, dist[q]);
            dmax2 = dmax * dmax * (1.0 + 1e-9);
        }
    }
    if (knn_by_id)
        for (int q = 0; q < k; ++q) knn_by_id[(size_t)id * k + q] = nb[q] < 0 ? -1 : (long long)s.id[order[nb[q]]];
    if (!dir_by_id) return;
    double ox = 0.0, oy = 0.0;
    bool has = false;
    if (herding) {   // herding_interaction (:113-154)
        const double2 mv = __ldg(reinterpret_cast<const double2 *>(nbr + (size_t)t * rec + 2));
        double mpx = 0.0, mpy = 0.0, mvx = 0.0, mvy = 0.0;
        int num = 0;
        for (int q = 0; q < k; ++q) {
            if (nb[q] < 0) continue;
            const double2 op = __ldg(reinterpret_cast<const double2 *>(nbr + (size_t)nb[q] * rec));
            const double2 ov = __ldg(reinterpret_cast<const double2 *>(nbr + (size_t)nb[q] * rec + 2));
            if (is_heading_away(mp.x, mp.y, op.x, op.y, mv.x, mv.y, ov.x, ov.y, cos_phi)) {
                mpx += op.x; mpy += op.y; mvx += ov.x; mvy += ov.y; ++num;
            }
        }
        if (num > 0) {
            double e0x, e0y, e1x, e1y;
            normalize2(mpx / num - mp.x, mpy / num - mp.y, e0x, e0y);
            normalize2(mvx, mvy, e1x, e1y);
            normalize2(weight_position * e0x + (1 - weight_position) * e1x, weight_position * e0y + (1 - weight_position) * e1y, ox, oy);
            has = true;
            s.target[slot] = NO_TARGET;                    // collective_motion.py:274
        }
    }
    dir_by_id[2 * id] = ox; dir_by_id[2 * id + 1] = oy; has_by_id[id] = has;
}

// ---- leader_follower_interaction_brute + the tails of leader_follower_interaction / ..._with_herding_interaction ----------------
// (collective_motion.py:157-289).  One thread per slot.  Leaders are visited in ascending (distance, position in the id-sorted
// leader list); only leaders within `sight` matter (the reference `continue`s the others), so every scan prefilters on the
// squared distance.  Reads of other agents' target go to a snapshot taken before the launch (the reference mutates in place,
// which is the same thing as long as no agent is leader and follower at once).
__global__ void k_leader_follower(Soa s, int n, const double *__restrict__ obs, int n_obs, const int *__restrict__ leader_ids, int n_leaders,
                                  const int *__restrict__ slot_of_id, const long long *__restrict__ target_by_id, int n_ids,
                                  const uint8_t *__restrict__ is_follower, long long *__restrict__ index_leader,
                                  const long long *__restrict__ familiar_exit, double sight, double cos_phi, double weight_position,
                                  const double *__restrict__ dir_herding, const uint8_t *__restrict__ has_direction, double weight_direction,
                                  double *__restrict__ direction_by_id) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n || s.id[t] < 0) return;
    const int id = s.id[t];
    const bool follower = is_follower[id];
    double lx = 0.0, ly = 0.0;          // new_direction of the brute part
    bool has_strategy = false;
    if (follower) {
        const double px = s(PX, t), py = s(PY, t), vx = s(VX, t), vy = s(VY, t);
        const double pre = sight * sight * (1.0 + 1e-9);
        int behind_obstacle = 0, heading_away = 0;
        double prev_d = -1.0;
        int prev_k = -1;
        while (true) {
            double best_d = 0.0;
            int best_k = -1;
            for (int q = 0; q < n_leaders; ++q) {
                const int ls = slot_of_id[__ldg(leader_ids + q)];
                const double dx = px - s(PX, ls), dy = py - s(PY, ls);
                if (!(dx * dx + dy * dy <= pre)) continue;
                const double d = hypot(dx, dy);
                if (d > sight) continue;
                if (!(d > prev_d || (d == prev_d && q > prev_k))) continue;        // already visited
                if (best_k < 0 || d < best_d) { best_d = d; best_k = q; }          // ties: smaller q first
            }
            if (best_k < 0) break;
            prev_d = best_d; prev_k = best_k;
            const int j = __ldg(leader_ids + best_k), ls = slot_of_id[j];
            const double qx = s(PX, ls), qy = s(PY, ls);
            if (is_obstacle_between_points(px, py, qx, qy, obs, n_obs)) {
                const long long leader = index_leader[id];
                if (leader != NO_LEADER && leader == j) {                          // keep following the leader we remember
                    ++behind_obstacle;
                    s.target[t] = target_by_id[leader];
                    has_strategy = true;
                    break;
                }
                continue;
            }
            const double ux = s(VX, ls), uy = s(VY, ls);
            if (is_heading_away(px, py, qx, qy, vx, vy, ux, uy, cos_phi)) {
                ++heading_away;
                index_leader[id] = j;
                s.target[t] = NO_TARGET;
                double e0x, e0y, e1x, e1y;
                normalize2(qx - px, qy - py, e0x, e0y);
                normalize2(ux, uy, e1x, e1y);
                normalize2(weight_position * e0x + (1 - weight_position) * e1x, weight_position * e0y + (1 - weight_position) * e1y, lx, ly);
                has_strategy = true;
                break;
            }
        }
        if (behind_obstacle == 0 && heading_away == 0) {
            const long long leader = index_leader[id];
            if (leader != NO_LEADER && leader >= 0 && leader < n_ids) { s.target[t] = target_by_id[leader]; has_strategy = true; }
        }
    }
    double ox = lx, oy = ly;
    bool has_dir = false;
    if (dir_herding) {      // ..._with_herding_interaction (:283-289)
        has_dir = has_direction[id];
        const double hx = dir_herding[2 * id], hy = dir_herding[2 * id + 1];
        normalize2(weight_direction * lx + (1 - weight_direction) * hx, weight_direction * ly + (1 - weight_direction) * hy, ox, oy);
    }
    if (follower && !(has_dir || has_strategy)) s.target[t] = familiar_exit[id];   // use familiar exits (:239-241, :284-285)
    direction_by_id[2 * id] = ox; direction_by_id[2 * id + 1] = oy;
    if (follower) { s(E0X, t) = ox; s(E0Y, t) = oy; }                              // logic.py:181-182, 220-221
}
