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

// geom2D.py:38-59 -- segments (x0, x1) and (y0, y1): 0 <= t0 / d <= 1 and 0 <= t1 / d <= 1.
// The quotients are only formed when a ratio is within 1e-14 of 0 or 1 (or tiny enough to underflow): away from those
// boundaries sign and magnitude comparisons decide exactly what the reference's divisions decide.
__device__ __forceinline__ int unit_interval_class(double t, double d) {   // 1: clearly inside (0, 1), 0: clearly outside, -1: look closer
    const double at = fabs(t), ad = fabs(d);
    if (!(at > 1e-150) || !(ad > 1e-150) || !(ad < 1e150) || !(at < 1e150)) return -1;   // the quotient can neither underflow nor overflow
    if ((t < 0.0) != (d < 0.0)) return 0;
    if (at < ad * (1.0 - 1e-14)) return 1;
    if (at > ad * (1.0 + 1e-14)) return 0;
    return -1;
}
__device__ __forceinline__ bool line_intersect(double x0x, double x0y, double x1x, double x1y, double y0x, double y0y, double y1x, double y1y) {
    const double ux = x1x - x0x, uy = x1y - x0y, vx = y1x - y0x, vy = y1y - y0y, bx = y0x - x0x, by = y0y - x0y;
    const double d = ux * vy - uy * vx;
    if (d == 0.0) return false;
    const double t0 = bx * vy - by * vx, t1 = bx * uy - by * ux;
    const int c0 = unit_interval_class(t0, d);
    if (c0 == 0) return false;
    const int c1 = unit_interval_class(t1, d);
    if (c1 == 0) return false;
    if (c0 == 1 && c1 == 1) return true;
    const double q0 = t0 / d, q1 = t1 / d;
    return 0.0 <= q0 && q0 <= 1.0 && 0.0 <= q1 && q1 <= 1.0;
}

// sensory_region.py:9-16 -- obstacle records: {p0x, p0y, p1x, p1y, ...} (SEG doubles each)
__device__ __noinline__ bool is_obstacle_between_points(double p0x, double p0y, double p1x, double p1y, const double *__restrict__ obs, int n_obs) {
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
// One thread per agent in cell order over the block list built with cell_size = sight; the table is maintained with the
// reference's replace-the-first-maximum rule (set_neighbor, :61-66).  `rec` doubles per neighbour record, {px, py, vx, vy} first.
//   ORDERED = true  (cdb_nearest_neighbors): the candidates of the 3x3 cells are visited in ascending sorted-slot order, which
//                    is the order in which the reference's pair iteration presents them to this agent, so the rows come out
//                    in the reference's own slot order.
//   ORDERED = false (the herding step): the block list is built with FINE cells (about the radius expected to hold 3k agents,
//                    chosen by the host from the crowd density; cell_size <= sight).  The search first looks inside a radius
//                    `cap` estimated from the local occupancy -- own cell first, then the surrounding cells, each skipped when
//                    its nearest point is not closer than the current k-th best -- and only if fewer than k visible
//                    neighbours turn up inside `cap` repeats out to `sight`.  Lines of sight are only tested when some wall
//                    comes within the search radius of the agent.  The result is the same SET of neighbours (the k nearest
//                    visible ones within sight); only the summation order of herding_interaction differs (<= 1e-15).
//                    `exact_cells` = 0 (clamped / fixed lattice, cell_size == sight: cell rectangles are not reliable)
//                    disables the pruning.
__device__ __forceinline__ bool segment_within(double px, double py, const double *__restrict__ seg, double r) {
    const double ax = seg[0], ay = seg[1], dx = seg[2] - ax, dy = seg[3] - ay;
    const double len2 = dx * dx + dy * dy;
    double tt = len2 > 0.0 ? ((px - ax) * dx + (py - ay) * dy) / len2 : 0.0;
    tt = fmin(fmax(tt, 0.0), 1.0);
    const double qx = ax + tt * dx - px, qy = ay + tt * dy - py;
    return !(qx * qx + qy * qy > r * r * (1.0 + 1e-9));     // NaNs count as "within"
}

template <bool ORDERED>
__global__ void k_herding(Soa s, int n_host, const int *n_dev, const double *__restrict__ nbr, int rec, const Grid *grid, double cell_size,
                          int exact_cells, const int *__restrict__ cell_sorted, const int *__restrict__ cell_start,
                          const int *__restrict__ cell_count, const int *__restrict__ order, const double *__restrict__ obs, int n_obs,
                          double sight, int k, const uint8_t *__restrict__ is_follower, int all_agents, double weight_position,
                          double cos_phi, long long *__restrict__ knn_by_id, double *__restrict__ dir_by_id, uint8_t *__restrict__ has_by_id) {
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
    double dmax, dmax2;
    auto init_table = [&](double cap) {
        for (int q = 0; q < k; ++q) { dist[q] = cap; nb[q] = -1; }
        dmax = cap; dmax2 = cap * cap * (1.0 + 1e-9);
    };
    init_table(sight);
    const double2 mp = __ldg(reinterpret_cast<const double2 *>(nbr + (size_t)t * rec));
    const int ny = (int)grid->ny, nxg = (int)grid->nx;
    const int c = cell_sorted[t];
    const int cx = c / ny, cy = c - cx * ny;
    const double mcx = floor(mp.x / sight), mcy = floor(mp.y / sight);
    bool walls = n_obs > 0;

    auto sweep = [&](int b, int e) {
        for (int u0 = b; u0 < e; u0 += 4) {
            double2 cp[4];
            bool near[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {                  // four independent loads + squared distances in flight
                const int u = u0 + j < e ? u0 + j : e - 1;
                cp[j] = __ldg(reinterpret_cast<const double2 *>(nbr + (size_t)u * rec));
                const double rx = mp.x - cp[j].x, ry = mp.y - cp[j].y;
                near[j] = u0 + j < e && u != t && !(rx * rx + ry * ry > dmax2);   // conservative against the current radius
            }
            if (!(near[0] || near[1] || near[2] || near[3])) continue;
#pragma unroll 1
            for (int j = 0; j < 4; ++j) {                  // in slot order, with the exact tests (one copy of the code: the
                if (!(j == 0 ? near[0] : j == 1 ? near[1] : j == 2 ? near[2] : near[3])) continue;   // kernel is fetch-bound otherwise)
                const int u = u0 + j;
                const double2 op = j == 0 ? cp[0] : j == 1 ? cp[1] : j == 2 ? cp[2] : cp[3];
                const double l = hypot(mp.x - op.x, mp.y - op.y);
                if (!(l < dmax)) continue;
                if (walls) {
                    // line of sight in the reference's pair orientation: i = the agent that comes first in the order
                    // (cell_x, cell_y, agent index) of the block list with cell_size = sight
                    bool me_first = t < u;
                    if (!ORDERED) {
                        const double ocx = floor(op.x / sight), ocy = floor(op.y / sight);
                        me_first = mcx != ocx ? mcx < ocx : (mcy != ocy ? mcy < ocy : id < s.id[order[u]]);
                    }
                    if (is_obstacle_between_points(me_first ? mp.x : op.x, me_first ? mp.y : op.y, me_first ? op.x : mp.x, me_first ? op.y : mp.y, obs, n_obs))
                        continue;
                }
                int arg = 0;                               // set_neighbor (:61-66): np.argmax = first maximum
                for (int q = 1; q < k; ++q) if (dist[q] > dist[arg]) arg = q;
                nb[arg] = u; dist[arg] = l;
                dmax = dist[0];
                for (int q = 1; q < k; ++q) dmax = fmax(dmax, dist[q]);
                dmax2 = dmax * dmax * (1.0 + 1e-9);
            }
        }
    };

    if (ORDERED) {
        const int ylo = cy > 0 ? cy - 1 : 0, yhi = cy + 1 < ny ? cy + 1 : ny - 1;
        for (int dx = -1; dx <= 1; ++dx) {
            const int x2 = cx + dx;
            if (x2 < 0 || x2 >= nxg) continue;
            sweep(cell_start[x2 * ny + ylo], cell_start[x2 * ny + yhi] + cell_count[x2 * ny + yhi]);
        }
    } else {
        // distances from this agent to the four edges of its cell (>= 0 up to rounding, hence the slack below)
        const double cs = cell_size;
        const double x0 = (double)(grid->ix_min + cx) * cs, y0 = (double)(grid->iy_min + cy) * cs;
        const double gxm = mp.x - x0, gxp = x0 + cs - mp.x, gym = mp.y - y0, gyp = y0 + cs - mp.y;
        const double slack = 1e-9 * (1.0 + fabs(mp.x) + fabs(mp.y));
        double cap = sight;
        if (exact_cells) {      // radius expected to hold ~3k agents, from the occupancy of the 3x3 cells around this one
            int cnt9 = 0;
            for (int dx = -1; dx <= 1; ++dx)
                for (int dy = -1; dy <= 1; ++dy) {
                    const int x2 = cx + dx, y2 = cy + dy;
                    if (x2 >= 0 && x2 < nxg && y2 >= 0 && y2 < ny) cnt9 += cell_count[x2 * ny + y2];
                }
            const double r0 = sqrt(3.0 * (double)k * 9.0 * cs * cs / (3.141592653589793 * (double)max(cnt9, 1)));
            if (r0 < sight) cap = r0;
        }
#pragma unroll 1
        for (int attempt = 0; attempt < 2; ++attempt) {
            init_table(cap);
            walls = false;      // can any wall cut a line of sight shorter than cap?
            for (int w = 0; w < n_obs && !walls; ++w) walls = segment_within(mp.x, mp.y, obs + (size_t)w * SEG, cap + slack);
            const int hw = (int)ceil(cap / cs);    // every point closer than cap lies within hw cells of the own one
            const int span = 2 * hw + 1;
#pragma unroll 1
            for (int q = -1; q < span * span; ++q) {       // q = -1: the own cell first (it tightens the radius most)
                const int dx = q < 0 ? 0 : q / span - hw, dy = q < 0 ? 0 : q % span - hw;
                if (q >= 0 && dx == 0 && dy == 0) continue;
                const int x2 = cx + dx, y2 = cy + dy;
                if (x2 < 0 || x2 >= nxg || y2 < 0 || y2 >= ny) continue;
                if (exact_cells && q >= 0) {
                    const double gx = dx < 0 ? gxm + (double)(-dx - 1) * cs : dx > 0 ? gxp + (double)(dx - 1) * cs : 0.0;
                    const double gy = dy < 0 ? gym + (double)(-dy - 1) * cs : dy > 0 ? gyp + (double)(dy - 1) * cs : 0.0;
                    const double gx0 = fmax(gx, 0.0), gy0 = fmax(gy, 0.0);
                    if (sqrt(gx0 * gx0 + gy0 * gy0) - slack >= dmax) continue;   // nothing in that cell is closer than the k-th best
                }
                sweep(cell_start[x2 * ny + y2], cell_start[x2 * ny + y2] + cell_count[x2 * ny + y2]);
            }
            bool full = true;
            for (int q = 0; q < k; ++q) full = full && nb[q] >= 0;
            if (full || !(cap < sight)) break;
            cap = sight;
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

// ---- leader hash grid: leaders binned by floor(p / sight) into a power-of-two table of linked lists (collisions only add
// candidates, which the distance test rejects), so a follower looks at the 9 buckets around its own cell instead of at all
// leaders.  lrec[q] = {px, py, vx, vy} of leader q (q = position in the id-sorted leader list).
__device__ __forceinline__ unsigned leader_bucket(long long ix, long long iy, int bits) {
    const unsigned long long h = (unsigned long long)ix * 0x9E3779B97F4A7C15ULL ^ (unsigned long long)iy * 0xC2B2AE3D27D4EB4FULL;
    return (unsigned)((h * 0xD6E8FEB86659FD93ULL) >> (64 - bits));
}
__global__ void k_leader_records(Soa s, const int *__restrict__ leader_ids, int n_leaders, const int *__restrict__ slot_of_id,
                                 double *__restrict__ lrec, double sight, int *__restrict__ lhead, int *__restrict__ lnext, int bits) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_leaders) return;
    const int ls = slot_of_id[leader_ids[q]];
    const double px = s(PX, ls), py = s(PY, ls);
    lrec[4 * q] = px; lrec[4 * q + 1] = py; lrec[4 * q + 2] = s(VX, ls); lrec[4 * q + 3] = s(VY, ls);
    if (!lhead) return;
    const double cx = floor(px / sight), cy = floor(py / sight);
    if (!(fabs(cx) < 4.0e18) || !(fabs(cy) < 4.0e18)) { lnext[q] = -1; return; }   // non-finite: nobody can see this leader
    lnext[q] = atomicExch(&lhead[leader_bucket((long long)cx, (long long)cy, bits)], q);
}

// ---- leader_follower_interaction_brute + the tails of leader_follower_interaction / ..._with_herding_interaction ----------------
// (collective_motion.py:157-289).  One thread per slot.  Leaders are visited in ascending (distance, position in the id-sorted
// leader list); only leaders within `sight` matter (the reference `continue`s the others), so every scan looks at the hash-grid
// buckets around the follower (or at all leaders when there is no grid) and prefilters on the squared distance.  Reads of other agents' target go to a snapshot taken before the launch (the reference mutates in place,
// which is the same thing as long as no agent is leader and follower at once).
__global__ void k_leader_follower(Soa s, int n, const double *__restrict__ obs, int n_obs, const int *__restrict__ leader_ids, int n_leaders,
                                  const double *__restrict__ lrec, const int *__restrict__ lhead, const int *__restrict__ lnext, int bits,
                                  const long long *__restrict__ target_by_id, int n_ids,
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
        long long cix = 0, ciy = 0;
        bool use_grid = lhead != nullptr;
        if (use_grid) {
            const double cx = floor(px / sight), cy = floor(py / sight);
            if (fabs(cx) < 4.0e18 && fabs(cy) < 4.0e18) { cix = (long long)cx; ciy = (long long)cy; } else use_grid = false;
        }
        while (true) {
            double best_d = 0.0;
            int best_k = -1;
            auto consider = [&](int q) {
                const double2 lp = __ldg(reinterpret_cast<const double2 *>(lrec + 4 * (size_t)q));
                const double dx = px - lp.x, dy = py - lp.y;
                if (!(dx * dx + dy * dy <= pre)) return;
                const double d = hypot(dx, dy);
                if (d > sight) return;
                if (!(d > prev_d || (d == prev_d && q > prev_k))) return;          // already visited
                if (best_k < 0 || d < best_d || (d == best_d && q < best_k)) { best_d = d; best_k = q; }   // ties: smaller q first
            };
            if (use_grid) {
                for (int gy = -1; gy <= 1; ++gy)
                    for (int gx = -1; gx <= 1; ++gx)
                        for (int q = __ldg(lhead + leader_bucket(cix + gx, ciy + gy, bits)); q >= 0; q = __ldg(lnext + q)) consider(q);
            } else {
                for (int q = 0; q < n_leaders; ++q) consider(q);
            }
            if (best_k < 0) break;
            prev_d = best_d; prev_k = best_k;
            const int j = __ldg(leader_ids + best_k);
            const double2 lp = __ldg(reinterpret_cast<const double2 *>(lrec + 4 * (size_t)best_k));
            const double2 lv = __ldg(reinterpret_cast<const double2 *>(lrec + 4 * (size_t)best_k + 2));
            const double qx = lp.x, qy = lp.y;
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
            const double ux = lv.x, uy = lv.y;
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

// =====================================================================================================================
// InsideDomain / TargetReached (reference simulation/logic.py:343-387), i.e.
// matplotlib.path.Path.contains_points with radius 0 -- crossing-number rule, see oracle_point_in_polygon.  Polygons are
// stored back to back as (x, y) pairs with an offsets table (n_polygons + 1 entries, in vertices), implicitly closed.
// =====================================================================================================================
__device__ __forceinline__ bool point_in_polygon(const double *__restrict__ v, int nv, double tx, double ty) {
    if (nv < 3) return false;
    bool inside = false;
    double vx0 = __ldg(v + 2 * (nv - 1)), vy0 = __ldg(v + 2 * (nv - 1) + 1);
    bool yflag0 = vy0 >= ty;
    for (int k = 0; k < nv; ++k) {
        const double vx1 = __ldg(v + 2 * k), vy1 = __ldg(v + 2 * k + 1);
        const bool yflag1 = vy1 >= ty;
        if (yflag0 != yflag1)
            if (((vy1 - ty) * (vx0 - vx1) >= (vx1 - tx) * (vy0 - vy1)) == yflag1) inside = !inside;
        yflag0 = yflag1; vx0 = vx1; vy0 = vy1;
    }
    return inside;
}

// InsideDomain.update: active[id] = contains(position); counter[0] += number of flags that changed
__global__ void k_inside_domain(Soa s, int n_host, const int *n_dev, const double *__restrict__ verts, int nv, uint8_t *__restrict__ active_by_id,
                                unsigned long long *counter) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = eff_n(n_host, n_dev);
    bool changed = false;
    if (t < n && s.id[t] >= 0) {
        const int id = s.id[t];
        const bool now = point_in_polygon(verts, nv, s(PX, t), s(PY, t));
        changed = (active_by_id[id] != 0) != now;
        active_by_id[id] = now;
    }
    const unsigned m = __ballot_sync(0xffffffffu, changed);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(counter, (unsigned long long)__popc(m));
}

// TargetReached.update: reached[p][id] |= contains_p(position); counts[p] = number of agents that ever reached polygon p
__global__ void k_target_reached(Soa s, int n_host, const int *n_dev, const double *__restrict__ verts, const int *__restrict__ offsets, int n_polygons,
                                 uint8_t *__restrict__ reached, long long stride, unsigned long long *counts) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = t < eff_n(n_host, n_dev) && s.id[t] >= 0;
    const int id = live ? s.id[t] : 0;
    const double px = live ? s(PX, t) : 0.0, py = live ? s(PY, t) : 0.0;
    for (int p = 0; p < n_polygons; ++p) {
        bool fresh = false;
        if (live && !reached[(size_t)p * stride + id]) {
            const int b = offsets[p], e = offsets[p + 1];
            if (point_in_polygon(verts + 2 * (size_t)b, e - b, px, py)) { reached[(size_t)p * stride + id] = 1; fresh = true; }
        }
        const unsigned m = __ballot_sync(0xffffffffu, fresh);
        if ((threadIdx.x & 31) == 0 && m) atomicAdd(&counts[p], (unsigned long long)__popc(m));
    }
}
