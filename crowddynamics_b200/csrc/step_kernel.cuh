// step_kernel.cuh -- the fused per-agent step kernel (navigation sample, orientation, adjusting, agent-agent,
// agent-obstacle, adaptive-dt velocity Verlet, reset) over cell-sorted state.
//
// Agent-agent part: one thread per cell-sorted agent ("target"), one warp per 32 consecutive targets, two phases.
//
//   phase 1 (classify)  every lane sweeps the three neighbour cell-columns of its own target, reading the candidates'
//                       packed neighbour records (48 B circular / first 64 B of 128 B three-circle; built by k_gather), and
//                       decides with ~25 branch-free fp64 operations -- no sqrt, div or hypot -- whether the pair can
//                       contribute a force at all:
//                         gate     d^2 <= (3 + R)^2 (1 + eps)            conservative form of  h < SIGTH_SOC
//                         contact  d^2 <= R^2 (1 + eps)                  conservative form of  h < 0
//                         social   disc > 0 and b > 0   (circular: the reference's own a, b, c, disc = b*b - a*c, computed
//                                  with its exact operation order, so this is exactly the necessary condition for a
//                                  non-zero social force, power_law.py:236-246);
//                                  three-circle: the same test on the bounding circles (R = sum of body extents), which
//                                  contains every one of the 9 part pairs of power_law.py:308-329.
//                       Survivors (3-6 of ~115 candidates at 1 agent/m^2) go to a per-lane list in shared memory.
//   phase 2 (evaluate)  the ragged per-lane lists are flattened and dealt out evenly over the 32 lanes; each lane runs the
//                       exact reference pair arithmetic (hypot gate, tau, gradient, exp, truncation, contact, torque) for
//                       whichever (target, neighbour) it was dealt, and the target lanes add up their results in list
//                       order -- deterministic, independent of how the work was dealt.
// A pair dropped in phase 1 contributes exactly (0, 0) in the reference too, so results equal evaluating every candidate.
#pragma once
#include "kernels.cuh"
#include "pair_kernels.cuh"

constexpr int STEP_THREADS = 128;
constexpr int DT_LOG_SLOTS = 1024;
constexpr int STEP_WARPS = STEP_THREADS / 32;
constexpr int LCAP = 16;     // survivor list entries per lane between two flushes
#ifndef STEP_CHUNK
#define STEP_CHUNK 4
#endif
constexpr int CHUNK = STEP_CHUNK;     // candidates classified between two list-capacity checks

// strips: an agent whose new position lies in a cell_size column outside the owned range is handed to the neighbour --
// the finish kernel appends its whole new state to that side's migrant message and vacates the slot (id = -1)
// Per-agent flags of the host-visible state nodes (InsideDomain: active; TargetReached: reached_by of up to FLAG_POLYGONS
// polygons), kept in arrays indexed by GLOBAL agent id on every rank.  A rank's entries are authoritative for the agents it
// owns; when an agent migrates its flags travel in the upper bits of the id slot of the migrant record (the id itself takes
// the low 32 bits of the 53-bit mantissa) and are written into the receiver's arrays.
constexpr int FLAG_POLYGONS = 20;
struct AgentFlags {
    uint8_t *active;         // [n_global] or nullptr
    uint8_t *reached;        // [np][stride] or nullptr
    long long stride;
    int np;
};
__device__ __forceinline__ double pack_id_flags(int id, const AgentFlags &f) {
    unsigned long long bits = 0ULL;
    if (f.active && f.active[id]) bits |= 1ULL;
    if (f.reached) for (int p = 0; p < f.np; ++p) if (f.reached[(size_t)p * f.stride + id]) bits |= 2ULL << p;
    return (double)(((unsigned long long)(unsigned int)id) | (bits << 32));
}
__device__ __forceinline__ int unpack_id_flags(double v, const AgentFlags &f) {
    const unsigned long long w = (unsigned long long)v;
    const int id = (int)(unsigned int)(w & 0xffffffffULL);
    const unsigned long long bits = w >> 32;
    if (f.active) f.active[id] = (uint8_t)(bits & 1ULL);
    if (f.reached) for (int p = 0; p < f.np; ++p) f.reached[(size_t)p * f.stride + id] = (uint8_t)((bits >> (p + 1)) & 1ULL);
    return id;
}

struct MigrantArgs {
    int enabled;
    double cell_size;
    long long ix0;           // cell_size column of the local lattice origin
    int col_lo, col_hi, has_left, has_right;
    double *msg_left, *msg_right;
    long long cap;
    int *counters;           // [0] left, [1] right
    int *error;
    AgentFlags flags;        // flags that travel with a migrant (all null: none)
};

struct StepArgs {
    Soa in, out;             // out == in unless integrating (then the new state is written to the other buffer)
    const double *nbr;       // packed neighbour records of `in`
    const double *nbr_sweep; // compact 48 B sweep records {px, py, vx, vy, R, -} (circular: == nbr)
    double cell_size;
    int n;                   // targets (host-side bound)
    const int *n_dev;        // device-side exact count (nullptr: n is exact)
    const Grid *grid;
    const int *cell_sorted, *cell_start, *cell_count;
    const int *order;        // sorted slot -> slot of `in` holding that agent (nullptr: `in` is physically in cell order)
    const NavField *nav;
    int n_nav;
    const double *obs;
    int n_obs;
    unsigned flags;
    double dt_min, dt_max;
    const unsigned long long *vmax;
    double *dt_out;          // [0] dt, [1] time_tot
    double *dt_log;          // ring of DT_LOG_SLOTS entries: this step's dt goes to slot step % DT_LOG_SLOTS (or nullptr)
    unsigned long long seed; // Fluctuation: Philox key (with the step index)
    const unsigned long long *step_ptr;   // device-side step index (advanced by k_step_advance after every step)
    PairBuf pb;              // k_finish: per-agent contributions written by k_pair_eval (pair_kernels.cuh)
    int n_planes;            // planes of the model (k_finish: a step that is not applied still moves them to `out`)
    MigrantArgs mig;         // k_finish in strip mode
    int reach;               // k_step: cell columns / rows swept on either side of the target's cell (1, or 2 on the finer lattice)
    // resident-order steps (k_finish): see ChainState in kernels.cuh
    ChainState *chain;       // non-null: reduce max |v|, max v0, max |dx| of the new state and write the next step's records
    double *rec_nbr, *rec_sweep;   // where those records go (the buffers this step's sweep / evaluation read)
    int inplace;             // 1: the agents keep their slots (in == out, order == nullptr): only mutable planes are written
};

struct WarpSmem {
    int list[LCAP * 32];     // entry k of lane l at list[k * 32 + l]
    int pre[33];             // exclusive prefix of the per-lane counts, pre[32] = total
    double res[3][32];       // results of one dealt round
};

// ---- exact three-circle pair in the reference's (i, j) orientation, register-only (no dynamic indexing) ---------------
__device__ __forceinline__ void pair_three_exact(const Three &I, const Three &J, bool me_is_i, const ThreePar &me,
                                                 double &fx, double &fy, double &torque) {
    const double jx[3] = {J.x0, J.x1, J.x2}, jy[3] = {J.y0, J.y1, J.y2}, rj[3] = {J.rt, J.rs, J.rs};
    // distance_three_circles (distance.py:55-105): strict '<', first wins, order torso, left, right
    double h_min = nan(""), sx = 0.0, sy = 0.0, sd = 0.0;   // selected x, y, d (normal = x / d computed once, same value)
    int i_min = 0, j_min = 0;
#pragma unroll 1
    for (int pi = 0; pi < 3; ++pi) {
        const double xi = sel3(pi, I.x0, I.x1, I.x2), yi = sel3(pi, I.y0, I.y1, I.y2), rip = pi == 0 ? I.rt : I.rs;
#pragma unroll
        for (int pj = 0; pj < 3; ++pj) {
            double x = xi - jx[pj], y = yi - jy[pj];
            double d = hypot(x, y);
            double h = d - (rip + rj[pj]);
            if (h < h_min || isnan(h_min)) { h_min = h; sx = x; sy = y; sd = d; i_min = pi; j_min = pj; }
        }
    }
    if (!(h_min < SIGTH_SOC)) return;
    double nx = 0.0, ny = 0.0;
    if (sd != 0.0) { nx = sx / sd; ny = sy / sd; }
    const double vx = I.vx - J.vx, vy = I.vy - J.vy;
    const double a = vx * vx + vy * vy;
    double fsx = 0.0, fsy = 0.0;
    if (a != 0.0) {
        // smallest time-to-collision over the 9 part pairs with the reference's selection rule (power_law.py:308-329):
        // `isnan(tau) or 0 < tau_new < tau`.  tau_new = (b - d) / a is monotone in its numerator, so the division is only
        // needed when the numerator is positive and smaller than the selected one -- same decisions, fewer divisions.
        double tau = nan(""), num_sel = 0.0, b_min = 0.0, d_min = 0.0;
        int contact_i = 0, contact_j = 0;
#pragma unroll 1
        for (int pi = 0; pi < 3; ++pi) {
            const double xi = sel3(pi, I.x0, I.x1, I.x2), yi = sel3(pi, I.y0, I.y1, I.y2), rip = pi == 0 ? I.rt : I.rs;
#pragma unroll
            for (int pj = 0; pj < 3; ++pj) {
                double x = xi - jx[pj], y = yi - jy[pj];
                double r_tot = rip + rj[pj];
                double b = -(x * vx + y * vy);
                double c = (x * x + y * y) - r_tot * r_tot;
                double disc = b * b - a * c;
                if (!(disc > 0.0)) continue;         // sqrt gives NaN (disc < 0 or NaN) or 0
                double dd = sqrt(disc);
                double num = b - dd;
                bool take = false;
                double tau_new = 0.0;
                if (isnan(tau)) { tau_new = num / a; take = true; }
                else if (num > 0.0 && num < num_sel) { tau_new = num / a; take = 0.0 < tau_new && tau_new < tau; }
                if (take) { tau = tau_new; num_sel = num; b_min = b; d_min = dd; contact_i = pi; contact_j = pj; }
            }
        }
        if (!(isnan(tau) || tau <= 0.0)) {
            // shoulder displacement of the contacting parts: 0 for the torso, +o for left, -o for right
            const double oix = sel3(contact_i, 0.0, I.ox, 0.0 - I.ox), oiy = sel3(contact_i, 0.0, I.oy, 0.0 - I.oy);
            const double ojx = sel3(contact_j, 0.0, J.ox, 0.0 - J.ox), ojy = sel3(contact_j, 0.0, J.oy, 0.0 - J.oy);
            double xr = I.x0 - J.x0, yr = I.y0 - J.y0;
            double ox = oix - ojx, oy = oiy - ojy;
            double gx = (vx - (a * (xr + 2 * ox) + b_min * vx) / d_min) / a;   // power_law.py:131-149
            double gy = (vy - (a * (yr + 2 * oy) + b_min * vy) / d_min) / a;
            double mag = magnitude(tau, me.tau_0);
            double mk = -me.mass * me.k_soc;
            fsx = mk * gx * mag;
            fsy = mk * gy * mag;
            if (!me_is_i) { fsx = 0.0 - fsx; fsy = 0.0 - fsy; }   // force_j[:] -= ... (power_law.py:358)
            truncate2(fsx, fsy, F_SOC_MAX);
        }
    }
    if (h_min < 0.0) {
        double cx, cy;
        force_contact(h_min, nx, ny, vx, vy, ny, -nx, me.mu, me.kappa, me.damping, cx, cy);
        if (me_is_i) { fsx += cx; fsy += cy; } else { fsx -= cx; fsy -= cy; }
    }
    double mx, my;   // moment arms, distance.py:102-103
    if (me_is_i) {    // x0[i_min] + r0[i_min] n - x0[0]
        const double mir = i_min == 0 ? I.rt : I.rs;
        mx = sel3(i_min, I.x0, I.x1, I.x2) + mir * nx - I.x0; my = sel3(i_min, I.y0, I.y1, I.y2) + mir * ny - I.y0;
    } else {          // x0[j_min] - r1[j_min] n - x1[0]   (sic: agent i's part, distance.py:103)
        const double mjr = j_min == 0 ? J.rt : J.rs;
        mx = sel3(j_min, I.x0, I.x1, I.x2) - mjr * nx - J.x0; my = sel3(j_min, I.y0, I.y1, I.y2) - mjr * ny - J.y0;
    }
    fx += fsx; fy += fsy;
    torque += mx * fsy - my * fsx;
}

__device__ __forceinline__ void sel_three(bool first, const Three &a, const Three &b, Three &o) {
    o.x0 = first ? a.x0 : b.x0; o.y0 = first ? a.y0 : b.y0; o.x1 = first ? a.x1 : b.x1; o.y1 = first ? a.y1 : b.y1;
    o.x2 = first ? a.x2 : b.x2; o.y2 = first ? a.y2 : b.y2; o.rt = first ? a.rt : b.rt; o.rs = first ? a.rs : b.rs;
    o.vx = first ? a.vx : b.vx; o.vy = first ? a.vy : b.vy; o.ox = first ? a.ox : b.ox; o.oy = first ? a.oy : b.oy;
}

// =====================================================================================================================
#ifndef STEP_MINB_CIRC
#define STEP_MINB_CIRC 8
#endif
#ifndef STEP_MINB_THREE
#define STEP_MINB_THREE 4
#endif
// Variant 2: agent-agent interactions classified and evaluated inside this kernel, from every agent's side (full stencil).
// Kept as an independent cross-check of the once-per-pair pipeline (pair_kernels.cuh + finish_kernel.cuh), which adds the same
// per-pair numbers in the same order.
template <int MODEL>
__global__ void __launch_bounds__(STEP_THREADS, MODEL == 0 ? STEP_MINB_CIRC : STEP_MINB_THREE) k_step(const StepArgs A) {
    __shared__ WarpSmem s_warp[STEP_WARPS];
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    WarpSmem &W = s_warp[threadIdx.x >> 5];
    const int t0 = t - lane;             // first target of this warp
    const bool active = t < eff_n(A.n, A.n_dev);
    const Soa &s = A.in;
    const unsigned FULL = 0xffffffffu;
    const int tt = active ? t : 0;       // inactive lanes read slot 0 and never write
    const int oo = A.order ? A.order[tt] : tt;   // where this agent's planes live in `in`
    constexpr int REC = MODEL == 0 ? REC_CIRC : REC_THREE;

    // ---- Navigation, Orientation, Adjusting (logic.py:149-165,258-261,89-94) --------------------------------------------
    double e0x = s(E0X, oo), e0y = s(E0Y, oo);
    double fx = s(FX, oo), fy = s(FY, oo), tq = 0.0, phi0 = 0.0;
    if (MODEL == 1) { tq = s(TORQUE, oo); phi0 = s(PHI0, oo); }
    if (A.flags & CDB_STEP_FLUCTUATION)      // logic.py:78-86, first in the reference's post-order
        fluctuation(A.seed, *A.step_ptr, s.id[oo], s(MASS, oo), s(STD_RAND_FORCE, oo), MODEL == 1 ? s(INERTIA, oo) : 0.0,
                    MODEL == 1 ? s(STD_RAND_TORQUE, oo) : 0.0, MODEL == 1, fx, fy, tq);
    {
        const double px = s(PX, oo), py = s(PY, oo);
        if (A.flags & CDB_STEP_NAVIGATION) navigation_sample(A.nav, A.n_nav, s.target[oo], px, py, e0x, e0y);
        if (MODEL == 1 && (A.flags & CDB_STEP_ORIENTATION)) phi0 = atan2(e0y, e0x);
        if (A.flags & CDB_STEP_ADJUSTING) {
            double ax, ay;
            adjust_force(s(MASS, oo), s(TAU_ADJ, oo), s(V0, oo), e0x, e0y, s(VX, oo), s(VY, oo), ax, ay);
            fx += ax; fy += ay;
            if (MODEL == 1) tq += adjust_torque(s(INERTIA, oo), s(TAU_ROT, oo), phi0, s(PHI, oo), s(OMEGA0, oo), s(OMEGA, oo));
        }
    }

    // ---- AgentAgentInteractions (interactions.py:191-205) -----------------------------------------------------------------
    if (A.flags & CDB_STEP_AGENT_AGENT) {
        const double *__restrict__ nbr = A.nbr;
        // what phase 1 needs of the target: centre, velocity, radius (circular) / body extent (three-circle)
        double mpx, mpy, mvx, mvy, mr;
        {
            const double *r = A.nbr_sweep + (size_t)tt * REC_CIRC;
            const double2 p = ldg2(r), v = ldg2(r + 2);
            mpx = p.x; mpy = p.y; mvx = v.x; mvy = v.y; mr = __ldg(r + 4);
        }
        const int ny = (int)A.grid->ny, nxg = (int)A.grid->nx;
        const int c = A.cell_sorted[tt];
        const int cx = c / ny, cy = c - cx * ny;
        const int reach = A.reach;
        const int ylo = max(cy - reach, 0), yhi = min(cy + reach, ny - 1);
        int cnt = 0;

        auto flush = [&]() {
#ifdef SWEEP_ONLY
            fx += (double)cnt; cnt = 0; return;
#endif
            // flatten the ragged lists: pre[l] = first flattened index of lane l
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { int w = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += w; }
            const int total = __shfl_sync(FULL, incl, 31);
            const int mypre = incl - cnt;
            W.pre[lane] = mypre;
            if (lane == 31) W.pre[32] = total;
            __syncwarp();
            for (int r0 = 0; r0 < total; r0 += 32) {
                const int f = r0 + lane;
                double rfx = 0.0, rfy = 0.0, rtq = 0.0;
                if (f < total) {
                    int l = 0;   // largest l with pre[l] <= f (lanes with empty lists share a prefix value: skipped by '<=')
#pragma unroll
                    for (int o = 16; o; o >>= 1) if (W.pre[l + o] <= f) l += o;
                    const int u = W.list[(f - W.pre[l]) * 32 + l];
                    const int tg = t0 + l;    // the target this entry belongs to
                    const int og = A.order ? A.order[tg] : tg;
                    const ThreePar par = {s(MASS, og), s(K_SOC, og), s(TAU_0, og), s(MU, og), s(KAPPA, og), s(DAMPING, og)};
                    if (MODEL == 0) {
                        const double *rm = nbr + (size_t)tg * REC, *ro = nbr + (size_t)u * REC;
                        const double2 p = ldg2(rm), v = ldg2(rm + 2), po = ldg2(ro), vo = ldg2(ro + 2);
                        const CircMe me = {p.x, p.y, v.x, v.y, __ldg(rm + 4), par.mass, par.k_soc, par.tau_0, par.mu, par.kappa, par.damping};
                        pair_circular(me, po.x, po.y, vo.x, vo.y, __ldg(ro + 4), rfx, rfy);
                    } else {
                        // reference pair orientation: i is the lexicographically smaller (cell_x, cell_y, agent index).
                        // The TRUE cell coordinates floor(p / c) (last two slots of the neighbour record) are compared, not
                        // the flat ids of the search lattice, so the convention does not depend on how the lattice was chosen
                        // (padded, fixed, clamped, per-strip).  Agents of the same true cell share a flat cell, inside which
                        // the slots are ordered by agent index.
                        const double2 ca = ldg2(nbr + (size_t)tg * REC + 14), cb = ldg2(nbr + (size_t)u * REC + 14);
                        const bool me_is_i = ca.x != cb.x ? ca.x < cb.x : (ca.y != cb.y ? ca.y < cb.y :
                                             __ldg(nbr + (size_t)tg * REC + 7) < __ldg(nbr + (size_t)u * REC + 7));   // agent index
                        Three I, J;     // loaded straight into their roles: one inlined copy of the pair arithmetic, no selects
                        load_three_rec(nbr, me_is_i ? tg : u, I);
                        load_three_rec(nbr, me_is_i ? u : tg, J);
                        pair_three_exact(I, J, me_is_i, par, rfx, rfy, rtq);
                    }
                }
                W.res[0][lane] = rfx; W.res[1][lane] = rfy;
                if (MODEL == 1) W.res[2][lane] = rtq;
                __syncwarp();
                // every target adds up its own entries of this round, in list order
                const int lo = max(mypre, r0) - r0, hi = min(mypre + cnt, r0 + 32) - r0;
                const int span = __reduce_max_sync(FULL, hi - lo);
                for (int k = 0; k < span; ++k)
                    if (lo + k < hi) {
                        fx += W.res[0][lo + k]; fy += W.res[1][lo + k];
                        if (MODEL == 1) tq += W.res[2][lo + k];
                    }
                __syncwarp();
            }
            cnt = 0;
        };

        // classify until some list is nearly full or all three cell-columns are done, then evaluate; repeat.
        // (one flush call site: the pair arithmetic is inlined exactly once)
        int dx = -reach, b = 0, e = 0, maxlen = 0, k0 = 0;
        bool have_col = false, done = false;
        do {
            while (true) {
                if (!have_col) {
                    if (dx > reach) { done = true; break; }
                    const int x2 = cx + dx;
                    b = 0; e = 0;
                    if (active && x2 >= 0 && x2 < nxg) {
                        b = A.cell_start[x2 * ny + ylo];
                        e = A.cell_start[x2 * ny + yhi] + A.cell_count[x2 * ny + yhi];
                    }
                    maxlen = __reduce_max_sync(FULL, e - b);
                    k0 = 0;
                    have_col = true;
                }
                if (k0 >= maxlen) { have_col = false; ++dx; continue; }
                if (__any_sync(FULL, cnt > LCAP - CHUNK)) break;
                bool keep[CHUNK];
#pragma unroll
                for (int kk = 0; kk < CHUNK; ++kk) {
                    const int u = b + k0 + kk;
                    const bool inr = u < e;
                    const double *r = A.nbr_sweep + (size_t)(inr ? u : tt) * REC_CIRC;
                    const double2 p = ldg2(r), v = ldg2(r + 2);
                    const double ro = __ldg(r + 4);
                    const double x = mpx - p.x, y = mpy - p.y;
                    const double R = mr + ro;                    // r_tot (circular) / sum of body extents (three-circle)
                    const double d2 = x * x + y * y;
                    const double lim = SIGTH_SOC + R;
                    const double RR = R * R;
                    const bool gate = d2 <= lim * lim * (1.0 + PREFILTER_EPS);
                    const bool contact = d2 <= RR * (1.0 + PREFILTER_EPS);
                    const double vx = mvx - v.x, vy = mvy - v.y;
                    const double a = vx * vx + vy * vy;
                    const double bb = -(x * vx + y * vy);
                    bool social;
                    if (MODEL == 0) {
                        const double cc = d2 - RR;
                        const double disc = bb * bb - a * cc;
                        social = disc > 0.0 && bb > 0.0;
                    } else {
                        // bounding circles (inflated by BOUND_EPS so that rounding in the exact per-part discriminants
                        // cannot matter): no real root for them => none for any part pair; all part pairs receding
                        // (b_k <= b + R |v| <= 0) => no positive time-to-collision
                        const double Rs = R * (1.0 + BOUND_EPS), RRs = Rs * Rs;
                        const double disc = bb * bb - a * (d2 - RRs);
                        social = disc >= 0.0 && (bb >= 0.0 || bb * bb <= RRs * a);
                    }
                    keep[kk] = inr && u != t && gate && (social || contact);
                }
#pragma unroll
                for (int kk = 0; kk < CHUNK; ++kk)
                    if (keep[kk]) { W.list[cnt * 32 + lane] = b + k0 + kk; ++cnt; }
                k0 += CHUNK;
            }
            flush();
        } while (!done);
    }
    if (!active) return;

    // ---- own state for the wall and integrator parts ------------------------------------------------------------------------
    const double px = s(PX, oo), py = s(PY, oo), vx = s(VX, oo), vy = s(VY, oo), mass = s(MASS, oo);
    const double mu = s(MU, oo), kappa = s(KAPPA, oo), damping = s(DAMPING, oo);

    // ---- AgentObstacleInteractions (interactions.py:208-214) ---------------------------------------------------------------
    if ((A.flags & CDB_STEP_AGENT_OBSTACLE) && A.n_obs > 0) {
        if (MODEL == 0) {
            walls_circular(px, py, s(RADIUS, oo), vx, vy, ContactValues{mu, kappa, damping}, A.obs, A.n_obs, fx, fy);
        } else {
            walls_three_circle(px, py, s(LSX, oo), s(LSY, oo), s(RSX, oo), s(RSY, oo), s(R_T, oo), s(R_S, oo), vx, vy, ContactValues{mu, kappa, damping},
                               A.obs, A.n_obs, fx, fy, tq);
        }
    }

    const Soa &o = A.out;
    const bool rst = A.flags & CDB_STEP_RESET;
    if (!(A.flags & CDB_STEP_INTEGRATOR)) {
        // node-wise use: publish what the selected nodes wrote, in place
        o(E0X, oo) = e0x; o(E0Y, oo) = e0y;
        o(FX, oo) = rst ? 0.0 : fx; o(FY, oo) = rst ? 0.0 : fy;
        if (MODEL == 1) { o(PHI0, oo) = phi0; o(TORQUE, oo) = rst ? 0.0 : tq; }
        return;
    }

    // ---- Integrator (integrator.py:209-256) + Reset (logic.py:59-64); new state goes to the other buffer --------------------
    // Every load is issued BEFORE the first store: `in` and `out` are not restrict-qualified (they are the same buffer in
    // the node-wise path), so a load placed after a store would have to wait for it -- one DRAM round trip per plane.
    const double dt = adaptive_timestep(A.vmax, A.dt_min, A.dt_max);
    const double fpx = s(FPX, oo), fpy = s(FPY, oo);
    const double c_radius = s(RADIUS, oo), c_v0 = s(V0, oo), c_tau_adj = s(TAU_ADJ, oo), c_k_soc = s(K_SOC, oo), c_tau_0 = s(TAU_0, oo);
    const double c_srf = s(STD_RAND_FORCE, oo);
    const int c_id = s.id[oo];
    const long long c_target = s.target[oo];
    double inertia = 0.0, r_ts = 0.0, w = 0.0, phi = 0.0, tq_prev = 0.0, c_r_t = 0.0, c_r_s = 0.0, c_omega0 = 0.0, c_tau_rot = 0.0, c_srt = 0.0;
    if (MODEL == 1) {
        inertia = s(INERTIA, oo); r_ts = s(R_TS, oo); w = s(OMEGA, oo); phi = s(PHI, oo); tq_prev = s(TORQUE_PREV, oo);
        c_r_t = s(R_T, oo); c_r_s = s(R_S, oo); c_omega0 = s(OMEGA0, oo); c_tau_rot = s(TAU_ROT, oo); c_srt = s(STD_RAND_TORQUE, oo);
    }
    if (t == 0) {
        A.dt_out[0] = dt; A.dt_out[1] += dt;
        if (A.dt_log) A.dt_log[*A.step_ptr % DT_LOG_SLOTS] = dt;
    }
    double nvx = vx, nvy = vy, npx = px, npy = py;
    verlet(fx, fpx, mass, dt, nvx, npx);
    verlet(fy, fpy, mass, dt, nvy, npy);
    double ox = 0.0, oy = 0.0;
    if (MODEL == 1) {
        verlet(tq, tq_prev, inertia, dt, w, phi);
        phi = wrap_to_pi(phi);
        ox = sin(phi) * r_ts; oy = -cos(phi) * r_ts;   // shoulders(), agents.py:473-486
    }
    o(PX, t) = npx; o(PY, t) = npy; o(VX, t) = nvx; o(VY, t) = nvy;
    o(E0X, t) = e0x; o(E0Y, t) = e0y;
    o(FX, t) = rst ? 0.0 : fx; o(FY, t) = rst ? 0.0 : fy;
    o(FPX, t) = fx; o(FPY, t) = fy;
    o(RADIUS, t) = c_radius; o(MASS, t) = mass; o(V0, t) = c_v0; o(TAU_ADJ, t) = c_tau_adj;
    o(K_SOC, t) = c_k_soc; o(TAU_0, t) = c_tau_0; o(MU, t) = mu; o(KAPPA, t) = kappa; o(DAMPING, t) = damping;
    o(STD_RAND_FORCE, t) = c_srf;
    o.id[t] = c_id;
    o.target[t] = c_target;
    if (MODEL == 1) {
        o(LSX, t) = npx - ox; o(LSY, t) = npy - oy; o(RSX, t) = npx + ox; o(RSY, t) = npy + oy;
        o(R_T, t) = c_r_t; o(R_S, t) = c_r_s; o(R_TS, t) = r_ts; o(INERTIA, t) = inertia; o(OMEGA0, t) = c_omega0;
        o(PHI, t) = phi; o(OMEGA, t) = w; o(PHI0, t) = phi0;
        o(TORQUE, t) = rst ? 0.0 : tq; o(TORQUE_PREV, t) = tq; o(TAU_ROT, t) = c_tau_rot;
        o(STD_RAND_TORQUE, t) = c_srt;
    }
}
