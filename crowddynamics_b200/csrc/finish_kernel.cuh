// finish_kernel.cuh -- last kernel of a step of the once-per-pair pipeline (variant 3): everything of the step that is
// per agent.  Fluctuation, navigation sample, orientation, adjusting (logic.py:78-94,149-165,258-261), the sum of the
// agent's pair contributions written by k_pair_eval (pair_kernels.cuh) in ascending partner order, agent-obstacle
// (interactions.py:208-214), adaptive-dt velocity Verlet + shoulders (integrator.py:209-256, agents.py:473-486) and reset
// (logic.py:59-64).  The new state goes to the other ping-pong buffer in cell order (the physical re-sort is a by-product).
//
// This kernel is bound by memory LATENCY, not bandwidth or arithmetic (ncu, round 2: 71 % of the stall samples of its first
// version were long-scoreboard with DRAM at 30 %), so it is written in phases: (0) the three index loads, (1) every plane of
// the agent in one batch of independent loads, (2) the dependent gathers (navigation field, contributions) and the
// arithmetic, (3) the stores.  `in` and `out` are not restrict-qualified (node-wise use passes the same buffer), hence no
// load may follow a store.
#pragma once
#include "kernels.cuh"
#include "pair_kernels.cuh"
#include "step_kernel.cuh"

#ifndef FIN_THREADS_N
#define FIN_THREADS_N 256
#endif
#ifndef FIN_MINB_THREE
#define FIN_MINB_THREE 2
#endif
#ifndef FIN_MINB_CIRC
#define FIN_MINB_CIRC 3
#endif
constexpr int FIN_THREADS = FIN_THREADS_N;

#ifndef FIN_MINB_THREE_INPLACE
#define FIN_MINB_THREE_INPLACE 3
#endif
#ifndef FIN_THREADS_INPLACE_N
#define FIN_THREADS_INPLACE_N FIN_THREADS_N
#endif
constexpr int FIN_THREADS_INPLACE = FIN_THREADS_INPLACE_N;
#ifndef FIN_MINB_CIRC_INPLACE
#define FIN_MINB_CIRC_INPLACE 4
#endif
// INPLACE: the agents keep their slots (resident-order steps, A.inplace == 1) -- a separate instantiation, so that the constants
// the other variant carries from `in` to `out` cost neither loads nor registers here
template <int MODEL, bool INPLACE>
__global__ void __launch_bounds__(INPLACE ? FIN_THREADS_INPLACE : FIN_THREADS, MODEL == 0 ? (INPLACE ? FIN_MINB_CIRC_INPLACE : FIN_MINB_CIRC) : (INPLACE ? FIN_MINB_THREE_INPLACE : FIN_MINB_THREE))
k_finish(const StepArgs A) {
    const int t0 = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = t0 < eff_n(A.n, A.n_dev);
    // resident-order steps end with block-wide reductions: there every thread stays (idle ones shadow slot 0, store nothing)
    if (!live && !A.chain) return;
    const int t = live ? t0 : 0;
    const Soa &s = A.in;
    const Soa &o = A.out;
    const bool pairs = A.flags & CDB_STEP_AGENT_AGENT;
    // ---- phase 0: where the agent's planes and contributions live ---------------------------------------------------------
    const int oo = A.order ? A.order[t] : t;
    const int n_con = pairs ? A.pb.fill[t] : 0;
    const int con_off = pairs ? A.pb.off[t] : 0;
    if (pairs && pairs_overflowed(A.pb)) {
        // the pair list did not hold this step's pairs (or the search lattice had gone stale): the step is NOT applied (the
        // host grows the list / rebuilds the block list and repeats it); an integrating step that moves the agents to the
        // other buffer still has to leave the unchanged state there, because the host makes that buffer current
        if ((A.flags & CDB_STEP_INTEGRATOR) && !INPLACE && live) {
            for (int k = 0; k < A.n_planes; ++k) o(k, t) = s(k, oo);
            o.id[t] = s.id[oo];
            o.target[t] = s.target[oo];
        }
        return;
    }
    // ---- phase 1: the agent (one batch of independent loads) --------------------------------------------------------------
    const double px = s(PX, oo), py = s(PY, oo), vx = s(VX, oo), vy = s(VY, oo);
    double e0x = s(E0X, oo), e0y = s(E0Y, oo), fx = s(FX, oo), fy = s(FY, oo);
    const double fpx = s(FPX, oo), fpy = s(FPY, oo);
    // constants the step only carries along (k_soc, tau_0 are in the pair parameters; the contact parameters matter for an
    // agent that overlaps a wall; the fluctuation scales for that node) are not even read when the agents keep their slots
    constexpr bool carry = !INPLACE;
    const bool fluct = A.flags & CDB_STEP_FLUCTUATION;
    const double radius = s(RADIUS, oo), mass = s(MASS, oo), v0 = s(V0, oo), tau_adj = s(TAU_ADJ, oo);
    const double k_soc = carry ? s(K_SOC, oo) : 0.0, tau_0 = carry ? s(TAU_0, oo) : 0.0;
    const double mu = carry ? s(MU, oo) : 0.0, kappa = carry ? s(KAPPA, oo) : 0.0, damping = carry ? s(DAMPING, oo) : 0.0;
    const double srf = carry || fluct ? s(STD_RAND_FORCE, oo) : 0.0;
    const int id = s.id[oo];
    const long long target = s.target[oo];
    double lsx = 0, lsy = 0, rsx = 0, rsy = 0, r_t = 0, r_s = 0, r_ts = 0, inertia = 0, omega0 = 0, phi = 0, w = 0, phi0 = 0, tq = 0,
           tq_prev = 0, tau_rot = 0, srt = 0;
    if (MODEL == 1) {
        lsx = s(LSX, oo); lsy = s(LSY, oo); rsx = s(RSX, oo); rsy = s(RSY, oo);
        r_t = s(R_T, oo); r_s = s(R_S, oo); r_ts = s(R_TS, oo); inertia = s(INERTIA, oo); omega0 = s(OMEGA0, oo);
        phi = s(PHI, oo); w = s(OMEGA, oo); phi0 = s(PHI0, oo); tq = s(TORQUE, oo); tq_prev = s(TORQUE_PREV, oo);
        tau_rot = s(TAU_ROT, oo); srt = carry || fluct ? s(STD_RAND_TORQUE, oo) : 0.0;
    }
    // ---- phase 2: the nodes, in the reference's post-order ----------------------------------------------------------------
    if (A.flags & CDB_STEP_FLUCTUATION) fluctuation(A.seed, *A.step_ptr, id, mass, srf, inertia, srt, MODEL == 1, fx, fy, tq);
    if (A.flags & CDB_STEP_NAVIGATION) navigation_sample(A.nav, A.n_nav, target, px, py, e0x, e0y);
    if (MODEL == 1 && (A.flags & CDB_STEP_ORIENTATION)) phi0 = atan2(e0y, e0x);
    if (A.flags & CDB_STEP_ADJUSTING) {
        double ax, ay;
        adjust_force(mass, tau_adj, v0, e0x, e0y, vx, vy, ax, ay);
        fx += ax; fy += ay;
        if (MODEL == 1) tq += adjust_torque(inertia, tau_rot, phi0, phi, omega0, w);
    }
    if (n_con > 0) gather_contributions(A.pb.cres + (size_t)con_off * 4, n_con, MODEL == 1, fx, fy, tq);
    if ((A.flags & CDB_STEP_AGENT_OBSTACLE) && A.n_obs > 0) {
        auto contact = [&](double &m, double &k, double &d) {
            if (carry) { m = mu; k = kappa; d = damping; }
            else { m = s(MU, oo); k = s(KAPPA, oo); d = s(DAMPING, oo); }      // in place: no store to these planes anywhere
        };
        if (MODEL == 0) walls_circular(px, py, radius, vx, vy, contact, A.obs, A.n_obs, fx, fy);
        else walls_three_circle(px, py, lsx, lsy, rsx, rsy, r_t, r_s, vx, vy, contact, A.obs, A.n_obs, fx, fy, tq);
    }
    const bool rst = A.flags & CDB_STEP_RESET;
    if (!(A.flags & CDB_STEP_INTEGRATOR)) {
        // node-wise use: publish what the selected nodes wrote, in place
        o(E0X, oo) = e0x; o(E0Y, oo) = e0y;
        o(FX, oo) = rst ? 0.0 : fx; o(FY, oo) = rst ? 0.0 : fy;
        if (MODEL == 1) { o(PHI0, oo) = phi0; o(TORQUE, oo) = rst ? 0.0 : tq; }
        return;
    }
    const double dt = adaptive_timestep(A.vmax, A.dt_min, A.dt_max);
    double nvx = vx, nvy = vy, npx = px, npy = py;
    verlet(fx, fpx, mass, dt, nvx, npx);
    verlet(fy, fpy, mass, dt, nvy, npy);
    double ox = 0.0, oy = 0.0;
    if (MODEL == 1) {
        verlet(tq, tq_prev, inertia, dt, w, phi);
        phi = wrap_to_pi(phi);
        ox = sin(phi) * r_ts; oy = -cos(phi) * r_ts;   // shoulders(), agents.py:473-486
    }
    // ---- phase 3: the new state, in cell order (in place: only what the step changes) -------------------------------------
    if (t0 == 0) {
        A.dt_out[0] = dt; A.dt_out[1] += dt;
        if (A.dt_log) A.dt_log[*A.step_ptr % DT_LOG_SLOTS] = dt;
    }
    if (live) {
        o(PX, t) = npx; o(PY, t) = npy; o(VX, t) = nvx; o(VY, t) = nvy;
        o(E0X, t) = e0x; o(E0Y, t) = e0y;
        o(FX, t) = rst ? 0.0 : fx; o(FY, t) = rst ? 0.0 : fy;
        o(FPX, t) = fx; o(FPY, t) = fy;
        if (!INPLACE) {
            o(RADIUS, t) = radius; o(MASS, t) = mass; o(V0, t) = v0; o(TAU_ADJ, t) = tau_adj;
            o(K_SOC, t) = k_soc; o(TAU_0, t) = tau_0; o(MU, t) = mu; o(KAPPA, t) = kappa; o(DAMPING, t) = damping;
            o(STD_RAND_FORCE, t) = srf;
            o.target[t] = target;
        }
        if (MODEL == 1) {
            o(LSX, t) = npx - ox; o(LSY, t) = npy - oy; o(RSX, t) = npx + ox; o(RSY, t) = npy + oy;
            o(PHI, t) = phi; o(OMEGA, t) = w; o(PHI0, t) = phi0;
            o(TORQUE, t) = rst ? 0.0 : tq; o(TORQUE_PREV, t) = tq;
            if (!INPLACE) {
                o(R_T, t) = r_t; o(R_S, t) = r_s; o(R_TS, t) = r_ts; o(INERTIA, t) = inertia; o(OMEGA0, t) = omega0;
                o(TAU_ROT, t) = tau_rot; o(STD_RAND_TORQUE, t) = srt;
            }
        }
    }
    if (A.chain) {
        // ---- resident-order steps: the next step sweeps THESE slots again, so its neighbour records are written here -- the
        // same values k_records would derive from the stored state -- and the maxima adaptive_timestep needs of the new
        // state plus the largest displacement of this step are reduced block-wide
        if (live) {
            if (MODEL == 0) {
                double2 *r = reinterpret_cast<double2 *>(A.rec_nbr + (size_t)t * REC_CIRC);
                r[0] = make_double2(npx, npy); r[1] = make_double2(nvx, nvy);
                r[2] = make_double2(radius, radius * (1.0 + 1e-12));   // (constant, but a 48 B record left partly unwritten costs
                                                                      // the L2 a read-modify-write of its straddling sectors)
            } else {
                // body extent: a rigid body's is constant, r_ts + r_s around the centre (the stored shoulder positions differ
                // from centre -+ o by one rounding of the coordinate, which the sweep's 1e-9 inflation of the extent covers)
                const double nlx = npx - ox, nly = npy - oy, nrx = npx + ox, nry = npy + oy;
                const double ext = fmax(r_t, r_ts * (1.0 + 1e-12) + r_s) * (1.0 + 1e-12);
                double2 *r = reinterpret_cast<double2 *>(A.rec_nbr + (size_t)t * REC_THREE);
                r[0] = make_double2(npx, npy); r[1] = make_double2(nvx, nvy);
                if (!INPLACE) { r[2] = make_double2(ext, r_t); r[3] = make_double2(r_s, (double)id); }   // kept slot: unchanged
                r[4] = make_double2(nlx, nly); r[5] = make_double2(nrx, nry);
                r[6] = make_double2(ox, oy);
                r[7] = make_double2(floor(npx / A.cell_size), floor(npy / A.cell_size));
                double2 *q = reinterpret_cast<double2 *>(A.rec_sweep + (size_t)t * REC_CIRC);
                q[0] = make_double2(npx, npy); q[1] = make_double2(nvx, nvy);
                q[2] = make_double2(ext, ext * (1.0 + 1e-9));          // whole 48 B record: no partially written sectors
            }
        }
        const double ninf = -__longlong_as_double(0x7ff0000000000000LL);
        const double sp = live ? hypot(nvx, nvy) : 0.0;
        unsigned long long vm = ordered_bits(sp > 0.0 ? sp : 0.0);               // NaN speeds are skipped (k_cell_count)
        unsigned long long vb = !live ? ordered_bits(ninf) : (isnan(v0) ? 0xffffffffffffffffULL : ordered_bits(v0));
        double dd = live ? hypot(npx - px, npy - py) : 0.0;
        if (isnan(dd)) dd = -ninf;                                              // unknown drift: never trust the kept order
        unsigned long long db = (unsigned long long)__double_as_longlong(dd);
        constexpr int BT = INPLACE ? FIN_THREADS_INPLACE : FIN_THREADS;
        __shared__ unsigned long long s_red[3][BT / 32];
        vm = warp_max_u64(vm); vb = warp_max_u64(vb); db = warp_max_u64(db);
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        if (lane == 0) { s_red[0][warp] = vm; s_red[1][warp] = vb; s_red[2][warp] = db; }
        __syncthreads();
        if (warp == 0) {
            constexpr int NW = BT / 32;
            vm = lane < NW ? s_red[0][lane] : 0ULL; vb = lane < NW ? s_red[1][lane] : 0ULL; db = lane < NW ? s_red[2][lane] : 0ULL;
            vm = warp_max_u64(vm); vb = warp_max_u64(vb); db = warp_max_u64(db);
            if (lane == 0) { atomicMax(&A.chain->vmax_next[0], vm); atomicMax(&A.chain->vmax_next[1], vb); atomicMax(&A.chain->disp_step, db); }
        }
        if (live && !INPLACE) o.id[t] = id;
        return;
    }
    if (!live) return;
    // ---- strips: did the agent leave the owned columns?  Then its new state goes into the neighbour's migrant message ------
    int side = -1;
    if (A.mig.enabled) {
        const double col = floor(npx / A.mig.cell_size) - (double)A.mig.ix0;
        if (A.mig.has_left && col < (double)A.mig.col_lo) side = 0;
        else if (A.mig.has_right && col > (double)A.mig.col_hi) side = 1;
    }
    if (side >= 0) {
        const int k = atomicAdd(&A.mig.counters[side], 1);
        if (k >= A.mig.cap) { atomicExch(A.mig.error, ERR_CELL_RANGE + 3); side = -1; }
        else {
            double *d = (side == 0 ? A.mig.msg_left : A.mig.msg_right) + MSG_HEADER + (size_t)k * (A.n_planes + 2);
            d[PX] = npx; d[PY] = npy; d[VX] = nvx; d[VY] = nvy; d[E0X] = e0x; d[E0Y] = e0y;
            d[FX] = rst ? 0.0 : fx; d[FY] = rst ? 0.0 : fy; d[FPX] = fx; d[FPY] = fy;
            d[RADIUS] = radius; d[MASS] = mass; d[V0] = v0; d[TAU_ADJ] = tau_adj; d[K_SOC] = k_soc; d[TAU_0] = tau_0;
            d[MU] = mu; d[KAPPA] = kappa; d[DAMPING] = damping; d[STD_RAND_FORCE] = srf;
            if (MODEL == 1) {
                d[LSX] = npx - ox; d[LSY] = npy - oy; d[RSX] = npx + ox; d[RSY] = npy + oy;
                d[R_T] = r_t; d[R_S] = r_s; d[R_TS] = r_ts; d[INERTIA] = inertia; d[OMEGA0] = omega0;
                d[PHI] = phi; d[OMEGA] = w; d[PHI0] = phi0; d[TORQUE] = rst ? 0.0 : tq; d[TORQUE_PREV] = tq; d[TAU_ROT] = tau_rot;
                d[STD_RAND_TORQUE] = srt;
            }
            d[A.n_planes] = pack_id_flags(id, A.mig.flags);
            d[A.n_planes + 1] = (double)target;
        }
    }
    o.id[t] = side >= 0 ? -1 : id;
}
