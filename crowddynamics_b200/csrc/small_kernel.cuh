// small_kernel.cuh -- the launch-bound end of the path: a crowd of at most SMALL_MAX agents (the reference's own example
// simulations: Hallway, 50 agents, examples/simulations.py:74-163; README.md:39-50) advanced by ONE thread block that keeps the
// crowd in shared memory and runs MANY steps per launch.  The general pipeline costs ~16 dependent launches of single-CTA
// kernels per step for such a crowd (45 us / step even replayed as a CUDA graph); here a step costs two block barriers.
//
// A group of G = 1 .. 16 threads per agent (G the largest power of two with n G <= 512 circular / 256 three-circle threads): every thread of a group carries the
// agent's state and runs the per-agent nodes redundantly (same inputs, same results); the pair loop is split G ways and its
// partial sums are combined by a butterfly, which gives every thread of the group the same bits -- the single-thread
// critical path of a step is what a small crowd costs.  Per step, in the reference's post-order (logic.py:59-165,258-261):
//   fluctuation, navigation sample, orientation, adjusting                      (registers)
//   every agent's kinematics published to shared memory, barrier
//   agent-agent: all other agents in ascending slot order; a cheap conservative gate (the sweep test of pair_kernels.cuh
//   without its time-to-collision part) in front of the exact one-sided evaluation (pair_circular / pair_three_exact, the
//   functions of the one-phase cross-check kernels).  No block list: with 3 + 2 max R < cell_size every pair that can exert a
//   force lies in adjacent cells of ANY lattice, and pairs in adjacent cells that are out of range contribute exactly zero, so
//   the all-pairs loop adds the same terms as core/interactions.py:191-205 (the host only takes this path when the condition
//   holds); the (i, j) orientation of a three-circle pair comes from the true cells floor(p / cell_size) and the agent index,
//   as everywhere else
//   agent-obstacle, block-wide maxima for adaptive_timestep, velocity Verlet, shoulders, reset
#pragma once
#include "kernels.cuh"
#include "pair_kernels.cuh"
#include "step_kernel.cuh"

constexpr int SMALL_MAX = 256;          // agents
// threads of the one block: the circular kernel needs half the registers and takes twice the threads (more threads per agent)
template <int MODEL> struct SmallThreads { static constexpr int value = MODEL == 0 ? 512 : 256; };

struct SmallArgs {
    Soa s;                  // the state, updated in place (slot order is irrelevant here and left alone)
    int n;
    const NavField *nav;
    int n_nav;
    const double *obs;
    int n_obs;
    unsigned flags;
    double cell_size, dt_min, dt_max;
    double *dt_out;         // [0] dt, [1] time_tot
    double *dt_log;         // ring of DT_LOG_SLOTS entries (or nullptr)
    unsigned long long seed;
    unsigned long long *step_ptr;   // device-side step index, advanced here
    int n_steps;
    int group;              // threads per agent (power of two, <= 16, n * group <= threads of the block)
};

template <int MODEL>
__global__ void __launch_bounds__(SmallThreads<MODEL>::value, 1) k_small_steps(const SmallArgs A) {
    constexpr int THREADS = SmallThreads<MODEL>::value;
    __shared__ double s_rec[SMALL_MAX][MODEL == 0 ? 5 : 16];     // kinematics of every agent for the pair loop
    __shared__ unsigned long long s_vm[THREADS / 32], s_v0[THREADS / 32];
    __shared__ double s_dt;
    const int t = threadIdx.x;
    const int G = A.group, ag = t / G, sub = t - ag * G;      // agent of this thread, its place in the agent's group
    const bool live = ag < A.n && A.s.id[ag < A.n ? ag : 0] >= 0;
    const int tt = ag < A.n ? ag : 0;
    const Soa &s = A.s;
    // ---- the agent, once -------------------------------------------------------------------------------------------------
    double px = s(PX, tt), py = s(PY, tt), vx = s(VX, tt), vy = s(VY, tt), e0x = s(E0X, tt), e0y = s(E0Y, tt);
    double fx = s(FX, tt), fy = s(FY, tt), fpx = s(FPX, tt), fpy = s(FPY, tt);
    const double radius = s(RADIUS, tt), mass = s(MASS, tt), v0 = s(V0, tt), tau_adj = s(TAU_ADJ, tt), k_soc = s(K_SOC, tt),
                 tau_0 = s(TAU_0, tt), mu = s(MU, tt), kappa = s(KAPPA, tt), damping = s(DAMPING, tt), srf = s(STD_RAND_FORCE, tt);
    const int id = s.id[tt];
    const long long target = s.target[tt];
    double lsx = 0, lsy = 0, rsx = 0, rsy = 0, r_t = 0, r_s = 0, r_ts = 0, inertia = 0, omega0 = 0, phi = 0, w = 0, phi0 = 0, tq = 0,
           tq_prev = 0, tau_rot = 0, srt = 0;
    if (MODEL == 1) {
        lsx = s(LSX, tt); lsy = s(LSY, tt); rsx = s(RSX, tt); rsy = s(RSY, tt);
        r_t = s(R_T, tt); r_s = s(R_S, tt); r_ts = s(R_TS, tt); inertia = s(INERTIA, tt); omega0 = s(OMEGA0, tt);
        phi = s(PHI, tt); w = s(OMEGA, tt); phi0 = s(PHI0, tt); tq = s(TORQUE, tt); tq_prev = s(TORQUE_PREV, tt);
        tau_rot = s(TAU_ROT, tt); srt = s(STD_RAND_TORQUE, tt);
    }
    const double ext = MODEL == 0 ? radius : fmax(r_t, fmax(hypot(lsx - px, lsy - py), hypot(rsx - px, rsy - py)) + r_s);
    unsigned long long step = *A.step_ptr;
    const bool rst = A.flags & CDB_STEP_RESET;
    for (int it = 0; it < A.n_steps; ++it, ++step) {
        // ---- per-agent nodes --------------------------------------------------------------------------------------------
        if (A.flags & CDB_STEP_FLUCTUATION) fluctuation(A.seed, step, id, mass, srf, inertia, srt, MODEL == 1, fx, fy, tq);
        if (A.flags & CDB_STEP_NAVIGATION) navigation_sample(A.nav, A.n_nav, target, px, py, e0x, e0y);
        if (MODEL == 1 && (A.flags & CDB_STEP_ORIENTATION)) phi0 = atan2(e0y, e0x);
        if (A.flags & CDB_STEP_ADJUSTING) {
            double ax, ay;
            adjust_force(mass, tau_adj, v0, e0x, e0y, vx, vy, ax, ay);
            fx += ax; fy += ay;
            if (MODEL == 1) tq += adjust_torque(inertia, tau_rot, phi0, phi, omega0, w);
        }
        // ---- publish, then the pair loop -------------------------------------------------------------------------------------
        double *me = s_rec[tt];
        if (sub == 0 && ag < A.n) {
            me[0] = px; me[1] = py; me[2] = vx; me[3] = vy; me[4] = live ? ext : -1.0;      // ext < 0: empty slot
            if (MODEL == 1) {
                me[5] = r_t; me[6] = r_s; me[7] = (double)id;
                me[8] = lsx; me[9] = lsy; me[10] = rsx; me[11] = rsy;
                me[12] = r_ts * sin(phi); me[13] = r_ts * -cos(phi);                         // power_law.py:338-350
                me[14] = floor(px / A.cell_size); me[15] = floor(py / A.cell_size);
            }
        }
        __syncthreads();
        double qx = 0.0, qy = 0.0, qt = 0.0;       // this thread's share of the agent's pair contributions
        if (live && (A.flags & CDB_STEP_AGENT_AGENT)) {
            const double lim_t = SIGTH_SOC * (1.0 + BOUND_EPS) + ext * (1.0 + BOUND_EPS);
            for (int u = sub; u < A.n; u += G) {
                const double *o = s_rec[u];
                const double eo = o[4];
                if (u == ag || eo < 0.0) continue;
                const double x = px - o[0], y = py - o[1];
                const double lim = lim_t + eo * (1.0 + BOUND_EPS);
                if (!(x * x + y * y <= lim * lim)) continue;        // conservative form of h < SIGTH_SOC (bounding circles)
                if (MODEL == 0) {
                    const CircMe cm = {px, py, vx, vy, radius, mass, k_soc, tau_0, mu, kappa, damping};
                    pair_circular(cm, o[0], o[1], o[2], o[3], eo, qx, qy);
                } else {
                    // reference pair orientation: i is the lexicographically smaller (cell_x, cell_y, agent index)
                    const bool me_is_i = me[14] != o[14] ? me[14] < o[14] : (me[15] != o[15] ? me[15] < o[15] : me[7] < o[7]);
                    Three M, O, I, J;
                    M.x0 = px; M.y0 = py; M.x1 = lsx; M.y1 = lsy; M.x2 = rsx; M.y2 = rsy; M.rt = r_t; M.rs = r_s;
                    M.vx = vx; M.vy = vy; M.ox = me[12]; M.oy = me[13];
                    O.x0 = o[0]; O.y0 = o[1]; O.x1 = o[8]; O.y1 = o[9]; O.x2 = o[10]; O.y2 = o[11]; O.rt = o[5]; O.rs = o[6];
                    O.vx = o[2]; O.vy = o[3]; O.ox = o[12]; O.oy = o[13];
                    sel_three(me_is_i, M, O, I);
                    sel_three(me_is_i, O, M, J);
                    const ThreePar par = {mass, k_soc, tau_0, mu, kappa, damping};
                    pair_three_exact(I, J, me_is_i, par, qx, qy, qt);
                }
            }
        }
        for (int o = G >> 1; o; o >>= 1) {          // butterfly inside the group: identical totals in all its threads
            qx += __shfl_xor_sync(0xffffffffu, qx, o); qy += __shfl_xor_sync(0xffffffffu, qy, o);
            if (MODEL == 1) qt += __shfl_xor_sync(0xffffffffu, qt, o);
        }
        fx += qx; fy += qy; tq += qt;
        if (live && (A.flags & CDB_STEP_AGENT_OBSTACLE) && A.n_obs > 0) {
            if (MODEL == 0) walls_circular(px, py, radius, vx, vy, ContactValues{mu, kappa, damping}, A.obs, A.n_obs, fx, fy);
            else walls_three_circle(px, py, lsx, lsy, rsx, rsy, r_t, r_s, vx, vy, ContactValues{mu, kappa, damping}, A.obs, A.n_obs, fx, fy, tq);
        }
        // ---- adaptive dt: the two maxima of integrator.py:81-90 over the block ----------------------------------------------
        if (A.flags & CDB_STEP_INTEGRATOR) {
            const double ninf = -__longlong_as_double(0x7ff0000000000000LL);
            const double sp = live ? hypot(vx, vy) : 0.0;
            unsigned long long vm = ordered_bits(sp > 0.0 ? sp : 0.0);          // NaN speeds are skipped, as in the reference loop
            unsigned long long vb = !live ? ordered_bits(ninf) : (isnan(v0) ? 0xffffffffffffffffULL : ordered_bits(v0));
            vm = warp_max_u64(vm); vb = warp_max_u64(vb);
            if ((t & 31) == 0) { s_vm[t >> 5] = vm; s_v0[t >> 5] = vb; }
            __syncthreads();
            if (t == 0) {
                unsigned long long m[2] = {ordered_bits(0.0), ordered_bits(ninf)};
                for (int k = 0; k < (int)(blockDim.x >> 5); ++k) { m[0] = s_vm[k] > m[0] ? s_vm[k] : m[0]; m[1] = s_v0[k] > m[1] ? s_v0[k] : m[1]; }
                const double dt = adaptive_timestep(m, A.dt_min, A.dt_max);
                s_dt = dt;
                A.dt_out[0] = dt; A.dt_out[1] += dt;
                if (A.dt_log) A.dt_log[step % DT_LOG_SLOTS] = dt;
            }
            __syncthreads();
            const double dt = s_dt;
            verlet(fx, fpx, mass, dt, vx, px);
            verlet(fy, fpy, mass, dt, vy, py);
            fpx = fx; fpy = fy;
            if (MODEL == 1) {
                verlet(tq, tq_prev, inertia, dt, w, phi);
                phi = wrap_to_pi(phi);
                tq_prev = tq;
                const double ox = sin(phi) * r_ts, oy = -cos(phi) * r_ts;       // shoulders(), agents.py:473-486
                lsx = px - ox; lsy = py - oy; rsx = px + ox; rsy = py + oy;
            }
        } else {
            __syncthreads();        // the records are rewritten at the top of the next step
        }
        if (rst) { fx = 0.0; fy = 0.0; tq = 0.0; }
    }
    if (t == 0) *A.step_ptr = step;
    if (!live || sub != 0) return;
    const int w_ = ag;
    // ---- the state, once ---------------------------------------------------------------------------------------------------
    s(PX, w_) = px; s(PY, w_) = py; s(VX, w_) = vx; s(VY, w_) = vy; s(E0X, w_) = e0x; s(E0Y, w_) = e0y;
    s(FX, w_) = fx; s(FY, w_) = fy; s(FPX, w_) = fpx; s(FPY, w_) = fpy;
    if (MODEL == 1) {
        s(LSX, w_) = lsx; s(LSY, w_) = lsy; s(RSX, w_) = rsx; s(RSY, w_) = rsy;
        s(PHI, w_) = phi; s(OMEGA, w_) = w; s(PHI0, w_) = phi0; s(TORQUE, w_) = tq; s(TORQUE_PREV, w_) = tq_prev;
    }
}
