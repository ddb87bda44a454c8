// step_kernel.cuh -- the fused per-agent step kernel (navigation sample, orientation, adjusting, agent-agent,
// agent-obstacle, adaptive-dt velocity Verlet, reset) over cell-sorted SoA state.
//
// Agent-agent part, one thread per cell-sorted agent ("me"), two phases per warp:
//   phase 1 (classify)  sweep the three neighbour cell-columns; for every candidate decide with a handful of fp64
//                       operations and NO sqrt/div/hypot whether the pair can contribute a force at all:
//                         - conservative sight gate     d^2 <= (3 + r_tot)^2 (1 + eps)
//                         - the reference's own time-to-collision quantities a, b, c, disc = b*b - a*c, computed with the
//                           reference's exact operation order, so "disc > 0 and b > 0" is EXACTLY the necessary condition for
//                           its social force to be non-zero (power_law.py:236-246), or
//                         - conservative contact test   d^2 <= r_tot^2 (1 + eps)
//                       survivors (about 3-5 of ~115 candidates at 1 agent/m^2) are appended to a per-lane list in shared
//                       memory;
//   phase 2 (evaluate)  every lane runs the exact reference pair arithmetic (gate with hypot, tau, gradient, exp, truncation,
//                       contact, torque) on its short list.
// A pair dropped in phase 1 contributes exactly (0, 0) in the reference as well, so results are identical to evaluating
// every candidate; the expensive, divergent part runs on ~4 % of the candidates instead of all of them.
#pragma once
#include "kernels.cuh"

constexpr int STEP_THREADS = 128;
constexpr int LCAP = 32;     // survivor list entries per lane
constexpr int CHUNK = 8;     // candidates classified between two list-capacity checks
#define PREFILTER_EPS 1e-12

struct StepArgs {
    Soa in, out;             // out == in unless integrating (then the new state is written to the other buffer)
    int n;
    const Grid *grid;
    const int *cell_sorted, *cell_start, *cell_count;
    const NavField *nav;
    int n_nav;
    const double *obs;
    int n_obs;
    unsigned flags;
    double dt_min, dt_max;
    const unsigned long long *vmax;
    double *dt_out;          // [0] dt, [1] time_tot
    double *dt_log;          // where to log this step's dt (or nullptr)
};

// ---- exact three-circle pair in the reference's (i, j) orientation, register-only (no dynamic indexing) ---------------
struct Three {               // kinematics of one three-circle agent as the pair kernels need them
    double x[3], y[3];       // torso, left shoulder, right shoulder centres
    double rt, rs;           // torso / shoulder radius
    double vx, vy;
    double ox, oy;           // r_ts * (sin(phi), -cos(phi)): shoulder displacement (power_law.py:338-350, agents.py:483-484)
};

__device__ __forceinline__ void pair_three_exact(const Three &I, const Three &J, bool me_is_i, const ThreePar &me,
                                                 double &fx, double &fy, double &torque) {
    const double ri[3] = {I.rt, I.rs, I.rs}, rj[3] = {J.rt, J.rs, J.rs};
    // distance_three_circles (distance.py:55-105): strict '<', first wins, order torso, left, right
    double h_min = nan(""), sx = 0.0, sy = 0.0, sd = 0.0;   // selected x, y, d (normal = x / d computed once, same value)
    double mix = 0.0, miy = 0.0, mir = 0.0;                  // x0[i_min], r0[i_min]
    double qx = 0.0, qy = 0.0, mjr = 0.0;                    // x0[j_min] (the :103 quirk), r1[j_min]
#pragma unroll
    for (int pi = 0; pi < 3; ++pi)
#pragma unroll
        for (int pj = 0; pj < 3; ++pj) {
            double x = I.x[pi] - J.x[pj], y = I.y[pi] - J.y[pj];
            double d = hypot(x, y);
            double h = d - (ri[pi] + rj[pj]);
            if (h < h_min || isnan(h_min)) {
                h_min = h; sx = x; sy = y; sd = d;
                mix = I.x[pi]; miy = I.y[pi]; mir = ri[pi];
                qx = I.x[pj]; qy = I.y[pj]; mjr = rj[pj];
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
        double tau = nan(""), num_sel = 0.0, b_min = 0.0, d_min = 0.0, oix = 0.0, oiy = 0.0, ojx = 0.0, ojy = 0.0;
#pragma unroll
        for (int pi = 0; pi < 3; ++pi)
#pragma unroll
            for (int pj = 0; pj < 3; ++pj) {
                double x = I.x[pi] - J.x[pj], y = I.y[pi] - J.y[pj];
                double r_tot = ri[pi] + rj[pj];
                double b = -(x * vx + y * vy);
                double c = (x * x + y * y) - r_tot * r_tot;
                double disc = b * b - a * c;
                if (!(disc > 0.0)) continue;         // sqrt gives NaN (disc < 0 or NaN) or 0
                double dd = sqrt(disc);
                double num = b - dd;
                bool take;
                double tau_new;
                if (isnan(tau)) { tau_new = num / a; take = true; }
                else if (num > 0.0 && num < num_sel) { tau_new = num / a; take = 0.0 < tau_new && tau_new < tau; }
                else { tau_new = 0.0; take = false; }
                if (take) {
                    tau = tau_new; num_sel = num; b_min = b; d_min = dd;
                    // shoulder displacement of the contacting parts: 0 for the torso, +o for left, -o for right
                    oix = pi == 0 ? 0.0 : (pi == 1 ? I.ox : 0.0 - I.ox);
                    oiy = pi == 0 ? 0.0 : (pi == 1 ? I.oy : 0.0 - I.oy);
                    ojx = pj == 0 ? 0.0 : (pj == 1 ? J.ox : 0.0 - J.ox);
                    ojy = pj == 0 ? 0.0 : (pj == 1 ? J.oy : 0.0 - J.oy);
                }
            }
        if (!(isnan(tau) || tau <= 0.0)) {
            double xr = I.x[0] - J.x[0], yr = I.y[0] - J.y[0];
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
    if (me_is_i) { mx = mix + mir * nx - I.x[0]; my = miy + mir * ny - I.y[0]; }
    else { mx = qx - mjr * nx - J.x[0]; my = qy - mjr * ny - J.y[0]; }
    fx += fsx; fy += fsy;
    torque += mx * fsy - my * fsx;
}

__device__ __forceinline__ void load_three(const Soa &s, int ox_plane, int u, Three &k) {
    k.x[0] = __ldg(&s(PX, u)); k.y[0] = __ldg(&s(PY, u));
    k.x[1] = __ldg(&s(LSX, u)); k.y[1] = __ldg(&s(LSY, u));
    k.x[2] = __ldg(&s(RSX, u)); k.y[2] = __ldg(&s(RSY, u));
    k.rt = __ldg(&s(R_T, u)); k.rs = __ldg(&s(R_S, u));
    k.vx = __ldg(&s(VX, u)); k.vy = __ldg(&s(VY, u));
    k.ox = __ldg(&s(ox_plane, u)); k.oy = __ldg(&s(ox_plane + 1, u));
}

// derived planes of the three-circle model (filled by k_gather every step)
enum { D_OX = NP_THREE, D_OY, D_EXT, NP_THREE_ALL };

// ---- phase-1 classifiers ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool classify_circular(const CircMe &me, const Soa &s, int u) {
    const double x = me.px - __ldg(&s(PX, u)), y = me.py - __ldg(&s(PY, u));
    const double r_tot = me.r + __ldg(&s(RADIUS, u));
    const double d2 = x * x + y * y;
    const double lim = SIGTH_SOC + r_tot;
    if (!(d2 <= lim * lim * (1.0 + PREFILTER_EPS))) return false;
    const double vx = me.vx - __ldg(&s(VX, u)), vy = me.vy - __ldg(&s(VY, u));
    const double a = vx * vx + vy * vy;
    const double b = -(x * vx + y * vy);
    const double rr = r_tot * r_tot;
    const double c = d2 - rr;
    const double disc = b * b - a * c;
    return (disc > 0.0 && b > 0.0) || d2 <= rr * (1.0 + PREFILTER_EPS);
}

__device__ __forceinline__ bool classify_three(const Three &me, double me_ext, bool me_is_i, const Soa &s, int u) {
    const double x = me.x[0] - __ldg(&s(PX, u)), y = me.y[0] - __ldg(&s(PY, u));
    const double e_tot = me_ext + __ldg(&s(D_EXT, u));
    const double d2 = x * x + y * y;
    const double lim = SIGTH_SOC + e_tot;
    if (!(d2 <= lim * lim * (1.0 + PREFILTER_EPS))) return false;
    if (d2 <= e_tot * e_tot * (1.0 + PREFILTER_EPS)) return true;          // may touch: contact branch possible
    const double vx = me.vx - __ldg(&s(VX, u)), vy = me.vy - __ldg(&s(VY, u));
    const double a = vx * vx + vy * vy;
    if (a == 0.0) return false;
    const double ux[3] = {__ldg(&s(PX, u)), __ldg(&s(LSX, u)), __ldg(&s(RSX, u))};
    const double uy[3] = {__ldg(&s(PY, u)), __ldg(&s(LSY, u)), __ldg(&s(RSY, u))};
    const double urt = __ldg(&s(R_T, u)), urs = __ldg(&s(R_S, u));
    const double mr[3] = {me.rt, me.rs, me.rs}, ur[3] = {urt, urs, urs};
    // The reference's loop (power_law.py:308-329) lets the FIRST part pair with a real, non-zero discriminant fix the sign
    // of tau for good, so the social force is non-zero only if that first pair has b - d > 0, for which b > 0 is necessary.
    // (x_rel, v_rel) -> (-x_rel, -v_rel) leaves b, c, disc bitwise unchanged, so only the enumeration order depends on who is i.
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const int hi = k / 3, lo = k % 3;
        const double mx = me_is_i ? me.x[hi] : me.x[lo], my = me_is_i ? me.y[hi] : me.y[lo];
        const double mrr = me_is_i ? mr[hi] : mr[lo];
        const double oxx = me_is_i ? ux[lo] : ux[hi], oyy = me_is_i ? uy[lo] : uy[hi];
        const double orr = me_is_i ? ur[lo] : ur[hi];
        const double xr = mx - oxx, yr = my - oyy;
        const double r_tot = me_is_i ? (mrr + orr) : (orr + mrr);
        const double b = -(xr * vx + yr * vy);
        const double c = (xr * xr + yr * yr) - r_tot * r_tot;
        const double disc = b * b - a * c;
        if (disc > 0.0) return b > 0.0;
    }
    return false;
}

// =====================================================================================================================
template <int MODEL>
__global__ void __launch_bounds__(STEP_THREADS) k_step(const StepArgs A) {
    extern __shared__ int s_list[];
    const int t = blockIdx.x * STEP_THREADS + threadIdx.x;
    const int lane = threadIdx.x & 31;
    int *list = s_list + (threadIdx.x >> 5) * (LCAP * 32) + lane;   // entry k of this lane at list[k * 32]
    const bool active = t < A.n;
    const Soa &s = A.in;
    const unsigned FULL = 0xffffffffu;
    const int tt = active ? t : 0;     // inactive lanes read slot 0 and never write

    // ---- own state -----------------------------------------------------------------------------------------------------
    const double px = s(PX, tt), py = s(PY, tt), vx = s(VX, tt), vy = s(VY, tt);
    const double mass = s(MASS, tt);
    const ThreePar par = {mass, s(K_SOC, tt), s(TAU_0, tt), s(MU, tt), s(KAPPA, tt), s(DAMPING, tt)};
    double e0x = s(E0X, tt), e0y = s(E0Y, tt);
    double fx = s(FX, tt), fy = s(FY, tt), tq = 0.0, phi0 = 0.0;
    CircMe cme;
    Three tme;
    double t_ext = 0.0;
    if (MODEL == 0) {
        cme = CircMe{px, py, vx, vy, s(RADIUS, tt), mass, par.k_soc, par.tau_0, par.mu, par.kappa, par.damping};
    } else {
        load_three(s, D_OX, tt, tme);
        t_ext = s(D_EXT, tt);
        tq = s(TORQUE, tt);
        phi0 = s(PHI0, tt);
    }

    // ---- Navigation, Orientation, Adjusting (logic.py:149-165,258-261,89-94) --------------------------------------------
    if (A.flags & CDB_STEP_NAVIGATION) navigation_sample(A.nav, A.n_nav, s.target[tt], px, py, e0x, e0y);
    if (MODEL == 1 && (A.flags & CDB_STEP_ORIENTATION)) phi0 = atan2(e0y, e0x);
    if (A.flags & CDB_STEP_ADJUSTING) {
        double ax, ay;
        adjust_force(mass, s(TAU_ADJ, tt), s(V0, tt), e0x, e0y, vx, vy, ax, ay);
        fx += ax; fy += ay;
        if (MODEL == 1) tq += adjust_torque(s(INERTIA, tt), s(TAU_ROT, tt), phi0, s(PHI, tt), s(OMEGA0, tt), s(OMEGA, tt));
    }

    // ---- AgentAgentInteractions (interactions.py:191-205) -----------------------------------------------------------------
    if (A.flags & CDB_STEP_AGENT_AGENT) {
        const Grid g = *A.grid;
        const int ny = (int)g.ny, nxg = (int)g.nx;
        const int c = A.cell_sorted[tt];
        const int cx = c / ny, cy = c - cx * ny;
        const int ylo = cy > 0 ? cy - 1 : 0, yhi = cy + 1 < ny ? cy + 1 : ny - 1;
        int cnt = 0;

        auto flush = [&]() {
            const int m = __reduce_max_sync(FULL, cnt);
            for (int k = 0; k < m; ++k) {
                if (k < cnt) {
                    const int u = list[k * 32];
                    if (MODEL == 0) {
                        pair_circular(cme, __ldg(&s(PX, u)), __ldg(&s(PY, u)), __ldg(&s(VX, u)), __ldg(&s(VY, u)),
                                      __ldg(&s(RADIUS, u)), fx, fy);
                    } else {
                        Three other;
                        load_three(s, D_OX, u, other);
                        // reference pair orientation: i is the lexicographically smaller (cell_x, cell_y, agent index),
                        // which is exactly the order of the cell-sorted slots
                        if (t < u) pair_three_exact(tme, other, true, par, fx, fy, tq);
                        else pair_three_exact(other, tme, false, par, fx, fy, tq);
                    }
                }
            }
            cnt = 0;
        };

#pragma unroll 1
        for (int dx = -1; dx <= 1; ++dx) {
            const int x2 = cx + dx;
            int b = 0, e = 0;
            if (active && x2 >= 0 && x2 < nxg) {
                b = A.cell_start[x2 * ny + ylo];
                e = A.cell_start[x2 * ny + yhi] + A.cell_count[x2 * ny + yhi];
            }
            const int maxlen = __reduce_max_sync(FULL, e - b);
#pragma unroll 1
            for (int k0 = 0; k0 < maxlen; k0 += CHUNK) {
                if (__any_sync(FULL, cnt > LCAP - CHUNK)) flush();
#pragma unroll
                for (int kk = 0; kk < CHUNK; ++kk) {
                    const int u = b + k0 + kk;
                    if (u < e && u != t) {
                        const bool keep = MODEL == 0 ? classify_circular(cme, s, u) : classify_three(tme, t_ext, t < u, s, u);
                        if (keep) { list[cnt * 32] = u; ++cnt; }
                    }
                }
            }
        }
        flush();
    }
    if (!active) return;

    // ---- AgentObstacleInteractions (interactions.py:208-214) ---------------------------------------------------------------
    if ((A.flags & CDB_STEP_AGENT_OBSTACLE) && A.n_obs > 0) {
        if (MODEL == 0) {
            walls_circular(px, py, cme.r, vx, vy, par.mu, par.kappa, par.damping, A.obs, A.n_obs, fx, fy);
        } else {
            const double x[3][2] = {{tme.x[0], tme.y[0]}, {tme.x[1], tme.y[1]}, {tme.x[2], tme.y[2]}};
            const double r[3] = {tme.rt, tme.rs, tme.rs};
            walls_three_circle(x, r, vx, vy, par.mu, par.kappa, par.damping, A.obs, A.n_obs, fx, fy, tq);
        }
    }

    const Soa &o = A.out;
    if (!(A.flags & CDB_STEP_INTEGRATOR)) {
        // node-wise use: publish what the selected nodes wrote, in place
        o(E0X, t) = e0x; o(E0Y, t) = e0y;
        const bool rst = A.flags & CDB_STEP_RESET;
        o(FX, t) = rst ? 0.0 : fx; o(FY, t) = rst ? 0.0 : fy;
        if (MODEL == 1) { o(PHI0, t) = phi0; o(TORQUE, t) = rst ? 0.0 : tq; }
        return;
    }

    // ---- Integrator (integrator.py:209-256) + Reset (logic.py:59-64); new state goes to the other buffer --------------------
    const double dt = adaptive_timestep(A.vmax, A.dt_min, A.dt_max);
    if (t == 0) {
        A.dt_out[0] = dt; A.dt_out[1] += dt;
        if (A.dt_log) *A.dt_log = dt;
    }
    double nvx = vx, nvy = vy, npx = px, npy = py;
    verlet(fx, s(FPX, t), mass, dt, nvx, npx);
    verlet(fy, s(FPY, t), mass, dt, nvy, npy);
    const bool rst = A.flags & CDB_STEP_RESET;
    o(PX, t) = npx; o(PY, t) = npy; o(VX, t) = nvx; o(VY, t) = nvy;
    o(E0X, t) = e0x; o(E0Y, t) = e0y;
    o(FX, t) = rst ? 0.0 : fx; o(FY, t) = rst ? 0.0 : fy;
    o(FPX, t) = fx; o(FPY, t) = fy;
    o(RADIUS, t) = s(RADIUS, t); o(MASS, t) = mass; o(V0, t) = s(V0, t); o(TAU_ADJ, t) = s(TAU_ADJ, t);
    o(K_SOC, t) = par.k_soc; o(TAU_0, t) = par.tau_0; o(MU, t) = par.mu; o(KAPPA, t) = par.kappa; o(DAMPING, t) = par.damping;
    o.id[t] = s.id[t];
    o.target[t] = s.target[t];
    if (MODEL == 1) {
        const double inertia = s(INERTIA, t), r_ts = s(R_TS, t);
        double w = s(OMEGA, t), phi = s(PHI, t);
        verlet(tq, s(TORQUE_PREV, t), inertia, dt, w, phi);
        phi = wrap_to_pi(phi);
        const double ox = sin(phi) * r_ts, oy = -cos(phi) * r_ts;   // shoulders(), agents.py:473-486
        o(LSX, t) = npx - ox; o(LSY, t) = npy - oy; o(RSX, t) = npx + ox; o(RSY, t) = npy + oy;
        o(R_T, t) = tme.rt; o(R_S, t) = tme.rs; o(R_TS, t) = r_ts; o(INERTIA, t) = inertia; o(OMEGA0, t) = s(OMEGA0, t);
        o(PHI, t) = phi; o(OMEGA, t) = w; o(PHI0, t) = phi0;
        o(TORQUE, t) = rst ? 0.0 : tq; o(TORQUE_PREV, t) = tq; o(TAU_ROT, t) = s(TAU_ROT, t);
    }
}
