// pair_kernels.cuh -- agent-agent interactions evaluated ONCE per unordered pair, as the reference does
// (core/interactions.py:53-70,75-104 update both agents of a pair from one evaluation):
//
//   k_sweep      (classify)  one thread per cell-sorted agent ("target") sweeps only its FORWARD half stencil -- the rest of
//                            its own cell column range and the next cell column(s) -- so every unordered pair of adjacent
//                            cells is tested exactly once (the block list's own pair enumeration, core/block_list.py:28-52 /
//                            cell_lists.iter_nearest_neighbors).  The ~25-operation branch-free fp64 test of step_kernel.cuh
//                            decides whether the pair can contribute a force at all; survivors are compacted warp-wide
//                            (ballot + popc) into a staging buffer in shared memory and appended, coalesced, to a global pair
//                            list.  Per-agent counts of listed pairs are kept with integer atomics.
//   k_pair_alloc             gives every agent a private region of the contribution array (warp-aggregated allocation; the
//                            regions need not be ordered, only disjoint).
//   k_pair_eval  (evaluate)  one thread per listed pair runs the exact reference arithmetic once, in the reference's (i, j)
//                            orientation, and produces BOTH agents' force / torque -- bit-identical to evaluating the pair
//                            from each side (circular: x -> -x is an exact negation; three-circle: the shared quantities
//                            h_min, tau, gradient are the same numbers for both sides).  Each side's result goes to that
//                            agent's region, tagged with the partner's position in cell order.
//   k_step<M, 1> (gather)    adds an agent's contributions in ascending partner order -- the order in which the one-kernel
//                            variant (step_kernel.cuh, PAIRS = 0) meets them -- so results are run-to-run reproducible and
//                            bit-identical to that variant, whatever order the atomics resolved in.
//
// The pair list has a fixed capacity.  If a step finds more pairs than fit, NOTHING of that step is applied (k_pair_eval and
// the gather kernel see ctr[0] > cap and leave the state untouched, the device step counter does not advance); the host
// notices at its next synchronisation point, grows the buffers and re-issues the missing steps (crowd_b200.cu: settle()).
#pragma once
#include "kernels.cuh"

constexpr int SW_THREADS = 128;
constexpr int SW_WARPS = SW_THREADS / 32;
constexpr int SW_CHUNK = 4;                  // candidates classified per lane between two staging-capacity checks
constexpr int SW_STAGE = 256;                // staged pairs per warp
constexpr int GHOST_KEY_SHIFT = 1 << 30;     // ghosts of the left neighbour strip sort before every owned agent

#define PREFILTER_EPS 1e-12
#define BOUND_EPS 1e-9

// packed neighbour records (doubles per agent)
constexpr int REC_CIRC = 6;     // px py vx vy r -
constexpr int REC_THREE = 16;   // px py vx vy ext rt rs id | lsx lsy rsx rsy ox oy cell_x cell_y

struct PairBuf {
    int2 *pairs;                 // [cap] (target slot, candidate slot), target < candidate in cell order
    double *cres;                // [2 * cap][4] contributions {partner key (int64 bits), fx, fy, torque}
    int *cnt;                    // per slot: listed pairs the agent takes part in
    int *off;                    // per slot: first entry of the agent's region of cres
    int *fill;                   // per slot: entries written so far (== contributions to add)
    unsigned long long *ctr;     // [0] pairs found this step (may exceed cap => step not applied), [1] entries allocated
    long long cap;
    int *fatal;                  // strips: an unapplied step cannot be repeated (the other ranks went on) => device error flag
};

__device__ __forceinline__ bool pairs_overflowed(const PairBuf &pb) { return pb.ctr[0] > (unsigned long long)pb.cap; }

struct SweepArgs {
    const double *nbr_sweep;     // 48 B sweep records {px, py, vx, vy, R, -} in cell order (+ ghost tail)
    int n;                       // owned targets (host-side bound)
    const int *n_dev;            // device-side exact count (nullptr: n is exact)
    const Grid *grid;
    const int *cell_sorted, *cell_start, *cell_count;
    int ghost_base;              // first slot of the left ghost columns (strips; targets n .. n + n_ghost map there), or -1
    int n_ghost;                 // launch bound for the left ghost targets
    int ghost_cells;             // cells of the left ghost block (its columns come first in the lattice)
    int reach;                   // forward columns / rows swept around the target's cell: 1 on the cell_size lattice, 2 on the
                                 // twice finer search lattice (chosen by the host when 3 + 2 max R < cell_size, see build_block_list)
    PairBuf pb;
    const ChainState *chain;     // resident-order steps: drift bound since the block list was built (nullptr: list is fresh)
    double drift_limit;          // the sweep is complete while chain->disp_acc <= drift_limit (half the slack of the cells)
};

__device__ __forceinline__ double2 ldg2(const double *p) { return __ldg(reinterpret_cast<const double2 *>(p)); }

#ifndef PAIR_PREFETCH
#define PAIR_PREFETCH 1      // k_pair_eval<three_circle>: 0 none, 1 next pair's records -> L1, 2 -> L2 (measured equal to L1)
#endif
__device__ __forceinline__ void prefetch_line(const void *p) {
#if PAIR_PREFETCH == 2
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#endif
}

// The classification of one candidate: can the pair exert a force at all?  The stored radii / body extents are inflated by
// SWEEP_EPS (slot 5 of the sweep record), which makes every comparison conservative under any rounding: a pair the exact
// arithmetic of k_pair_eval would give a force is never dropped.  Fused multiply-adds are fine here (this is a filter, not
// the reference arithmetic): ~24 fp64 operations, branch-free.
template <int MODEL>
__device__ __forceinline__ bool sweep_keep(double mpx, double mpy, double mvx, double mvy, double mr, double lim_t,
                                           const double2 p, const double2 v, double ro) {
    const double x = mpx - p.x, y = mpy - p.y;
    const double d2 = fma(x, x, y * y);
    const double lim = lim_t + ro;               // >= (3 + R)(1 + eps): conservative form of h < SIGTH_SOC
    const bool gate = d2 <= lim * lim;
    const double R = mr + ro;                    // inflated r_tot (circular) / sum of body extents (three-circle)
    const double RR = R * R;
    const double cc = d2 - RR;
    const bool contact = cc <= 1e-9;             // conservative form of h < 0
    const double vx = mvx - v.x, vy = mvy - v.y;
    const double a = fma(vx, vx, vy * vy);
    const double bb = -fma(x, vx, y * vy);
    bool social;
    if (MODEL == 0) {
        // necessary for a non-zero social force (power_law.py:236-246): a real time-to-collision (b^2 - a c > 0) that is
        // positive (b > 0); margins cover the different rounding of the reference's own evaluation order
        social = bb > -1e-12 && fma(bb * (1.0 + PREFILTER_EPS), bb, -(a * cc)) > 0.0;
    } else {
        // bounding circles: no real root for them => none for any of the 9 part pairs (power_law.py:308-329);
        // all part pairs receding (b_k <= b + R |v| <= 0) => no positive time-to-collision
        social = fma(bb, bb, -(a * cc)) >= 0.0 && (bb >= -1e-12 || bb * bb <= RR * a);
    }
    return gate && (social || contact);
}

template <int MODEL>
__global__ void __launch_bounds__(SW_THREADS, 8) k_sweep(const SweepArgs A) {
    __shared__ int2 s_stage[SW_WARPS][SW_STAGE];
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    int2 *stage = s_stage[threadIdx.x >> 5];
    const int g = blockIdx.x * SW_THREADS + threadIdx.x;
    const int n_own = eff_n(A.n, A.n_dev);
    const int ny = (int)A.grid->ny, nxg = (int)A.grid->nx;
    if (A.chain && !(A.chain->disp_acc <= A.drift_limit)) {
        // some agent may have drifted out of the cells this sweep relies on: the step is NOT applied (same protocol as a
        // pair list that is too small -- every later kernel of the step sees the counter above the capacity); the host
        // rebuilds the block list and repeats it
        if (g == 0) A.pb.ctr[0] = CHAIN_STALE;
        return;
    }
    // targets: owned agents, then (strips) the ghosts of the left neighbour, which only pair with owned candidates
    int t = -1;
    bool ghost = false;
    if (g < A.n) { if (g < n_own) t = g; }
    else if (A.ghost_base >= 0 && A.ghost_cells > 0) {      // no left neighbour: no ghost targets (and no cell -1 to look at)
        const int k = g - A.n;
        const int n_gl = A.cell_start[A.ghost_cells - 1] + A.cell_count[A.ghost_cells - 1] - A.ghost_base;
        if (k < n_gl) { t = A.ghost_base + k; ghost = true; }
    }
    const bool active = t >= 0;
    const int tt = active ? t : 0;
    // the stored radii / body extents are inflated by SWEEP_EPS (slot 5 of the sweep record), which makes every comparison
    // below conservative under any rounding: a pair the exact arithmetic of k_pair_eval would give a force is never dropped
    constexpr double SWEEP_EPS = MODEL == 0 ? PREFILTER_EPS : BOUND_EPS;
    double mpx, mpy, mvx, mvy, mr, lim_t;
    {
        const double *r = A.nbr_sweep + (size_t)tt * REC_CIRC;
        const double2 p = ldg2(r), v = ldg2(r + 2);
        mpx = p.x; mpy = p.y; mvx = v.x; mvy = v.y; mr = __ldg(r + 5);
        lim_t = SIGTH_SOC * (1.0 + SWEEP_EPS) + mr;
    }
    const int c = A.cell_sorted[tt];
    const int cx = c / ny, cy = c - cx * ny;
    const int reach = A.reach;
    const int ylo = max(cy - reach, 0), yhi = min(cy + reach, ny - 1);
    const int col_min = (int)A.grid->cx_lo;      // ghost targets: candidates in owned columns only
    int nst = 0, mine = 0;
    const int own_limit = A.ghost_base >= 0 ? A.ghost_base : 0x7fffffff;   // slots beyond are ghosts: nothing is stored for them

    auto flush = [&]() {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(&A.pb.ctr[0], (unsigned long long)nst);
        base = __shfl_sync(FULL, base, 0);
        __syncwarp();
        for (int i = lane; i < nst; i += 32) {
            const unsigned long long p = base + i;
            if (p < (unsigned long long)A.pb.cap) {
                const int2 e = stage[i];
                A.pb.pairs[p] = e;
                if (e.y < own_limit) atomicAdd(&A.pb.cnt[e.y], 1);
            }
        }
        __syncwarp();
        nst = 0;
    };

    for (int dx = 0; dx <= reach; ++dx) {
        const int x2 = cx + dx;
        int b = 0, e = 0;
        if (active && x2 < nxg && !(ghost && x2 < col_min)) {
            if (dx == 0) { b = t + 1; e = A.cell_start[x2 * ny + yhi] + A.cell_count[x2 * ny + yhi]; }
            else { b = A.cell_start[x2 * ny + ylo]; e = A.cell_start[x2 * ny + yhi] + A.cell_count[x2 * ny + yhi]; }
        }
        const int maxlen = __reduce_max_sync(FULL, e - b);
        for (int k0 = 0; k0 < maxlen; k0 += SW_CHUNK) {
            bool keep[SW_CHUNK];
#pragma unroll
            for (int kk = 0; kk < SW_CHUNK; ++kk) {
                const int u = b + k0 + kk;
                const bool inr = u < e;
                const double *r = A.nbr_sweep + (size_t)(inr ? u : tt) * REC_CIRC;
                const double2 p = ldg2(r), v = ldg2(r + 2);
                const double ro = __ldg(r + 5);
                keep[kk] = inr && sweep_keep<MODEL>(mpx, mpy, mvx, mvy, mr, lim_t, p, v, ro);
            }
#pragma unroll
            for (int kk = 0; kk < SW_CHUNK; ++kk) {
                const unsigned m = __ballot_sync(FULL, keep[kk]);
                if (keep[kk]) { stage[nst + __popc(m & ((1u << lane) - 1u))] = make_int2(t, b + k0 + kk); ++mine; }
                nst += __popc(m);
            }
            if (nst > SW_STAGE - 32 * SW_CHUNK) flush();
        }
    }
    if (nst) flush();
    if (mine && !ghost) atomicAdd(&A.pb.cnt[t], mine);
}

// ---- k_sweep_staged: the same classification with the candidates staged in shared memory by TMA bulk copies ------------------
// The 128 targets of a CTA are consecutive slots of the cell order, so for each forward column dx the union of their candidate
// ranges is (the hull of) ONE contiguous range of sweep records, ~128 + a few cells' worth: three `cp.async.bulk` copies
// (global -> shared, completion on an mbarrier) fetch everything the CTA will classify, and the inner loop reads LDS
// instead of waiting on L1 / L2 (ncu, round 2: `long_scoreboard` was the top stall of k_sweep at 5.2 warps per issue).  A
// column whose hull does not fit SWS_CAP records (strong density gradients) is swept from global memory as before; the
// choice is CTA-uniform.  Single device only: ghost targets of strips keep the plain kernel.
#ifndef SWS_CAP
#define SWS_CAP 192                          // staged records per forward column (48 B each)
#endif
#ifndef SWS_MINB
#define SWS_MINB 6                           // 8 KB pair staging + 3 * SWS_CAP * 48 B of records per CTA
#endif
constexpr int SWS_COLS = 3;                  // forward columns 0 .. reach, reach <= 2

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// 1-D bulk copy global -> shared (TMA unit; 16-byte aligned addresses and size), completion counted on `bar`
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long *bar, unsigned parity) {
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}

template <int MODEL>
__global__ void __launch_bounds__(SW_THREADS, SWS_MINB) k_sweep_staged(const SweepArgs A) {
    __shared__ int2 s_stage[SW_WARPS][SW_STAGE];
    __shared__ __align__(128) double s_rec[SWS_COLS][SWS_CAP * REC_CIRC];
    __shared__ __align__(8) unsigned long long s_bar[SWS_COLS];
    __shared__ int s_lo[SWS_COLS][SW_WARPS], s_hi[SWS_COLS][SW_WARPS];
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int2 *stage = s_stage[warp];
    const int g = blockIdx.x * SW_THREADS + threadIdx.x;
    const int n_own = eff_n(A.n, A.n_dev);
    const int ny = (int)A.grid->ny, nxg = (int)A.grid->nx;
    if (A.chain && !(A.chain->disp_acc <= A.drift_limit)) {      // stale search lattice: step not applied (see k_sweep)
        if (g == 0) A.pb.ctr[0] = CHAIN_STALE;
        return;
    }
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < SWS_COLS; ++k) mbar_init(&s_bar[k], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const bool active = g < n_own;
    const int t = active ? g : -1, tt = active ? g : 0;
    constexpr double SWEEP_EPS = MODEL == 0 ? PREFILTER_EPS : BOUND_EPS;
    double mpx, mpy, mvx, mvy, mr, lim_t;
    {
        const double *r = A.nbr_sweep + (size_t)tt * REC_CIRC;
        const double2 p = ldg2(r), v = ldg2(r + 2);
        mpx = p.x; mpy = p.y; mvx = v.x; mvy = v.y; mr = __ldg(r + 5);
        lim_t = SIGTH_SOC * (1.0 + SWEEP_EPS) + mr;
    }
    const int c = A.cell_sorted[tt];
    const int cx = c / ny, cy = c - cx * ny;
    const int reach = A.reach;
    const int ylo = max(cy - reach, 0), yhi = min(cy + reach, ny - 1);
    // candidate range of this target in forward column dx (empty: b == e)
    auto range = [&](int dx, int &b, int &e) {
        b = 0; e = 0;
        const int x2 = cx + dx;
        if (active && dx <= reach && x2 < nxg) {
            const int last = x2 * ny + yhi;
            e = A.cell_start[last] + A.cell_count[last];
            b = dx == 0 ? t + 1 : A.cell_start[x2 * ny + ylo];
        }
    };
    // ---- hull of the CTA's ranges per column, then one bulk copy per column -------------------------------------------------
#pragma unroll
    for (int dx = 0; dx < SWS_COLS; ++dx) {
        int b, e;
        range(dx, b, e);
        const int lo = __reduce_min_sync(FULL, e > b ? b : 0x7fffffff);
        const int hi = __reduce_max_sync(FULL, e > b ? e : 0);
        if (lane == 0) { s_lo[dx][warp] = lo; s_hi[dx][warp] = hi; }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int dx = 0; dx < SWS_COLS; ++dx) {
            int lo = 0x7fffffff, hi = 0;
#pragma unroll
            for (int w = 0; w < SW_WARPS; ++w) { lo = min(lo, s_lo[dx][w]); hi = max(hi, s_hi[dx][w]); }
            const int len = hi - lo;
            if (len > 0 && len <= SWS_CAP) {
                const unsigned bytes = (unsigned)len * (REC_CIRC * (unsigned)sizeof(double));
                mbar_expect_tx(&s_bar[dx], bytes);
                bulk_g2s(s_rec[dx], A.nbr_sweep + (size_t)lo * REC_CIRC, bytes, &s_bar[dx]);
            }
        }
    }
    int nst = 0, mine = 0;
    auto flush = [&]() {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(&A.pb.ctr[0], (unsigned long long)nst);
        base = __shfl_sync(FULL, base, 0);
        __syncwarp();
        for (int i = lane; i < nst; i += 32) {
            const unsigned long long p = base + i;
            if (p < (unsigned long long)A.pb.cap) {
                const int2 e = stage[i];
                A.pb.pairs[p] = e;
                atomicAdd(&A.pb.cnt[e.y], 1);
            }
        }
        __syncwarp();
        nst = 0;
    };
    // sweep of [b, e) with the records read through `load(u, p, v, ro)`; lanes past their range shadow record `idle`
    auto sweep = [&](int b, int e, int idle, auto load) {
        const int maxlen = __reduce_max_sync(FULL, e - b);
        for (int k0 = 0; k0 < maxlen; k0 += SW_CHUNK) {
            bool keep[SW_CHUNK];
#pragma unroll
            for (int kk = 0; kk < SW_CHUNK; ++kk) {
                const int u = b + k0 + kk;
                const bool inr = u < e;
                double2 p, v;
                double ro;
                load(inr ? u : idle, p, v, ro);
                keep[kk] = inr && sweep_keep<MODEL>(mpx, mpy, mvx, mvy, mr, lim_t, p, v, ro);
            }
#pragma unroll
            for (int kk = 0; kk < SW_CHUNK; ++kk) {
                const unsigned m = __ballot_sync(FULL, keep[kk]);
                if (keep[kk]) { stage[nst + __popc(m & ((1u << lane) - 1u))] = make_int2(t, b + k0 + kk); ++mine; }
                nst += __popc(m);
            }
            if (nst > SW_STAGE - 32 * SW_CHUNK) flush();
        }
    };
    for (int dx = 0; dx <= reach; ++dx) {
        int lo = 0x7fffffff, hi = 0;
#pragma unroll
        for (int w = 0; w < SW_WARPS; ++w) { lo = min(lo, s_lo[dx][w]); hi = max(hi, s_hi[dx][w]); }
        const int len = hi - lo;
        if (len <= 0) continue;                                  // no candidates for any target of the CTA
        int b, e;
        range(dx, b, e);
        if (len <= SWS_CAP) {
            unsigned spins = 0;
            while (!mbar_try_wait(&s_bar[dx], 0)) { if (++spins > (1u << 26)) asm volatile("trap;"); }
            const double *base = s_rec[dx];
            sweep(b, e, lo, [&](int u, double2 &p, double2 &v, double &ro) {
                const double *r = base + (u - lo) * REC_CIRC;
                p = *reinterpret_cast<const double2 *>(r); v = *reinterpret_cast<const double2 *>(r + 2); ro = r[5];
            });
        } else {
            sweep(b, e, tt, [&](int u, double2 &p, double2 &v, double &ro) {
                const double *r = A.nbr_sweep + (size_t)u * REC_CIRC;
                p = ldg2(r); v = ldg2(r + 2); ro = __ldg(r + 5);
            });
        }
    }
    if (nst) flush();
    if (mine) atomicAdd(&A.pb.cnt[t], mine);
}

// private region of the contribution array for every agent (order of the regions is irrelevant); one atomic per block
__global__ void __launch_bounds__(256) k_pair_alloc(PairBuf pb, int n_slots) {
    __shared__ int s_warp[8];
    __shared__ unsigned long long s_base;
    if (pairs_overflowed(pb)) {
        if (pb.fatal && blockIdx.x == 0 && threadIdx.x == 0) atomicExch(pb.fatal, ERR_PAIR_OVERFLOW);
        return;
    }
    const int a = blockIdx.x * 256 + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = a < n_slots ? pb.cnt[a] : 0;
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int w = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += w; }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const int ws = lane < 8 ? s_warp[lane] : 0;
        int wi = ws;
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) { const int w = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += w; }
        if (lane < 8) s_warp[lane] = wi - ws;
        if (lane == 7) s_base = wi > 0 ? atomicAdd(&pb.ctr[1], (unsigned long long)wi) : 0ULL;
    }
    __syncthreads();
    if (a < n_slots) { pb.off[a] = (int)s_base + s_warp[warp] + incl - c; pb.fill[a] = 0; }
}

// ---- both sides of a circular pair (interactions.py:53-70, distance.py:19-47, power_law.py:215-259) ---------------------
// Every quantity of the (u, t) evaluation is either equal to or the exact negation of the (t, u) one (differences, products
// of two negated factors, hypot), so both results are bit-identical to two one-sided evaluations by pair_circular().
struct PairPar { double mk, tau_0; };          // -mass * k_soc, tau_0
struct ContactPar { double mu, kappa, damping; };

template <typename LoadContact>
__device__ __forceinline__ void pair_circular_both(double pxt, double pyt, double vxt, double vyt, double rt, const PairPar &pt,
                                                   double pxu, double pyu, double vxu, double vyu, double ru, const PairPar &pu,
                                                   LoadContact load_contact, double ft[2], double fu[2]) {
    ft[0] = ft[1] = fu[0] = fu[1] = 0.0;
    const double x = pxt - pxu, y = pyt - pyu;
    const double d = hypot(x, y);
    const double r_tot = rt + ru;
    const double h = d - r_tot;
    if (!(h < SIGTH_SOC)) return;
    const double vx = vxt - vxu, vy = vyt - vyu;
    const double a = vx * vx + vy * vy;
    const double b = -(x * vx + y * vy);
    const double c = (x * x + y * y) - r_tot * r_tot;
    const double dd = sqrt(b * b - a * c);
    if (!(isnan(dd) || dd == 0.0 || a == 0.0)) {
        const double tau = (b - dd) / a;
        if (!(tau <= 0.0 || tau > TAU_MAX)) {
            const double gx = (vx - (vx * b + x * a) / dd) / a;   // power_law.py:107-126
            const double gy = (vy - (vy * b + y * a) / dd) / a;
            const double mag_t = magnitude(tau, pt.tau_0);
            const double mag_u = pu.tau_0 == pt.tau_0 ? mag_t : magnitude(tau, pu.tau_0);
            ft[0] = pt.mk * gx * mag_t; ft[1] = pt.mk * gy * mag_t;
            fu[0] = pu.mk * (0.0 - gx) * mag_u; fu[1] = pu.mk * (0.0 - gy) * mag_u;
            truncate2(ft[0], ft[1], F_SOC_MAX);
            truncate2(fu[0], fu[1], F_SOC_MAX);
        }
    }
    if (h < 0.0) {
        double nx = 0.0, ny = 0.0;
        if (d != 0.0) { nx = x / d; ny = y / d; }
        ContactPar ct, cu;
        load_contact(ct, cu);
        double cx, cy;
        force_contact(h, nx, ny, vx, vy, ny, -nx, ct.mu, ct.kappa, ct.damping, cx, cy);   // t = rotate270(n)
        ft[0] += cx; ft[1] += cy;
        force_contact(h, 0.0 - nx, 0.0 - ny, 0.0 - vx, 0.0 - vy, 0.0 - ny, nx, cu.mu, cu.kappa, cu.damping, cx, cy);
        fu[0] += cx; fu[1] += cy;
    }
}

// ---- both sides of a three-circle pair in the reference's (i, j) orientation (interactions.py:75-104, distance.py:55-105,
//      power_law.py:264-363): the body of pair_three_exact() of step_kernel.cuh with the per-side tail run twice ------------
#ifndef PAIR_HMIN_RANKED
#define PAIR_HMIN_RANKED 1      // distance_three_circles: rank the nine part pairs in fp32, exact hypot for the possible minima only
#endif
struct Three {               // kinematics of one three-circle agent as the pair kernels need them
    double x0, y0, x1, y1, x2, y2;   // torso, left shoulder, right shoulder centres
    double rt, rs;           // torso / shoulder radius
    double vx, vy;
    double ox, oy;           // r_ts * (sin(phi), -cos(phi)): shoulder displacement (power_law.py:338-350, agents.py:483-484)
};

__device__ __forceinline__ double sel3(int k, double a, double b, double c) { return k == 0 ? a : (k == 1 ? b : c); }

__device__ __forceinline__ void load_three_rec(const double *__restrict__ nbr, int u, Three &k) {
    const double *r = nbr + (size_t)u * REC_THREE;
    const double2 a = ldg2(r), b = ldg2(r + 2), c = ldg2(r + 4), d = ldg2(r + 6), e = ldg2(r + 8), f = ldg2(r + 10), g = ldg2(r + 12);
    k.x0 = a.x; k.y0 = a.y; k.vx = b.x; k.vy = b.y; k.rt = c.y; k.rs = d.x;
    k.x1 = e.x; k.y1 = e.y; k.x2 = f.x; k.y2 = f.y; k.ox = g.x; k.oy = g.y;
}

template <typename LoadContact>
__device__ __forceinline__ void pair_three_both(const Three &I, const Three &J, const PairPar &pi, const PairPar &pj,
                                                LoadContact load_contact, double fi[3], double fj[3]) {
    fi[0] = fi[1] = fi[2] = fj[0] = fj[1] = fj[2] = 0.0;
    const double jx[3] = {J.x0, J.x1, J.x2}, jy[3] = {J.y0, J.y1, J.y2}, rj[3] = {J.rt, J.rs, J.rs};
    // distance_three_circles (distance.py:55-105): h = hypot(x, y) - (r_i + r_j) over the nine part pairs, strict '<', first
    // wins, order torso, left, right.  The nine fp64 hypot are most of this kernel's instructions, and all but one lose:
    // the part pairs are RANKED in fp32 first (error of h below 4e-7 (d_max + radii); the margin is five times that), and
    // the exact hypot runs only for those that can attain the minimum, in the reference's order with the reference's rule.
    // Every part pair whose exact h equals the exact minimum is a candidate and every other candidate loses the strict
    // comparison, so h_min, the winning pair and its (x, y, d) are the very numbers the nine-fold loop produces
    // (validated on the host against that loop, 24 M pairs incl. exact ties: scripts/validate_hmin_ranking.c).
    double h_min = nan(""), sx = 0.0, sy = 0.0, sd = 0.0;
    int i_min = 0, j_min = 0;
#if PAIR_HMIN_RANKED
    {
        const float rit = (float)I.rt, ris = (float)I.rs, rjt = (float)J.rt, rjs = (float)J.rs;
        float ha[9], ha_lo = 3.0e38f, d_hi = 0.0f, chk = 0.0f;
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            const int pi_ = k / 3, pj_ = k % 3;
            const double x = (pi_ == 0 ? I.x0 : (pi_ == 1 ? I.x1 : I.x2)) - jx[pj_];
            const double y = (pi_ == 0 ? I.y0 : (pi_ == 1 ? I.y1 : I.y2)) - jy[pj_];
            const float df = __fsqrt_rn((float)fma(x, x, y * y));
            ha[k] = df - ((pi_ == 0 ? rit : ris) + (pj_ == 0 ? rjt : rjs));
            ha_lo = fminf(ha_lo, ha[k]); d_hi = fmaxf(d_hi, df); chk += ha[k];
        }
        const float thr = ha_lo + 2.0f * (1e-6f * (d_hi + rit + ris + rjt + rjs) + 1e-15f);
        unsigned cand = 0;
#pragma unroll
        for (int k = 0; k < 9; ++k) cand |= ha[k] <= thr ? (1u << k) : 0u;
        if (!(fabsf(chk) < 1e30f)) cand = 0x1ffu;       // NaN / overflow somewhere: all nine, exactly the reference's loop
        while (cand) {
            const int k = __ffs((int)cand) - 1;
            cand &= cand - 1;
            const int pi_ = (k >= 3) + (k >= 6), pj_ = k - 3 * pi_;
            const double x = sel3(pi_, I.x0, I.x1, I.x2) - sel3(pj_, J.x0, J.x1, J.x2);
            const double y = sel3(pi_, I.y0, I.y1, I.y2) - sel3(pj_, J.y0, J.y1, J.y2);
            const double d = hypot(x, y);
            const double h = d - ((pi_ == 0 ? I.rt : I.rs) + (pj_ == 0 ? J.rt : J.rs));
            if (h < h_min || isnan(h_min)) { h_min = h; sx = x; sy = y; sd = d; i_min = pi_; j_min = pj_; }
        }
    }
#else
#pragma unroll 1
    for (int pi_ = 0; pi_ < 3; ++pi_) {
        const double xi = sel3(pi_, I.x0, I.x1, I.x2), yi = sel3(pi_, I.y0, I.y1, I.y2), rip = pi_ == 0 ? I.rt : I.rs;
#pragma unroll
        for (int pj_ = 0; pj_ < 3; ++pj_) {
            const double x = xi - jx[pj_], y = yi - jy[pj_];
            const double d = hypot(x, y);
            const double h = d - (rip + rj[pj_]);
            if (h < h_min || isnan(h_min)) { h_min = h; sx = x; sy = y; sd = d; i_min = pi_; j_min = pj_; }
        }
    }
#endif
    if (!(h_min < SIGTH_SOC)) return;
    double nx = 0.0, ny = 0.0;
    if (sd != 0.0) { nx = sx / sd; ny = sy / sd; }
    const double vx = I.vx - J.vx, vy = I.vy - J.vy;
    const double a = vx * vx + vy * vy;
    double fix = 0.0, fiy = 0.0, fjx = 0.0, fjy = 0.0;
    if (a != 0.0) {
        // smallest time-to-collision over the 9 part pairs with the reference's selection rule (power_law.py:308-329):
        // `isnan(tau) or 0 < tau_new < tau` with tau_new = (b - d) / a.  `a` is the same for all nine, so the rule is decided
        // on the numerators: num < num_sel (1 - 1e-15) implies fl(num / a) < fl(num_sel / a) whenever both quotients are
        // normal numbers (each is within 2^-53 relative of the exact value), and num > 0 then implies tau_new > 0.  Only
        // near-ties and extreme magnitudes take the two divisions; the selected tau itself is formed once, after the loop --
        // the same quotient of the same operands the reference forms at selection time.
        const bool a_ok = a > 1e-100 && a < 1e100;
        bool have = false;
        double num_sel = 0.0, b_min = 0.0, d_min = 0.0;
        int contact_i = 0, contact_j = 0;
#pragma unroll 1
        for (int pi_ = 0; pi_ < 3; ++pi_) {
            const double xi = sel3(pi_, I.x0, I.x1, I.x2), yi = sel3(pi_, I.y0, I.y1, I.y2), rip = pi_ == 0 ? I.rt : I.rs;
#pragma unroll
            for (int pj_ = 0; pj_ < 3; ++pj_) {
                const double x = xi - jx[pj_], y = yi - jy[pj_];
                const double r_tot = rip + rj[pj_];
                const double b = -(x * vx + y * vy);
                const double c = (x * x + y * y) - r_tot * r_tot;
                const double disc = b * b - a * c;
                if (!(disc > 0.0)) continue;         // sqrt gives NaN (disc < 0 or NaN) or 0
                const double dd = sqrt(disc);
                const double num = b - dd;
                bool take = !have;
                if (have && num > 0.0 && num < num_sel) {
                    if (a_ok && num > 1e-200 && num_sel < 1e100 && num < num_sel * (1.0 - 1e-15)) take = true;
                    else { const double tau_new = num / a; take = 0.0 < tau_new && tau_new < num_sel / a; }
                }
                if (take) { have = true; num_sel = num; b_min = b; d_min = dd; contact_i = pi_; contact_j = pj_; }
            }
        }
        const double tau = have ? num_sel / a : nan("");
        if (!(isnan(tau) || tau <= 0.0)) {
            // shoulder displacement of the contacting parts: 0 for the torso, +o for left, -o for right
            const double oix = sel3(contact_i, 0.0, I.ox, 0.0 - I.ox), oiy = sel3(contact_i, 0.0, I.oy, 0.0 - I.oy);
            const double ojx = sel3(contact_j, 0.0, J.ox, 0.0 - J.ox), ojy = sel3(contact_j, 0.0, J.oy, 0.0 - J.oy);
            const double xr = I.x0 - J.x0, yr = I.y0 - J.y0;
            const double ox = oix - ojx, oy = oiy - ojy;
            const double gx = (vx - (a * (xr + 2 * ox) + b_min * vx) / d_min) / a;   // power_law.py:131-149
            const double gy = (vy - (a * (yr + 2 * oy) + b_min * vy) / d_min) / a;
            const double mag_i = magnitude(tau, pi.tau_0);
            const double mag_j = pj.tau_0 == pi.tau_0 ? mag_i : magnitude(tau, pj.tau_0);
            fix = pi.mk * gx * mag_i; fiy = pi.mk * gy * mag_i;
            fjx = 0.0 - pj.mk * gx * mag_j; fjy = 0.0 - pj.mk * gy * mag_j;   // force_j[:] -= ... (power_law.py:358)
            truncate2(fix, fiy, F_SOC_MAX);
            truncate2(fjx, fjy, F_SOC_MAX);
        }
    }
    if (h_min < 0.0) {
        ContactPar ci, cj;
        load_contact(ci, cj);
        double cx, cy;
        force_contact(h_min, nx, ny, vx, vy, ny, -nx, ci.mu, ci.kappa, ci.damping, cx, cy);
        fix += cx; fiy += cy;
        force_contact(h_min, nx, ny, vx, vy, ny, -nx, cj.mu, cj.kappa, cj.damping, cx, cy);
        fjx -= cx; fjy -= cy;
    }
    // moment arms, distance.py:102-103:  i: x0[i_min] + r0[i_min] n - x0[0];  j: x0[j_min] - r1[j_min] n - x1[0]  (sic)
    const double mir = i_min == 0 ? I.rt : I.rs, mjr = j_min == 0 ? J.rt : J.rs;
    const double mix = sel3(i_min, I.x0, I.x1, I.x2) + mir * nx - I.x0, miy = sel3(i_min, I.y0, I.y1, I.y2) + mir * ny - I.y0;
    const double mjx = sel3(j_min, I.x0, I.x1, I.x2) - mjr * nx - J.x0, mjy = sel3(j_min, I.y0, I.y1, I.y2) - mjr * ny - J.y0;
    fi[0] = fix; fi[1] = fiy; fi[2] = mix * fiy - miy * fix;
    fj[0] = fjx; fj[1] = fjy; fj[2] = mjx * fjy - mjy * fjx;
}

struct EvalArgs {
    const double *nbr;           // packed neighbour records in cell order (+ ghost tail)
    const double2 *par;          // {-mass * k_soc, tau_0} in cell order (+ ghost tail)
    Soa in;                      // planes (contact parameters, read through `order` only when a pair overlaps)
    const int *order;            // sorted slot -> plane slot (nullptr: planes are in cell order)
    int ghost_base;              // slots >= ghost_base are ghosts: their side of a pair is not stored (INT_MAX: none)
    int ghost_left_end;          // ghost_base <= slot < ghost_left_end: left ghost column (sorts first)
    PairBuf pb;
};

__device__ __forceinline__ void store_contribution(const PairBuf &pb, int a, int partner_key, double fx, double fy, double tq) {
    const int k = atomicAdd(&pb.fill[a], 1);
    double *e = pb.cres + ((size_t)pb.off[a] + k) * 4;
    reinterpret_cast<double2 *>(e)[0] = make_double2(__longlong_as_double((long long)partner_key), fx);
    reinterpret_cast<double2 *>(e)[1] = make_double2(fy, tq);
}

template <int MODEL>
#ifndef EVAL_MINB_THREE
#define EVAL_MINB_THREE 4
#endif
#ifndef EVAL_MINB_CIRC
#define EVAL_MINB_CIRC 8
#endif
__global__ void __launch_bounds__(128, MODEL == 0 ? EVAL_MINB_CIRC : EVAL_MINB_THREE) k_pair_eval(const EvalArgs A) {
    if (pairs_overflowed(A.pb)) return;
    const long long np = (long long)A.pb.ctr[0];
    const Soa &s = A.in;
    // The kernel waits on memory, not on arithmetic (ncu, round 2: long_scoreboard 4 warps per issue at 16 warps per SM, L1 hit
    // rate 53 %): the pair two iterations ahead is read while this one is evaluated, and for three-circle agents the records
    // of the next pair (one 128-byte line each) are requested into L1 -- a prefetch holds no registers.  Measured (1 M agents,
    // profiles/prefetch_ab_r4c/): three-circle 0.153 -> 0.1445 ms; the cheap circular evaluation got 5 % slower with it.
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    int2 e0 = p < np ? A.pb.pairs[p] : make_int2(0, 0);
    int2 e1 = p + stride < np ? A.pb.pairs[p + stride] : make_int2(0, 0);
    for (; p < np; p += stride) {
        const int2 e = e0;
        e0 = e1;
#if PAIR_PREFETCH
        if (MODEL == 1 && p + stride < np) {
            constexpr int REC = MODEL == 0 ? REC_CIRC : REC_THREE;
            prefetch_line(A.nbr + (size_t)e0.x * REC); prefetch_line(A.nbr + (size_t)e0.y * REC);
            prefetch_line(A.par + e0.x); prefetch_line(A.par + e0.y);
        }
#endif
        if (p + 2 * stride < np) e1 = A.pb.pairs[p + 2 * stride];
        const int t = e.x, u = e.y;
        const double2 qt = __ldg(A.par + t), qu = __ldg(A.par + u);
        const PairPar pt = {qt.x, qt.y}, pu = {qu.x, qu.y};
        double ft[3] = {0.0, 0.0, 0.0}, fu[3] = {0.0, 0.0, 0.0};
        if (MODEL == 0) {
            const double *rt = A.nbr + (size_t)t * REC_CIRC, *ru = A.nbr + (size_t)u * REC_CIRC;
            const double2 p0 = ldg2(rt), v0 = ldg2(rt + 2), p1 = ldg2(ru), v1 = ldg2(ru + 2);
            auto contact = [&](ContactPar &ct, ContactPar &cu) {
                ct = ContactPar{0.0, 0.0, 0.0}; cu = ct;
                if (t < A.ghost_base) { const int o = A.order ? A.order[t] : t; ct = ContactPar{s(MU, o), s(KAPPA, o), s(DAMPING, o)}; }
                if (u < A.ghost_base) { const int o = A.order ? A.order[u] : u; cu = ContactPar{s(MU, o), s(KAPPA, o), s(DAMPING, o)}; }
            };
            pair_circular_both(p0.x, p0.y, v0.x, v0.y, __ldg(rt + 4), pt, p1.x, p1.y, v1.x, v1.y, __ldg(ru + 4), pu, contact, ft, fu);
        } else {
            // reference pair orientation: i is the lexicographically smaller (cell_x, cell_y, agent index).  The TRUE cell
            // coordinates floor(p / c) (last two slots of the neighbour record) are compared, not the flat ids of the search
            // lattice, so the convention does not depend on how the lattice was chosen (padded, fixed, clamped, per-strip).
            const double2 ca = ldg2(A.nbr + (size_t)t * REC_THREE + 14), cb = ldg2(A.nbr + (size_t)u * REC_THREE + 14);
            // Inside one true cell the agent index decides (slot 7 of the record; on the finer search lattice two agents of a
            // cell_size cell may sit in different search cells, so the slot order does not tell).
            const bool t_is_i = ca.x != cb.x ? ca.x < cb.x : (ca.y != cb.y ? ca.y < cb.y :
                                __ldg(A.nbr + (size_t)t * REC_THREE + 7) < __ldg(A.nbr + (size_t)u * REC_THREE + 7));
            const int si = t_is_i ? t : u, sj = t_is_i ? u : t;
            Three I, J;
            load_three_rec(A.nbr, si, I);
            load_three_rec(A.nbr, sj, J);
            auto contact = [&](ContactPar &ci, ContactPar &cj) {
                ci = ContactPar{0.0, 0.0, 0.0}; cj = ci;
                if (si < A.ghost_base) { const int o = A.order ? A.order[si] : si; ci = ContactPar{s(MU, o), s(KAPPA, o), s(DAMPING, o)}; }
                if (sj < A.ghost_base) { const int o = A.order ? A.order[sj] : sj; cj = ContactPar{s(MU, o), s(KAPPA, o), s(DAMPING, o)}; }
            };
            // ONE call site: with two (one per orientation) the warp splits and each half runs the whole evaluation at half
            // occupancy of its lanes (ncu, round 2: 17.5 of 32 lanes active throughout pair_three_both)
            const PairPar pI = t_is_i ? pt : pu, pJ = t_is_i ? pu : pt;
            double fI[3], fJ[3];
            pair_three_both(I, J, pI, pJ, contact, fI, fJ);
#pragma unroll
            for (int k = 0; k < 3; ++k) { ft[k] = t_is_i ? fI[k] : fJ[k]; fu[k] = t_is_i ? fJ[k] : fI[k]; }
        }
        // a side whose result is exactly zero adds nothing: not stored
        const int key_t = (t >= A.ghost_base && t < A.ghost_left_end) ? t - GHOST_KEY_SHIFT : t;
        const int key_u = (u >= A.ghost_base && u < A.ghost_left_end) ? u - GHOST_KEY_SHIFT : u;
        if (t < A.ghost_base && (ft[0] != 0.0 || ft[1] != 0.0 || ft[2] != 0.0)) store_contribution(A.pb, t, key_u, ft[0], ft[1], ft[2]);
        if (u < A.ghost_base && (fu[0] != 0.0 || fu[1] != 0.0 || fu[2] != 0.0)) store_contribution(A.pb, u, key_t, fu[0], fu[1], fu[2]);
    }
}

// adds the `n` contributions at `e` in ascending partner order (selection: the lists are a handful of entries long)
__device__ __forceinline__ void gather_contributions(const double *__restrict__ e, int n, bool torque, double &fx, double &fy, double &tq) {
    long long last = -0x7fffffffffffffffLL - 1;
    for (int r = 0; r < n; ++r) {
        long long best = 0x7fffffffffffffffLL;
        int bi = 0;
        for (int k = 0; k < n; ++k) {
            const long long key = __double_as_longlong(e[(size_t)k * 4]);
            if (key > last && key < best) { best = key; bi = k; }
        }
        const double2 v0 = *reinterpret_cast<const double2 *>(e + (size_t)bi * 4), v1 = *reinterpret_cast<const double2 *>(e + (size_t)bi * 4 + 2);
        fx += v0.y; fy += v1.x;
        if (torque) tq += v1.y;
        last = best;
    }
}
