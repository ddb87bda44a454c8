// strip_kernels.cuh -- halo / migrant packing for the spatial strip decomposition.
// A halo message carries one cell column: header {count, ny, -, -}, ny int32 per-cell counts, then `count` packed neighbour
// records in cell order.  A migrant message carries whole agents: header {count, -, -, -}, then per agent n_planes doubles
// + global id + target.
#pragma once
#include "kernels.cuh"
#include "step_kernel.cuh"

// ---- one-sided exchange over peer memory (NVLink): the producer writes the message straight into the consumer's receive
// buffer and then publishes a sequence number in the consumer's memory; the consumer's kernel spins on that number.
// Single buffers suffice: halo and migrant messages alternate, and each one can only be produced after the other one of the
// previous phase was consumed ("Multi-GPU").
__device__ __forceinline__ void signal_store(unsigned long long *flag, unsigned long long seq) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag), "l"(seq) : "memory");
}
__device__ __forceinline__ unsigned long long signal_load(const unsigned long long *flag) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
    return v;
}
// all threads of the block return once *flag >= seq (flag == nullptr: nothing to wait for)
__device__ __forceinline__ void signal_wait(const unsigned long long *flag, unsigned long long seq) {
    if (flag) {
        if (threadIdx.x == 0) while (signal_load(flag) < seq) __nanosleep(64);
        __syncthreads();
    }
}
// last block of a grid to arrive publishes the flag (done: a zero-initialised counter, reset for the next use)
__device__ __forceinline__ void signal_when_grid_done(unsigned int *done, unsigned long long *flag, unsigned long long seq) {
    if (!flag) return;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int prev = atomicAdd(done, 1u);
        if (prev == gridDim.x - 1) { *done = 0u; signal_store(flag, seq); }
    }
}

__host__ __device__ inline long long halo_counts_doubles(long long ny) { return (ny + 1) / 2; }

// column `col` of the local lattice -> message.  One block; the column's agents are contiguous in cell order.
__global__ void k_halo_pack(const double *__restrict__ nbr, int rec, const int *__restrict__ cell_start,
                            const int *__restrict__ cell_count, int col, int ny, double *__restrict__ msg, long long cap, int *error,
                            unsigned int *done, unsigned long long *peer_flag, unsigned long long seq) {
    const int b = cell_start[col * ny];
    const int e = cell_start[col * ny + ny - 1] + cell_count[col * ny + ny - 1];
    int count = e - b;
    if (count > cap) { if (threadIdx.x == 0 && blockIdx.x == 0) atomicExch(error, ERR_CELL_RANGE + 1); count = (int)cap; }
    if (blockIdx.x == 0 && threadIdx.x == 0) { msg[0] = (double)count; msg[1] = (double)ny; msg[2] = 0.0; msg[3] = 0.0; }
    int *counts = reinterpret_cast<int *>(msg + MSG_HEADER);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ny; i += gridDim.x * blockDim.x) counts[i] = cell_count[col * ny + i];
    double *dst = msg + MSG_HEADER + halo_counts_doubles(ny);
    const double *src = nbr + (size_t)b * rec;
    for (long long i = blockIdx.x * blockDim.x + threadIdx.x; i < (long long)count * rec; i += (long long)gridDim.x * blockDim.x)
        dst[i] = src[i];
    signal_when_grid_done(done, peer_flag, seq);
}

// message -> ghost column `col` of the local lattice; ghost records live at slots [base, base + count) of nbr.
// Block 0 scans the per-cell counts into the cell tables; every block copies a share of the records (the copy does not
// depend on the scan: ghosts arrive in cell order, so record k simply goes to slot base + k).
__global__ void k_halo_unpack(const double *__restrict__ msg, int rec, double *__restrict__ nbr, double *__restrict__ nbr_sweep,
                              int *__restrict__ cell_sorted,
                              int *__restrict__ cell_start, int *__restrict__ cell_count, int col, int ny, int base, long long cap,
                              int *error, double2 *__restrict__ par, const unsigned long long *flag, unsigned long long seq) {
    __shared__ int warp_sums[32];
    __shared__ int carry_s;
    signal_wait(flag, seq);
    const int *counts = reinterpret_cast<const int *>(msg + MSG_HEADER);
    const int count = (int)msg[0];
    if ((int)msg[1] != ny || count > cap) { if (threadIdx.x == 0 && blockIdx.x == 0) atomicExch(error, ERR_CELL_RANGE + 2); return; }
    const double *src = msg + MSG_HEADER + halo_counts_doubles(ny);
    double *dst = nbr + (size_t)base * rec;
    const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
    for (long long i = tid; i < (long long)count * rec; i += nth) dst[i] = src[i];
    if (nbr_sweep != nbr)     // three-circle: compact sweep records of the ghosts {px, py, vx, vy, extent, inflated extent}
        for (long long i = tid; i < (long long)count * 6; i += nth) {
            const long long a = i / 6, f = i % 6;
            nbr_sweep[(size_t)base * 6 + i] = f < 5 ? src[a * rec + f] : src[a * rec + 4] * (1.0 + 1e-9);
        }
    // pair parameters of ghosts are never used for a stored result, but k_pair_eval reads them: keep them defined
    if (par) for (long long i = tid; i < count; i += nth) par[base + i] = make_double2(0.0, 1.0);
    if (blockIdx.x != 0) return;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base_i = 0; base_i < ny; base_i += blockDim.x) {
        const int i = base_i + threadIdx.x;
        const int v = i < ny ? counts[i] : 0;
        int incl = v;
        for (int o = 1; o < 32; o <<= 1) { int w = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += w; }
        if (lane == 31) warp_sums[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int ws = warp_sums[lane], wi = ws;
            for (int o = 1; o < 32; o <<= 1) { int w = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += w; }
            warp_sums[lane] = wi - ws;
        }
        __syncthreads();
        const int excl = carry_s + warp_sums[warp] + incl - v;
        if (i < ny) {
            cell_start[col * ny + i] = base + excl;
            cell_count[col * ny + i] = v;
            for (int k = 0; k < v; ++k) cell_sorted[base + excl + k] = col * ny + i;
        }
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry_s = excl + v;
        __syncthreads();
    }
}

// no neighbour on this side / nothing received: empty ghost column
__global__ void k_ghost_clear(int *cell_start, int *cell_count, int col, int ny, int base) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ny; i += gridDim.x * blockDim.x) {
        cell_start[col * ny + i] = base; cell_count[col * ny + i] = 0;
    }
}

__global__ void k_counters_zero(int *c, int n) { if (threadIdx.x < n) c[threadIdx.x] = 0; }

// device-side bookkeeping of the strip step (no host round trip): after the integrating step the live agents are compact
// in [0, live); after absorbing migrants the slot count grows by the number appended
__global__ void k_counts_after_step(DevCounts *c) { if (threadIdx.x == 0) c->slots = c->live; }
// host_bound: the slot count the host will assume for its next launches (it only synchronises every few steps); more
// arrivals than that in one step would be silently skipped by those launches => device error instead
__global__ void k_counts_after_absorb(DevCounts *c, const int *counters, long long host_bound, int *error) {
    if (threadIdx.x == 0) {
        c->slots += counters[2];
        if (host_bound >= 0 && c->slots > host_bound) atomicExch(error, ERR_CELL_RANGE + 5);
    }
}
__global__ void k_counts_set(DevCounts *c, int slots) { if (threadIdx.x == 0) { c->slots = slots; c->live = slots; } }

// agents whose cell column left the owned range [col_lo, col_hi] move to the neighbour: append the whole agent to the
// message of that side and vacate the slot (id = -1).  counters: [0] left, [1] right.
__global__ void k_migrants_pack(Soa s, int n_host, const int *n_dev, int n_planes, double cell_size, long long ix_min_local, int col_lo,
                                int col_hi, int has_left, int has_right, double *__restrict__ msg_left, double *__restrict__ msg_right,
                                long long cap, int *counters, int *error, AgentFlags flags) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= eff_n(n_host, n_dev) || s.id[i] < 0) return;
    const double col = floor(s(PX, i) / cell_size) - (double)ix_min_local;
    int side = -1;
    if (has_left && col < (double)col_lo) side = 0;
    else if (has_right && col > (double)col_hi) side = 1;
    if (side < 0) return;
    const int k = atomicAdd(&counters[side], 1);
    if (k >= cap) { atomicExch(error, ERR_CELL_RANGE + 3); return; }
    double *dst = (side == 0 ? msg_left : msg_right) + MSG_HEADER + (size_t)k * (n_planes + 2);
    for (int p = 0; p < n_planes; ++p) dst[p] = s(p, i);
    dst[n_planes] = pack_id_flags(s.id[i], flags);
    dst[n_planes + 1] = (double)s.target[i];
    s.id[i] = -1;
}

__global__ void k_migrants_header(double *msg_left, double *msg_right, const int *counters, long long cap,
                                  unsigned long long *flag_left, unsigned long long *flag_right, unsigned long long seq) {
    if (threadIdx.x == 0) {
        if (msg_left) { msg_left[0] = (double)min((long long)counters[0], cap); msg_left[1] = msg_left[2] = msg_left[3] = 0.0; }
        if (msg_right) { msg_right[0] = (double)min((long long)counters[1], cap); msg_right[1] = msg_right[2] = msg_right[3] = 0.0; }
        // (the migrant records were written by the previous kernel of this stream: complete, made visible by the fence)
        if (flag_left) signal_store(flag_left, seq);
        if (flag_right) signal_store(flag_right, seq);
    }
}

// append received migrants after the current slots; counters[2] = number appended so far
__global__ void k_migrants_unpack(const double *__restrict__ msg, Soa s, int n_slots_host, const int *n_slots_dev, int n_planes,
                                  long long capacity, int *counters, int *error, const unsigned long long *flag, unsigned long long seq,
                                  AgentFlags flags) {
    signal_wait(flag, seq);
    const int m = (int)msg[0];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const int n_slots = n_slots_dev ? *n_slots_dev : n_slots_host;
    const int slot = n_slots + atomicAdd(&counters[2], 1);
    if (slot >= capacity) { atomicExch(error, ERR_CELL_RANGE + 4); return; }
    const double *src = msg + MSG_HEADER + (size_t)i * (n_planes + 2);
    for (int p = 0; p < n_planes; ++p) s(p, slot) = src[p];
    s.id[slot] = unpack_id_flags(src[n_planes], flags);
    s.target[slot] = (long long)src[n_planes + 1];
}

// Adaptive dt across strips: the two maxima travel as {max |v|, max v0, NaN flag of the first, NaN flag of the second}.  The
// reference's np.max propagates NaN (integrator.py:25-60); an all_reduce(MAX) over NCCL / gloo does not promise to, so a NaN is
// replaced by a neutral value plus a flag that survives the MAX and is turned back into NaN on import: every rank sees the
// same dt whatever the collective does with NaN payloads.
__global__ void k_vmax_export(const unsigned long long *vmax, double *out) {
    if (threadIdx.x == 0) {
        const double ninf = -__longlong_as_double(0x7ff0000000000000LL);
        const double a = from_ordered_bits(vmax[0]);
        const double b = vmax[1] == 0xffffffffffffffffULL ? nan("") : (vmax[1] == 0ULL ? ninf : from_ordered_bits(vmax[1]));
        out[0] = isnan(a) ? 0.0 : a;
        out[1] = isnan(b) ? ninf : b;
        out[2] = isnan(a) ? 1.0 : 0.0;
        out[3] = isnan(b) ? 1.0 : 0.0;
    }
}
__global__ void k_vmax_import(const double *in, unsigned long long *vmax) {
    if (threadIdx.x == 0) {
        const double a = in[2] > 0.0 ? nan("") : in[0];
        vmax[0] = ordered_bits(a);
        vmax[1] = in[3] > 0.0 || isnan(in[1]) ? 0xffffffffffffffffULL : ordered_bits(in[1]);
    }
}

__global__ void k_set_ids(Soa s, const long long *ids, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) s.id[i] = (int)ids[i];
}

// live agents, compacted in slot order: full records rebuilt from the planes (unmirrored fields zero) + ids
template <int MODEL>
__global__ void k_export_records(Soa s, int n, uint8_t *__restrict__ aos, long long *ids, int *counter) {
    constexpr int ITEM = MODEL == 0 ? 228 : 316;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n || s.id[t] < 0) return;
    const int k = atomicAdd(counter, 1);
    uint32_t *rec = reinterpret_cast<uint32_t *>(aos + (size_t)k * ITEM);
    for (int w = 0; w < ITEM / 4; ++w) rec[w] = 0u;
    const FieldMap *fm = MODEL == 0 ? c_fields_circ : c_fields_three;
    constexpr int NF = MODEL == 0 ? N_FIELDS_CIRC : N_FIELDS_THREE;
#pragma unroll 1
    for (int f = 0; f < NF; ++f) {
        double v = s(fm[f].plane, t);
        int w = fm[f].offset >> 2;
        rec[w] = (uint32_t)__double2loint(v);
        rec[w + 1] = (uint32_t)__double2hiint(v);
    }
    uint8_t *rb = reinterpret_cast<uint8_t *>(rec) + (MODEL == 0 ? 2 : 34);
    unsigned long long tv = (unsigned long long)s.target[t];
    for (int b = 0; b < 8; ++b) rb[b] = (uint8_t)(tv >> (8 * b));
    reinterpret_cast<uint8_t *>(rec)[MODEL == 0 ? 0 : 32] = 1;   // active
    ids[k] = s.id[t];
}
