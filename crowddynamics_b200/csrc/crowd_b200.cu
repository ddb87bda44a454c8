// crowd_b200.cu -- hand-written sm_100a CUDA implementation of the crowddynamics per-timestep agent update,
// behind the C ABI declared in include/crowd_b200.h.
//
// Compiled with -fmad=false: the reference's numba/LLVM code never contracts a*b+c, and several quantities
// (the time-to-collision discriminant b*b - a*c, skin distances d - r_tot) are differences of nearly equal numbers,
// so the pair kernels keep the reference's exact operation order and rounding.  Division and sqrt are IEEE
// correctly rounded in CUDA fp64, so per-pair results differ from the CPU only through hypot/exp/sin/cos (<= 1-2 ulp).
//
// File:line citations refer to /root/reference/crowddynamics/.
#include "crowd_b200.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#ifndef SWEEP_MODE_DEFAULT
#define SWEEP_MODE_DEFAULT 0     // classification sweep: 0 k_sweep, 1 k_sweep_staged (pair_kernels.cuh)
#endif

#include "kernels.cuh"
#include "step_kernel.cuh"
#include "finish_kernel.cuh"
#include "small_kernel.cuh"
#include "transfer_kernels.cuh"
#include "field_kernels.cuh"
#include "strip_kernels.cuh"
#include "collective_kernels.cuh"

namespace {

thread_local std::string g_err;

int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CK(call)                                                                                              \
    do {                                                                                                      \
        cudaError_t e_ = (call);                                                                              \
        if (e_ != cudaSuccess)                                                                                \
            return fail(CDB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)
#define CKS(call)                    \
    do {                             \
        int s_ = (call);             \
        if (s_ != CDB_OK) return s_; \
    } while (0)

template <typename T>
int dev_alloc(T **p, size_t count) {
    if (*p) cudaFree(*p);
    *p = nullptr;
    if (count == 0) count = 1;
    CK(cudaMalloc((void **)p, count * sizeof(T)));
    return CDB_OK;
}

constexpr int DT_LOG = DT_LOG_SLOTS;
constexpr int HALO_BLOCKS = 32;    // blocks of the halo pack / unpack kernels (128 measured the same: a column of a 1 M-agent strip is ~1 MB)
constexpr int STRIP_REFRESH = 16;
constexpr int PROFILE_MAX_STEPS = 4096;
constexpr int PROF_EVENTS = 6;   // per step: begin, block list, sweep + alloc, pair evaluation, step kernel, end

inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

}  // namespace

struct cdb_sim;
static cudaError_t sync_stream(cdb_sim *sim);   // cudaStreamSynchronize(sim->stream), counted (cdb_sync_count)

// every kernel launch of the library goes through here (counted: cdb_launch_count)
#define LAUNCH(sim, kernel, grid, block, smem, ...)                          \
    do {                                                                     \
        kernel<<<(grid), (block), (smem), (sim)->stream>>>(__VA_ARGS__);     \
        (sim)->launches++;                                                   \
    } while (0)

struct cdb_sim {
    int device = 0;
    int sm_count = 148;
    int model = 0;
    int64_t itemsize = 0;
    int n_planes = 0;         // planes that mirror record fields
    int n_alloc_planes = 0;   // planes allocated (== n_planes)
    int sweep_mode = SWEEP_MODE_DEFAULT;   // 0 k_sweep, 1 k_sweep_staged (env CROWD_B200_SWEEP = plain | staged overrides)
    int variant = 3;          // agent-agent: 1 = one-phase reference kernels, 2 = two-phase inside the fused kernel (every pair
                              // evaluated from both sides), 3 = half-stencil sweep + pair list + one evaluation per pair
    bool auto_lattice_valid = false;
    int auto_lattice_age = 0;
    int64_t capacity = 0;   // agent slots allocated (incl. room for ghosts / migrants)
    int64_t n = 0;          // slots in use in `cur` (agents currently held, incl. slots vacated by migrants)
    int64_t n_sorted = 0;   // live agents in the current block list
    bool perm_valid = false; // block list refers to `cur` through d_order (planes not physically sorted)
    cudaStream_t stream = nullptr;
    bool own_stream = false;

    Soa cur{}, alt{};       // ping-pong SoA state; `cur` is authoritative
    uint8_t *d_aos = nullptr;      // device image of the packed host records (n * itemsize)
    int64_t aos_capacity = 0;
    uint8_t *h_bounce = nullptr;   // pinned host bounce buffer for field-wise downloads
    int64_t bounce_bytes = 0;
    std::vector<std::pair<void *, size_t>> registered;   // host ranges pinned by cdb_host_register
    int *d_rec_slot = nullptr;     // record index -> slot (field-masked uploads into a re-sorted state)
    int64_t rec_slot_cap = 0;
    int64_t h2d_bytes = 0, d2h_bytes = 0;   // bytes moved over PCIe by the transfer entry points (cdb_transfer_stats)
    int64_t syncs = 0;             // blocking host synchronisations so far (cdb_sync_count)
    // asynchronous snapshots (cdb_snapshot_* / cdb_scalars_*): side stream, two pinned slots each
    cudaStream_t side = nullptr;
    cudaEvent_t ev_main = nullptr, ev_snap[2] = {nullptr, nullptr}, ev_scal[2] = {nullptr, nullptr}, ev_pairs = nullptr;
    uint8_t *d_snap[2] = {nullptr, nullptr}, *h_snap[2] = {nullptr, nullptr};
    int64_t snap_bytes[2] = {0, 0}, snap_n[2] = {0, 0};
    int snap_next = 0, scal_next = 0;
    unsigned long long *d_scal = nullptr, *h_scal[2] = {nullptr, nullptr};   // [0] dt bits, [1] time_tot bits, [2] inside-domain changes, [3..] target counts
    int64_t scal_n[2] = {0, 0};
    bool defer_sync = false;       // cdb_set_deferred_sync: cdb_step does not wait for its own pair-count check
    int64_t pairs_known_found = -1; int pairs_known_age = 1 << 30;
    bool pairs_inflight = false;   // a non-blocking pair-count check has been issued and not read yet
    int64_t iterations_at_check = 0;   // sim->iterations when that check was queued
    uint32_t last_flags = 0; double last_cell_size = 0.0, last_dt_min = 0.0, last_dt_max = 0.0;   // of the last cdb_step call (catch_up)

    // block list
    bool lattice_fixed = false;
    Grid grid{};             // host copy
    Grid *d_grid = nullptr;
    double cell_size = 0.0;
    double cell_size_lattice = 0.0;
    int fine = 1;                    // the current tables bin on cell_size / fine (1, or 2: the twice finer search lattice)
    int fine_lattice = 1;            // refinement the kept padded lattice was derived for
    double scale_lattice = 1.0;      // ... and the widening of its cells (resident-order steps)
    int fine_request = 0;            // cdb_set_search_refinement: 0 automatic, 1 / 2 forced (2 only where it is valid)
    double ext_max = 1e300;          // bound on any agent's radius / body extent (k_ext_max at upload)
    unsigned long long *d_extmax = nullptr, *h_extmax = nullptr;
    bool tables_valid = false;
    int64_t cell_capacity = 0;
    int *d_cell_count = nullptr, *d_cell_start = nullptr, *d_cell_fill = nullptr;
    int *d_cell_of_slot = nullptr;   // flat cell of each slot (of `cur`, valid after build)
    int *d_order_tmp = nullptr, *d_order = nullptr;
    double *d_nbr = nullptr;         // packed neighbour records of the cell-sorted state
    double *d_nbr_sweep = nullptr;   // three-circle: compact 48 B records for the phase-1 sweep (circular: alias of d_nbr)
    int *d_scan_partials = nullptr;
    // once-per-pair evaluation (variant 3, pair_kernels.cuh)
    PairBuf pb{};                    // device pointers + capacity
    double2 *d_par = nullptr;        // {-mass * k_soc, tau_0} in cell order (+ ghost tail)
    unsigned long long *h_pctr = nullptr;   // pinned: [0] pairs found by the last step, [1] entries allocated, [2] device step counter,
                                            // [3] bits of ChainState::disp_last
    bool pairs_pending = false;      // steps were issued whose pair count has not been checked yet
    int64_t pair_cap_request = 0;    // cdb_set_pair_capacity (0: automatic)
    int64_t pair_overflows = 0;      // steps that had to be repeated after growing the list
    long long *d_bbox = nullptr;     // min ix, max ix, min iy, max iy
    long long *h_bbox = nullptr;     // pinned

    // resident-order steps (fused steps of variant 3 on one device, see ChainState in kernels.cuh): the block list is rebuilt
    // every `rebuild_every` steps only; in between the agents keep their slots and the step works in place
    bool chain_enabled = true;       // cdb_set_rebuild_policy
    bool policy_explicit = false;    // ... has been called (strips keep block lists only on request: every rank must agree)
    double skin_frac = 0.10;         // search cells are (1 + skin_frac) * cell_size / fine wide
    int rebuild_max = 16;            // upper bound of the rebuild interval (1: rebuild every step, the round-1 behaviour)
    int rebuild_every = 1;           // current interval, adapted to the observed per-step displacement
    int since_rebuild = 0;           // steps issued on the current block list
    bool chain_step = false;         // the step being issued runs in resident-order mode (set by issue_step)
    bool chain_inplace = false;      // ... on the kept order, in place
    bool chain_valid = false;        // `cur` is in the slot order of the tables; records and ChainState describe it
    double chain_cell_size = 0.0, chain_scale = 1.0;
    uint64_t chain_version = 0;      // state_version the kept order belongs to (uploads, node-wise calls ... end it)
    int64_t chain_min_agents = 16384;   // smaller crowds are launch-bound: they keep the CUDA-graph path
    int64_t chain_rebuilds = 0, chain_kept = 0, chain_stale = 0;   // statistics (cdb_get_rebuild_stats)
    ChainState *d_chain = nullptr;
    double drift_limit = 0.0;        // of the current block list: (coverage of the search cells - interaction range) / 2

    // a kept step is the same launch sequence on the same buffers every time: replayed as a CUDA graph (one per buffer parity)
    struct KeptGraph { const void *cur, *cells, *pairs; long long cap; uint32_t flags; double cell_size, dt_min, dt_max; const void *log;
                       int64_t n; uint64_t version; cudaGraphExec_t exec; int64_t launches; } kept_graph[4] = {};
    int kept_graph_next = 0;
    bool kept_graph_ok = true;       // false once this stream has refused a capture
    int64_t kept_graph_max_agents = 200000;   // larger crowds are not launch-bound
    int64_t small_max = SMALL_MAX;   // crowds up to this size are stepped by one thread block (small_kernel.cuh); 0: never

    // obstacles / navigation
    double *d_obstacles = nullptr;
    int64_t n_obstacles = 0;
    std::vector<NavField> nav;       // host copy (device pointers inside)
    NavField *d_nav = nullptr;
    int n_nav = 0;

    // reductions / time
    unsigned long long *d_vmax = nullptr;   // ordered-uint64 encodings: [0] max |v|, [1] max v0
    double *d_dt = nullptr;                 // [0] last dt, [1] time_tot
    double *h_dt = nullptr;                 // pinned double[2]
    double *d_dt_log = nullptr;             // dt of the last DT_LOG fused steps
    int *d_error = nullptr;
    int *h_error = nullptr;                 // pinned
    unsigned long long *d_pair_count = nullptr;
    int64_t iterations = 0;
    unsigned long long *d_stepctr = nullptr;   // device mirror of `iterations` (read by the kernels, advanced after every step)
    // CUDA graph of two consecutive fused steps (the ping-pong buffers are back in place after two)
    bool use_graphs = true;
    cudaGraphExec_t graph_exec = nullptr;
    int64_t graph_launches = 0;      // kernel launches inside one replay
    uint64_t state_version = 0;      // bumped by everything that invalidates a captured graph
    struct GraphKey { uint32_t flags; double cell_size, dt_min, dt_max; bool log; int64_t n; long long ncell, nx, ny; uint64_t version; const void *cur, *cells; } graph_key{};
    unsigned long long seed = 0x9E3779B97F4A7C15ULL;   // Fluctuation
    unsigned long long fluct_calls = 0;               // per-node Fluctuation calls (keeps successive calls independent)

    // strip decomposition
    bool strip = false;
    int has_left = 0, has_right = 0;
    // one-sided exchange over peer memory (cdb_strip_exchange_*): my receive buffers [0] from the left, [1] from the right
    double *x_halo_in[2] = {nullptr, nullptr}, *x_mig_in[2] = {nullptr, nullptr};
    unsigned long long *x_flags = nullptr;       // [0] halo from left, [1] halo from right, [2] migrants from left, [3] from right
    unsigned int *x_done = nullptr;              // arrival counters of my two halo-pack grids
    double *p_halo[2] = {nullptr, nullptr}, *p_mig[2] = {nullptr, nullptr};   // where I write: the neighbours' receive buffers
    unsigned long long *p_flags[2] = {nullptr, nullptr};                      // the neighbours' flag arrays
    void *p_opened[2][5] = {{nullptr}};          // IPC mappings to close
    bool x_connected = false;
    int strip_fine = 1;              // search refinement of the strip lattice (fixed by cdb_set_strip)
    double strip_scale = 1.0;        // widening of the strip's search cells (1 + skin when block lists are kept, cdb_set_strip)
    int strip_kind = 0;              // cdb_strip_set_kind: what the step being issued is (see include/crowd_b200.h)
    long long strip_ix0 = 0, strip_col_lo = 0, strip_col_hi = 0;   // cell_size columns: lattice origin, owned range
    int64_t halo_cap = 0, mig_cap = 0;
    int64_t n_dead = 0;              // slots vacated by migrants (dropped at the next sort)
    int64_t n_global = 0;            // agents of the WHOLE crowd (cdb_strip_set_global_agents): size of the per-id flag arrays
    int *d_counters = nullptr;       // [0] migrants left, [1] migrants right, [2] appended
    DevCounts *d_counts = nullptr;   // device-side slot / live counts: in strip mode the host only keeps upper bounds in n
    DevCounts *h_counts = nullptr;   // pinned
    bool dev_counts = false;
    int steps_since_refresh = 0;
    int *h_counters = nullptr;       // pinned

    // collective motion (section 8(f) rank 4): States fields by original agent index, scratch of the steering kernels
    int64_t states_n = 0;            // agents the arrays below describe (0: cdb_set_states not called yet)
    int64_t states_cap = 0;
    uint8_t *d_is_leader = nullptr, *d_is_follower = nullptr, *d_has_a = nullptr, *d_has_detected = nullptr;
    long long *d_index_leader = nullptr, *d_familiar_exit = nullptr, *d_target_by_id = nullptr, *d_detected = nullptr, *d_knn = nullptr;
    int64_t knn_cap = 0;
    int *d_slot_of_id = nullptr, *d_leader_ids = nullptr;
    int64_t n_leaders = 0;
    double *d_lrec = nullptr;        // {px, py, vx, vy} per leader
    int *d_lhead = nullptr, *d_lnext = nullptr;   // leader hash grid (linked lists)
    int64_t lhead_cap = 0;
    double *d_dir_a = nullptr, *d_direction = nullptr, *d_doors = nullptr;
    int64_t doors_cap = 0;
    bool direction_valid = false, detection_valid = false;
    // host-visible state nodes (rank 3): polygons [0] domain, [1] targets; flags by original agent index
    double *d_poly_xy[2] = {nullptr, nullptr};
    int *d_poly_off[2] = {nullptr, nullptr};
    int64_t n_polygons[2] = {0, 0};
    int64_t poly_nv[2] = {0, 0};     // vertices in total
    uint8_t *d_active = nullptr, *d_reached = nullptr;
    int64_t active_n = 0, reached_cap = 0, reached_n = 0;
    unsigned long long *d_poly_counts = nullptr;     // [0] InsideDomain changes of the last call, [1 + p] reached counts

    // instrumentation: kernel launch counter and CUDA-event timing of the step phases
    int64_t launches = 0;
    bool profiling = false;
    std::vector<cudaEvent_t> ev_pool;    // PROF_EVENTS events per profiled step
    size_t ev_used = 0;
};

static cudaError_t sync_stream(cdb_sim *sim) { sim->syncs++; return cudaStreamSynchronize(sim->stream); }

namespace {

int alloc_soa(Soa &s, int n_planes, int64_t capacity) {
    s.stride = (capacity + 31) / 32 * 32;
    s.np = n_planes;
    CKS(dev_alloc(&s.p, (size_t)n_planes * s.stride));
    CKS(dev_alloc(&s.id, s.stride));
    CKS(dev_alloc(&s.target, s.stride));
    return CDB_OK;
}

void free_soa(Soa &s) {
    cudaFree(s.p); cudaFree(s.id); cudaFree(s.target);
    s = Soa{};
}

int alloc_ghost_tail(cdb_sim *sim);
int ensure_pairs(cdb_sim *sim, int64_t cap);

int ensure_capacity(cdb_sim *sim, int64_t n) {
    if (n <= sim->capacity && sim->cur.p) return CDB_OK;
    int64_t cap = n < 1024 ? 1024 : n;
    free_soa(sim->cur); free_soa(sim->alt);
    CKS(alloc_soa(sim->cur, sim->n_alloc_planes, cap));
    CKS(alloc_soa(sim->alt, sim->n_alloc_planes, cap));
    sim->capacity = cap;
    CKS(dev_alloc(&sim->d_order, cap));
    return alloc_ghost_tail(sim);
}

// buffers that also hold the ghost agents of the two neighbour strips (slots [capacity, capacity + 2 * halo_cap))
int alloc_ghost_tail(cdb_sim *sim) {
    const size_t cap = (size_t)(sim->capacity + 2 * sim->halo_cap);
    CKS(dev_alloc(&sim->d_cell_of_slot, cap));
    CKS(dev_alloc(&sim->d_order_tmp, cap));
    CKS(dev_alloc(&sim->d_nbr, cap * (sim->model == CDB_MODEL_CIRCULAR ? REC_CIRC : REC_THREE)));
    if (sim->model == CDB_MODEL_CIRCULAR) sim->d_nbr_sweep = sim->d_nbr;
    else CKS(dev_alloc(&sim->d_nbr_sweep, cap * REC_CIRC));
    CKS(dev_alloc(&sim->d_par, cap));
    CKS(dev_alloc(&sim->pb.cnt, cap));
    CKS(dev_alloc(&sim->pb.off, cap));
    CKS(dev_alloc(&sim->pb.fill, cap));
    return CDB_OK;
}

// pair list + contribution array for `cap` pairs (contents are per-step scratch: nothing to preserve)
int ensure_pairs(cdb_sim *sim, int64_t cap) {
    if (cap <= sim->pb.cap && sim->pb.pairs) return CDB_OK;
    CK(sync_stream(sim));
    CKS(dev_alloc(&sim->pb.pairs, (size_t)cap));
    CKS(dev_alloc(&sim->pb.cres, (size_t)cap * 8));
    sim->pb.cap = cap;
    sim->state_version++;
    return CDB_OK;
}

int ensure_cells(cdb_sim *sim, int64_t ncell) {
    if (ncell <= sim->cell_capacity) return CDB_OK;
    int64_t cap = ncell + ncell / 4 + 1024;
    CKS(dev_alloc(&sim->d_cell_count, cap));
    CKS(dev_alloc(&sim->d_cell_start, cap + 1));
    CKS(dev_alloc(&sim->d_cell_fill, cap));
    CKS(dev_alloc(&sim->d_scan_partials, cap / SCAN_TILE + 2));
    sim->cell_capacity = cap;
    return CDB_OK;
}

int check_device_error(cdb_sim *sim) {
    CK(cudaMemcpyAsync(sim->h_error, sim->d_error, sizeof(int), cudaMemcpyDeviceToHost, sim->stream));
    CK(sync_stream(sim));
    int e = *sim->h_error;
    if (e == 0) return CDB_OK;
    CK(cudaMemsetAsync(sim->d_error, 0, sizeof(int), sim->stream));
    if (e == ERR_NONFINITE) return fail(CDB_ERR_INVALID_VALUE, "non-finite agent position encountered on device");
    if (e == ERR_CELL_RANGE) return fail(CDB_ERR_CAPACITY, "cell lattice too large (more than 2^31 cells)");
    if (e == ERR_CELL_RANGE + 5) return fail(CDB_ERR_CAPACITY, "more migrants arrived in one step than the host-side slot bound allows; ask for the exact count (cdb_strip_count / n_out) after bursts");
    if (e == ERR_PAIR_OVERFLOW) return fail(CDB_ERR_CAPACITY, "pair list overflow in strip mode (a strip step cannot be repeated): raise cdb_set_pair_capacity");
    return fail(CDB_ERR_CUDA, "device error flag %d", e);
}

// ---- instrumentation -----------------------------------------------------------------------------------------------
int prof_mark(cdb_sim *sim) {
    if (!sim->profiling) return CDB_OK;
    if (sim->ev_used == sim->ev_pool.size()) {
        cudaEvent_t e;
        CK(cudaEventCreate(&e));
        sim->ev_pool.push_back(e);
    }
    CK(cudaEventRecord(sim->ev_pool[sim->ev_used++], sim->stream));
    return CDB_OK;
}

// ---- block list --------------------------------------------------------------------------------------------------
// Search lattice refinement for the pair search.  Cells of cell_size / 2 (exact in binary: floor(p / (c/2)) >> 1 ==
// floor(p / c)) with a reach of 2 cells cover every pair closer than cell_size, all of which lie in the same or in adjacent
// cell_size cells -- so when no pair can interact beyond cell_size (3 + 2 max R < cell_size; true for every body type of the
// reference with its default cell_size = 3.6) the swept area drops from 9 c^2 to 6.25 c^2 with the same pair set.  The debug
// exports always report the cell_size lattice (core/block_list.py:28-52).
int search_refinement(const cdb_sim *sim, double cell_size) {
    if (sim->strip) return sim->strip_fine;
    if (sim->fine_request == 1 || sim->lattice_fixed || sim->variant == 1) return 1;
    // automatic: both models (round 2: with one call site per pair in k_pair_eval the less coherent pair list of the finer
    // lattice no longer costs the three-circle evaluation more than the sweep gains: 0.684 vs 0.711 ms / step at 1 M agents)
    return (SIGTH_SOC + 2.0 * sim->ext_max) * (1.0 + 1e-9) < cell_size ? 2 : 1;
}

int build_block_list(cdb_sim *sim, double cell_size, bool padded_lattice = false, unsigned long long *vmax = nullptr, bool physical = true,
                     int fine = 1, double scale = 1.0) {
    if (!(cell_size > 0.0) || !std::isfinite(cell_size)) return fail(CDB_ERR_INVALID_VALUE, "cell_size must be > 0");
    // bin size of the search lattice; scale > 1 (resident-order steps): wider cells, so that the list stays complete while the
    // agents drift -- the pair set is decided by distances, not by the lattice, whenever 3 + 2 max R < cell_size
    const double cs = cell_size * scale / fine;
    sim->chain_valid = false;                    // whoever rebuilds the tables ends a run of resident-order steps
    const int64_t n = sim->n;                   // slots in use, including the ones vacated by migrants (upper bound with dev_counts)
    const int64_t live = sim->dev_counts ? sim->n : sim->n - sim->n_dead;
    const int *slots_dev = sim->dev_counts ? &sim->d_counts->slots : nullptr;
    int *live_dev = sim->dev_counts ? &sim->d_counts->live : nullptr;
    sim->cell_size = cell_size;
    sim->tables_valid = false;
    cudaStream_t st = sim->stream;
    if (n == 0 && !sim->strip) {
        sim->grid = Grid{0, 0, 0, 0, 0, 0, 0};
        CK(cudaMemcpyAsync(sim->d_grid, &sim->grid, sizeof(Grid), cudaMemcpyHostToDevice, st));
        sim->tables_valid = true;
        sim->n_sorted = 0;
        sim->perm_valid = false;
        return CDB_OK;
    }
    const int T = 256;
    // padded_lattice (fused steps): derive the lattice from the bounding box once, pad it by one cell and keep it for the
    // next steps -- cells stay anchored at multiples of cell_size and agents that leave the lattice are binned into its
    // border cells (adjacency is preserved by clamping), so neighbour sets and forces do not depend on this choice.
    const bool reuse = padded_lattice && sim->auto_lattice_valid && sim->cell_size_lattice == cell_size && sim->fine_lattice == fine &&
                       sim->scale_lattice == scale && sim->auto_lattice_age < 64;
    sim->fine = fine;
    if (reuse) sim->auto_lattice_age++;
    if (!sim->lattice_fixed && !reuse) {
        LAUNCH(sim, k_bbox_init, 1, 32, 0, sim->d_bbox);
        LAUNCH(sim, k_bbox, (cdiv(n, T * 4) < 1184 ? cdiv(n, T * 4) : 1184), T, 0, sim->cur, (int)n, cs, sim->d_bbox, sim->d_error);
        CK(cudaMemcpyAsync(sim->h_bbox, sim->d_bbox, 4 * sizeof(long long), cudaMemcpyDeviceToHost, st));
        CK(sync_stream(sim));
        CKS(check_device_error(sim));
        long long x0 = sim->h_bbox[0], x1 = sim->h_bbox[1], y0 = sim->h_bbox[2], y1 = sim->h_bbox[3];
        long long nx = x1 - x0 + 1, ny = y1 - y0 + 1;
        if (nx <= 0 || ny <= 0 || (double)nx * (double)ny > 2.0e9)
            return fail(CDB_ERR_CAPACITY, "block list of %lld x %lld cells is too large", nx, ny);
        if (padded_lattice) { x0 -= fine; y0 -= fine; nx += 2 * fine; ny += 2 * fine; }
        sim->grid = Grid{x0, y0, nx, ny, nx * ny, 0, nx - 1};
        sim->auto_lattice_valid = padded_lattice;
        sim->auto_lattice_age = 0;
        sim->cell_size_lattice = cell_size;
        sim->fine_lattice = fine;
        sim->scale_lattice = scale;
        CK(cudaMemcpyAsync(sim->d_grid, &sim->grid, sizeof(Grid), cudaMemcpyHostToDevice, st));
    }
    const int64_t ncell = sim->grid.ncell;
    CKS(ensure_cells(sim, ncell));
    CK(cudaMemsetAsync(sim->d_cell_count, 0, ncell * sizeof(int), st));
    CK(cudaMemsetAsync(sim->d_cell_fill, 0, ncell * sizeof(int), st));
    if (n > 0) LAUNCH(sim, k_cell_count, cdiv(n, T), T, 0, sim->cur, (int)n, cs, sim->d_grid, sim->d_cell_of_slot, sim->d_cell_count, sim->d_error, vmax, slots_dev);
    // exclusive scan count -> start
    const int nblk = cdiv(ncell, SCAN_TILE);
    LAUNCH(sim, k_scan_tiles, nblk, SCAN_THREADS, 0, sim->d_cell_count, sim->d_cell_start, (int)ncell, sim->d_scan_partials);
    LAUNCH(sim, k_scan_partials, 1, 1024, 0, sim->d_scan_partials, nblk);
    LAUNCH(sim, k_scan_add, nblk, SCAN_THREADS, 0, sim->d_cell_start, (int)ncell, sim->d_scan_partials, sim->d_cell_count, live_dev);
    if (n > 0) LAUNCH(sim, k_scatter, cdiv(n, T), T, 0, sim->d_cell_of_slot, (int)n, sim->d_cell_start, sim->d_cell_fill, sim->d_order_tmp, slots_dev);
    if (live > 0) LAUNCH(sim, k_rank_fix, cdiv(live, T), T, 0, sim->d_order_tmp, (int)live, sim->cur.id, sim->d_cell_of_slot, sim->d_cell_start,
                                         sim->d_cell_count, sim->d_order, live_dev);
    if (physical) {
        if (live > 0) LAUNCH(sim, k_gather, cdiv(live, T), T, 0, sim->cur, sim->alt, (int)live, sim->n_planes, sim->model,
                             sim->d_order, sim->d_cell_of_slot, sim->d_order_tmp, sim->d_nbr, sim->d_nbr_sweep, cell_size);
        // d_order_tmp now holds the flat cell of every *sorted* slot
        std::swap(sim->cur, sim->alt);
        std::swap(sim->d_cell_of_slot, sim->d_order_tmp);
        sim->n = live;
        sim->n_dead = 0;
        sim->perm_valid = false;
    } else {
        if (live > 0) LAUNCH(sim, k_records, cdiv(live, T), T, 0, sim->cur, (int)live, live_dev, sim->model, sim->d_order, sim->d_cell_of_slot,
                             sim->d_order_tmp, sim->d_nbr, sim->d_nbr_sweep, cell_size, sim->d_par);
        std::swap(sim->d_cell_of_slot, sim->d_order_tmp);   // d_cell_of_slot: flat cell per sorted slot; d_order_tmp: per plane slot
        sim->perm_valid = true;
    }
    CK(cudaGetLastError());
    sim->n_sorted = live;
    sim->tables_valid = true;
    return CDB_OK;
}

int launch_reduce_vmax(cdb_sim *sim) {
    LAUNCH(sim, k_vmax_init, 1, 32, 0, sim->d_vmax);
    if (sim->n > 0) {
        int blocks = cdiv(sim->n, 256 * 4);
        if (blocks > 1184) blocks = 1184;
        LAUNCH(sim, k_vmax, blocks, 256, 0, sim->cur, (int)sim->n, sim->d_vmax);
    }
    CK(cudaGetLastError());
    return CDB_OK;
}

int node_reset(cdb_sim *sim) {
    if (sim->n) LAUNCH(sim, k_reset, cdiv(sim->n, 256), 256, 0, sim->cur, (int)sim->n, sim->model);
    CK(cudaGetLastError());
    return CDB_OK;
}
int node_navigation(cdb_sim *sim) {
    if (sim->n && sim->n_nav) LAUNCH(sim, k_navigation, cdiv(sim->n, 256), 256, 0, sim->cur, (int)sim->n, sim->d_nav, sim->n_nav);
    CK(cudaGetLastError());
    return CDB_OK;
}
int node_orientation(cdb_sim *sim) {
    if (sim->n && sim->model == CDB_MODEL_THREE_CIRCLE) LAUNCH(sim, k_orientation, cdiv(sim->n, 256), 256, 0, sim->cur, (int)sim->n);
    CK(cudaGetLastError());
    return CDB_OK;
}
int node_adjust(cdb_sim *sim) {
    if (sim->n) LAUNCH(sim, k_adjust, cdiv(sim->n, 256), 256, 0, sim->cur, (int)sim->n, sim->model);
    CK(cudaGetLastError());
    return CDB_OK;
}
StepArgs step_args(cdb_sim *sim, unsigned flags, double dt_min, double dt_max, double *dt_log) {
    StepArgs a{};
    a.in = sim->cur;
    a.out = (flags & CDB_STEP_INTEGRATOR) ? sim->alt : sim->cur;
    a.nbr = sim->d_nbr;
    a.nbr_sweep = sim->d_nbr_sweep;
    a.cell_size = sim->cell_size;
    const bool listed = (flags & CDB_STEP_AGENT_AGENT) != 0;
    a.n = (int)(listed ? sim->n_sorted : sim->n);
    a.order = listed && sim->perm_valid ? sim->d_order : nullptr;
    a.n_dev = listed && sim->dev_counts ? &sim->d_counts->live : (sim->dev_counts ? &sim->d_counts->slots : nullptr);
    a.grid = sim->d_grid;
    a.cell_sorted = sim->d_cell_of_slot; a.cell_start = sim->d_cell_start; a.cell_count = sim->d_cell_count;
    a.nav = sim->d_nav; a.n_nav = sim->n_nav;
    a.obs = sim->d_obstacles; a.n_obs = (int)sim->n_obstacles;
    a.flags = flags;
    a.dt_min = dt_min; a.dt_max = dt_max;
    a.vmax = sim->d_vmax; a.dt_out = sim->d_dt; a.dt_log = dt_log;
    a.seed = sim->seed; a.step_ptr = sim->d_stepctr;
    return a;
}

inline bool use_pairs(const cdb_sim *sim) { return sim->variant == 3; }

// allocations of variant 3, made OUTSIDE stream capture (before steps are issued): 8 pairs per agent to start with, grown
// by settle_pairs() when a step gets close to or beyond the capacity
int prepare_pairs(cdb_sim *sim) {
    if (!use_pairs(sim)) return CDB_OK;
    if (!sim->pb.ctr) {
        CKS(dev_alloc(&sim->pb.ctr, 2));
        CK(cudaMallocHost((void **)&sim->h_pctr, 4 * sizeof(unsigned long long)));
        sim->h_pctr[0] = sim->h_pctr[1] = sim->h_pctr[2] = sim->h_pctr[3] = 0;
    }
    if (sim->pair_cap_request > 0) return ensure_pairs(sim, sim->pair_cap_request);
    const int64_t want = std::max<int64_t>((sim->strip || sim->defer_sync ? 16 : 8) * std::max(sim->n, sim->capacity / 2), 1 << 16);
    if (!sim->pb.pairs || sim->pb.cap < want / 2) CKS(ensure_pairs(sim, want));
    return CDB_OK;
}

// variant 3: classify every unordered pair of adjacent cells once, evaluate the survivors once (pair_kernels.cuh)
int launch_pairs(cdb_sim *sim) {
    const int64_t n = sim->n_sorted;
    if (n <= 0) return CDB_OK;
    if (!sim->pb.pairs || !sim->pb.ctr) return fail(CDB_ERR_STATE, "pair buffers not prepared");
    cudaStream_t st = sim->stream;
    CK(cudaMemsetAsync(sim->pb.ctr, 0, 2 * sizeof(unsigned long long), st));
    CK(cudaMemsetAsync(sim->pb.cnt, 0, (size_t)n * sizeof(int), st));
    SweepArgs a{};
    a.nbr_sweep = sim->d_nbr_sweep;
    a.n = (int)n;
    a.n_dev = sim->dev_counts ? &sim->d_counts->live : nullptr;
    a.grid = sim->d_grid;
    a.cell_sorted = sim->d_cell_of_slot; a.cell_start = sim->d_cell_start; a.cell_count = sim->d_cell_count;
    a.ghost_base = -1; a.n_ghost = 0; a.ghost_cells = 0;
    if (sim->strip) {
        a.ghost_base = (int)sim->capacity;
        a.n_ghost = sim->has_left ? (int)sim->halo_cap : 0;
        a.ghost_cells = (int)(sim->grid.cx_lo * sim->grid.ny);
    }
    a.reach = sim->fine;
    a.pb = sim->pb;
    a.pb.fatal = sim->strip ? sim->d_error : nullptr;
    a.chain = sim->chain_step ? sim->d_chain : nullptr;
    a.drift_limit = sim->drift_limit;
    const int blocks = cdiv(n + a.n_ghost, SW_THREADS);
    if (sim->sweep_mode == 1 && !sim->strip) {
        // candidates staged in shared memory by TMA bulk copies (k_sweep_staged); strips have ghost targets: plain kernel
        if (sim->model == CDB_MODEL_CIRCULAR) LAUNCH(sim, k_sweep_staged<0>, blocks, SW_THREADS, 0, a);
        else LAUNCH(sim, k_sweep_staged<1>, blocks, SW_THREADS, 0, a);
    } else if (sim->model == CDB_MODEL_CIRCULAR) LAUNCH(sim, k_sweep<0>, blocks, SW_THREADS, 0, a);
    else LAUNCH(sim, k_sweep<1>, blocks, SW_THREADS, 0, a);
    LAUNCH(sim, k_pair_alloc, cdiv(n, 256), 256, 0, a.pb, (int)n);
    CKS(prof_mark(sim));
    EvalArgs e{};
    e.nbr = sim->d_nbr; e.par = sim->d_par; e.in = sim->cur;
    e.order = sim->perm_valid ? sim->d_order : nullptr;
    e.ghost_base = 0x7fffffff; e.ghost_left_end = 0x7fffffff;
    if (sim->strip) { e.ghost_base = (int)sim->capacity; e.ghost_left_end = (int)(sim->capacity + sim->halo_cap); }
    e.pb = sim->pb;
    // persistent grid-stride launch: the number of pairs is only known on the device
    const int eval_blocks = (int)std::min<int64_t>(cdiv(std::max<int64_t>(4 * n, 128), 128), (int64_t)sim->sm_count * (sim->model == CDB_MODEL_CIRCULAR ? 16 : 8));
    if (sim->model == CDB_MODEL_CIRCULAR) LAUNCH(sim, k_pair_eval<0>, eval_blocks, 128, 0, e);
    else LAUNCH(sim, k_pair_eval<1>, eval_blocks, 128, 0, e);
    CKS(prof_mark(sim));
    CK(cudaGetLastError());
    sim->pairs_pending = true;
    return CDB_OK;
}

// the fused kernel; requires a current block list when CDB_STEP_AGENT_AGENT is selected
int launch_step_kernel(cdb_sim *sim, unsigned flags, double dt_min, double dt_max, double *dt_log, const MigrantArgs *mig = nullptr) {
    const bool pairs = use_pairs(sim);
    if (pairs && (flags & CDB_STEP_AGENT_AGENT)) CKS(launch_pairs(sim));
    else { CKS(prof_mark(sim)); CKS(prof_mark(sim)); }
    StepArgs a = step_args(sim, flags, dt_min, dt_max, dt_log);
    a.pb = sim->pb;
    a.n_planes = sim->n_planes;
    a.reach = sim->fine;
    a.mig = MigrantArgs{};
    if (mig && pairs && (flags & CDB_STEP_INTEGRATOR)) a.mig = *mig;
    a.chain = nullptr; a.rec_nbr = nullptr; a.rec_sweep = nullptr; a.inplace = 0;
    if (sim->chain_step) {
        a.chain = sim->d_chain; a.rec_nbr = sim->d_nbr; a.rec_sweep = sim->d_nbr_sweep;
        if (sim->chain_inplace) { a.inplace = 1; a.out = sim->cur; a.order = nullptr; }
    }
    const int smem = 0;
    if (a.n > 0) {
        if (pairs) {
            const int ft = a.inplace ? FIN_THREADS_INPLACE : FIN_THREADS;
            const int fb = cdiv(a.n, ft);
            if (sim->model == CDB_MODEL_CIRCULAR) {
                if (a.inplace) LAUNCH(sim, (k_finish<0, true>), fb, ft, smem, a);
                else LAUNCH(sim, (k_finish<0, false>), fb, ft, smem, a);
            } else {
                if (a.inplace) LAUNCH(sim, (k_finish<1, true>), fb, ft, smem, a);
                else LAUNCH(sim, (k_finish<1, false>), fb, ft, smem, a);
            }
        } else {
            if (sim->model == CDB_MODEL_CIRCULAR) LAUNCH(sim, k_step<0>, cdiv(a.n, STEP_THREADS), STEP_THREADS, smem, a);
            else LAUNCH(sim, k_step<1>, cdiv(a.n, STEP_THREADS), STEP_THREADS, smem, a);
        }
    } else if (flags & CDB_STEP_INTEGRATOR) {
        LAUNCH(sim, k_integrate, 1, 32, 0, sim->cur, 0, sim->model, dt_min, dt_max, sim->d_vmax, sim->d_dt);
    }
    CK(cudaGetLastError());
    if (flags & CDB_STEP_INTEGRATOR) {
        if (!a.inplace) {
            std::swap(sim->cur, sim->alt);
            if (a.order) {   // the step wrote the live agents compacted, in cell order
                sim->n = a.n; sim->n_dead = 0;
                if (sim->dev_counts) LAUNCH(sim, k_counts_after_step, 1, 32, 0, sim->d_counts);
            }
        }
        sim->perm_valid = false;
        sim->tables_valid = false;
    }
    return CDB_OK;
}

// Host side of the "a step whose pairs did not fit is not applied" protocol: one synchronisation; *overflow says whether
// the most recent step found more pairs than the list holds (then the list has been grown), *dev_steps is the number of
// integrating steps the device has really applied.
int analyze_pairs(cdb_sim *sim, bool *overflow, int64_t *dev_steps) {
    const int64_t found = (int64_t)sim->h_pctr[0];
    if (dev_steps) *dev_steps = (int64_t)sim->h_pctr[2];
    // resident-order steps: size the rebuild interval so that the drift bound (sum of the per-step maximum displacements)
    // stays below the slack of the search cells, with a margin for crowds that speed up
    {
        double last; memcpy(&last, &sim->h_pctr[3], sizeof(last));
        if (sim->drift_limit > 0.0 && last > 0.0 && std::isfinite(last)) {
            const double k = sim->drift_limit / (1.25 * last);
            sim->rebuild_every = (int)std::max(1.0, std::min((double)sim->rebuild_max, std::floor(k)));
        } else if (!(last >= 0.0) || !std::isfinite(last)) {
            sim->rebuild_every = 1;
        }
    }
    if ((unsigned long long)found >= CHAIN_STALE) {
        // a step on the kept order found the search lattice stale and was not applied: rebuild, and more often from now on
        *overflow = true;
        sim->chain_valid = false;
        sim->chain_stale++;
        sim->rebuild_every = std::max(1, sim->rebuild_every / 2);
        return CDB_OK;
    }
    sim->pairs_known_found = found;
    sim->pairs_known_age = 0;
    if (found > sim->pb.cap) {
        *overflow = true;
        sim->chain_valid = false;
        sim->pair_overflows++;
        sim->pair_cap_request = 0;
        CKS(ensure_pairs(sim, std::max<int64_t>(2 * sim->pb.cap, found + found / 2)));
    } else if (found > sim->pb.cap / 2 && sim->pair_cap_request == 0) {
        CKS(ensure_pairs(sim, 2 * sim->pb.cap));     // the crowd is getting denser: grow before it overflows
    }
    return CDB_OK;
}

// blocking read of the pair counters, the device step counter and the last displacement (one synchronisation)
int read_counters(cdb_sim *sim) {
    CK(cudaMemcpyAsync(sim->h_pctr, sim->pb.ctr, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, sim->stream));
    CK(cudaMemcpyAsync(sim->h_pctr + 2, sim->d_stepctr, sizeof(unsigned long long), cudaMemcpyDeviceToHost, sim->stream));
    CK(cudaMemcpyAsync(sim->h_pctr + 3, &sim->d_chain->disp_last, sizeof(double), cudaMemcpyDeviceToHost, sim->stream));
    return CDB_OK;
}

// deferred mode: pick up a check whose copy has completed since it was issued (never waits).  *overflow also reports a step
// that was refused before the check was queued (fewer steps applied than the host had issued by then); the caller then
// synchronises and reads the counters again.
int harvest_pairs(cdb_sim *sim, bool *overflow, int64_t *dev_steps) {
    if (!sim->pairs_inflight) return CDB_OK;
    if (cudaEventQuery(sim->ev_pairs) != cudaSuccess) { cudaGetLastError(); return CDB_OK; }
    sim->pairs_inflight = false;
    int64_t seen = sim->iterations_at_check;
    CKS(analyze_pairs(sim, overflow, &seen));
    if (seen < sim->iterations_at_check) *overflow = true;
    if (*overflow) {
        // what has been applied by NOW (steps issued after the check may or may not have been)
        CKS(read_counters(sim));
        CK(sync_stream(sim));
        if (dev_steps) *dev_steps = (int64_t)sim->h_pctr[2];
        sim->chain_valid = false;
        sim->pairs_pending = false;
    }
    return CDB_OK;
}

int settle_pairs(cdb_sim *sim, bool *overflow, int64_t *dev_steps, bool may_defer = false) {
    *overflow = false;
    if (dev_steps) *dev_steps = sim->iterations;
    if (!sim->pairs_pending || !sim->pb.ctr) return CDB_OK;
    if (sim->defer_sync && may_defer) {
        // Deferred mode (cdb_set_deferred_sync): do not wait for this call's own check while the list is known to have
        // ample room -- a recent check found it at most a quarter full, and a crowd cannot get four times denser within a
        // few steps (agents move about a centimetre per step).  The check is read once its copy has completed.
        CKS(harvest_pairs(sim, overflow, dev_steps));
        if (*overflow) return CDB_OK;
        const bool safe = sim->pairs_known_found >= 0 && sim->pairs_known_found <= sim->pb.cap / 4 && sim->pairs_known_age < 16;
        if (safe) {
            if (!sim->pairs_inflight) {
                CKS(read_counters(sim));
                CK(cudaEventRecord(sim->ev_pairs, sim->stream));
                sim->pairs_inflight = true;
                sim->iterations_at_check = sim->iterations;
            }
            sim->pairs_known_age++;
            return CDB_OK;
        }
    }
    if (sim->pairs_inflight) { CK(sync_stream(sim)); sim->pairs_inflight = false; }     // its slot is reused below
    CKS(read_counters(sim));
    CK(sync_stream(sim));
    sim->pairs_pending = false;
    return analyze_pairs(sim, overflow, dev_steps);
}

int launch_agent_agent(cdb_sim *sim) {
    if (sim->n == 0) return CDB_OK;
    if (sim->variant != 1) return launch_step_kernel(sim, CDB_STEP_AGENT_AGENT, 0.0, 0.0, nullptr);
    const int T = 128;
    if (sim->model == CDB_MODEL_CIRCULAR)
        LAUNCH(sim, k_agent_agent_circular_v1, cdiv(sim->n, T), T, 0, sim->cur, (int)sim->n, sim->d_grid, sim->d_cell_of_slot,
                                                                          sim->d_cell_start, sim->d_cell_count);
    else
        LAUNCH(sim, k_agent_agent_three_circle_v1, cdiv(sim->n, T), T, 0, sim->cur, (int)sim->n, sim->d_grid, sim->d_cell_of_slot,
                                                                              sim->d_cell_start, sim->d_cell_count);
    CK(cudaGetLastError());
    return CDB_OK;
}
int node_agent_agent(cdb_sim *sim, double cell_size) {
    CKS(prepare_pairs(sim));
    CKS(build_block_list(sim, cell_size, false, nullptr, sim->variant == 1, search_refinement(sim, cell_size)));
    for (int attempt = 0; attempt < 8; ++attempt) {
        CKS(launch_agent_agent(sim));
        if (!use_pairs(sim)) return CDB_OK;
        bool overflow = false;
        CKS(settle_pairs(sim, &overflow, nullptr));
        if (!overflow) return CDB_OK;      // otherwise nothing was applied: the list has been grown, run the node again
    }
    return fail(CDB_ERR_CAPACITY, "pair list keeps overflowing");
}
int node_agent_obstacle(cdb_sim *sim) {
    if (sim->n && sim->n_obstacles)
        LAUNCH(sim, k_agent_obstacle, cdiv(sim->n, 128), 128, 0, sim->cur, (int)sim->n, sim->model, sim->d_obstacles, (int)sim->n_obstacles);
    CK(cudaGetLastError());
    return CDB_OK;
}
int node_integrate(cdb_sim *sim, double dt_min, double dt_max) {
    CKS(launch_reduce_vmax(sim));
    LAUNCH(sim, k_integrate, (sim->n ? cdiv(sim->n, 256) : 1), 256, 0, sim->cur, (int)sim->n, sim->model, dt_min, dt_max, sim->d_vmax, sim->d_dt);
    CK(cudaGetLastError());
    sim->tables_valid = false;   // positions moved
    return CDB_OK;
}

}  // namespace

// =====================================================================================================================
extern "C" {

const char *cdb_last_error(void) { return g_err.c_str(); }
int cdb_version(void) { return 100; }

int cdb_device_count(int *count) {
    if (!count) return fail(CDB_ERR_INVALID_VALUE, "count is NULL");
    CK(cudaGetDeviceCount(count));
    return CDB_OK;
}

int cdb_measure_fp64_peak(int device, double *tflops) {
    if (!tflops) return fail(CDB_ERR_INVALID_VALUE, "tflops is NULL");
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    double *d = nullptr;
    CK(cudaMalloc((void **)&d, sizeof(double)));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 15;
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
        CK(cudaEventRecord(e0));
        k_dfma_peak<<<blocks, threads>>>(d, iters, 1.0 + rep);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        const double flops = 2.0 * 8.0 * (double)iters * blocks * threads;
        if (rep > 0 && flops / (ms * 1e-3) > best) best = flops / (ms * 1e-3);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    *tflops = best * 1e-12;
    return CDB_OK;
}

int cdb_create(int device, int model, int64_t capacity, cdb_sim **out) {
    if (!out) return fail(CDB_ERR_INVALID_VALUE, "out is NULL");
    *out = nullptr;
    if (model != CDB_MODEL_CIRCULAR && model != CDB_MODEL_THREE_CIRCLE) return fail(CDB_ERR_INVALID_TYPE, "unknown agent model %d", model);
    if (capacity < 0) return fail(CDB_ERR_INVALID_VALUE, "negative capacity");
    CK(cudaSetDevice(device));
    cdb_sim *sim = new cdb_sim();
    sim->device = device;
    cudaDeviceGetAttribute(&sim->sm_count, cudaDevAttrMultiProcessorCount, device);
    sim->model = model;
    sim->itemsize = model == CDB_MODEL_CIRCULAR ? 228 : 316;
    sim->n_planes = model == CDB_MODEL_CIRCULAR ? NP_CIRC : NP_THREE;
    sim->n_alloc_planes = sim->n_planes;
    if (const char *e = getenv("CROWD_B200_SWEEP")) sim->sweep_mode = strcmp(e, "staged") == 0 ? 1 : (strcmp(e, "plain") == 0 ? 0 : sim->sweep_mode);
    if (sim->sweep_mode) {     // 36 KB of static shared memory per CTA: ask for the largest carve-out so that SWS_MINB CTAs fit an SM
        if (model == CDB_MODEL_CIRCULAR) cudaFuncSetAttribute(k_sweep_staged<0>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        else cudaFuncSetAttribute(k_sweep_staged<1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaGetLastError();
    }
    int rc = [&]() -> int {
        CK(cudaStreamCreateWithFlags(&sim->stream, cudaStreamNonBlocking));
        sim->own_stream = true;
        CKS(ensure_capacity(sim, capacity));
        CKS(dev_alloc(&sim->d_grid, 1));
        CKS(dev_alloc(&sim->d_bbox, 4));
        CKS(dev_alloc(&sim->d_vmax, 2));
        CKS(dev_alloc(&sim->d_dt, 2));
        CKS(dev_alloc(&sim->d_dt_log, DT_LOG));
        CKS(dev_alloc(&sim->d_stepctr, 1));
        CKS(dev_alloc(&sim->d_chain, 1));
        CK(cudaMemset(sim->d_chain, 0, sizeof(ChainState)));
        CK(cudaStreamCreateWithFlags(&sim->side, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&sim->ev_main, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&sim->ev_pairs, cudaEventDisableTiming));
        for (int k = 0; k < 2; ++k) {
            CK(cudaEventCreateWithFlags(&sim->ev_snap[k], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&sim->ev_scal[k], cudaEventDisableTiming));
            CK(cudaMallocHost((void **)&sim->h_scal[k], (2 + 1 + 64) * sizeof(unsigned long long)));
        }
        CKS(dev_alloc(&sim->d_scal, 2 + 1 + 64));
        CKS(dev_alloc(&sim->d_extmax, 1));
        CK(cudaMallocHost((void **)&sim->h_extmax, sizeof(unsigned long long)));
        CK(cudaMemset(sim->d_stepctr, 0, sizeof(unsigned long long)));
        CKS(dev_alloc(&sim->d_error, 1));
        CKS(dev_alloc(&sim->d_pair_count, 1));
        CKS(dev_alloc(&sim->d_counters, 4));
        CKS(dev_alloc(&sim->d_counts, 1));
        CK(cudaMallocHost((void **)&sim->h_counts, sizeof(DevCounts)));
        CK(cudaMallocHost((void **)&sim->h_counters, 4 * sizeof(int)));
        CKS(dev_alloc(&sim->d_nav, MAX_NAV_TARGETS));
        CK(cudaMemset(sim->d_error, 0, sizeof(int)));
        CK(cudaMemset(sim->d_counters, 0, 4 * sizeof(int)));
        CK(cudaMemset(sim->d_counts, 0, sizeof(DevCounts)));
        CK(cudaMemset(sim->d_dt, 0, 2 * sizeof(double)));
        CK(cudaMallocHost((void **)&sim->h_bbox, 4 * sizeof(long long)));
        CK(cudaMallocHost((void **)&sim->h_dt, 2 * sizeof(double)));
        CK(cudaMallocHost((void **)&sim->h_error, sizeof(int)));
        return CDB_OK;
    }();
    if (rc != CDB_OK) { cdb_destroy(sim); return rc; }
    *out = sim;
    return CDB_OK;
}

int cdb_destroy(cdb_sim *sim) {
    if (!sim) return CDB_OK;
    cudaSetDevice(sim->device);
    if (sim->stream) cudaStreamSynchronize(sim->stream);
    free_soa(sim->cur); free_soa(sim->alt);
    cudaFree(sim->d_aos); cudaFreeHost(sim->h_bounce); cudaFree(sim->d_rec_slot);
    for (auto &r : sim->registered) cudaHostUnregister(r.first);
    for (int k = 0; k < 2; ++k) {
        for (int j = 0; j < 5; ++j) if (sim->p_opened[k][j]) cudaIpcCloseMemHandle(sim->p_opened[k][j]);
        cudaFree(sim->x_halo_in[k]); cudaFree(sim->x_mig_in[k]);
    }
    cudaFree(sim->x_flags); cudaFree(sim->x_done);
    if (sim->side) { cudaStreamSynchronize(sim->side); cudaStreamDestroy(sim->side); }
    for (int k = 0; k < 2; ++k) {
        if (sim->ev_snap[k]) cudaEventDestroy(sim->ev_snap[k]);
        if (sim->ev_scal[k]) cudaEventDestroy(sim->ev_scal[k]);
        cudaFree(sim->d_snap[k]); cudaFreeHost(sim->h_snap[k]); cudaFreeHost(sim->h_scal[k]);
    }
    if (sim->ev_main) cudaEventDestroy(sim->ev_main);
    if (sim->ev_pairs) cudaEventDestroy(sim->ev_pairs);
    cudaFree(sim->d_scal); cudaFree(sim->d_chain);
    for (auto &g : sim->kept_graph) if (g.exec) cudaGraphExecDestroy(g.exec);
    cudaFree(sim->d_grid); cudaFree(sim->d_cell_count); cudaFree(sim->d_cell_start); cudaFree(sim->d_cell_fill);
    cudaFree(sim->d_cell_of_slot); cudaFree(sim->d_order_tmp); cudaFree(sim->d_order); if (sim->d_nbr_sweep != sim->d_nbr) cudaFree(sim->d_nbr_sweep); cudaFree(sim->d_nbr); cudaFree(sim->d_scan_partials);
    cudaFree(sim->d_bbox); cudaFreeHost(sim->h_bbox);
    cudaFree(sim->d_par); cudaFree(sim->pb.pairs); cudaFree(sim->pb.cres); cudaFree(sim->pb.cnt); cudaFree(sim->pb.off); cudaFree(sim->pb.fill);
    cudaFree(sim->pb.ctr); cudaFreeHost(sim->h_pctr);
    cudaFree(sim->d_obstacles);
    for (auto &f : sim->nav) { cudaFree((void *)f.U); cudaFree((void *)f.V); }
    cudaFree(sim->d_is_leader); cudaFree(sim->d_is_follower); cudaFree(sim->d_has_a); cudaFree(sim->d_has_detected);
    cudaFree(sim->d_index_leader); cudaFree(sim->d_familiar_exit); cudaFree(sim->d_target_by_id); cudaFree(sim->d_detected);
    cudaFree(sim->d_knn); cudaFree(sim->d_slot_of_id); cudaFree(sim->d_leader_ids); cudaFree(sim->d_dir_a); cudaFree(sim->d_direction);
    cudaFree(sim->d_poly_xy[0]); cudaFree(sim->d_poly_xy[1]); cudaFree(sim->d_poly_off[0]); cudaFree(sim->d_poly_off[1]);
    cudaFree(sim->d_active); cudaFree(sim->d_reached); cudaFree(sim->d_poly_counts);
    cudaFree(sim->d_doors); cudaFree(sim->d_lrec); cudaFree(sim->d_lhead); cudaFree(sim->d_lnext);
    cudaFree(sim->d_nav); cudaFree(sim->d_vmax); cudaFree(sim->d_dt); cudaFree(sim->d_dt_log); cudaFreeHost(sim->h_dt);
    for (auto e : sim->ev_pool) cudaEventDestroy(e);
    if (sim->graph_exec) cudaGraphExecDestroy(sim->graph_exec);
    cudaFree(sim->d_stepctr); cudaFree(sim->d_extmax); cudaFreeHost(sim->h_extmax);
    cudaFree(sim->d_error); cudaFreeHost(sim->h_error); cudaFree(sim->d_pair_count);
    cudaFree(sim->d_counters); cudaFreeHost(sim->h_counters); cudaFree(sim->d_counts); cudaFreeHost(sim->h_counts);
    if (sim->own_stream && sim->stream) cudaStreamDestroy(sim->stream);
    delete sim;
    return CDB_OK;
}

int cdb_set_stream(cdb_sim *sim, void *cuda_stream) {
    if (!sim) return fail(CDB_ERR_INVALID_VALUE, "sim is NULL");
    sim->state_version++;
    CK(cudaSetDevice(sim->device));
    CK(sync_stream(sim));
    if (sim->own_stream) { cudaStreamDestroy(sim->stream); sim->own_stream = false; }
    sim->stream = (cudaStream_t)cuda_stream;
    return CDB_OK;
}

// Deferred synchronisation (cdb_set_deferred_sync) leaves the check of the last cdb_step call unread.  Entry points that hand
// state to the host read it first and, if the device refused steps (pair list too small, search lattice stale), issue them
// again with the parameters of that call -- the host never sees a state that is behind the step count it was told.
static int catch_up(cdb_sim *sim) {
    if (!sim->defer_sync || !(sim->pairs_pending || sim->pairs_inflight) || !sim->pb.ctr) return CDB_OK;
    if (sim->pairs_inflight) { CK(sync_stream(sim)); sim->pairs_inflight = false; }
    CKS(read_counters(sim));
    CK(sync_stream(sim));
    sim->pairs_pending = false;
    bool overflow = false;
    int64_t dev_steps = sim->iterations;
    CKS(analyze_pairs(sim, &overflow, &dev_steps));
    const int64_t missing = sim->iterations - dev_steps;
    if (!overflow && missing <= 0) return CDB_OK;
    if (!overflow) { sim->chain_stale++; sim->rebuild_every = std::max(1, sim->rebuild_every / 2); }
    sim->chain_valid = false;
    sim->iterations = dev_steps;
    sim->defer_sync = false;             // the repeated steps are checked before this returns
    const int rc = cdb_step(sim, sim->last_flags, sim->last_cell_size, sim->last_dt_min, sim->last_dt_max, missing, nullptr);
    sim->defer_sync = true;
    return rc;
}

int cdb_synchronize(cdb_sim *sim) {
    if (!sim) return fail(CDB_ERR_INVALID_VALUE, "sim is NULL");
    CK(cudaSetDevice(sim->device));
    CKS(catch_up(sim));
    CK(sync_stream(sim));
    return CDB_OK;
}

int64_t cdb_num_agents(const cdb_sim *sim) { return sim ? sim->n : -1; }

}  // extern "C" (helpers follow)

namespace {
// selected fields of a record as 32-bit words in record order (transfer_kernels.cuh)
WordMap word_map(const cdb_sim *sim, uint32_t mask, int *bytes) {
    int nf = 0;
    const FieldMap *fm = host_field_map(sim->model, &nf);
    std::vector<std::pair<int, int>> sel;   // (offset, plane)
    for (int f = 0; f < nf; ++f) if (fm[f].bit & mask) sel.emplace_back(fm[f].offset, fm[f].plane);
    std::sort(sel.begin(), sel.end());
    WordMap m{};
    for (const auto &e : sel) {
        for (int h = 0; h < 2; ++h) {
            m.word[m.n_words] = (short)(e.first / 4 + h); m.plane[m.n_words] = (short)e.second; m.hi[m.n_words] = (unsigned char)h;
            ++m.n_words;
        }
    }
    *bytes = 4 * m.n_words;
    return m;
}

// device-visible alias of a pinned / registered host range (nullptr: pageable memory)
void *mapped_host(const void *p) {
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (a.type != cudaMemoryTypeHost || !a.devicePointer) return nullptr;
    return a.devicePointer;
}
}  // namespace

extern "C" {

int cdb_upload_agents_aos(cdb_sim *sim, const void *agents, int64_t n, int64_t itemsize) {
    if (!sim) return fail(CDB_ERR_INVALID_VALUE, "sim is NULL");
    sim->state_version++;
    if (itemsize != sim->itemsize) return fail(CDB_ERR_INVALID_TYPE, "agent itemsize %lld does not match the model (%lld)", (long long)itemsize, (long long)sim->itemsize);
    if (n < 0 || (n > 0 && !agents)) return fail(CDB_ERR_INVALID_VALUE, "bad agents buffer");
    if (n > 2000000000LL) return fail(CDB_ERR_CAPACITY, "too many agents");
    CK(cudaSetDevice(sim->device));
    CKS(ensure_capacity(sim, n));
    if (n * itemsize > sim->aos_capacity) {
        CKS(dev_alloc(&sim->d_aos, (size_t)(n * itemsize + 16)));
        sim->aos_capacity = n * itemsize;
    }
    sim->n = n;
    sim->n_dead = 0;
    sim->tables_valid = false;
    sim->auto_lattice_valid = false;
    if (n == 0) return CDB_OK;
    CK(cudaMemcpyAsync(sim->d_aos, agents, n * itemsize, cudaMemcpyHostToDevice, sim->stream));
    sim->h2d_bytes += n * itemsize;
    const int smem = AOS_REC_PER_BLOCK * (int)itemsize;
    if (sim->model == CDB_MODEL_CIRCULAR)
        LAUNCH(sim, k_unpack_aos<0>, cdiv(n, AOS_REC_PER_BLOCK), AOS_REC_PER_BLOCK, smem, sim->d_aos, (int)n, sim->cur);
    else
        LAUNCH(sim, k_unpack_aos<1>, cdiv(n, AOS_REC_PER_BLOCK), AOS_REC_PER_BLOCK, smem, sim->d_aos, (int)n, sim->cur);
    CK(cudaGetLastError());
    if (sim->dev_counts) LAUNCH(sim, k_counts_set, 1, 32, 0, sim->d_counts, (int)n);
    // bound on the radii / body extents: decides whether the finer search lattice is valid (search_refinement)
    CK(cudaMemsetAsync(sim->d_extmax, 0, sizeof(unsigned long long), sim->stream));
    LAUNCH(sim, k_ext_max, (cdiv(n, 1024) < 1184 ? cdiv(n, 1024) : 1184), 256, 0, sim->cur, (int)n, sim->model, sim->d_extmax);
    CK(cudaMemcpyAsync(sim->h_extmax, sim->d_extmax, sizeof(unsigned long long), cudaMemcpyDeviceToHost, sim->stream));
    // the host buffer may be pageable and reused by the caller right away
    CK(sync_stream(sim));
    {
        const unsigned long long k = *sim->h_extmax;
        const unsigned long long b = (k & 0x8000000000000000ULL) ? (k & 0x7fffffffffffffffULL) : ~k;
        double e; memcpy(&e, &b, sizeof(e));
        sim->ext_max = k == 0 ? 0.0 : e;
    }
    return CDB_OK;
}

int cdb_download_agents_aos(cdb_sim *sim, void *agents, int64_t n, int64_t itemsize, uint32_t field_mask) {
    if (!sim) return fail(CDB_ERR_INVALID_VALUE, "sim is NULL");
    if (itemsize != sim->itemsize) return fail(CDB_ERR_INVALID_TYPE, "agent itemsize %lld does not match the model (%lld)", (long long)itemsize, (long long)sim->itemsize);
    if (n != sim->n) return fail(CDB_ERR_INVALID_VALUE, "host array has %lld agents, device holds %lld", (long long)n, (long long)sim->n);
    if (n == 0) return CDB_OK;
    if (!agents) return fail(CDB_ERR_INVALID_VALUE, "agents is NULL");
    CK(cudaSetDevice(sim->device));
    CKS(catch_up(sim));
    const uint32_t mask = field_mask & CDB_F_ALL_MUTABLE;
    const uint32_t pack_mask = (field_mask & CDB_F_WHOLE_RECORD) ? (uint32_t)CDB_F_ALL_MUTABLE : mask;
    if (!(field_mask & CDB_F_WHOLE_RECORD)) {
        // selected fields straight into the caller's records when they are pinned / registered (zero-copy over PCIe)
        void *hm = mapped_host(agents);
        int bytes = 0;
        const WordMap m = word_map(sim, mask, &bytes);
        if (hm && m.n_words > 0) {
            const long long threads = (long long)n * m.n_words;
            LAUNCH(sim, k_fields_to_host, cdiv(threads, 256), 256, 0, sim->cur, (int)n, (uint32_t *)hm, (int)(itemsize / 4), m);
            CK(cudaGetLastError());
            CK(sync_stream(sim));
            sim->d2h_bytes += (int64_t)bytes * n;
            return check_device_error(sim);
        }
        if (m.n_words == 0) return check_device_error(sim);
    }
    sim->d2h_bytes += n * itemsize;
    if (sim->model == CDB_MODEL_CIRCULAR)
        LAUNCH(sim, k_pack_aos<0>, cdiv(n, 128), 128, 0, sim->cur, (int)n, sim->d_aos, pack_mask);
    else
        LAUNCH(sim, k_pack_aos<1>, cdiv(n, 128), 128, 0, sim->cur, (int)n, sim->d_aos, pack_mask);
    CK(cudaGetLastError());
    if (field_mask & CDB_F_WHOLE_RECORD) {
        CK(cudaMemcpyAsync(agents, sim->d_aos, n * itemsize, cudaMemcpyDeviceToHost, sim->stream));
        CK(sync_stream(sim));
        return check_device_error(sim);
    }
    if (n * itemsize > sim->bounce_bytes) {
        if (sim->h_bounce) cudaFreeHost(sim->h_bounce);
        sim->h_bounce = nullptr;
        CK(cudaMallocHost((void **)&sim->h_bounce, n * itemsize));
        sim->bounce_bytes = n * itemsize;
    }
    CK(cudaMemcpyAsync(sim->h_bounce, sim->d_aos, n * itemsize, cudaMemcpyDeviceToHost, sim->stream));
    CK(sync_stream(sim));
    // merge the selected fields into the caller's records (plumbing: byte copies only); adjacent selected fields are
    // copied as one run per record (position .. force_prev are contiguous in the record)
    int nf = 0;
    const FieldMap *fm = host_field_map(sim->model, &nf);
    std::vector<std::pair<int, int>> runs;   // (offset, bytes), sorted by offset
    {
        std::vector<int> offs;
        for (int f = 0; f < nf; ++f) if (fm[f].bit & mask) offs.push_back(fm[f].offset);
        std::sort(offs.begin(), offs.end());
        for (int o : offs) {
            if (!runs.empty() && runs.back().first + runs.back().second == o) runs.back().second += 8;
            else runs.emplace_back(o, 8);
        }
    }
    uint8_t *dst = (uint8_t *)agents;
    const uint8_t *src = sim->h_bounce;
    for (int64_t i = 0; i < n; ++i, dst += itemsize, src += itemsize)
        for (const auto &r : runs) memcpy(dst + r.first, src + r.first, (size_t)r.second);
    return check_device_error(sim);
}

// ---- asynchronous snapshots: nothing here makes the host wait for the step that is being copied --------------------------
int cdb_set_deferred_sync(cdb_sim *sim, int enable) {
    if (!sim) return fail(CDB_ERR_INVALID_VALUE, "sim is NULL");
    sim->defer_sync = enable != 0;
    return CDB_OK;
}

int64_t cdb_sync_count(const cdb_sim *sim) { return sim ? sim->syncs : -1; }

int cdb_snapshot_begin(cdb_sim *sim, int64_t *slot_out) {
    if (!sim || !slot_out) return fail(CDB_ERR_INVALID_VALUE, "sim / slot_out is NULL");
    CK(cudaSetDevice(sim->device));
    if (sim->strip) return fail(CDB_ERR_STATE, "snapshots are not available in strip mode");
    const int slot = sim->snap_next;
    const int64_t n = sim->n, bytes = n * sim->itemsize;
    if (bytes > sim->snap_bytes[slot]) {
        // (re)allocation: the slot may still be in flight from an earlier snapshot
        if (sim->ev_snap[slot]) { sim->syncs++; CK(cudaEventSynchronize(sim->ev_snap[slot])); }
        cudaFree(sim->d_snap[slot]); cudaFreeHost(sim->h_snap[slot]);
        sim->d_snap[slot] = nullptr; sim->h_snap[slot] = nullptr;
        CKS(dev_alloc(&sim->d_snap[slot], (size_t)bytes + 16));
        CK(cudaMallocHost((void **)&sim->h_snap[slot], (size_t)bytes + 16));
        sim->snap_bytes[slot] = bytes;
    }
    sim->snap_n[slot] = n;
    if (n > 0) {
        // whole records: the uploaded image (constants, States bytes) with every mutable field, target, active and the
        // follower fields as the device has them NOW -- what cdb_download_agents_aos + cdb_get_states + cdb_get_active return
        CK(cudaMemcpyAsync(sim->d_snap[slot], sim->d_aos, bytes, cudaMemcpyDeviceToDevice, sim->stream));
        const uint8_t *active = sim->d_active && sim->active_n == n ? sim->d_active : nullptr;
        const bool states = sim->states_n == n;
        if (sim->model == CDB_MODEL_CIRCULAR)
            LAUNCH(sim, k_snapshot_records<0>, cdiv(n, 128), 128, 0, sim->cur, (int)n, sim->d_snap[slot], active, states ? sim->d_is_follower : nullptr,
                   states ? sim->d_index_leader : nullptr);
        else
            LAUNCH(sim, k_snapshot_records<1>, cdiv(n, 128), 128, 0, sim->cur, (int)n, sim->d_snap[slot], active, states ? sim->d_is_follower : nullptr,
                   states ? sim->d_index_leader : nullptr);
        CK(cudaGetLastError());
        CK(cudaEventRecord(sim->ev_main, sim->stream));
        CK(cudaStreamWaitEvent(sim->side, sim->ev_main, 0));
        CK(cudaMemcpyAsync(sim->h_snap[slot], sim->d_snap[slot], bytes, cudaMemcpyDeviceToHost, sim->side));
        sim->d2h_bytes += bytes;
    }
    CK(cudaEventRecord(sim->ev_snap[slot], sim->side));
    *slot_out = slot;
    sim->snap_next ^= 1;
    return CDB_OK;
}

int cdb_snapshot_wait(cdb_sim *sim, int64_t slot, const void **records, int64_t *n) {
    if (!sim || slot < 0 || slot > 1 || !records) return fail(CDB_ERR_INVALID_VALUE, "bad snapshot slot");
    CK(cudaSetDevice(sim->device));
    if (cudaEventQuery(sim->ev_snap[slot]) != cudaSuccess) { cudaGetLastError(); sim->syncs++; CK(cudaEventSynchronize(sim->ev_snap[slot])); }
    *records = sim->h_snap[slot];
    if (n) *n = sim->snap_n[slot];
    return CDB_OK;
}

// the scalars host-side bookkeeping wants every update: dt, time_tot, InsideDomain's change count, TargetReached's counts
int cdb_scalars_begin(cdb_sim *sim, int64_t *slot_out) {
    if (!sim || !slot_out) return fail(CDB_ERR_INVALID_VALUE, "sim / slot_out is NULL");
    CK(cudaSetDevice(sim->device));
    const int slot = sim->scal_next;
    const int64_t np = std::min<int64_t>(sim->n_polygons[CDB_POLY_TARGETS], 64);
    CK(cudaMemcpyAsync(sim->d_scal, sim->d_dt, 2 * sizeof(double), cudaMemcpyDeviceToDevice, sim->stream));
    if (sim->d_poly_counts) CK(cudaMemcpyAsync(sim->d_scal + 2, sim->d_poly_counts, (size_t)(1 + np) * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, sim->stream));
    CK(cudaEventRecord(sim->ev_main, sim->stream));
    CK(cudaStreamWaitEvent(sim->side, sim->ev_main, 0));
    CK(cudaMemcpyAsync(sim->h_scal[slot], sim->d_scal, (size_t)(3 + np) * sizeof(unsigned long long), cudaMemcpyDeviceToHost, sim->side));
    CK(cudaEventRecord(sim->ev_scal[slot], sim->side));
    sim->scal_n[slot] = 3 + np;
    *slot_out = slot;
    sim->scal_next ^= 1;
    return CDB_OK;
}

int cdb_scalars_wait(cdb_sim *sim, int64_t slot, double *dt, double *time_tot, int64_t *inside_changes, int64_t *target_counts, int64_t n_targets) {
    if (!sim || slot < 0 || slot > 1) return fail(CDB_ERR_INVALID_VALUE, "bad scalars slot");
    CK(cudaSetDevice(sim->device));
    if (cudaEventQuery(sim->ev_scal[slot]) != cudaSuccess) { cudaGetLastError(); sim->syncs++; CK(cudaEventSynchronize(sim->ev_scal[slot])); }
    const unsigned long long *h = sim->h_scal[slot];
    if (dt) memcpy(dt, &h[0], sizeof(double));
    if (time_tot) memcpy(time_tot, &h[1], sizeof(double));
    if (inside_changes) *inside_changes = (int64_t)h[2];
    for (int64_t p = 0; target_counts && p < n_targets && 3 + p < sim->scal_n[slot]; ++p) target_counts[p] = (int64_t)h[3 + p];
    return CDB_OK;
}

int cdb_host_register(cdb_sim *sim, void *agents, int64_t n, int64_t itemsize) {
    if (!sim) return fail(CDB_ERR_INVALID_VALUE, "sim is NULL");
    if (!agents || n <= 0) return CDB_OK;
    CK(cudaSetDevice(sim->device));
    if (mapped_host(agents)) return CDB_OK;          // already pinned (cudaHostAlloc / registered before)
    CK(cudaHostRegister(agents, (size_t)(n * itemsize), cudaHostRegisterMapped));
    sim->registered.emplace_back(agents, (size_t)(n * itemsize));
    return CDB_OK;
}

int cdb_host_unregister(cdb_sim *sim, void *agents) {
    if (!sim) return fail(CDB_ERR_INVALID_VALUE, "sim is NULL");
    CK(cudaSetDevice(sim->device));
    CK(sync_stream(sim));
    for (size_t k = 0; k < sim->registered.size(); ++k)
        if (sim->registered[k].first == agents) {
            cudaHostUnregister(agents);
            sim->registered.erase(sim->registered.begin() + k);
            return CDB_OK;
        }
    return CDB_OK;
}

int cdb_upload_agents_fields(cdb_sim *sim, const void *agents, int64_t n, int64_t itemsize, uint32_t field_mask) {
    if (!sim) return fail(CDB_ERR_INVALID_VALUE, "sim is NULL");
    if (itemsize != sim->itemsize) return fail(CDB_ERR_INVALID_TYPE, "agent itemsize %lld does not match the model (%lld)", (long long)itemsize, (long long)sim->itemsize);
    if (n != sim->n || sim->n_dead != 0 || sim->strip)
        return fail(CDB_ERR_STATE, "cdb_upload_agents_fields updates the %lld agents uploaded before (got %lld)", (long long)sim->n, (long long)n);
    const uint32_t mask = field_mask & CDB_F_ALL_MUTABLE;
    if (n == 0 || mask == 0) return CDB_OK;
    if (!agents) return fail(CDB_ERR_INVALID_VALUE, "agents is NULL");
    CK(cudaSetDevice(sim->device));
    sim->state_version++;
    int bytes = 0;
    const WordMap m = word_map(sim, mask, &bytes);
    void *hm = mapped_host(agents);
    // the planes may have been re-sorted by fused steps since the upload: record index -> slot
    if (n > sim->rec_slot_cap) { CKS(dev_alloc(&sim->d_rec_slot, (size_t)n)); sim->rec_slot_cap = n; }
    LAUNCH(sim, k_slot_of_id, cdiv(n, 256), 256, 0, sim->cur, (int)n, sim->d_rec_slot);
    if (hm && bytes <= TRANSFER_ZERO_COPY_MAX) {
        const long long threads = (long long)n * (m.n_words / 2);
        LAUNCH(sim, k_fields_from_host, cdiv(threads, 256), 256, 0, (const uint32_t *)hm, (int)(itemsize / 4), sim->cur, (int)n, sim->d_rec_slot, m);
        sim->h2d_bytes += (int64_t)bytes * n;
    } else {
        // whole-record DMA into the device image (55 GB/s from pinned memory), then the same field kernel from device memory
        CK(cudaMemcpyAsync(sim->d_aos, agents, n * itemsize, cudaMemcpyHostToDevice, sim->stream));
        const long long threads = (long long)n * (m.n_words / 2);
        LAUNCH(sim, k_fields_from_host, cdiv(threads, 256), 256, 0, (const uint32_t *)sim->d_aos, (int)(itemsize / 4), sim->cur, (int)n, sim->d_rec_slot, m);
        sim->h2d_bytes += n * itemsize;
    }
    CK(cudaGetLastError());
    sim->tables_valid = false;
    if (mask & (CDB_F_POSITION | CDB_F_SHOULDERS)) {
        sim->auto_lattice_valid = false;
        if (sim->model == CDB_MODEL_THREE_CIRCLE && (mask & (CDB_F_POSITION | CDB_F_SHOULDERS))) {
            // new positions / shoulders can change the body extents the search refinement relies on
            CK(cudaMemsetAsync(sim->d_extmax, 0, sizeof(unsigned long long), sim->stream));
            LAUNCH(sim, k_ext_max, (cdiv(n, 1024) < 1184 ? cdiv(n, 1024) : 1184), 256, 0, sim->cur, (int)n, sim->model, sim->d_extmax);
            CK(cudaMemcpyAsync(sim->h_extmax, sim->d_extmax, sizeof(unsigned long long), cudaMemcpyDeviceToHost, sim->stream));
            CK(sync_stream(sim));
            const unsigned long long k = *sim->h_extmax;
            const unsigned long long b = (k & 0x8000000000000000ULL) ? (k & 0x7fffffffffffffffULL) : ~k;
            double e; memcpy(&e, &b, sizeof(e));
            sim->ext_max = k == 0 ? 0.0 : e;
        }
    }
    CK(sync_stream(sim));      // the host buffer may be reused by the caller right away
    return CDB_OK;
}

int cdb_transfer_stats(cdb_sim *sim, int64_t *h2d_bytes, int64_t *d2h_bytes, int reset) {
    if (!sim) return fail(CDB_ERR_INVALID_VALUE, "sim is NULL");
    if (h2d_bytes) *h2d_bytes = sim->h2d_bytes;
    if (d2h_bytes) *d2h_bytes = sim->d2h_bytes;
    if (reset) sim->h2d_bytes = sim->d2h_bytes = 0;
    return CDB_OK;
}

int cdb_set_obstacles(cdb_sim *sim, const double *segments, int64_t n_segments) {
    if (!sim) return fail(CDB_ERR_INVALID_VALUE, "sim is NULL");
    sim->state_version++;
    if (n_segments < 0 || (n_segments > 0 && !segments)) return fail(CDB_ERR_INVALID_VALUE, "bad obstacle buffer");
    CK(cudaSetDevice(sim->device));
    CK(sync_stream(sim));
    CKS(dev_alloc(&sim->d_obstacles, (size_t)n_segments * SEG));
    sim->n_obstacles = n_segments;
    if (n_segments) {
        double *raw = nullptr;
        CKS(dev_alloc(&raw, (size_t)n_segments * 4));
        CK(cudaMemcpy(raw, segments, n_segments * 4 * sizeof(double), cudaMemcpyHostToDevice));
        LAUNCH(sim, k_obstacle_prep, cdiv(n_segments, 128), 128, 0, raw, (int)n_segments, sim->d_obstacles);
        CK(sync_stream(sim));
        cudaFree(raw);
    }
    return CDB_OK;
}

int cdb_set_navigation_field(cdb_sim *sim, int64_t target, const double *U, const double *V, int64_t ny, int64_t nx,
                             double minx, double miny, double step) {
    if (!sim) return fail(CDB_ERR_INVALID_VALUE, "sim is NULL");
    sim->state_version++;
    if (target < 0 || target >= MAX_NAV_TARGETS) return fail(CDB_ERR_INVALID_VALUE, "target index %lld out of range [0, %d)", (long long)target, MAX_NAV_TARGETS);
    if (ny < 0 || nx < 0 || ((ny * nx) > 0 && (!U || !V))) return fail(CDB_ERR_INVALID_VALUE, "bad navigation field");
    CK(cudaSetDevice(sim->device));
    CK(sync_stream(sim));
    if ((int64_t)sim->nav.size() <= target) sim->nav.resize(target + 1, NavField{});
    NavField &f = sim->nav[target];
    cudaFree((void *)f.U); cudaFree((void *)f.V);
    f = NavField{};
    double *dU = nullptr, *dV = nullptr;
    CKS(dev_alloc(&dU, (size_t)(ny * nx)));
    CKS(dev_alloc(&dV, (size_t)(ny * nx)));
    if (ny * nx > 0) {
        CK(cudaMemcpy(dU, U, ny * nx * sizeof(double), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dV, V, ny * nx * sizeof(double), cudaMemcpyHostToDevice));
    }
    f.U = dU; f.V = dV; f.ny = ny; f.nx = nx; f.minx = minx; f.miny = miny; f.step = step; f.valid = 1;
    sim->n_nav = (int)sim->nav.size();
    CK(cudaMemcpy(sim->d_nav, sim->nav.data(), sim->nav.size() * sizeof(NavField), cudaMemcpyHostToDevice));
    return CDB_OK;
}

// ---- navigation-field construction (field_kernels.cuh) -------------------------------------------------------------------
namespace {
struct FieldScratch {
    uint8_t *target = nullptr, *obst = nullptr, *mask = nullptr, *state = nullptr, *dirmask = nullptr, *act[2] = {nullptr, nullptr};
    double *Tt = nullptr, *To = nullptr, *U = nullptr, *V = nullptr, *seg_t = nullptr, *seg_o = nullptr;
    unsigned *changed = nullptr;
    ~FieldScratch() {
        cudaFree(target); cudaFree(obst); cudaFree(mask); cudaFree(state); cudaFree(dirmask); cudaFree(act[0]); cudaFree(act[1]);
        cudaFree(Tt); cudaFree(To); cudaFree(U); cudaFree(V); cudaFree(seg_t); cudaFree(seg_o); cudaFree(changed);
    }
};

// iterate the upwind update to its fixed point; *rounds_out = rounds launched
int eikonal_solve(cdb_sim *sim, FieldScratch &f, double *T, const uint8_t *raster, const uint8_t *mask, int ny, int nx, double h, int *rounds_out) {
    cudaStream_t st = sim->stream;
    const int tiles_x = cdiv(nx, ET), tiles_y = cdiv(ny, ET);
    const long long cells = (long long)ny * nx;
    CK(cudaMemsetAsync(f.act[0], 0, (size_t)tiles_x * tiles_y, st));
    LAUNCH(sim, k_eik_init, cdiv(cells, 256), 256, 0, raster, mask, ny, nx, h, T, f.state, f.act[0], tiles_x, tiles_y);
    const int max_rounds = 64 * (tiles_x + tiles_y) + 1024;
    unsigned h_changed = 1;
    int r = 0, cur = 0;
    CK(cudaMemsetAsync(f.changed, 0, sizeof(unsigned), st));
    for (; r < max_rounds; ++r) {
        CK(cudaMemsetAsync(f.act[cur ^ 1], 0, (size_t)tiles_x * tiles_y, st));
        LAUNCH(sim, k_eik_fim, dim3(tiles_x, tiles_y), dim3(ET, ET), 0, T, f.state, ny, nx, h, f.act[cur], f.act[cur ^ 1], tiles_x, tiles_y, f.changed);
        cur ^= 1;
        if ((r & 15) == 15) {
            CK(cudaMemcpyAsync(&h_changed, f.changed, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
            CK(sync_stream(sim));
            if (h_changed == 0) { ++r; break; }
            CK(cudaMemsetAsync(f.changed, 0, sizeof(unsigned), st));
        }
    }
    CK(cudaGetLastError());
    if (h_changed != 0) return fail(CDB_ERR_STATE, "eikonal solver did not converge in %d rounds", max_rounds);
    LAUNCH(sim, k_eik_sign, cdiv(cells, 256), 256, 0, T, raster, f.state, cells);
    if (rounds_out) *rounds_out = r;
    return CDB_OK;
}
}  // namespace

int cdb_build_navigation_field(cdb_sim *sim, int64_t target, const double *target_segments, int64_t n_target_segments,
                               const double *obstacle_segments, int64_t n_obstacle_segments, int64_t ny, int64_t nx, double minx,
                               double miny, double step, double radius, double strength, double *distance_map_out, double *U_out,
                               double *V_out, int64_t *rounds_out) {
    if (!sim) return fail(CDB_ERR_INVALID_VALUE, "sim is NULL");
    if (target < 0 || target >= MAX_NAV_TARGETS) return fail(CDB_ERR_INVALID_VALUE, "target index %lld out of range [0, %d)", (long long)target, MAX_NAV_TARGETS);
    if (ny <= 0 || nx <= 0 || (double)ny * (double)nx > 2.0e9 || !(step > 0.0)) return fail(CDB_ERR_INVALID_VALUE, "bad grid");
    if (n_target_segments <= 0 || !target_segments) return fail(CDB_ERR_INVALID_VALUE, "a navigation field needs target geometry");
    if (n_obstacle_segments < 0 || (n_obstacle_segments > 0 && !obstacle_segments)) return fail(CDB_ERR_INVALID_VALUE, "bad obstacle buffer");
    CK(cudaSetDevice(sim->device));
    sim->state_version++;
    cudaStream_t st = sim->stream;
    CK(sync_stream(sim));
    const long long cells = (long long)ny * nx;
    const int tiles = cdiv(nx, ET) * cdiv(ny, ET);
    const bool walls = n_obstacle_segments > 0;
    FieldScratch f;
    CKS(dev_alloc(&f.target, (size_t)cells)); CKS(dev_alloc(&f.state, (size_t)cells)); CKS(dev_alloc(&f.dirmask, (size_t)cells));
    CKS(dev_alloc(&f.act[0], (size_t)tiles)); CKS(dev_alloc(&f.act[1], (size_t)tiles)); CKS(dev_alloc(&f.changed, 1));
    CKS(dev_alloc(&f.Tt, (size_t)cells)); CKS(dev_alloc(&f.U, (size_t)cells)); CKS(dev_alloc(&f.V, (size_t)cells));
    CKS(dev_alloc(&f.seg_t, (size_t)n_target_segments * 4));
    CK(cudaMemcpyAsync(f.seg_t, target_segments, n_target_segments * 4 * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(f.target, 0, (size_t)cells, st));
    LAUNCH(sim, k_raster_lines, cdiv(n_target_segments, 64), 64, 0, f.seg_t, (int)n_target_segments, minx, miny, step, (int)ny, (int)nx, f.target);
    if (walls) {
        CKS(dev_alloc(&f.obst, (size_t)cells)); CKS(dev_alloc(&f.mask, (size_t)cells)); CKS(dev_alloc(&f.To, (size_t)cells));
        CKS(dev_alloc(&f.seg_o, (size_t)n_obstacle_segments * 4));
        CK(cudaMemcpyAsync(f.seg_o, obstacle_segments, n_obstacle_segments * 4 * sizeof(double), cudaMemcpyHostToDevice, st));
        CK(cudaMemsetAsync(f.obst, 0, (size_t)cells, st));
        LAUNCH(sim, k_raster_lines, cdiv(n_obstacle_segments, 64), 64, 0, f.seg_o, (int)n_obstacle_segments, minx, miny, step, (int)ny, (int)nx, f.obst);
        LAUNCH(sim, k_buffer_mask, cdiv(cells, 256), 256, 0, f.seg_o, (int)n_obstacle_segments, radius, minx, miny, step, (int)ny, (int)nx, f.mask);
    }
    // shortest_path: distance to the targets around the buffered obstacles, its normalised gradient, buffer zone filled
    int rounds_t = 0, rounds_o = 0;
    CKS(eikonal_solve(sim, f, f.Tt, f.target, f.mask, (int)ny, (int)nx, step, &rounds_t));
    LAUNCH(sim, k_direction_map, cdiv(cells, 256), 256, 0, f.Tt, f.mask, (int)ny, (int)nx, f.U, f.V, f.dirmask);
    double *dU = nullptr, *dV = nullptr;          // the installed maps (owned by the sim afterwards)
    CKS(dev_alloc(&dU, (size_t)cells));
    if (dev_alloc(&dV, (size_t)cells) != CDB_OK) { cudaFree(dU); return CDB_ERR_CUDA; }
    if (walls) {
        const int reach = (int)std::ceil(radius / step) + 3;
        LAUNCH(sim, k_fill_missing, cdiv(cells, 256), 256, 0, f.dirmask, f.obst, (int)ny, (int)nx, reach, f.U, f.V, dU, dV);
        // obstacle_handling: distance from the walls (no mask), blend within `radius`, normalise
        CKS(eikonal_solve(sim, f, f.To, f.obst, nullptr, (int)ny, (int)nx, step, &rounds_o));
        LAUNCH(sim, k_obstacle_handling, cdiv(cells, 256), 256, 0, f.To, (int)ny, (int)nx, radius, strength, dU, dV);
    } else {
        CK(cudaMemcpyAsync(dU, f.U, cells * sizeof(double), cudaMemcpyDeviceToDevice, st));
        CK(cudaMemcpyAsync(dV, f.V, cells * sizeof(double), cudaMemcpyDeviceToDevice, st));
    }
    CK(cudaGetLastError());
    if (distance_map_out) CK(cudaMemcpyAsync(distance_map_out, f.Tt, cells * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (U_out) CK(cudaMemcpyAsync(U_out, dU, cells * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (V_out) CK(cudaMemcpyAsync(V_out, dV, cells * sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(sync_stream(sim));
    if (rounds_out) *rounds_out = rounds_t + rounds_o;
    // install as the navigation field of `target` (what cdb_set_navigation_field does with host maps)
    if ((int64_t)sim->nav.size() <= target) sim->nav.resize(target + 1, NavField{});
    NavField &nf = sim->nav[target];
    cudaFree((void *)nf.U); cudaFree((void *)nf.V);
    nf = NavField{};
    nf.U = dU; nf.V = dV; nf.ny = ny; nf.nx = nx; nf.minx = minx; nf.miny = miny; nf.step = step; nf.valid = 1;
    sim->n_nav = (int)sim->nav.size();
    CK(cudaMemcpy(sim->d_nav, sim->nav.data(), sim->nav.size() * sizeof(NavField), cudaMemcpyHostToDevice));
    return CDB_OK;
}

int cdb_clear_navigation(cdb_sim *sim) {
    if (!sim) return fail(CDB_ERR_INVALID_VALUE, "sim is NULL");
    sim->state_version++;
    CK(cudaSetDevice(sim->device));
    CK(sync_stream(sim));
    for (auto &f : sim->nav) { cudaFree((void *)f.U); cudaFree((void *)f.V); }
    sim->nav.clear();
    sim->n_nav = 0;
    return CDB_OK;
}

#define SIM_ENTRY()                                               \
    if (!sim) return fail(CDB_ERR_INVALID_VALUE, "sim is NULL"); \
    CK(cudaSetDevice(sim->device))
// node-wise entry points that launch work on the buffers a captured pair of steps refers to: the graph is re-captured
#define SIM_ENTRY_NODE() \
    SIM_ENTRY();         \
    sim->state_version++

int cdb_reset(cdb_sim *sim) { SIM_ENTRY_NODE(); return node_reset(sim); }
int cdb_set_seed(cdb_sim *sim, uint64_t seed) { SIM_ENTRY(); sim->seed = seed; sim->fluct_calls = 0; sim->state_version++; return CDB_OK; }
int cdb_fluctuation(cdb_sim *sim) {
    SIM_ENTRY_NODE();
    // key the stream on a private call counter in the high half, so that node-wise calls never reuse a fused step's stream
    const unsigned long long step = (1ULL << 63) | sim->fluct_calls++;
    if (sim->n) LAUNCH(sim, k_fluctuation, cdiv(sim->n, 256), 256, 0, sim->cur, (int)sim->n, sim->model, sim->seed, step);
    CK(cudaGetLastError());
    return CDB_OK;
}
int cdb_navigation(cdb_sim *sim) { SIM_ENTRY_NODE(); return node_navigation(sim); }
int cdb_orientation(cdb_sim *sim) { SIM_ENTRY_NODE(); return node_orientation(sim); }
int cdb_adjust(cdb_sim *sim) { SIM_ENTRY_NODE(); return node_adjust(sim); }
int cdb_agent_agent(cdb_sim *sim, double cell_size) { SIM_ENTRY_NODE(); return node_agent_agent(sim, cell_size); }
int cdb_agent_obstacle(cdb_sim *sim) { SIM_ENTRY_NODE(); return node_agent_obstacle(sim); }

int cdb_integrate(cdb_sim *sim, double dt_min, double dt_max, double *dt_out) {
    SIM_ENTRY_NODE();
    CKS(node_integrate(sim, dt_min, dt_max));
    if (dt_out) {
        CK(cudaMemcpyAsync(sim->h_dt, sim->d_dt, 2 * sizeof(double), cudaMemcpyDeviceToHost, sim->stream));
        CK(sync_stream(sim));
        *dt_out = sim->h_dt[0];
    }
    return CDB_OK;
}

// Resident-order steps apply to whole fused steps of the once-per-pair pipeline on one device, for crowds large enough that
// the step is not launch-bound, and only where the pair set does not depend on the lattice (3 + 2 max R < cell_size).
static bool chain_usable(const cdb_sim *sim, uint32_t flags, double cell_size) {
    return sim->chain_enabled && sim->rebuild_max > 1 && sim->skin_frac > 0.0 && sim->variant == 3 && !sim->strip && !sim->lattice_fixed &&
           (flags & CDB_STEP_AGENT_AGENT) && (flags & CDB_STEP_INTEGRATOR) && sim->n >= sim->chain_min_agents && sim->n_dead == 0 &&
           !sim->dev_counts && sim->pb.ctr && (SIGTH_SOC + 2.0 * sim->ext_max) * (1.0 + 1e-9) < cell_size;
}

// the launches of a step on the kept order (resident-order steps), in stream order
static int issue_kept_launches(cdb_sim *sim, uint32_t flags, double dt_min, double dt_max, double *log) {
    LAUNCH(sim, k_chain_begin, 1, 32, 0, sim->d_chain, sim->d_vmax, 0);
    CKS(prof_mark(sim));                 // (the phase marks of issue_step; no-ops unless profiling, which rules the graph out)
    sim->chain_step = true;
    sim->chain_inplace = true;
    const int rc = launch_step_kernel(sim, flags, dt_min, dt_max, log);
    sim->chain_step = false;
    sim->chain_inplace = false;
    CKS(rc);
    LAUNCH(sim, k_chain_end, 1, 32, 0, sim->d_chain, sim->pb.ctr, (long long)sim->pb.cap);
    CKS(prof_mark(sim));
    LAUNCH(sim, k_step_advance, 1, 32, 0, sim->d_stepctr, sim->pb.ctr, (long long)sim->pb.cap);
    CKS(prof_mark(sim));
    CK(cudaGetLastError());
    return CDB_OK;
}

static bool kept_graph_usable(const cdb_sim *sim) {
    return sim->use_graphs && sim->kept_graph_ok && !sim->profiling && sim->n <= sim->kept_graph_max_agents && sim->stream != nullptr &&
           sim->stream != cudaStreamLegacy && sim->stream != cudaStreamPerThread;
}

// a kept step as one graph launch; falls back to plain launches where the stream cannot be captured
static int issue_kept_step(cdb_sim *sim, uint32_t flags, double cell_size, double dt_min, double dt_max, double *log) {
    if (!kept_graph_usable(sim)) return issue_kept_launches(sim, flags, dt_min, dt_max, log);
    cdb_sim::KeptGraph *hit = nullptr;
    for (auto &g : sim->kept_graph)
        if (g.exec && g.cur == sim->cur.p && g.cells == sim->d_cell_of_slot && g.pairs == sim->pb.pairs && g.cap == sim->pb.cap && g.flags == flags &&
            g.cell_size == cell_size && g.dt_min == dt_min && g.dt_max == dt_max && g.log == log && g.n == sim->n_sorted && g.version == sim->state_version)
            hit = &g;
    if (!hit) {
        cdb_sim::KeptGraph &g = sim->kept_graph[sim->kept_graph_next];
        sim->kept_graph_next = (sim->kept_graph_next + 1) % 4;
        if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
        const int64_t launches0 = sim->launches;
        cudaGraph_t graph = nullptr;
        if (cudaStreamBeginCapture(sim->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
            cudaGetLastError();
            sim->kept_graph_ok = false;
            return issue_kept_launches(sim, flags, dt_min, dt_max, log);
        }
        const int rc = issue_kept_launches(sim, flags, dt_min, dt_max, log);      // recorded, not run
        cudaError_t e = cudaStreamEndCapture(sim->stream, &graph);
        if (rc == CDB_OK && e == cudaSuccess) e = cudaGraphInstantiate(&g.exec, graph, 0);
        if (graph) cudaGraphDestroy(graph);
        g.launches = sim->launches - launches0;
        sim->launches = launches0;
        if (rc != CDB_OK || e != cudaSuccess) {
            cudaGetLastError();
            g.exec = nullptr;
            if (rc != CDB_OK) return rc;
            sim->kept_graph_ok = false;
            return issue_kept_launches(sim, flags, dt_min, dt_max, log);
        }
        g.cur = sim->cur.p; g.cells = sim->d_cell_of_slot; g.pairs = sim->pb.pairs; g.cap = sim->pb.cap; g.flags = flags; g.cell_size = cell_size;
        g.dt_min = dt_min; g.dt_max = dt_max; g.log = log; g.n = sim->n_sorted; g.version = sim->state_version;
        hit = &g;
    }
    CK(cudaGraphLaunch(hit->exec, sim->stream));
    sim->launches += hit->launches;
    // what launch_step_kernel / launch_pairs note on the host for a step on the kept order
    sim->pairs_pending = true;
    sim->perm_valid = false;
    sim->tables_valid = false;
    return CDB_OK;
}

// launches of ONE step of the selected nodes (no host synchronisation unless the search lattice has to be re-derived)
static int issue_step(cdb_sim *sim, uint32_t flags, double cell_size, double dt_min, double dt_max, bool log_dt) {
    double *log = log_dt && (flags & CDB_STEP_INTEGRATOR) ? sim->d_dt_log : nullptr;
    CKS(prof_mark(sim));
    if (sim->variant == 1) {
        // node-by-node path with the one-phase kernels (kept as an independent cross-check of the fused kernel)
        if ((flags & CDB_STEP_FLUCTUATION) && sim->n)
            LAUNCH(sim, k_fluctuation, cdiv(sim->n, 256), 256, 0, sim->cur, (int)sim->n, sim->model, sim->seed, (unsigned long long)sim->iterations);
        if (flags & CDB_STEP_NAVIGATION) CKS(node_navigation(sim));
        if (flags & CDB_STEP_ORIENTATION) CKS(node_orientation(sim));
        if (flags & CDB_STEP_ADJUSTING) CKS(node_adjust(sim));
        if (flags & CDB_STEP_AGENT_AGENT) CKS(build_block_list(sim, cell_size));
        CKS(prof_mark(sim));
        if (flags & CDB_STEP_AGENT_AGENT) CKS(launch_agent_agent(sim));
        CKS(prof_mark(sim)); CKS(prof_mark(sim)); CKS(prof_mark(sim));
        if (flags & CDB_STEP_AGENT_OBSTACLE) CKS(node_agent_obstacle(sim));
        if (flags & CDB_STEP_INTEGRATOR) {
            CKS(node_integrate(sim, dt_min, dt_max));
            if (log) CK(cudaMemcpyAsync(log + (sim->iterations % DT_LOG), sim->d_dt, sizeof(double), cudaMemcpyDeviceToDevice, sim->stream));
        }
        if (flags & CDB_STEP_RESET) CKS(node_reset(sim));
    } else {
        const bool need_vmax = flags & CDB_STEP_INTEGRATOR;
        const bool chain = chain_usable(sim, flags, cell_size);
        sim->chain_step = chain;
        sim->chain_inplace = false;
        if (chain) {
            // resident-order step: rebuild the block list only every `rebuild_every` steps (ChainState in kernels.cuh)
            const double scale = 1.0 + sim->skin_frac;
            const bool keep = sim->chain_valid && sim->chain_version == sim->state_version && sim->chain_cell_size == cell_size &&
                              sim->chain_scale == scale && sim->since_rebuild < sim->rebuild_every && sim->n_sorted == sim->n;
            if (keep) {
                // the whole step is a fixed launch sequence on fixed buffers
                sim->chain_step = false;
                sim->chain_kept++;
                CKS(issue_kept_step(sim, flags, cell_size, dt_min, dt_max, log));
                sim->since_rebuild++;
                sim->chain_version = sim->state_version;
                sim->iterations++;
                return CDB_OK;
            } else {
                LAUNCH(sim, k_chain_begin, 1, 32, 0, sim->d_chain, sim->d_vmax, 1);
                LAUNCH(sim, k_vmax_init, 1, 32, 0, sim->d_vmax);
                CKS(build_block_list(sim, cell_size, true, sim->d_vmax, false, search_refinement(sim, cell_size), scale));
                sim->since_rebuild = 0;
                sim->chain_rebuilds++;
                sim->chain_cell_size = cell_size;
                sim->chain_scale = scale;
                // every pair closer than the interaction range is swept while both agents together have drifted less than
                // the slack between the cells' reach (cell_size * scale) and that range
                sim->drift_limit = 0.5 * (cell_size * scale - (SIGTH_SOC + 2.0 * sim->ext_max) * (1.0 + 1e-9)) * (1.0 - 1e-9);
            }
        } else if (flags & CDB_STEP_AGENT_AGENT) {
            if (need_vmax) LAUNCH(sim, k_vmax_init, 1, 32, 0, sim->d_vmax);
            CKS(build_block_list(sim, cell_size, true, need_vmax ? sim->d_vmax : nullptr, false, search_refinement(sim, cell_size)));
        } else if (need_vmax) {
            LAUNCH(sim, k_vmax_init, 1, 32, 0, sim->d_vmax);
            if (sim->n > 0) LAUNCH(sim, k_vmax, (cdiv(sim->n, 1024) < 1184 ? cdiv(sim->n, 1024) : 1184), 256, 0, sim->cur, (int)sim->n, sim->d_vmax);
        }
        CKS(prof_mark(sim));
        const int rc = launch_step_kernel(sim, flags, dt_min, dt_max, log);
        sim->chain_step = false;
        CKS(rc);
        if (chain) {
            LAUNCH(sim, k_chain_end, 1, 32, 0, sim->d_chain, sim->pb.ctr, (long long)sim->pb.cap);
            sim->since_rebuild++;
            sim->chain_valid = true;
            sim->chain_version = sim->state_version;
        }
        CKS(prof_mark(sim));
    }
    const bool checked = use_pairs(sim) && sim->variant != 1 && (flags & CDB_STEP_AGENT_AGENT) && sim->n_sorted > 0;
    LAUNCH(sim, k_step_advance, 1, 32, 0, sim->d_stepctr, checked ? sim->pb.ctr : nullptr, (long long)sim->pb.cap);
    CKS(prof_mark(sim));
    sim->iterations++;
    return CDB_OK;
}

// Two consecutive fused steps as one CUDA graph: after two steps the ping-pong buffers and the swapped index arrays are
// back in their roles, so the same executable graph replays for every following pair as long as nothing it captured by
// value changes (agent count, lattice shape, flags, parameters, buffers -- sim->graph_key).
static bool graph_usable(cdb_sim *sim, uint32_t flags, double cell_size) {
    if (!sim->use_graphs || sim->variant == 1 || sim->profiling || sim->strip || sim->n <= 0) return false;
    // the legacy / per-thread default streams cannot be captured (e.g. torch's default current stream)
    if (sim->stream == nullptr || sim->stream == cudaStreamLegacy || sim->stream == cudaStreamPerThread) return false;
    if (!(flags & CDB_STEP_AGENT_AGENT) || !(flags & CDB_STEP_INTEGRATOR)) return false;
    if (chain_usable(sim, flags, cell_size)) return false;      // large crowds: resident-order steps instead
    if (sim->lattice_fixed) return sim->cell_capacity >= sim->grid.ncell;
    return sim->auto_lattice_valid && sim->cell_size_lattice == cell_size && sim->auto_lattice_age + 2 <= 64 &&
           sim->fine_lattice == search_refinement(sim, cell_size) && sim->cell_capacity >= sim->grid.ncell;
}

static int run_graph_pair(cdb_sim *sim, uint32_t flags, double cell_size, double dt_min, double dt_max, bool log_dt) {
    const cdb_sim::GraphKey key{flags, cell_size, dt_min, dt_max, log_dt, sim->n, sim->grid.ncell, sim->grid.nx, sim->grid.ny,
                                sim->state_version, sim->cur.p, sim->d_cell_of_slot};
    const cdb_sim::GraphKey &k0 = sim->graph_key;
    const bool same = sim->graph_exec && k0.flags == key.flags && k0.cell_size == key.cell_size && k0.dt_min == key.dt_min &&
                      k0.dt_max == key.dt_max && k0.log == key.log && k0.n == key.n && k0.ncell == key.ncell && k0.nx == key.nx &&
                      k0.ny == key.ny && k0.version == key.version && k0.cur == key.cur && k0.cells == key.cells;
    if (same) {
        CK(cudaGraphLaunch(sim->graph_exec, sim->stream));
        sim->launches += sim->graph_launches;
        sim->iterations += 2;
        if (!sim->lattice_fixed) sim->auto_lattice_age += 2;
        return CDB_OK;
    }
    if (sim->graph_exec) { cudaGraphExecDestroy(sim->graph_exec); sim->graph_exec = nullptr; }
    const int64_t launches0 = sim->launches;
    cudaGraph_t graph = nullptr;
    if (cudaStreamBeginCapture(sim->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
        // this stream cannot be captured: keep working with plain launches
        cudaGetLastError();
        sim->use_graphs = false;
        CKS(issue_step(sim, flags, cell_size, dt_min, dt_max, log_dt));
        return issue_step(sim, flags, cell_size, dt_min, dt_max, log_dt);
    }
    // The two issue_step calls below only RECORD device work but advance the host bookkeeping (step counter, lattice age,
    // ping-pong and index-array roles).  If the capture cannot be turned into an executable graph, that bookkeeping is put
    // back and the two steps are issued as plain launches, so host and device never disagree about what has run.
    const int64_t it0 = sim->iterations, n0 = sim->n, dead0 = sim->n_dead, sorted0 = sim->n_sorted;
    const int age0 = sim->auto_lattice_age;
    const Soa cur0 = sim->cur, alt0 = sim->alt;
    int *const cos0 = sim->d_cell_of_slot, *const tmp0 = sim->d_order_tmp;
    const bool perm0 = sim->perm_valid, tables0 = sim->tables_valid, pending0 = sim->pairs_pending;
    const size_t ev0 = sim->ev_used;
    int rc = issue_step(sim, flags, cell_size, dt_min, dt_max, log_dt);
    if (rc == CDB_OK) rc = issue_step(sim, flags, cell_size, dt_min, dt_max, log_dt);
    cudaError_t e = cudaStreamEndCapture(sim->stream, &graph);
    if (rc == CDB_OK && e == cudaSuccess) {
        sim->graph_launches = sim->launches - launches0;
        e = cudaGraphInstantiate(&sim->graph_exec, graph, 0);
    }
    if (graph) cudaGraphDestroy(graph);
    if (rc != CDB_OK || e != cudaSuccess) {
        cudaGetLastError();
        sim->graph_exec = nullptr;
        sim->iterations = it0; sim->n = n0; sim->n_dead = dead0; sim->n_sorted = sorted0; sim->auto_lattice_age = age0;
        sim->cur = cur0; sim->alt = alt0; sim->d_cell_of_slot = cos0; sim->d_order_tmp = tmp0;
        sim->perm_valid = perm0; sim->tables_valid = tables0; sim->pairs_pending = pending0; sim->ev_used = ev0;
        sim->launches = launches0;
        if (rc != CDB_OK) return rc;                 // an argument / allocation error of the step itself: report it
        sim->use_graphs = false;                     // capture or instantiation refused: plain launches from now on
        CKS(issue_step(sim, flags, cell_size, dt_min, dt_max, log_dt));
        return issue_step(sim, flags, cell_size, dt_min, dt_max, log_dt);
    }
    sim->graph_key = key;
    // the capture issued the two steps on the host side (counters, buffer roles); run them
    CK(cudaGraphLaunch(sim->graph_exec, sim->stream));
    return CDB_OK;
}

// Small crowds: the whole crowd in one thread block, many steps per launch (small_kernel.cuh).  The all-pairs loop adds the
// same terms as the block list only where the pair set does not depend on the lattice (3 + 2 max R < cell_size).
static bool small_usable(const cdb_sim *sim, uint32_t flags, double cell_size) {
    return sim->small_max > 0 && sim->n > 0 && sim->n <= sim->small_max && sim->n <= SMALL_MAX && sim->variant == 3 && !sim->strip &&
           !sim->lattice_fixed && !sim->profiling && sim->n_dead == 0 && cell_size > 0.0 && std::isfinite(cell_size) &&
           (!(flags & CDB_STEP_AGENT_AGENT) || (SIGTH_SOC + 2.0 * sim->ext_max) * (1.0 + 1e-9) < cell_size);
}

static int issue_small_steps(cdb_sim *sim, uint32_t flags, double cell_size, double dt_min, double dt_max, bool log_dt, int64_t m) {
    SmallArgs a{};
    a.s = sim->cur; a.n = (int)sim->n;
    a.nav = sim->d_nav; a.n_nav = sim->n_nav;
    a.obs = sim->d_obstacles; a.n_obs = (int)sim->n_obstacles;
    a.flags = flags; a.cell_size = cell_size; a.dt_min = dt_min; a.dt_max = dt_max;
    a.dt_out = sim->d_dt; a.dt_log = log_dt && (flags & CDB_STEP_INTEGRATOR) ? sim->d_dt_log : nullptr;
    a.seed = sim->seed; a.step_ptr = sim->d_stepctr; a.n_steps = (int)m;
    const int threads = sim->model == CDB_MODEL_CIRCULAR ? SmallThreads<0>::value : SmallThreads<1>::value;
    a.group = 1;
    while (a.group < 16 && 2 * a.group * sim->n <= threads) a.group *= 2;
    if (sim->model == CDB_MODEL_CIRCULAR) LAUNCH(sim, k_small_steps<0>, 1, threads, 0, a);
    else LAUNCH(sim, k_small_steps<1>, 1, threads, 0, a);
    CK(cudaGetLastError());
    sim->cell_size = cell_size;
    sim->iterations += m;
    sim->tables_valid = false;
    sim->perm_valid = false;
    sim->chain_valid = false;
    return CDB_OK;
}

int cdb_step(cdb_sim *sim, uint32_t flags, double cell_size, double dt_min, double dt_max, int64_t n_steps, double *dt_out) {
    SIM_ENTRY();
    if (n_steps < 0) return fail(CDB_ERR_INVALID_VALUE, "negative n_steps");
    sim->last_flags = flags; sim->last_cell_size = cell_size; sim->last_dt_min = dt_min; sim->last_dt_max = dt_max;
    const bool log_dt = dt_out && (flags & CDB_STEP_INTEGRATOR);
    int64_t k = 0, copied = 0;      // steps done / dt values already returned
    const bool prof_saved = sim->profiling;
    int regrown = 0;
    if (flags & CDB_STEP_AGENT_AGENT) CKS(prepare_pairs(sim));
    if (sim->defer_sync && sim->pairs_inflight) {
        // a check of an earlier call may have completed meanwhile: steps it reports as not applied are repeated first
        bool overflow = false;
        int64_t dev_steps = sim->iterations;
        CKS(harvest_pairs(sim, &overflow, &dev_steps));
        if (overflow) {
            const int64_t missing = sim->iterations - dev_steps;
            sim->iterations = dev_steps;
            k -= missing;
        }
    }
    while (k < n_steps) {
        sim->profiling = prof_saved && sim->ev_used + 2 * PROF_EVENTS <= (size_t)PROFILE_MAX_STEPS * PROF_EVENTS;
        // a pair must not straddle the end of the dt ring (its first half would be overwritten before it is read back)
        const bool ring_ok = !log_dt || (sim->iterations % DT_LOG) != DT_LOG - 1;
        if (small_usable(sim, flags, cell_size)) {
            int64_t m = std::min<int64_t>(n_steps - k, 8192);
            if (log_dt) m = std::min<int64_t>(m, DT_LOG - (sim->iterations % DT_LOG));    // up to the end of the dt ring
            CKS(issue_small_steps(sim, flags, cell_size, dt_min, dt_max, log_dt, m));
            k += m;
        } else if (n_steps - k >= 2 && ring_ok && graph_usable(sim, flags, cell_size)) {
            CKS(run_graph_pair(sim, flags, cell_size, dt_min, dt_max, log_dt));
            k += 2;
        } else {
            CKS(issue_step(sim, flags, cell_size, dt_min, dt_max, log_dt));
            k += 1;
        }
        sim->profiling = prof_saved;
        // Long calls on kept block lists: look at the counters every 16 steps WITHOUT waiting (the copy queued at step 16 j + 8 is
        // read once its event has completed), so that the rebuild interval follows a crowd that speeds up and refused steps are
        // noticed within a few steps instead of at the next synchronisation point.
        if (!sim->defer_sync && sim->pairs_pending && n_steps - k > 8 && chain_usable(sim, flags, cell_size)) {
            if (sim->pairs_inflight) {
                bool refused = false;
                int64_t dev_now = sim->iterations;
                CKS(harvest_pairs(sim, &refused, &dev_now));
                if (refused) {
                    if (++regrown > 64) return fail(CDB_ERR_CAPACITY, "steps keep being refused (pair list overflow / stale search lattice)");
                    const int64_t missing = sim->iterations - dev_now;
                    sim->chain_valid = false;
                    sim->iterations = dev_now;
                    k -= missing;
                }
            } else if ((k & 15) == 8) {
                CKS(read_counters(sim));
                CK(cudaEventRecord(sim->ev_pairs, sim->stream));
                sim->pairs_inflight = true;
                sim->iterations_at_check = sim->iterations;
            }
        }
        // Synchronisation points: the dt ring is about to wrap, the call is complete, or (variant 3) every 64 steps, so
        // that steps the device did not apply because the pair list overflowed are noticed and repeated early.
        const bool ring_point = log_dt && sim->iterations % DT_LOG == 0;
        const bool pairs_point = sim->pairs_pending && (k == n_steps || (k & 63) == 0);
        // resident-order steps start with a rebuild interval of 1: look at the displacement of the first steps early
        const bool chain_probe = sim->pairs_pending && k == 2 && sim->rebuild_every == 1 && chain_usable(sim, flags, cell_size);
        if (ring_point || pairs_point || chain_probe || (log_dt && k == n_steps)) {
            bool overflow = false;
            int64_t dev_steps = sim->iterations;
            CKS(settle_pairs(sim, &overflow, &dev_steps, !log_dt && k == n_steps));
            // The device counts the steps it really applied.  Steps issued after an overflowing one ran on the unchanged
            // state and were not applied either; a step refused because its search lattice had gone stale may be followed
            // by applied ones (the next scheduled rebuild clears the condition), so the count decides, not the last step.
            const int64_t missing = sim->iterations - dev_steps;
            if (overflow || missing > 0) {
                if (++regrown > 64) return fail(CDB_ERR_CAPACITY, "steps keep being refused (pair list overflow / stale search lattice)");
                if (!overflow) { sim->chain_stale++; sim->rebuild_every = std::max(1, sim->rebuild_every / 2); }
                sim->chain_valid = false;
                sim->iterations = dev_steps;
                k -= missing;
            }
            if (log_dt && k > copied) {
                // slots of the steps not yet returned: they end at slot (iterations - 1) % DT_LOG and do not wrap
                const int64_t cnt = k - copied;
                const int64_t first_slot = (sim->iterations - cnt) % DT_LOG;
                CK(cudaMemcpyAsync(dt_out + copied, sim->d_dt_log + first_slot, cnt * sizeof(double), cudaMemcpyDeviceToHost, sim->stream));
                CK(sync_stream(sim));
                copied = k;
            }
        }
    }
    return CDB_OK;
}

int cdb_set_search_refinement(cdb_sim *sim, int refinement) {
    SIM_ENTRY();
    if (refinement < 0 || refinement > 2) return fail(CDB_ERR_INVALID_VALUE, "refinement must be 0 (automatic), 1 or 2");
    sim->fine_request = refinement;
    sim->tables_valid = false;
    sim->auto_lattice_valid = false;
    sim->state_version++;
    return CDB_OK;
}

int cdb_get_ext_max(cdb_sim *sim, double *ext_max) {
    SIM_ENTRY();
    if (!ext_max) return fail(CDB_ERR_INVALID_VALUE, "ext_max is NULL");
    *ext_max = sim->ext_max;
    return CDB_OK;
}

int cdb_set_pair_capacity(cdb_sim *sim, int64_t pairs) {
    SIM_ENTRY();
    if (pairs < 0) return fail(CDB_ERR_INVALID_VALUE, "negative pair capacity");
    sim->pair_cap_request = pairs;
    if (pairs > 0) {
        CK(sync_stream(sim));
        sim->pb.cap = 0;                // shrink as well as grow: reallocate at exactly this size
        CKS(ensure_pairs(sim, pairs));
    }
    return CDB_OK;
}

int cdb_get_pair_stats(cdb_sim *sim, int64_t *capacity, int64_t *found_last, int64_t *overflows) {
    SIM_ENTRY();
    bool overflow = false;
    CKS(settle_pairs(sim, &overflow, nullptr));
    if (capacity) *capacity = sim->pb.cap;
    if (found_last) *found_last = sim->h_pctr ? (int64_t)sim->h_pctr[0] : 0;
    if (overflows) *overflows = sim->pair_overflows;
    return CDB_OK;
}

int cdb_set_rebuild_policy(cdb_sim *sim, double skin_fraction, int64_t max_interval, int64_t min_agents) {
    SIM_ENTRY();
    if (!(skin_fraction >= 0.0) || skin_fraction > 1.0 || max_interval < 1 || min_agents < 0)
        return fail(CDB_ERR_INVALID_VALUE, "rebuild policy: 0 <= skin_fraction <= 1, max_interval >= 1, min_agents >= 0");
    sim->skin_frac = skin_fraction;
    sim->rebuild_max = (int)std::min<int64_t>(max_interval, 1024);
    sim->rebuild_every = std::min(sim->rebuild_every, sim->rebuild_max);
    sim->chain_min_agents = min_agents;
    sim->chain_enabled = max_interval > 1 && skin_fraction > 0.0;
    sim->policy_explicit = true;
    sim->chain_valid = false;
    sim->auto_lattice_valid = false;
    sim->state_version++;
    return CDB_OK;
}

int cdb_get_rebuild_stats(cdb_sim *sim, int64_t *rebuilds, int64_t *kept, int64_t *stale, int64_t *interval) {
    SIM_ENTRY();
    if (rebuilds) *rebuilds = sim->chain_rebuilds;
    if (kept) *kept = sim->chain_kept;
    if (stale) *stale = sim->chain_stale;
    if (interval) *interval = sim->rebuild_every;
    return CDB_OK;
}

int cdb_set_small_crowd_max(cdb_sim *sim, int64_t max_agents) {
    SIM_ENTRY();
    if (max_agents < 0) return fail(CDB_ERR_INVALID_VALUE, "negative size");
    sim->small_max = std::min<int64_t>(max_agents, SMALL_MAX);
    return CDB_OK;
}

int cdb_set_graphs(cdb_sim *sim, int enable) {
    SIM_ENTRY();
    sim->use_graphs = enable != 0;
    return CDB_OK;
}

int cdb_set_variant(cdb_sim *sim, int variant) {
    SIM_ENTRY();
    if (variant < 1 || variant > 3) return fail(CDB_ERR_INVALID_VALUE, "unknown kernel variant %d", variant);
    sim->variant = variant;
    sim->state_version++;
    return CDB_OK;
}

int64_t cdb_launch_count(const cdb_sim *sim) { return sim ? sim->launches : -1; }

int cdb_profile_enable(cdb_sim *sim, int enable) {
    SIM_ENTRY();
    sim->profiling = enable != 0;
    sim->ev_used = 0;
    return CDB_OK;
}

int cdb_profile_read_phases(cdb_sim *sim, double ms[5], int64_t *steps) {
    SIM_ENTRY();
    if (!ms) return fail(CDB_ERR_INVALID_VALUE, "ms is NULL");
    CK(sync_stream(sim));
    for (int j = 0; j < PROF_EVENTS - 1; ++j) ms[j] = 0.0;
    const size_t n = sim->ev_used / PROF_EVENTS;
    for (size_t k = 0; k < n; ++k)
        for (int j = 0; j < PROF_EVENTS - 1; ++j) {
            float t = 0.f;
            CK(cudaEventElapsedTime(&t, sim->ev_pool[PROF_EVENTS * k + j], sim->ev_pool[PROF_EVENTS * k + j + 1]));
            ms[j] += t;
        }
    if (steps) *steps = (int64_t)n;
    sim->ev_used = 0;
    return CDB_OK;
}

int cdb_profile_read(cdb_sim *sim, double ms[3], int64_t *steps) {
    double p[PROF_EVENTS - 1];
    CKS(cdb_profile_read_phases(sim, p, steps));
    ms[0] = p[0]; ms[1] = p[1] + p[2] + p[3]; ms[2] = p[4];
    return CDB_OK;
}

int cdb_get_time(cdb_sim *sim, double *time_tot, int64_t *iterations) {
    SIM_ENTRY();
    CKS(catch_up(sim));
    CK(cudaMemcpyAsync(sim->h_dt, sim->d_dt, 2 * sizeof(double), cudaMemcpyDeviceToHost, sim->stream));
    CK(sync_stream(sim));
    if (time_tot) *time_tot = sim->h_dt[1];
    if (iterations) *iterations = sim->iterations;
    return check_device_error(sim);
}

// ---- block list exports -------------------------------------------------------------------------------------------
int cdb_build_block_list(cdb_sim *sim, double cell_size) { SIM_ENTRY_NODE(); return build_block_list(sim, cell_size, false, nullptr, false); }

// the exports describe the cell_size lattice of the reference; tables built on the finer search lattice are rebuilt
static int coarse_tables(cdb_sim *sim) {
    if (sim->tables_valid && sim->fine != 1) return build_block_list(sim, sim->cell_size, false, nullptr, false, 1);
    return CDB_OK;
}

int cdb_get_grid(cdb_sim *sim, int64_t grid[4]) {
    SIM_ENTRY();
    CKS(coarse_tables(sim));
    if (!sim->tables_valid) return fail(CDB_ERR_STATE, "no block list has been built for the current positions");
    grid[0] = sim->grid.ix_min; grid[1] = sim->grid.iy_min; grid[2] = sim->grid.nx; grid[3] = sim->grid.ny;
    return CDB_OK;
}

int cdb_get_cell_ids(cdb_sim *sim, int64_t *cell_of_agent, int64_t n) {
    SIM_ENTRY();
    CKS(coarse_tables(sim));
    if (!sim->tables_valid) return fail(CDB_ERR_STATE, "no block list has been built for the current positions");
    if (n != sim->n) return fail(CDB_ERR_INVALID_VALUE, "size mismatch");
    if (n == 0) return CDB_OK;
    long long *d = nullptr;
    CKS(dev_alloc(&d, (size_t)n));
    LAUNCH(sim, k_export_cell_ids, cdiv(n, 256), 256, 0, sim->cur.id, sim->perm_valid ? sim->d_order : nullptr, sim->d_cell_of_slot, (int)n, d);
    CK(cudaMemcpyAsync(cell_of_agent, d, n * sizeof(long long), cudaMemcpyDeviceToHost, sim->stream));
    CK(sync_stream(sim));
    cudaFree(d);
    return CDB_OK;
}

int cdb_get_cell_tables(cdb_sim *sim, int64_t *points_indices, int64_t n, int64_t *cells_count, int64_t *cells_offset, int64_t n_cells) {
    SIM_ENTRY();
    CKS(coarse_tables(sim));
    if (!sim->tables_valid) return fail(CDB_ERR_STATE, "no block list has been built for the current positions");
    if (n != sim->n || n_cells != sim->grid.ncell) return fail(CDB_ERR_INVALID_VALUE, "size mismatch (n %lld vs %lld, cells %lld vs %lld)", (long long)n, (long long)sim->n, (long long)n_cells, (long long)sim->grid.ncell);
    if (n == 0) return CDB_OK;
    long long *d = nullptr;
    const int64_t m = n > n_cells ? n : n_cells;
    CKS(dev_alloc(&d, (size_t)m));
    LAUNCH(sim, k_widen, cdiv(n, 256), 256, 0, sim->cur.id, sim->perm_valid ? sim->d_order : nullptr, (int)n, d);
    CK(cudaMemcpyAsync(points_indices, d, n * sizeof(long long), cudaMemcpyDeviceToHost, sim->stream));
    LAUNCH(sim, k_widen, cdiv(n_cells, 256), 256, 0, sim->d_cell_count, nullptr, (int)n_cells, d);
    CK(cudaMemcpyAsync(cells_count, d, n_cells * sizeof(long long), cudaMemcpyDeviceToHost, sim->stream));
    LAUNCH(sim, k_widen, cdiv(n_cells, 256), 256, 0, sim->d_cell_start, nullptr, (int)n_cells, d);
    CK(cudaMemcpyAsync(cells_offset, d, n_cells * sizeof(long long), cudaMemcpyDeviceToHost, sim->stream));
    CK(sync_stream(sim));
    cudaFree(d);
    return CDB_OK;
}

int cdb_get_neighbor_pairs(cdb_sim *sim, int64_t *pairs, int64_t cap, int64_t *count) {
    SIM_ENTRY();
    CKS(coarse_tables(sim));
    if (!sim->tables_valid) return fail(CDB_ERR_STATE, "no block list has been built for the current positions");
    if (!count) return fail(CDB_ERR_INVALID_VALUE, "count is NULL");
    *count = 0;
    if (sim->n == 0) return CDB_OK;
    long long *d = nullptr;
    CKS(dev_alloc(&d, (size_t)(cap > 0 ? 2 * cap : 1)));
    CK(cudaMemsetAsync(sim->d_pair_count, 0, sizeof(unsigned long long), sim->stream));
    LAUNCH(sim, k_export_pairs, cdiv(sim->n, 128), 128, 0, sim->cur.id, sim->perm_valid ? sim->d_order : nullptr, (int)sim->n, sim->d_grid, sim->d_cell_of_slot, sim->d_cell_start,
                                                             sim->d_cell_count, d, cap > 0 && pairs ? cap : 0, sim->d_pair_count);
    unsigned long long c = 0;
    CK(cudaMemcpyAsync(&c, sim->d_pair_count, sizeof(c), cudaMemcpyDeviceToHost, sim->stream));
    CK(sync_stream(sim));
    *count = (int64_t)c;
    if (pairs && cap > 0) {
        int64_t m = (int64_t)c < cap ? (int64_t)c : cap;
        CK(cudaMemcpy(pairs, d, 2 * m * sizeof(long long), cudaMemcpyDeviceToHost));
    }
    cudaFree(d);
    return CDB_OK;
}

// ---- fixed lattice / strips -----------------------------------------------------------------------------------------
int cdb_set_lattice(cdb_sim *sim, int64_t ix_min, int64_t iy_min, int64_t nx, int64_t ny) {
    SIM_ENTRY();
    if (nx <= 0 || ny <= 0 || (double)nx * (double)ny > 2.0e9) return fail(CDB_ERR_INVALID_VALUE, "bad lattice shape");
    CK(sync_stream(sim));
    sim->grid = Grid{ix_min, iy_min, nx, ny, nx * ny, 0, nx - 1};
    sim->lattice_fixed = true;
    sim->tables_valid = false;
    sim->state_version++;
    CK(cudaMemcpy(sim->d_grid, &sim->grid, sizeof(Grid), cudaMemcpyHostToDevice));
    CKS(ensure_cells(sim, sim->grid.ncell));
    return CDB_OK;
}
int cdb_clear_lattice(cdb_sim *sim) {
    SIM_ENTRY();
    sim->lattice_fixed = false;
    sim->tables_valid = false;
    sim->state_version++;
    return CDB_OK;
}

int cdb_set_strip(cdb_sim *sim, int64_t ix_min, int64_t iy_min, int64_t nx_owned, int64_t ny, int has_left, int has_right,
                  int64_t halo_cap, int64_t migrant_cap) {
    SIM_ENTRY();
    if (nx_owned <= 0 || ny <= 0 || halo_cap < 0 || migrant_cap < 0) return fail(CDB_ERR_INVALID_VALUE, "bad strip shape");
    CK(sync_stream(sim));
    sim->strip = true;
    sim->has_left = has_left ? 1 : 0;
    sim->has_right = has_right ? 1 : 0;
    sim->halo_cap = halo_cap;
    sim->mig_cap = migrant_cap;
    const int64_t nx = nx_owned + sim->has_left + sim->has_right;
    // the strip lattice in search cells (cell_size / f): every cell_size column is f columns wide; ghost blocks first / last
    const int f = sim->fine_request == 2 && sim->variant != 1 ? 2 : 1;
    sim->strip_fine = f;
    // kept block lists (cdb_set_rebuild_policy before cdb_set_strip): the columns the caller partitions are columns of the
    // WIDENED cells, cell_size * (1 + skin) -- every rank bins, owns and hands over agents on that lattice
    sim->strip_scale = sim->policy_explicit && sim->chain_enabled && sim->rebuild_max > 1 && sim->skin_frac > 0.0 && sim->variant == 3 ? 1.0 + sim->skin_frac : 1.0;
    sim->strip_kind = 0;
    sim->chain_valid = false;
    sim->strip_ix0 = ix_min - sim->has_left; sim->strip_col_lo = sim->has_left; sim->strip_col_hi = nx - 1 - sim->has_right;
    sim->grid = Grid{f * (ix_min - sim->has_left), f * iy_min, f * nx, f * ny, f * nx * f * ny, f * sim->has_left, f * (nx - sim->has_right) - 1};
    sim->lattice_fixed = true;
    sim->tables_valid = false;
    CK(cudaMemcpy(sim->d_grid, &sim->grid, sizeof(Grid), cudaMemcpyHostToDevice));
    sim->dev_counts = true;
    sim->steps_since_refresh = 0;
    LAUNCH(sim, k_counts_set, 1, 32, 0, sim->d_counts, (int)sim->n);
    CK(sync_stream(sim));
    return alloc_ghost_tail(sim);
}

int cdb_set_agent_ids(cdb_sim *sim, const int64_t *ids, int64_t n) {
    SIM_ENTRY();
    if (n != sim->n) return fail(CDB_ERR_INVALID_VALUE, "id count does not match the uploaded agents");
    if (n == 0) return CDB_OK;
    long long *d = nullptr;
    CKS(dev_alloc(&d, (size_t)n));
    CK(cudaMemcpyAsync(d, ids, n * sizeof(long long), cudaMemcpyHostToDevice, sim->stream));
    LAUNCH(sim, k_set_ids, cdiv(n, 256), 256, 0, sim->cur, d, (int)n);
    CK(sync_stream(sim));
    cudaFree(d);
    return CDB_OK;
}

int64_t cdb_halo_buffer_doubles(const cdb_sim *sim) {
    if (!sim) return -1;
    const int rec = sim->model == CDB_MODEL_CIRCULAR ? REC_CIRC : REC_THREE;
    return MSG_HEADER + halo_counts_doubles(sim->grid.ny * sim->strip_fine) + sim->halo_cap * rec;
}
int64_t cdb_migrant_buffer_doubles(const cdb_sim *sim) { return sim ? MSG_HEADER + sim->mig_cap * (sim->n_planes + 2) : -1; }

// the flag arrays that travel with migrating agents (strip mode: indexed by global agent id, see AgentFlags)
static AgentFlags strip_flags(const cdb_sim *sim) {
    AgentFlags f{};
    if (!sim->strip || sim->n_global <= 0) return f;
    if (sim->d_active && sim->active_n == sim->n_global) f.active = sim->d_active;
    if (sim->d_reached && sim->reached_n == sim->n_global && sim->n_polygons[CDB_POLY_TARGETS] > 0) {
        f.reached = sim->d_reached; f.stride = sim->n_global; f.np = (int)sim->n_polygons[CDB_POLY_TARGETS];
    }
    return f;
}

static int strip_begin_impl(cdb_sim *sim, uint32_t flags, double cell_size, double *halo_left_out, double *halo_right_out, bool direct) {
    if (!sim->strip) return fail(CDB_ERR_STATE, "cdb_set_strip has not been called");
    if (sim->variant == 1) return fail(CDB_ERR_STATE, "the strip decomposition needs kernel variant 2 or 3");
    if (direct && !sim->x_connected) return fail(CDB_ERR_STATE, "cdb_strip_exchange_connect_* has not been called");
    const int kind = sim->strip_kind;
    const bool chain = kind != 0, kept = kind == 2 || kind == 3;
    if (chain && (sim->strip_scale <= 1.0 || !use_pairs(sim) || !(flags & CDB_STEP_AGENT_AGENT) || !(flags & CDB_STEP_INTEGRATOR)))
        return fail(CDB_ERR_STATE, "kept block lists need cdb_set_rebuild_policy before cdb_set_strip and whole integrating steps");
    if (sim->strip_scale > 1.0 && !((SIGTH_SOC + 2.0 * sim->ext_max) * (1.0 + 1e-9) < cell_size))
        return fail(CDB_ERR_STATE, "widened search cells need 3 + 2 max R < cell_size");
    if (kept && !(sim->chain_valid && sim->chain_version == sim->state_version && sim->chain_cell_size == cell_size))
        return fail(CDB_ERR_STATE, "a kept step must follow a step of the same run of kept block lists");
    if (flags & CDB_STEP_AGENT_AGENT) CKS(prepare_pairs(sim));
    CKS(prof_mark(sim));
    if (!kept) LAUNCH(sim, k_vmax_init, 1, 32, 0, sim->d_vmax);
    LAUNCH(sim, k_counters_zero, 1, 32, 0, sim->d_counters, 4);      // migrant counters of this step
    if (chain) LAUNCH(sim, k_chain_begin, 1, 32, 0, sim->d_chain, sim->d_vmax, kept ? 0 : 1);
    const int f = sim->strip_fine;
    if (!kept) {
        CKS(build_block_list(sim, cell_size, false, sim->d_vmax, false, f, sim->strip_scale));
        if (chain) {
            sim->chain_rebuilds++;
            sim->chain_cell_size = cell_size;
            sim->drift_limit = 0.5 * (cell_size * sim->strip_scale - (SIGTH_SOC + 2.0 * sim->ext_max) * (1.0 + 1e-9)) * (1.0 - 1e-9);
        }
    } else {
        sim->tables_valid = true;        // of the kept order: same slots, same cells, same ghost columns
        sim->chain_kept++;
    }
    const int rec = sim->model == CDB_MODEL_CIRCULAR ? REC_CIRC : REC_THREE;
    // a halo message is one column of the strip lattice = f consecutive columns of the search lattice = f * ny consecutive
    // cells; on a kept step it is the SAME slice of slots, with the records k_finish wrote in place
    const int nyb = (int)sim->grid.ny * f;
    const unsigned long long seq = (unsigned long long)sim->iterations + 1;
    // direct exchange: two receive buffers per side, alternating with the step -- on kept steps no migrant handshake
    // separates my next halo from the neighbour's reading of the previous one
    const size_t par_off = direct ? (size_t)(seq & 1ULL) * (size_t)cdb_halo_buffer_doubles(sim) : 0;
    if (sim->has_left && halo_left_out)
        LAUNCH(sim, k_halo_pack, HALO_BLOCKS, 256, 0, sim->d_nbr, rec, sim->d_cell_start, sim->d_cell_count, (int)sim->strip_col_lo, nyb, halo_left_out + par_off,
               (long long)sim->halo_cap, sim->d_error, sim->x_done, direct ? sim->p_flags[0] + 1 : nullptr, seq);   // I am its right neighbour
    if (sim->has_right && halo_right_out)
        LAUNCH(sim, k_halo_pack, HALO_BLOCKS, 256, 0, sim->d_nbr, rec, sim->d_cell_start, sim->d_cell_count, (int)sim->strip_col_hi, nyb, halo_right_out + par_off,
               (long long)sim->halo_cap, sim->d_error, sim->x_done ? sim->x_done + 1 : nullptr, direct ? sim->p_flags[1] + 0 : nullptr, seq);
    CK(cudaGetLastError());
    return CDB_OK;
}

int cdb_strip_begin(cdb_sim *sim, uint32_t flags, double cell_size, double *halo_left_out, double *halo_right_out) {
    SIM_ENTRY();
    return strip_begin_impl(sim, flags, cell_size, halo_left_out, halo_right_out, false);
}

int cdb_strip_begin_direct(cdb_sim *sim, uint32_t flags, double cell_size, int send_halo) {
    SIM_ENTRY();
    return strip_begin_impl(sim, flags, cell_size, send_halo ? sim->p_halo[0] : nullptr, send_halo ? sim->p_halo[1] : nullptr, true);
}

int cdb_strip_export_vmax(cdb_sim *sim, double *dev_vmax4) {
    SIM_ENTRY();
    LAUNCH(sim, k_vmax_export, 1, 32, 0, sim->d_vmax, dev_vmax4);
    CK(cudaGetLastError());
    return CDB_OK;
}
int cdb_strip_import_vmax(cdb_sim *sim, const double *dev_vmax4) {
    SIM_ENTRY();
    LAUNCH(sim, k_vmax_import, 1, 32, 0, dev_vmax4, sim->d_vmax);
    CK(cudaGetLastError());
    return CDB_OK;
}

static int strip_finish_impl(cdb_sim *sim, uint32_t flags, double dt_min, double dt_max, const double *halo_left_in,
                             const double *halo_right_in, double *mig_left_out, double *mig_right_out, bool direct) {
    if (!sim->strip || !sim->tables_valid) return fail(CDB_ERR_STATE, "cdb_strip_begin must precede cdb_strip_finish");
    const int rec = sim->model == CDB_MODEL_CIRCULAR ? REC_CIRC : REC_THREE;
    const int ny = (int)sim->grid.ny * sim->strip_fine;      // cells of one cell_size column
    const int base_l = (int)sim->capacity, base_r = (int)(sim->capacity + sim->halo_cap);
    const unsigned long long seq = (unsigned long long)sim->iterations + 1;
    if (sim->has_left) {
        if (halo_left_in) LAUNCH(sim, k_halo_unpack, HALO_BLOCKS, 1024, 0, halo_left_in, rec, sim->d_nbr, sim->d_nbr_sweep, sim->d_cell_of_slot, sim->d_cell_start, sim->d_cell_count,
                                 0, ny, base_l, (long long)sim->halo_cap, sim->d_error, sim->d_par, direct ? sim->x_flags + 0 : nullptr, seq);
        else LAUNCH(sim, k_ghost_clear, 4, 256, 0, sim->d_cell_start, sim->d_cell_count, 0, ny, base_l);
    }
    if (sim->has_right) {
        const int col = (int)sim->strip_col_hi + 1;
        if (halo_right_in) LAUNCH(sim, k_halo_unpack, HALO_BLOCKS, 1024, 0, halo_right_in, rec, sim->d_nbr, sim->d_nbr_sweep, sim->d_cell_of_slot, sim->d_cell_start, sim->d_cell_count,
                                  col, ny, base_r, (long long)sim->halo_cap, sim->d_error, sim->d_par, direct ? sim->x_flags + 1 : nullptr, seq);
        else LAUNCH(sim, k_ghost_clear, 4, 256, 0, sim->d_cell_start, sim->d_cell_count, col, ny, base_r);
    }
    CKS(prof_mark(sim));
    const int kind = sim->strip_kind;
    const bool chain = kind != 0, kept = kind == 2 || kind == 3;
    const bool migrate = kind == 0 || kind == 3 || kind == 4;   // steps whose successor rebuilds the block list hand their leavers over
    // variant 3, integrating step (every step rebuilds): the finish kernel itself hands over the agents that left the owned columns
    const bool fused_migrants = use_pairs(sim) && (flags & CDB_STEP_INTEGRATOR) && sim->n_sorted > 0 && kind == 0;
    MigrantArgs mig{};
    mig.enabled = 1; mig.cell_size = sim->cell_size * sim->strip_scale; mig.ix0 = sim->strip_ix0;
    mig.col_lo = (int)sim->strip_col_lo; mig.col_hi = (int)sim->strip_col_hi; mig.has_left = sim->has_left; mig.has_right = sim->has_right;
    mig.msg_left = mig_left_out; mig.msg_right = mig_right_out; mig.cap = sim->mig_cap; mig.counters = sim->d_counters; mig.error = sim->d_error;
    mig.flags = strip_flags(sim);
    if (mig.has_left && !mig_left_out) mig.has_left = 0;
    if (mig.has_right && !mig_right_out) mig.has_right = 0;
    sim->chain_step = chain;
    sim->chain_inplace = kept;
    const int rc = launch_step_kernel(sim, flags, dt_min, dt_max, nullptr, fused_migrants ? &mig : nullptr);
    sim->chain_step = false;
    sim->chain_inplace = false;
    CKS(rc);
    if (chain) {
        LAUNCH(sim, k_chain_end, 1, 32, 0, sim->d_chain, sim->pb.ctr, (long long)sim->pb.cap);
        sim->chain_valid = true;
        sim->chain_version = sim->state_version;
    }
    CKS(prof_mark(sim));
    sim->iterations++;
    LAUNCH(sim, k_step_advance, 1, 32, 0, sim->d_stepctr, (const unsigned long long *)nullptr, 0LL);
    if (migrate) {
        if (sim->n > 0 && !fused_migrants)
            LAUNCH(sim, k_migrants_pack, cdiv(sim->n, 256), 256, 0, sim->cur, (int)sim->n, sim->dev_counts ? &sim->d_counts->slots : nullptr, sim->n_planes,
                   sim->cell_size * sim->strip_scale, sim->strip_ix0, (int)sim->strip_col_lo, (int)sim->strip_col_hi, sim->has_left, sim->has_right,
                   mig_left_out, mig_right_out, (long long)sim->mig_cap, sim->d_counters, sim->d_error, strip_flags(sim));
        LAUNCH(sim, k_migrants_header, 1, 32, 0, sim->has_left ? mig_left_out : nullptr, sim->has_right ? mig_right_out : nullptr, sim->d_counters,
               (long long)sim->mig_cap, direct && sim->has_left ? sim->p_flags[0] + 3 : nullptr, direct && sim->has_right ? sim->p_flags[1] + 2 : nullptr, seq);
        if (chain) sim->chain_valid = false;         // slots were vacated: the next step has to rebuild
    }
    CK(cudaGetLastError());
    CKS(prof_mark(sim));
    return CDB_OK;
}

int cdb_strip_finish(cdb_sim *sim, uint32_t flags, double dt_min, double dt_max, const double *halo_left_in,
                     const double *halo_right_in, double *mig_left_out, double *mig_right_out) {
    SIM_ENTRY();
    return strip_finish_impl(sim, flags, dt_min, dt_max, halo_left_in, halo_right_in, mig_left_out, mig_right_out, false);
}

int cdb_strip_finish_direct(cdb_sim *sim, uint32_t flags, double dt_min, double dt_max, int recv_halo) {
    SIM_ENTRY();
    if (!sim->x_connected) return fail(CDB_ERR_STATE, "cdb_strip_exchange_connect_* has not been called");
    const size_t par_off = (size_t)(((unsigned long long)sim->iterations + 1) & 1ULL) * (size_t)cdb_halo_buffer_doubles(sim);
    return strip_finish_impl(sim, flags, dt_min, dt_max, recv_halo ? sim->x_halo_in[0] + par_off : nullptr,
                             recv_halo ? sim->x_halo_in[1] + par_off : nullptr, sim->p_mig[0], sim->p_mig[1], true);
}

// exact counts back on the host (one sync): slots in use, slots vacated by the last step's migrants; also surfaces device errors
static int strip_refresh(cdb_sim *sim) {
    CK(cudaMemcpyAsync(sim->h_counters, sim->d_counters, 4 * sizeof(int), cudaMemcpyDeviceToHost, sim->stream));
    CK(cudaMemcpyAsync(sim->h_counts, sim->d_counts, sizeof(DevCounts), cudaMemcpyDeviceToHost, sim->stream));
    CKS(check_device_error(sim));    // synchronizes
    sim->n = sim->h_counts->slots;
    sim->n_dead = std::min<int64_t>(sim->h_counters[0], sim->mig_cap) + std::min<int64_t>(sim->h_counters[1], sim->mig_cap);
    sim->steps_since_refresh = 0;
    if (sim->pairs_pending) { bool overflow = false; CKS(settle_pairs(sim, &overflow, nullptr)); }   // grows the pair list early
    return CDB_OK;
}

static int strip_absorb_impl(cdb_sim *sim, const double *mig_left_in, const double *mig_right_in, int64_t *n_out, bool direct) {
    if (!sim->strip) return fail(CDB_ERR_STATE, "cdb_set_strip has not been called");
    const int g = cdiv(sim->mig_cap > 0 ? sim->mig_cap : 1, 128);
    const int *slots_dev = &sim->d_counts->slots;
    const unsigned long long seq = (unsigned long long)sim->iterations;      // the step cdb_strip_finish just completed
    if (sim->has_left && mig_left_in)
        LAUNCH(sim, k_migrants_unpack, g, 128, 0, mig_left_in, sim->cur, (int)sim->n, slots_dev, sim->n_planes, (long long)sim->capacity, sim->d_counters, sim->d_error,
               direct ? sim->x_flags + 2 : nullptr, seq, strip_flags(sim));
    if (sim->has_right && mig_right_in)
        LAUNCH(sim, k_migrants_unpack, g, 128, 0, mig_right_in, sim->cur, (int)sim->n, slots_dev, sim->n_planes, (long long)sim->capacity, sim->d_counters, sim->d_error,
               direct ? sim->x_flags + 3 : nullptr, seq, strip_flags(sim));
    // The exact counts stay on the device; the host keeps an upper bound for its launch sizes and only synchronises every
    // STRIP_REFRESH steps (or when the bound would not fit the allocation, or when the caller asks for the exact count).
    // The bound grows by what can plausibly arrive in one step (agents move ~1 cm per step: a few per km of border), not by
    // the capacity of the migrant messages; a burst beyond it raises a device error instead of being skipped silently.
    const int64_t grow = 2 * std::min<int64_t>(sim->mig_cap, std::max<int64_t>(64, sim->n / 2048));
    const int64_t bound = sim->n + grow;
    const bool refresh = ++sim->steps_since_refresh >= STRIP_REFRESH || bound > sim->capacity || n_out;
    LAUNCH(sim, k_counts_after_absorb, 1, 32, 0, sim->d_counts, sim->d_counters, refresh ? -1LL : (long long)bound, sim->d_error);
    CK(cudaGetLastError());
    sim->tables_valid = false;
    if (refresh) {
        CKS(strip_refresh(sim));
        if (n_out) *n_out = sim->n - sim->n_dead;
    } else {
        sim->n = bound;
    }
    return CDB_OK;
}

int cdb_strip_absorb(cdb_sim *sim, const double *mig_left_in, const double *mig_right_in, int64_t *n_out) {
    SIM_ENTRY();
    return strip_absorb_impl(sim, mig_left_in, mig_right_in, n_out, false);
}

int cdb_strip_absorb_direct(cdb_sim *sim, int64_t *n_out) {
    SIM_ENTRY();
    if (!sim->x_connected) return fail(CDB_ERR_STATE, "cdb_strip_exchange_connect_* has not been called");
    return strip_absorb_impl(sim, sim->x_mig_in[0], sim->x_mig_in[1], n_out, true);
}

// ---- one-sided exchange: receive buffers + flags, exported as CUDA IPC handles (one process per GPU) or raw pointers -------
int cdb_strip_exchange_alloc(cdb_sim *sim) {
    SIM_ENTRY();
    if (!sim->strip) return fail(CDB_ERR_STATE, "cdb_set_strip has not been called");
    const size_t hb = (size_t)cdb_halo_buffer_doubles(sim), mb = (size_t)cdb_migrant_buffer_doubles(sim);
    for (int k = 0; k < 2; ++k) {
        CKS(dev_alloc(&sim->x_halo_in[k], 2 * hb)); CKS(dev_alloc(&sim->x_mig_in[k], mb));     // halo: one buffer per step parity
        CK(cudaMemset(sim->x_halo_in[k], 0, 2 * hb * sizeof(double))); CK(cudaMemset(sim->x_mig_in[k], 0, mb * sizeof(double)));
    }
    CKS(dev_alloc(&sim->x_flags, 8)); CKS(dev_alloc(&sim->x_done, 4));
    CK(cudaMemset(sim->x_flags, 0, 8 * sizeof(unsigned long long))); CK(cudaMemset(sim->x_done, 0, 4 * sizeof(unsigned int)));
    return CDB_OK;
}

int64_t cdb_strip_exchange_handle_bytes(void) { return 5 * (int64_t)sizeof(cudaIpcMemHandle_t); }

int cdb_strip_exchange_handles(cdb_sim *sim, void *out) {
    SIM_ENTRY();
    if (!sim->x_flags || !out) return fail(CDB_ERR_STATE, "cdb_strip_exchange_alloc has not been called");
    cudaIpcMemHandle_t *h = (cudaIpcMemHandle_t *)out;
    CK(cudaIpcGetMemHandle(&h[0], sim->x_halo_in[0])); CK(cudaIpcGetMemHandle(&h[1], sim->x_halo_in[1]));
    CK(cudaIpcGetMemHandle(&h[2], sim->x_mig_in[0])); CK(cudaIpcGetMemHandle(&h[3], sim->x_mig_in[1]));
    CK(cudaIpcGetMemHandle(&h[4], sim->x_flags));
    return CDB_OK;
}

// left / right: the 5 handles of the neighbour's cdb_strip_exchange_handles (NULL: no neighbour on that side)
int cdb_strip_exchange_connect_ipc(cdb_sim *sim, const void *left, const void *right) {
    SIM_ENTRY();
    if (!sim->x_flags) return fail(CDB_ERR_STATE, "cdb_strip_exchange_alloc has not been called");
    const void *side[2] = {left, right};
    for (int k = 0; k < 2; ++k) {
        if (!side[k]) continue;
        const cudaIpcMemHandle_t *h = (const cudaIpcMemHandle_t *)side[k];
        for (int j = 0; j < 5; ++j) CK(cudaIpcOpenMemHandle(&sim->p_opened[k][j], h[j], cudaIpcMemLazyEnablePeerAccess));
        // I write into the buffer the neighbour reads from MY side: its "from right" buffers if it is my left neighbour
        sim->p_halo[k] = (double *)sim->p_opened[k][k == 0 ? 1 : 0];
        sim->p_mig[k] = (double *)sim->p_opened[k][k == 0 ? 3 : 2];
        sim->p_flags[k] = (unsigned long long *)sim->p_opened[k][4];
    }
    sim->x_connected = true;
    return CDB_OK;
}

// the same inside one process (all strips on one device, or peer-accessible devices): the neighbours' sims directly
int cdb_strip_exchange_connect_local(cdb_sim *sim, cdb_sim *left, cdb_sim *right) {
    SIM_ENTRY();
    if (!sim->x_flags) return fail(CDB_ERR_STATE, "cdb_strip_exchange_alloc has not been called");
    cdb_sim *side[2] = {left, right};
    for (int k = 0; k < 2; ++k) {
        if (!side[k]) continue;
        if (!side[k]->x_flags) return fail(CDB_ERR_STATE, "the neighbour has no exchange buffers");
        sim->p_halo[k] = side[k]->x_halo_in[k == 0 ? 1 : 0];
        sim->p_mig[k] = side[k]->x_mig_in[k == 0 ? 1 : 0];
        sim->p_flags[k] = side[k]->x_flags;
    }
    sim->x_connected = true;
    return CDB_OK;
}

int cdb_strip_set_kind(cdb_sim *sim, int kind) {
    SIM_ENTRY();
    if (!sim->strip) return fail(CDB_ERR_STATE, "cdb_set_strip has not been called");
    if (kind < 0 || kind > 4) return fail(CDB_ERR_INVALID_VALUE, "step kind must be 0 .. 4");
    if (kind != 0 && sim->strip_scale <= 1.0) return fail(CDB_ERR_STATE, "kept block lists need cdb_set_rebuild_policy before cdb_set_strip");
    sim->strip_kind = kind;
    return CDB_OK;
}

int cdb_strip_drift(cdb_sim *sim, double *disp_last, double *disp_acc, double *drift_limit) {
    SIM_ENTRY();
    ChainState h;
    CK(cudaMemcpyAsync(&h, sim->d_chain, sizeof(ChainState), cudaMemcpyDeviceToHost, sim->stream));
    CKS(check_device_error(sim));     // synchronizes; a sweep that found its block list stale surfaces here
    if (disp_last) *disp_last = h.disp_last;
    if (disp_acc) *disp_acc = h.disp_acc;
    if (drift_limit) *drift_limit = sim->drift_limit;
    return CDB_OK;
}

int cdb_strip_count(cdb_sim *sim, int64_t *n_out) {
    SIM_ENTRY();
    if (!sim->strip || !n_out) return fail(CDB_ERR_STATE, "cdb_set_strip has not been called");
    CKS(strip_refresh(sim));
    *n_out = sim->n - sim->n_dead;
    return CDB_OK;
}

int cdb_export_agents(cdb_sim *sim, void *agents, int64_t *ids, int64_t cap, int64_t *count) {
    SIM_ENTRY();
    if (!count) return fail(CDB_ERR_INVALID_VALUE, "count is NULL");
    if (sim->dev_counts) CKS(strip_refresh(sim));
    const int64_t live = sim->n - sim->n_dead;
    *count = live;
    if (live == 0) return CDB_OK;
    if (cap < live || !agents || !ids) return fail(CDB_ERR_CAPACITY, "export buffer holds %lld agents, %lld needed", (long long)cap, (long long)live);
    uint8_t *d_rec = nullptr;
    long long *d_ids = nullptr;
    CKS(dev_alloc(&d_rec, (size_t)(live * sim->itemsize + 16)));
    CKS(dev_alloc(&d_ids, (size_t)live));
    CK(cudaMemsetAsync(sim->d_counters + 3, 0, sizeof(int), sim->stream));
    if (sim->model == CDB_MODEL_CIRCULAR) LAUNCH(sim, k_export_records<0>, cdiv(sim->n, 128), 128, 0, sim->cur, (int)sim->n, d_rec, d_ids, sim->d_counters + 3);
    else LAUNCH(sim, k_export_records<1>, cdiv(sim->n, 128), 128, 0, sim->cur, (int)sim->n, d_rec, d_ids, sim->d_counters + 3);
    CK(cudaMemcpyAsync(agents, d_rec, live * sim->itemsize, cudaMemcpyDeviceToHost, sim->stream));
    CK(cudaMemcpyAsync(ids, d_ids, live * sizeof(long long), cudaMemcpyDeviceToHost, sim->stream));
    CK(sync_stream(sim));
    cudaFree(d_rec); cudaFree(d_ids);
    return CDB_OK;
}

}  // extern "C"


// =====================================================================================================================
// collective motion
// =====================================================================================================================
namespace {
int ensure_states(cdb_sim *sim, int64_t n) {
    if (n <= sim->states_cap) return CDB_OK;
    const size_t cap = (size_t)(n < 1024 ? 1024 : n);
    CKS(dev_alloc(&sim->d_is_leader, cap)); CKS(dev_alloc(&sim->d_is_follower, cap));
    CKS(dev_alloc(&sim->d_has_a, cap)); CKS(dev_alloc(&sim->d_has_detected, cap));
    CKS(dev_alloc(&sim->d_index_leader, cap)); CKS(dev_alloc(&sim->d_familiar_exit, cap));
    CKS(dev_alloc(&sim->d_target_by_id, cap)); CKS(dev_alloc(&sim->d_detected, cap));
    CKS(dev_alloc(&sim->d_slot_of_id, cap)); CKS(dev_alloc(&sim->d_leader_ids, cap));
    CKS(dev_alloc(&sim->d_dir_a, 2 * cap)); CKS(dev_alloc(&sim->d_direction, 2 * cap));
    CKS(dev_alloc(&sim->d_lrec, 4 * cap)); CKS(dev_alloc(&sim->d_lnext, cap));
    CK(cudaMemset(sim->d_is_leader, 0, cap)); CK(cudaMemset(sim->d_is_follower, 0, cap));
    CK(cudaMemset(sim->d_index_leader, 0xff, cap * sizeof(long long))); CK(cudaMemset(sim->d_familiar_exit, 0xff, cap * sizeof(long long)));
    sim->states_cap = (int64_t)cap;
    sim->states_n = 0; sim->n_leaders = 0;
    return CDB_OK;
}

int collective_entry(cdb_sim *sim, bool need_states) {
    if (sim->strip) return fail(CDB_ERR_STATE, "the collective-motion nodes are not available in strip mode");
    if (need_states && sim->states_n != sim->n) return fail(CDB_ERR_STATE, "cdb_set_states must describe the %lld agents on the device first", (long long)sim->n);
    CKS(ensure_states(sim, sim->n));
    return CDB_OK;
}

int launch_leader_follower(cdb_sim *sim, double sight, double phi, double w_leader, bool with_herding, double w_direction) {
    const int n = (int)sim->n;
    const int L = (int)sim->n_leaders;
    LAUNCH(sim, k_slot_map, cdiv(n, 256), 256, 0, sim->cur, n, sim->d_slot_of_id, sim->d_target_by_id);
    // hash grid over the leaders when there are enough of them to make the all-leaders scan expensive
    int bits = 0;
    const bool grid = L > 32 && sight > 0.0 && std::isfinite(sight);
    if (grid) {
        bits = 10;
        while ((1LL << bits) < 4LL * L && bits < 26) ++bits;
        if ((1LL << bits) > sim->lhead_cap) { CKS(dev_alloc(&sim->d_lhead, (size_t)1 << bits)); sim->lhead_cap = 1LL << bits; }
        CK(cudaMemsetAsync(sim->d_lhead, 0xff, sizeof(int) << bits, sim->stream));
    }
    if (L > 0)
        LAUNCH(sim, k_leader_records, cdiv(L, 128), 128, 0, sim->cur, sim->d_leader_ids, L, sim->d_slot_of_id, sim->d_lrec, sight,
               grid ? sim->d_lhead : nullptr, sim->d_lnext, bits);
    LAUNCH(sim, k_leader_follower, cdiv(n, 128), 128, 0, sim->cur, n, sim->d_obstacles, (int)sim->n_obstacles, sim->d_leader_ids, L,
           sim->d_lrec, grid ? sim->d_lhead : nullptr, sim->d_lnext, bits, sim->d_target_by_id, n, sim->d_is_follower, sim->d_index_leader,
           sim->d_familiar_exit, sight, std::cos(phi), w_leader, with_herding ? sim->d_dir_a : nullptr, with_herding ? sim->d_has_a : nullptr,
           w_direction, sim->d_direction);
    CK(cudaGetLastError());
    sim->direction_valid = true;
    sim->state_version++;
    return CDB_OK;
}

int launch_herding(cdb_sim *sim, double sight, int64_t k, bool all_agents, double w_position, double phi, bool with_direction) {
    if (!(sight > 0.0) || !std::isfinite(sight)) return fail(CDB_ERR_INVALID_VALUE, "sight must be > 0");
    double cell = sight;      // the reference's block list (collective_motion.py:262-267)
    if (!all_agents && !sim->lattice_fixed && sim->n > 0) {
        // the herding step searches fine cells: about the radius expected to hold 3k agents at the crowd's mean density
        LAUNCH(sim, k_bbox_init, 1, 32, 0, sim->d_bbox);
        LAUNCH(sim, k_bbox, (cdiv(sim->n, 1024) < 1184 ? cdiv(sim->n, 1024) : 1184), 256, 0, sim->cur, (int)sim->n, sight, sim->d_bbox, sim->d_error);
        CK(cudaMemcpyAsync(sim->h_bbox, sim->d_bbox, 4 * sizeof(long long), cudaMemcpyDeviceToHost, sim->stream));
        CK(sync_stream(sim));
        CKS(check_device_error(sim));
        const double area = (double)(sim->h_bbox[1] - sim->h_bbox[0] + 1) * (double)(sim->h_bbox[3] - sim->h_bbox[2] + 1) * sight * sight;
        const double r0 = std::sqrt(3.0 * (double)k * area / (3.141592653589793 * (double)sim->n));
        if (std::isfinite(r0)) cell = std::min(sight, std::max(r0, sight / 16.0));
    }
    CKS(build_block_list(sim, cell, false, nullptr, false));
    sim->auto_lattice_valid = false;     // the step kernel's padded lattice was replaced by this one
    const int64_t live = sim->n_sorted;
    if ((int64_t)live * k > sim->knn_cap) { CKS(dev_alloc(&sim->d_knn, (size_t)(live * k))); sim->knn_cap = live * k; }
    const int rec = sim->model == CDB_MODEL_CIRCULAR ? REC_CIRC : REC_THREE;
    if (live > 0) {
        // all_agents (cdb_nearest_neighbors): rows in the reference's slot order; the herding step: the pruned search
        const int exact_cells = sim->lattice_fixed ? 0 : 1;
        if (all_agents)
            LAUNCH(sim, k_herding<true>, cdiv(live, 128), 128, 0, sim->cur, (int)live, (const int *)nullptr, sim->d_nbr, rec, sim->d_grid, sight, exact_cells,
                   sim->d_cell_of_slot, sim->d_cell_start, sim->d_cell_count, sim->d_order, sim->d_obstacles, (int)sim->n_obstacles, sight, (int)k,
                   sim->d_is_follower, 1, w_position, std::cos(phi), sim->d_knn, with_direction ? sim->d_dir_a : nullptr,
                   with_direction ? sim->d_has_a : nullptr);
        else
            LAUNCH(sim, k_herding<false>, cdiv(live, 128), 128, 0, sim->cur, (int)live, (const int *)nullptr, sim->d_nbr, rec, sim->d_grid, cell, exact_cells,
                   sim->d_cell_of_slot, sim->d_cell_start, sim->d_cell_count, sim->d_order, sim->d_obstacles, (int)sim->n_obstacles, sight, (int)k,
                   sim->d_is_follower, 0, w_position, std::cos(phi), sim->d_knn, with_direction ? sim->d_dir_a : nullptr,
                   with_direction ? sim->d_has_a : nullptr);
    }
    CK(cudaGetLastError());
    sim->state_version++;
    return CDB_OK;
}
}  // namespace

int cdb_set_states(cdb_sim *sim, const int64_t *target, const uint8_t *is_leader, const uint8_t *is_follower,
                   const int64_t *index_leader, const int64_t *familiar_exit, int64_t n) {
    SIM_ENTRY();
    if (sim->strip) return fail(CDB_ERR_STATE, "the collective-motion nodes are not available in strip mode");
    if (n != sim->n) return fail(CDB_ERR_INVALID_VALUE, "cdb_set_states: n = %lld but the device holds %lld agents", (long long)n, (long long)sim->n);
    CKS(ensure_states(sim, n));
    CK(sync_stream(sim));
    if (n > 0) {
        if (is_follower) CK(cudaMemcpy(sim->d_is_follower, is_follower, n, cudaMemcpyHostToDevice));
        if (index_leader) CK(cudaMemcpy(sim->d_index_leader, index_leader, n * sizeof(int64_t), cudaMemcpyHostToDevice));
        if (familiar_exit) CK(cudaMemcpy(sim->d_familiar_exit, familiar_exit, n * sizeof(int64_t), cudaMemcpyHostToDevice));
        if (is_leader) {
            CK(cudaMemcpy(sim->d_is_leader, is_leader, n, cudaMemcpyHostToDevice));
            std::vector<int> ids;
            for (int64_t i = 0; i < n; ++i) if (is_leader[i]) ids.push_back((int)i);
            sim->n_leaders = (int64_t)ids.size();
            if (!ids.empty()) CK(cudaMemcpy(sim->d_leader_ids, ids.data(), ids.size() * sizeof(int), cudaMemcpyHostToDevice));
        }
        if (target) {
            CK(cudaMemcpy(sim->d_target_by_id, target, n * sizeof(int64_t), cudaMemcpyHostToDevice));
            LAUNCH(sim, k_target_scatter, cdiv(n, 256), 256, 0, sim->cur, (int)n, sim->d_target_by_id);
            CK(sync_stream(sim));
        }
    }
    sim->states_n = n;
    sim->state_version++;
    return CDB_OK;
}

int cdb_get_states(cdb_sim *sim, int64_t *target, uint8_t *is_follower, int64_t *index_leader, int64_t n) {
    SIM_ENTRY();
    CKS(collective_entry(sim, true));
    if (n != sim->n) return fail(CDB_ERR_INVALID_VALUE, "cdb_get_states: n = %lld but the device holds %lld agents", (long long)n, (long long)sim->n);
    if (n == 0) return CDB_OK;
    if (target) LAUNCH(sim, k_slot_map, cdiv(n, 256), 256, 0, sim->cur, (int)n, sim->d_slot_of_id, sim->d_target_by_id);
    CK(sync_stream(sim));
    if (target) CK(cudaMemcpy(target, sim->d_target_by_id, n * sizeof(int64_t), cudaMemcpyDeviceToHost));
    if (is_follower) CK(cudaMemcpy(is_follower, sim->d_is_follower, n, cudaMemcpyDeviceToHost));
    if (index_leader) CK(cudaMemcpy(index_leader, sim->d_index_leader, n * sizeof(int64_t), cudaMemcpyDeviceToHost));
    return CDB_OK;
}

int cdb_exit_detection(cdb_sim *sim, const double *center_door, int64_t n_doors, double detection_range, int apply) {
    SIM_ENTRY();
    CKS(collective_entry(sim, apply != 0));
    if (n_doors < 0 || (n_doors > 0 && !center_door)) return fail(CDB_ERR_INVALID_VALUE, "bad door buffer");
    if (n_doors > sim->doors_cap) { CKS(dev_alloc(&sim->d_doors, (size_t)(2 * n_doors))); sim->doors_cap = n_doors; }
    if (n_doors) CK(cudaMemcpyAsync(sim->d_doors, center_door, 2 * n_doors * sizeof(double), cudaMemcpyHostToDevice, sim->stream));
    if (sim->n)
        LAUNCH(sim, k_exit_detection, cdiv(sim->n, 128), 128, 0, sim->cur, (int)sim->n, sim->d_doors, (int)n_doors, sim->d_obstacles, (int)sim->n_obstacles,
               detection_range, sim->d_detected, sim->d_has_detected, sim->d_is_follower, apply ? 1 : 0);
    CK(cudaGetLastError());
    CK(sync_stream(sim));   // center_door is the caller's (pageable) memory
    sim->detection_valid = true;
    if (apply) sim->state_version++;
    return CDB_OK;
}

int cdb_get_exit_detection(cdb_sim *sim, int64_t *detected_exit, uint8_t *has_detected, int64_t n) {
    SIM_ENTRY();
    if (!sim->detection_valid || n != sim->n) return fail(CDB_ERR_STATE, "cdb_exit_detection must run first (on the same %lld agents)", (long long)sim->n);
    CK(sync_stream(sim));
    if (n && detected_exit) CK(cudaMemcpy(detected_exit, sim->d_detected, n * sizeof(int64_t), cudaMemcpyDeviceToHost));
    if (n && has_detected) CK(cudaMemcpy(has_detected, sim->d_has_detected, n, cudaMemcpyDeviceToHost));
    return CDB_OK;
}

int cdb_nearest_neighbors(cdb_sim *sim, double sight, int64_t k, int64_t *neighbors) {
    SIM_ENTRY();
    CKS(collective_entry(sim, false));
    if (k < 1 || k > CDB_KNN_MAX) return fail(CDB_ERR_INVALID_VALUE, "size_nearest_other = %lld outside [1, %d]", (long long)k, CDB_KNN_MAX);
    if (!neighbors && sim->n) return fail(CDB_ERR_INVALID_VALUE, "neighbors is NULL");
    if (sim->n == 0) return CDB_OK;
    CKS(launch_herding(sim, sight, k, true, 0.0, 0.0, false));
    CK(sync_stream(sim));
    CKS(check_device_error(sim));
    CK(cudaMemcpy(neighbors, sim->d_knn, (size_t)sim->n * k * sizeof(int64_t), cudaMemcpyDeviceToHost));
    return CDB_OK;
}

int cdb_leader_follower(cdb_sim *sim, double sight, double phi, double weight_position_leader) {
    SIM_ENTRY();
    CKS(collective_entry(sim, true));
    if (sim->n == 0) return CDB_OK;
    return launch_leader_follower(sim, sight, phi, weight_position_leader, false, 0.0);
}

int cdb_leader_follower_with_herding(cdb_sim *sim, double sight, int64_t size_nearest_other, double phi, double weight_position_herding,
                                     double weight_position_leader, double weight_direction_leader) {
    SIM_ENTRY();
    CKS(collective_entry(sim, true));
    if (size_nearest_other < 1 || size_nearest_other > CDB_KNN_MAX)
        return fail(CDB_ERR_INVALID_VALUE, "size_nearest_other = %lld outside [1, %d]", (long long)size_nearest_other, CDB_KNN_MAX);
    if (sim->n == 0) return CDB_OK;
    CKS(launch_herding(sim, sight, size_nearest_other, false, weight_position_herding, phi, true));
    return launch_leader_follower(sim, 20.0 /* sight_leader, collective_motion.py:255 */, phi, weight_position_leader, true, weight_direction_leader);
}

int cdb_get_direction(cdb_sim *sim, double *direction, int64_t n) {
    SIM_ENTRY();
    if (!sim->direction_valid || n != sim->n) return fail(CDB_ERR_STATE, "no direction computed for these %lld agents", (long long)sim->n);
    CK(sync_stream(sim));
    if (n && direction) CK(cudaMemcpy(direction, sim->d_direction, 2 * n * sizeof(double), cudaMemcpyDeviceToHost));
    return CDB_OK;
}


// =====================================================================================================================
// host-visible state nodes
// =====================================================================================================================
constexpr int MAX_POLYGONS = 1024;
int cdb_set_polygons(cdb_sim *sim, int which, const double *xy, const int64_t *offsets, int64_t n_polygons) {
    SIM_ENTRY();
    if (which != CDB_POLY_DOMAIN && which != CDB_POLY_TARGETS) return fail(CDB_ERR_INVALID_VALUE, "which must be CDB_POLY_DOMAIN or CDB_POLY_TARGETS");
    if (n_polygons < 0 || (n_polygons > 0 && (!xy || !offsets))) return fail(CDB_ERR_INVALID_VALUE, "bad polygon buffers");
    if (which == CDB_POLY_DOMAIN && n_polygons > 1) return fail(CDB_ERR_INVALID_VALUE, "the domain is one polygon");
    if (n_polygons > MAX_POLYGONS) return fail(CDB_ERR_CAPACITY, "at most %d target polygons", MAX_POLYGONS);
    std::vector<int> off((size_t)n_polygons + 1, 0);
    for (int64_t p = 0; p <= n_polygons && n_polygons > 0; ++p) {
        if (offsets[p] < 0 || offsets[p] > 0x7fffffff || (p > 0 && offsets[p] < offsets[p - 1]) || (p == 0 && offsets[0] != 0))
            return fail(CDB_ERR_INVALID_VALUE, "polygon offsets must start at 0 and be non-decreasing");
        off[(size_t)p] = (int)offsets[p];
    }
    const int64_t nv = n_polygons > 0 ? offsets[n_polygons] : 0;
    CK(sync_stream(sim));
    CKS(dev_alloc(&sim->d_poly_xy[which], (size_t)(2 * nv)));
    CKS(dev_alloc(&sim->d_poly_off[which], (size_t)n_polygons + 1));
    if (nv) CK(cudaMemcpy(sim->d_poly_xy[which], xy, 2 * nv * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(sim->d_poly_off[which], off.data(), off.size() * sizeof(int), cudaMemcpyHostToDevice));
    sim->n_polygons[which] = n_polygons;
    sim->poly_nv[which] = nv;
    if (!sim->d_poly_counts) {
        CKS(dev_alloc(&sim->d_poly_counts, (size_t)(1 + MAX_POLYGONS)));
        CK(cudaMemset(sim->d_poly_counts, 0, (1 + MAX_POLYGONS) * sizeof(unsigned long long)));
    }
    if (which == CDB_POLY_TARGETS) {       // new targets: nobody has reached them yet
        sim->reached_n = 0;
        CK(cudaMemset(sim->d_poly_counts + 1, 0, MAX_POLYGONS * sizeof(unsigned long long)));
    }
    return CDB_OK;
}

// Strip mode: the per-agent flag arrays of these nodes are indexed by GLOBAL agent id on every rank (n = the whole crowd,
// cdb_strip_set_global_agents); a rank's entries are authoritative for the agents it owns, a migrating agent takes its flags
// along (AgentFlags).  Counts are per rank: the caller sums them over the ranks.
static int64_t flag_ids(const cdb_sim *sim) { return sim->strip ? sim->n_global : sim->n; }

int cdb_strip_set_global_agents(cdb_sim *sim, int64_t n_global) {
    SIM_ENTRY();
    if (!sim->strip) return fail(CDB_ERR_STATE, "cdb_set_strip has not been called");
    if (n_global < 0 || n_global > 0x7fffffff) return fail(CDB_ERR_INVALID_VALUE, "bad global agent count");
    CK(sync_stream(sim));
    sim->n_global = n_global;
    sim->active_n = 0;
    sim->reached_n = 0;
    return CDB_OK;
}

int cdb_set_active(cdb_sim *sim, const uint8_t *active, int64_t n) {
    SIM_ENTRY();
    if (sim->strip && sim->n_global <= 0) return fail(CDB_ERR_STATE, "strip mode: cdb_strip_set_global_agents must be called first");
    const int64_t ids = flag_ids(sim);
    if (n != ids || (n > 0 && !active)) return fail(CDB_ERR_INVALID_VALUE, "cdb_set_active: n = %lld but the flags describe %lld agents", (long long)n, (long long)ids);
    CK(sync_stream(sim));
    if (n > sim->active_n || !sim->d_active) CKS(dev_alloc(&sim->d_active, (size_t)std::max<int64_t>(n, 1024)));
    if (n) CK(cudaMemcpy(sim->d_active, active, n, cudaMemcpyHostToDevice));
    sim->active_n = n;
    return CDB_OK;
}

int cdb_get_active(cdb_sim *sim, uint8_t *active, int64_t n) {
    SIM_ENTRY();
    if (!sim->d_active || n != sim->active_n || n != flag_ids(sim)) return fail(CDB_ERR_STATE, "cdb_set_active must describe the %lld agents first", (long long)flag_ids(sim));
    CK(sync_stream(sim));
    if (n && active) CK(cudaMemcpy(active, sim->d_active, n, cudaMemcpyDeviceToHost));
    return CDB_OK;
}

int cdb_inside_domain(cdb_sim *sim, int64_t *n_changed) {
    SIM_ENTRY();
    if (sim->n_polygons[CDB_POLY_DOMAIN] != 1) return fail(CDB_ERR_STATE, "cdb_set_polygons(CDB_POLY_DOMAIN) must be called first");
    if (!sim->d_active || sim->active_n != flag_ids(sim) || sim->active_n <= 0 && sim->n > 0)
        return fail(CDB_ERR_STATE, "cdb_set_active must describe the %lld agents first", (long long)flag_ids(sim));
    CK(cudaMemsetAsync(sim->d_poly_counts, 0, sizeof(unsigned long long), sim->stream));
    if (sim->n)
        LAUNCH(sim, k_inside_domain, cdiv(sim->n, 128), 128, 0, sim->cur, (int)sim->n, sim->dev_counts ? &sim->d_counts->slots : nullptr,
               sim->d_poly_xy[CDB_POLY_DOMAIN], (int)sim->poly_nv[CDB_POLY_DOMAIN], sim->d_active, sim->d_poly_counts);
    CK(cudaGetLastError());
    if (n_changed) {
        unsigned long long c = 0;
        CK(cudaMemcpyAsync(&c, sim->d_poly_counts, sizeof(c), cudaMemcpyDeviceToHost, sim->stream));
        CK(sync_stream(sim));
        *n_changed = (int64_t)c;
    }
    return CDB_OK;
}

int cdb_target_reached(cdb_sim *sim, int64_t *counts, int64_t n_polygons) {
    SIM_ENTRY();
    const int64_t np = sim->n_polygons[CDB_POLY_TARGETS];
    if (n_polygons != np) return fail(CDB_ERR_INVALID_VALUE, "n_polygons = %lld but %lld target polygons are set", (long long)n_polygons, (long long)np);
    if (sim->strip && sim->n_global <= 0) return fail(CDB_ERR_STATE, "strip mode: cdb_strip_set_global_agents must be called first");
    if (sim->strip && np > FLAG_POLYGONS) return fail(CDB_ERR_CAPACITY, "strip mode: at most %d target polygons (their flags travel with migrating agents)", FLAG_POLYGONS);
    const int64_t ids = flag_ids(sim);
    if (np > 0 && ids > 0) {
        if (sim->reached_n != ids || np * ids > sim->reached_cap) {     // first call for this crowd: nobody has arrived yet
            if (np * ids > sim->reached_cap) { CKS(dev_alloc(&sim->d_reached, (size_t)(np * ids))); sim->reached_cap = np * ids; }
            CK(cudaMemsetAsync(sim->d_reached, 0, (size_t)(np * ids), sim->stream));
            CK(cudaMemsetAsync(sim->d_poly_counts + 1, 0, np * sizeof(unsigned long long), sim->stream));
            sim->reached_n = ids;
        }
        if (sim->n > 0)
            LAUNCH(sim, k_target_reached, cdiv(sim->n, 128), 128, 0, sim->cur, (int)sim->n, sim->dev_counts ? &sim->d_counts->slots : nullptr,
                   sim->d_poly_xy[CDB_POLY_TARGETS], sim->d_poly_off[CDB_POLY_TARGETS], (int)np, sim->d_reached, (long long)ids, sim->d_poly_counts + 1);
        CK(cudaGetLastError());
    }
    if (counts && np > 0) {
        std::vector<unsigned long long> c((size_t)np, 0ULL);
        if (ids > 0) {
            CK(cudaMemcpyAsync(c.data(), sim->d_poly_counts + 1, np * sizeof(unsigned long long), cudaMemcpyDeviceToHost, sim->stream));
            CK(sync_stream(sim));
        }
        for (int64_t p = 0; p < np; ++p) counts[p] = (int64_t)c[(size_t)p];
    }
    return CDB_OK;
}

int cdb_get_target_reached(cdb_sim *sim, uint8_t *reached_by, int64_t n_polygons, int64_t n) {
    SIM_ENTRY();
    if (n_polygons != sim->n_polygons[CDB_POLY_TARGETS] || n != flag_ids(sim) || sim->reached_n != n)
        return fail(CDB_ERR_STATE, "cdb_target_reached must run first (on the same agents and polygons)");
    CK(sync_stream(sim));
    if (n_polygons * n > 0 && reached_by) CK(cudaMemcpy(reached_by, sim->d_reached, (size_t)(n_polygons * n), cudaMemcpyDeviceToHost));
    return CDB_OK;
}
