/* crowd_b200.h -- C ABI of the B200-native crowddynamics agent update (libcrowd_b200.so).
 *
 * This is the drop-in boundary for the reference's per-timestep hot path.  Every entry point replaces one
 * reference interface (cited as file:line under /root/reference/crowddynamics/); the Python logic nodes in
 * crowddynamics_b200/logic.py bind them with ctypes exactly as INTEGRATION.md shows for the reference's own
 * simulation/logic.py.  Plain pointers and sizes only: no torch, numpy or C++ types cross this boundary.
 *
 * Conventions
 *   - every function returns a cdb_status (0 = ok); cdb_last_error() gives the message of the last failure on the
 *     calling thread.  The Python side maps CDB_ERR_INVALID_TYPE -> InvalidType, everything else ->
 *     CrowdDynamicsException subclasses (reference exceptions.py:10-22; interactions.py:204-205,213-214).
 *   - agent records are the reference's packed structured dtypes: itemsize 228 (agent_type_circular) or 316
 *     (agent_type_three_circle), simulation/agents.py:447-457.  Any other itemsize is CDB_ERR_INVALID_TYPE.
 *   - all arithmetic is IEEE fp64 on the GPU; there is no CPU fallback anywhere behind this header.
 *   - a cdb_sim is bound to one CUDA device and used from one host thread at a time (the reference is single
 *     threaded, simulation/multiagent.py:51-55); the CUDA context is created lazily by cdb_create in the calling
 *     process, so a forked MultiAgentProcess (multiagent.py:58-100) must create its own sim.
 */
#ifndef CROWD_B200_H
#define CROWD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cdb_sim cdb_sim;

typedef enum {
    CDB_OK = 0,
    CDB_ERR_INVALID_TYPE = 1,  /* unknown agent dtype / itemsize                 -> InvalidType */
    CDB_ERR_INVALID_VALUE = 2, /* bad argument, non-finite positions, ...        -> InvalidValue */
    CDB_ERR_CUDA = 3,          /* CUDA runtime error                             -> DeviceError */
    CDB_ERR_CAPACITY = 4,      /* more agents / cells / halo entries than allocated */
    CDB_ERR_STATE = 5          /* call order (e.g. tables requested before a block list was built) */
} cdb_status;

enum { CDB_MODEL_CIRCULAR = 0, CDB_MODEL_THREE_CIRCLE = 1 };

/* Field mask bits for cdb_download_agents_aos: which record fields are written back to the host array. */
enum {
    CDB_F_POSITION = 1u << 0,
    CDB_F_VELOCITY = 1u << 1,
    CDB_F_TARGET_DIRECTION = 1u << 2,
    CDB_F_FORCE = 1u << 3,
    CDB_F_FORCE_PREV = 1u << 4,
    CDB_F_SHOULDERS = 1u << 5,          /* position_ls, position_rs */
    CDB_F_ORIENTATION = 1u << 6,
    CDB_F_ANGULAR_VELOCITY = 1u << 7,
    CDB_F_TARGET_ORIENTATION = 1u << 8,
    CDB_F_TORQUE = 1u << 9,
    CDB_F_TORQUE_PREV = 1u << 10,
    CDB_F_ALL_MUTABLE = (1u << 11) - 1,
    CDB_F_WHOLE_RECORD = 1u << 31       /* copy the full records (valid when the host did not touch the array since upload) */
};

/* Node selection bits for cdb_step -- one bit per replaced LogicNode, executed in the reference's post-order
 * (examples/simulations.py:123-136): navigation, orientation, adjusting, agent-agent, agent-obstacle, integrator, reset. */
enum {
    CDB_STEP_NAVIGATION = 1u << 0,
    CDB_STEP_ORIENTATION = 1u << 1,
    CDB_STEP_ADJUSTING = 1u << 2,
    CDB_STEP_AGENT_AGENT = 1u << 3,
    CDB_STEP_AGENT_OBSTACLE = 1u << 4,
    CDB_STEP_INTEGRATOR = 1u << 5,
    CDB_STEP_RESET = 1u << 6,
    CDB_STEP_ALL = (1u << 7) - 1,        /* the seven deterministic nodes */
    CDB_STEP_FLUCTUATION = 1u << 7       /* + Fluctuation (stochastic: distribution parity only), runs first like in the reference tree */
};

/* ---- library ---------------------------------------------------------------------------------------------------- */
const char *cdb_last_error(void);
int cdb_version(void);
int cdb_device_count(int *count);
/* DFMA throughput of the device in TFLOP/s (8 independent chains per thread): the fp64 roofline denominator. */
int cdb_measure_fp64_peak(int device, double *tflops);

/* ---- lifetime ---------------------------------------------------------------------------------------------------
 * Replaces: nothing in the reference (its state is the host array simulation.agents.array, agents.py:605-680);
 * the sim object is the device-resident mirror of that array in SoA layout. */
int cdb_create(int device, int model, int64_t capacity, cdb_sim **out);
int cdb_destroy(cdb_sim *sim);
/* Run all work of this sim on an existing CUDA stream (cudaStream_t, e.g. torch.cuda.current_stream().cuda_stream). */
int cdb_set_stream(cdb_sim *sim, void *cuda_stream);
int cdb_synchronize(cdb_sim *sim);
int64_t cdb_num_agents(const cdb_sim *sim);

/* ---- host <-> device of the reference data contract ---------------------------------------------------------------
 * agents: packed records exactly as simulation.agents.array (C-contiguous, asserted at agents.py:680).
 * upload converts AoS -> device SoA; download writes the fields selected by field_mask back into the host records
 * (in the original agent order, whatever the device-side cell ordering is). */
int cdb_upload_agents_aos(cdb_sim *sim, const void *agents, int64_t n, int64_t itemsize);
int cdb_download_agents_aos(cdb_sim *sim, void *agents, int64_t n, int64_t itemsize, uint32_t field_mask);
/* Field-granular traffic for the strict mode of the nodes ("upload dirty fields -> kernel -> download written fields"):
 *   cdb_host_register      pins + maps the host array (simulation.agents.array is ordinary pageable numpy memory) so that
 *                          kernels can read / write the selected fields of the packed records in place, over PCIe;
 *   cdb_upload_agents_fields  refreshes only the fields in field_mask of the n agents uploaded before (same array order),
 *                          wherever the fused steps have moved them on the device since;
 *   cdb_download_agents_aos   without CDB_F_WHOLE_RECORD writes only the fields in field_mask into the caller's records.
 * Masks of at most 120 B per agent travel by zero-copy kernels (time proportional to the bytes selected), larger ones by one
 * DMA of the whole records (55 GB/s); pageable, unregistered memory falls back to a pinned bounce buffer + host merge.
 * cdb_transfer_stats reports the bytes these entry points moved over PCIe. */
int cdb_host_register(cdb_sim *sim, void *agents, int64_t n, int64_t itemsize);
int cdb_host_unregister(cdb_sim *sim, void *agents);
int cdb_upload_agents_fields(cdb_sim *sim, const void *agents, int64_t n, int64_t itemsize, uint32_t field_mask);
int cdb_transfer_stats(cdb_sim *sim, int64_t *h2d_bytes, int64_t *d2h_bytes, int reset);

/* obstacles: (W, 4) doubles p0x,p0y,p1x,p1y == obstacle_type_linear (core/structures.py:6-9), the output format of
 * geom_to_linear_obstacles (core/geometry.py:99-102) that AgentObstacleInteractions.update builds (logic.py:122-130). */
int cdb_set_obstacles(cdb_sim *sim, const double *segments, int64_t n_segments);

/* Navigation field of one target: U, V of shape (ny, nx) indexed [iy, ix], origin (minx, miny), grid step -- the
 * (mgrid, direction_map) pair returned by Field.navigation_to_target (simulation/field.py:155-164). */
int cdb_set_navigation_field(cdb_sim *sim, int64_t target, const double *U, const double *V, int64_t ny, int64_t nx,
                             double minx, double miny, double step);
int cdb_clear_navigation(cdb_sim *sim);
/* Builds the navigation field of one target ON THE DEVICE and installs it like cdb_set_navigation_field -- what
 * Field.navigation_to_target (simulation/field.py:155-164) computes on the host with skfmm / shapely / skimage:
 * shortest_path (core/steering/quickest_path.py:184-197: eikonal distance to the target around the obstacles buffered by
 * `radius`, distance_map :54-117; normalised gradient, direction_map :144-163; fill_missing :168-181), the distance /
 * direction from the walls (core/steering/obstacle_handling.py:106+) and obstacle_handling (:15-74, `strength`).
 * Geometry: target and obstacle line segments (p0x, p0y, p1x, p1y), rasterised like draw_geom does for LineStrings
 * (core/geometry.py:105-116); grid of (ny, nx) points from (minx, miny) with spacing `step` (quickest_path.meshgrid :21-48).
 * Optional host outputs (may be NULL): the signed target distance map (NaN inside the buffered obstacles) and (U, V), each
 * (ny, nx) doubles; rounds_out = relaxation rounds of the eikonal solver.  Nothing of size (ny, nx) is needed on the host. */
int cdb_build_navigation_field(cdb_sim *sim, int64_t target, const double *target_segments, int64_t n_target_segments,
                               const double *obstacle_segments, int64_t n_obstacle_segments, int64_t ny, int64_t nx, double minx,
                               double miny, double step, double radius, double strength, double *distance_map_out, double *U_out,
                               double *V_out, int64_t *rounds_out);

/* ---- per-node entry points (one per replaced LogicNode.update) ------------------------------------------------------ */
int cdb_reset(cdb_sim *sim);                 /* Reset.update, logic.py:59-64 */
/* Fluctuation.update, logic.py:78-86 -> core/motion/fluctuation.py:14-62.  Counter-based RNG (Philox, key = seed and the
 * sim's iteration counter, counter = agent id); the reference uses numpy's unseeded global RNG, so parity is statistical. */
int cdb_set_seed(cdb_sim *sim, uint64_t seed);
int cdb_fluctuation(cdb_sim *sim);
int cdb_navigation(cdb_sim *sim);            /* Navigation.update sampling, logic.py:149-165 + navigation.py:60-78 + quickest_path.py:41-44 */
int cdb_orientation(cdb_sim *sim);           /* Orientation.update, logic.py:258-261 + steering/orientation.py:17-21 */
int cdb_adjust(cdb_sim *sim);                /* Adjusting.update, logic.py:89-94 + motion/adjusting.py:101-121 */
int cdb_agent_agent(cdb_sim *sim, double cell_size);   /* AgentAgentInteractions.update, logic.py:118-119 -> core/interactions.py:191-205 */
int cdb_agent_obstacle(cdb_sim *sim);        /* AgentObstacleInteractions.update, logic.py:122-130 -> core/interactions.py:208-214 */
int cdb_integrate(cdb_sim *sim, double dt_min, double dt_max, double *dt_out);   /* Integrator.update, logic.py:71-75 -> core/integrator.py:209-256 */

/* Fused resident step: n_steps iterations of the selected nodes without leaving the device (the body of
 * MultiAgentSimulation.update, simulation/multiagent.py:51-55, restricted to the replaced sub-tree).
 * dt_out (may be NULL) receives the n_steps time steps used; time_tot / iterations accumulate like simulation.data. */
int cdb_step(cdb_sim *sim, uint32_t node_flags, double cell_size, double dt_min, double dt_max, int64_t n_steps,
             double *dt_out);
int cdb_get_time(cdb_sim *sim, double *time_tot, int64_t *iterations);
/* cdb_step replays pairs of steps as one CUDA graph when nothing it captured changes (default on; not while profiling and
 * not in strip mode).  Results are identical either way; this only removes launch latency for small crowds. */
int cdb_set_graphs(cdb_sim *sim, int enable);
/* Crowds of at most max_agents agents (default and upper limit 256) are advanced by cdb_step with ONE thread block that keeps
 * the crowd in shared memory and runs all the requested steps in one launch (the reference's example simulations are of this
 * size: Hallway, 50 agents, examples/simulations.py:74-163).  All pairs are tested instead of a block list, which adds the
 * same terms whenever 3 + 2 max R < cell_size (otherwise, and with max_agents = 0, the general path is used).  Contributions
 * are added in ascending slot order: results equal the general path's up to summation order. */
int cdb_set_small_crowd_max(cdb_sim *sim, int64_t max_agents);

/* ---- asynchronous host-visible state (SaveSimulationData, logic.py:266-337 + io.py:19-45; the per-update scalars of
 * Integrator / InsideDomain / TargetReached, logic.py:71-75,343-387) -- a fully resident tree never waits for the device:
 *   cdb_snapshot_begin   queues, behind the work already issued, a copy of the WHOLE packed records (as
 *                        cdb_download_agents_aos + cdb_get_states + cdb_get_active would return them) into one of two pinned
 *                        host buffers, on a side stream; returns at once with the slot;
 *   cdb_snapshot_wait    pointer to that slot's records; blocks only if its copy has not finished (one update later it has);
 *   cdb_scalars_begin / cdb_scalars_wait   the same for {dt, time_tot, InsideDomain change count, TargetReached counts};
 *   cdb_set_deferred_sync  lets cdb_step return without waiting for its own pair-list check while the list is at most a
 *                        quarter full (see cdb_set_pair_capacity); the check is read when a later call finds it complete;
 *   cdb_sync_count       blocking host synchronisations the library has performed so far (the tests count them). */
int cdb_snapshot_begin(cdb_sim *sim, int64_t *slot_out);
int cdb_snapshot_wait(cdb_sim *sim, int64_t slot, const void **records, int64_t *n);
int cdb_scalars_begin(cdb_sim *sim, int64_t *slot_out);
int cdb_scalars_wait(cdb_sim *sim, int64_t slot, double *dt, double *time_tot, int64_t *inside_changes, int64_t *target_counts,
                     int64_t n_targets);
int cdb_set_deferred_sync(cdb_sim *sim, int enable);
int64_t cdb_sync_count(const cdb_sim *sim);

/* ---- resident-order steps (no reference counterpart; the reference re-bins every agent at every update,
 * core/interactions.py:191-205).  For crowds of at least `min_agents` agents cdb_step rebuilds its block list only every few steps:
 * the search cells are (1 + skin_fraction) times wider than the interaction range needs, the agents keep their slots in between,
 * the step then works in place (constant fields are not rewritten) and writes the next step's neighbour records itself.  A
 * device-side bound on how far any agent has drifted since the last rebuild guards every sweep; a step that finds the bound
 * exceeded is not applied, and the host rebuilds and repeats it (same protocol as a pair list that is too small).  The interval
 * adapts to the observed displacement, at most `max_interval` steps; max_interval = 1 rebuilds at every step.  Forces are the
 * same numbers up to the order in which an agent's pair contributions are added (<= 1e-15 relative).
 * Defaults: skin_fraction 0.10, max_interval 16, min_agents 16384.  In strip mode only on request (cdb_strip_set_kind below); not used with a fixed lattice, or when
 * 3 + 2 max R >= cell_size (then the pair set depends on the lattice itself). */
int cdb_set_rebuild_policy(cdb_sim *sim, double skin_fraction, int64_t max_interval, int64_t min_agents);
/* steps that rebuilt the block list / ran on the kept order / were refused as stale so far, and the current interval */
int cdb_get_rebuild_stats(cdb_sim *sim, int64_t *rebuilds, int64_t *kept, int64_t *stale, int64_t *interval);

/* ---- instrumentation (no reference counterpart) -------------------------------------------------------------------- */
/* agent-agent kernel variant:
 *   3 (default) = every unordered pair of the block list classified and evaluated ONCE, as the reference's pair loop does
 *       (core/interactions.py:69-70,100-104 update both agents from one evaluation): forward-half-stencil sweep -> pair
 *       list -> one evaluation per pair -> per-agent sums in a fixed order;
 *   2 = two-phase fused step kernel that evaluates every pair from both agents' sides (bit-identical results to 3);
 *   1 = one-phase kernels.  2 and 1 are kept as independent cross-checks. */
int cdb_set_variant(cdb_sim *sim, int variant);
/* Variant 3 keeps the listed pairs of one step in a buffer that grows on demand: a step that finds more pairs than fit is
 * not applied on the device and is transparently repeated after growing the buffer.  cdb_set_pair_capacity fixes the
 * capacity (0 = automatic, 8 pairs per agent to start with) -- a test hook for the repeat path; cdb_get_pair_stats reports
 * the capacity, the pairs listed by the most recent step and how many steps had to be repeated so far. */
int cdb_set_pair_capacity(cdb_sim *sim, int64_t pairs);
/* The pair search bins on cell_size / 2 with a reach of two cells whenever no pair can interact beyond cell_size
 * (3 + 2 max R < cell_size): the same pairs as the reference's block list with 31 % less area swept.  0 = automatic (default:
 * refined for circular agents, not for three-circle agents), 1 = always search on the cell_size lattice itself, 2 = refined
 * wherever it is valid; the block-list exports below report the cell_size lattice either way. */
int cdb_set_search_refinement(cdb_sim *sim, int refinement);
/* bound on the radius (circular) / body extent (three-circle) of the uploaded agents: refinement is valid when
 * 3 + 2 * ext_max < cell_size.  In strip mode the CALLER decides (cdb_set_search_refinement(2) before cdb_set_strip, using
 * the maximum over all ranks), because every rank must bin on the same lattice. */
int cdb_get_ext_max(cdb_sim *sim, double *ext_max);
int cdb_get_pair_stats(cdb_sim *sim, int64_t *capacity, int64_t *found_last, int64_t *overflows);
int64_t cdb_launch_count(const cdb_sim *sim);            /* kernels launched by this sim so far */
int cdb_profile_enable(cdb_sim *sim, int enable);        /* CUDA-event timing of the phases of cdb_step on the sim's stream */
/* ms[0] = per-agent nodes before + block list build, ms[1] = agent-agent kernel, ms[2] = obstacle + integrator + reset;
 * summed over the profiled steps since the last read (at most 4096 steps are recorded). */
int cdb_profile_read(cdb_sim *sim, double ms[3], int64_t *steps);
/* finer split of the same events: ms[0] = per-agent nodes before + block list, ms[1] = pair sweep (+ region allocation),
 * ms[2] = pair evaluation, ms[3] = step kernel (gather of the pair results, walls, integrator, reset), ms[4] = rest.
 * Variants 1 / 2 report their whole agent-agent (+ fused step) kernel in ms[3]. */
int cdb_profile_read_phases(cdb_sim *sim, double ms[5], int64_t *steps);

/* ---- block list: debug / parity exports (cell_lists.add_to_cells & iter_nearest_neighbors, call sites
 * core/interactions.py:191-205; spec core/block_list.py:28-52) ---------------------------------------------------- */
int cdb_build_block_list(cdb_sim *sim, double cell_size);
int cdb_get_grid(cdb_sim *sim, int64_t grid[4]);                       /* ix_min, iy_min, nx, ny */
int cdb_get_cell_ids(cdb_sim *sim, int64_t *cell_of_agent, int64_t n); /* flat cell id per agent, original order */
int cdb_get_cell_tables(cdb_sim *sim, int64_t *points_indices, int64_t n, int64_t *cells_count,
                        int64_t *cells_offset, int64_t n_cells);
/* Candidate pairs (same or adjacent cell), each unordered pair once as ordered (i, j) with i the agent that is
 * lexicographically smaller in (cell_x, cell_y, agent_index); pairs[2*k], pairs[2*k+1]; *count may exceed cap. */
int cdb_get_neighbor_pairs(cdb_sim *sim, int64_t *pairs, int64_t cap, int64_t *count);

/* ---- fixed lattice + strip decomposition (multi-GPU; no reference counterpart) ------------- */
/* Fix the cell lattice instead of deriving it from the bounding box every step (cells stay anchored at multiples of
 * cell_size, so cell coordinates are identical; agents outside are binned into the border cells). */
int cdb_set_lattice(cdb_sim *sim, int64_t ix_min, int64_t iy_min, int64_t nx, int64_t ny);
int cdb_clear_lattice(cdb_sim *sim);
/* Strip decomposition along x, aligned to cell columns.  This sim owns the nx_owned cell columns starting at global
 * column ix_min (rows iy_min .. iy_min + ny - 1 -- every rank uses the same rows); one ghost column is added on each side
 * that has a neighbour.  halo_cap / migrant_cap: record capacity of one halo / migrant message. */
int cdb_set_strip(cdb_sim *sim, int64_t ix_min, int64_t iy_min, int64_t nx_owned, int64_t ny, int has_left, int has_right,
                  int64_t halo_cap, int64_t migrant_cap);
/* Global agent indices of the uploaded agents (used for pair orientation and to identify agents after migration). */
int cdb_set_agent_ids(cdb_sim *sim, const int64_t *ids, int64_t n);
/* Message sizes in doubles.  Messages are plain device memory owned by the caller (e.g. torch tensors handed to NCCL
 * send/recv); they start with a 4-double header whose first entry is the record count. */
int64_t cdb_halo_buffer_doubles(const cdb_sim *sim);
int64_t cdb_migrant_buffer_doubles(const cdb_sim *sim);
/* One strip step is split around the two exchanges:
 *   begin   block list of the owned agents; packs the first / last owned cell column (neighbour records + per-cell
 *           counts) into halo_left_out / halo_right_out (NULL where there is no neighbour)
 *   -- caller exchanges halos, and MAX-reduces the FOUR doubles of export_vmax {max |v|, max v0, NaN flag, NaN flag}
 *      across ranks when dt_min != dt_max (NaN travels as a flag: a MAX collective need not propagate it) --
 *   finish  installs the received ghost columns, runs the fused step kernel on the owned agents, then packs the agents
 *           that left the strip into mig_left_out / mig_right_out
 *   -- caller exchanges migrants --
 *   absorb  appends the received migrants.  The exact agent counts stay on the device: the host keeps upper bounds for its
 *           launch sizes and synchronises only every 16th step -- or when n_out is non-NULL (*n_out = agents now owned). */
int cdb_strip_begin(cdb_sim *sim, uint32_t node_flags, double cell_size, double *halo_left_out, double *halo_right_out);
/* Kept block lists in strip mode (resident-order steps, see cdb_set_rebuild_policy; the policy has to be set BEFORE
 * cdb_set_strip, whose columns are then columns of the widened cells cell_size * (1 + skin_fraction)).  The caller -- every
 * rank alike -- announces what the step it is about to issue is:
 *   0  rebuilds the block list and hands its leavers over at its end (every step rebuilds: the default);
 *   1  rebuilds; the next step keeps the order: no migrant messages, no absorb;
 *   2  keeps the order of the last rebuild: begin only packs the halo records (same slice of slots, same per-cell counts),
 *      finish works in place; no migrant messages, no absorb;
 *   3  like 2, and the next step rebuilds: finish fills the migrant messages, the caller exchanges them and calls absorb;
 *   4  like 1, but the next step rebuilds as well (an interval of one step, with the drift bookkeeping of kept lists running).
 * An agent that drifts across the strip border between two rebuilds stays with its rank until then; the ghost column still
 * shows it every partner within reach because it is as wide as the widened cells.  All ranks must use the same sequence of
 * kinds; a step that finds its block list stale raises a device error (a strip cannot repeat a step on its own).
 * cdb_strip_drift (one synchronisation): largest displacement of the last step, drift bound since the last rebuild and its
 * limit -- what the caller sizes the common rebuild interval with. */
int cdb_strip_set_kind(cdb_sim *sim, int kind);
/* InsideDomain / TargetReached in strip mode (logic.py:343-387): their per-agent flags (`active`, `reached_by`) live in arrays
 * indexed by GLOBAL agent id on every rank -- n of cdb_set_active / cdb_get_active / cdb_get_target_reached is then the size of
 * the whole crowd, announced here.  A rank's entries are authoritative for the agents it currently owns (cdb_export_agents
 * lists them); a migrating agent takes its flags along inside the migrant message.  cdb_inside_domain / cdb_target_reached
 * return this rank's counts: the caller adds them up over the ranks.  At most 20 target polygons in strip mode. */
int cdb_strip_set_global_agents(cdb_sim *sim, int64_t n_global);
int cdb_strip_drift(cdb_sim *sim, double *disp_last, double *disp_acc, double *drift_limit);
int cdb_strip_export_vmax(cdb_sim *sim, double *dev_vmax4);
int cdb_strip_import_vmax(cdb_sim *sim, const double *dev_vmax4);
int cdb_strip_finish(cdb_sim *sim, uint32_t node_flags, double dt_min, double dt_max, const double *halo_left_in,
                     const double *halo_right_in, double *mig_left_out, double *mig_right_out);
int cdb_strip_absorb(cdb_sim *sim, const double *mig_left_in, const double *mig_right_in, int64_t *n_out);
int cdb_strip_count(cdb_sim *sim, int64_t *n_out);   /* agents currently owned (one host sync) */
/* One-sided exchange over NVLink peer memory instead of NCCL send / recv: every sim owns its receive buffers and a flag
 * array (cdb_strip_exchange_alloc), the neighbours map them (CUDA IPC handles between processes -- one process per GPU --
 * or directly inside one process) and the *_direct variants of begin / finish / absorb make the producer's kernels write
 * the halo / migrant messages straight into the consumer's buffers and publish a sequence number there, on which the
 * consumer's kernels spin: no collective kernel, no host-side handshake per step. */
int cdb_strip_exchange_alloc(cdb_sim *sim);
int64_t cdb_strip_exchange_handle_bytes(void);
int cdb_strip_exchange_handles(cdb_sim *sim, void *handles_out);
int cdb_strip_exchange_connect_ipc(cdb_sim *sim, const void *left_handles, const void *right_handles);
int cdb_strip_exchange_connect_local(cdb_sim *sim, cdb_sim *left, cdb_sim *right);
int cdb_strip_begin_direct(cdb_sim *sim, uint32_t node_flags, double cell_size, int send_halo);
int cdb_strip_finish_direct(cdb_sim *sim, uint32_t node_flags, double dt_min, double dt_max, int recv_halo);
int cdb_strip_absorb_direct(cdb_sim *sim, int64_t *n_out);
/* Live agents in device order: packed records rebuilt from the device state (record fields the kernels never touch are
 * zero) and their global ids. */
int cdb_export_agents(cdb_sim *sim, void *agents, int64_t *ids, int64_t cap, int64_t *count);

/* ---- collective motion: exit detection, herding, leader-follower ---------------------------
 * These nodes read / write the States fields of the agent records (simulation/agents.py:33-60).  `target` lives on the
 * device per agent (uploaded with the records); the others are set here as plain arrays indexed like the host array.
 * Not available in strip mode (CDB_ERR_STATE).  Any pointer of cdb_set_states / cdb_get_states may be NULL (= skip). */
#define CDB_KNN_MAX 32
/* replaces: reading agents['target' | 'is_leader' | 'is_follower' | 'index_leader' | 'familiar_exit'] */
int cdb_set_states(cdb_sim *sim, const int64_t *target, const uint8_t *is_leader, const uint8_t *is_follower,
                   const int64_t *index_leader, const int64_t *familiar_exit, int64_t n);
/* replaces: the in-place writes to agents['target' | 'is_follower' | 'index_leader'] */
int cdb_get_states(cdb_sim *sim, int64_t *target, uint8_t *is_follower, int64_t *index_leader, int64_t n);
/* ExitDetection.update (simulation/logic.py:237-256) -> exit_detection (core/evacuation.py:137-174): closest door centre in
 * range with a free line of sight (is_obstacle_between_points, core/sensory_region.py:9-16; line_intersect,
 * core/geom2D.py:38-59).  apply != 0 also does logic.py:253-255 (followers that detected an exit get it as target and stop
 * being followers).  cdb_get_exit_detection returns the two arrays exit_detection returns. */
int cdb_exit_detection(cdb_sim *sim, const double *center_door, int64_t n_doors, double detection_range, int apply);
int cdb_get_exit_detection(cdb_sim *sim, int64_t *detected_exit, uint8_t *has_detected, int64_t n);
/* find_nearest_neighbors (core/steering/collective_motion.py:69-110) over the block list with cell_size = sight
 * (:262-267): neighbors[n][k], -1 = missing, rows in the reference's own slot order.  1 <= k <= CDB_KNN_MAX. */
int cdb_nearest_neighbors(cdb_sim *sim, double sight, int64_t k, int64_t *neighbors);
/* LeaderFollower.update (logic.py:168-182) -> leader_follower_interaction (collective_motion.py:229-243): updates target /
 * index_leader of the followers and sets target_direction[is_follower] = direction[is_follower]. */
int cdb_leader_follower(cdb_sim *sim, double sight, double phi, double weight_position_leader);
/* LeaderFollowerWithHerding.update (logic.py:185-221) -> leader_follower_with_herding_interaction
 * (collective_motion.py:246-289; leaders are seen up to 20 m, :255). */
int cdb_leader_follower_with_herding(cdb_sim *sim, double sight, int64_t size_nearest_other, double phi,
                                     double weight_position_herding, double weight_position_leader,
                                     double weight_direction_leader);
/* the direction array (n x 2, all agents) the last of the two calls above computed = the reference functions' return value */
int cdb_get_direction(cdb_sim *sim, double *direction, int64_t n);

/* ---- host-visible state nodes -----------------------------------------------------------------
 * InsideDomain / TargetReached (simulation/logic.py:343-387) = matplotlib Path(vertices).contains_points(position).
 * Polygons: (x, y) vertex pairs back to back, offsets[n_polygons + 1] in vertices, implicitly closed (do not repeat the first
 * vertex).  which = CDB_POLY_DOMAIN takes exactly one polygon (field.domain.exterior), CDB_POLY_TARGETS one per measured
 * target; setting the target polygons clears their reached_by state.  In strip mode: see cdb_strip_set_global_agents. */
#define CDB_POLY_DOMAIN 0
#define CDB_POLY_TARGETS 1
int cdb_set_polygons(cdb_sim *sim, int which, const double *xy, const int64_t *offsets, int64_t n_polygons);
/* replaces: agents['active'] (simulation/agents.py:33-35) */
int cdb_set_active(cdb_sim *sim, const uint8_t *active, int64_t n);
int cdb_get_active(cdb_sim *sim, uint8_t *active, int64_t n);
/* InsideDomain.update (logic.py:351-357): active = inside(domain); *n_changed (may be NULL: no host sync) = np.sum(change) */
int cdb_inside_domain(cdb_sim *sim, int64_t *n_changed);
/* TargetReached.update (logic.py:383-387): reached_by |= inside(target p); counts[p] (may be NULL) = np.sum(reached_by) */
int cdb_target_reached(cdb_sim *sim, int64_t *counts, int64_t n_polygons);
int cdb_get_target_reached(cdb_sim *sim, uint8_t *reached_by, int64_t n_polygons, int64_t n);   /* [n_polygons][n] */

#ifdef __cplusplus
}
#endif
#endif /* CROWD_B200_H */
