"""Device-resident mirror of ``simulation.agents.array`` (host side of the C ABI).

``DeviceAgents`` owns one ``cdb_sim`` handle.  Two operating modes:

* strict   -- every node: upload the host array, run the kernel(s), download the fields the node writes.  Used by the
              functional API (``crowddynamics_b200.core.*``) and by logic nodes living in a tree with host-side nodes.
* resident -- state stays on the device between nodes / steps; the host array is refreshed on demand (``sync_host``).
"""
import ctypes as C

import numpy as np

from . import _lib
from .structures import model_of, as_obstacles, MODEL_CIRCULAR
from .exceptions import InvalidType, InvalidValue


def _ptr(a):
    return C.c_void_p(a.ctypes.data)


def _check_agents(agents):
    model = model_of(agents)
    if agents.ndim != 1 or not agents.flags.c_contiguous:
        raise InvalidValue('agents must be a 1-D C-contiguous structured array (reference agents.py:680)')
    return model


class DeviceAgents:
    def __init__(self, model, capacity=0, device=0, stream=None):
        self.lib = _lib.load()
        self.model = int(model)
        self.device = int(device)
        self.itemsize = 228 if self.model == MODEL_CIRCULAR else 316
        h = C.c_void_p()
        _lib.check(self.lib.cdb_create(self.device, self.model, int(capacity), C.byref(h)))
        self.handle = h
        self.n = 0
        if stream is not None:
            self.set_stream(stream)

    # -- lifetime -------------------------------------------------------------------------------------------------
    def close(self):
        if getattr(self, 'handle', None):
            self.lib.cdb_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, stream):
        """stream: raw cudaStream_t as int (``torch.cuda.current_stream().cuda_stream``)."""
        _lib.check(self.lib.cdb_set_stream(self.handle, C.c_void_p(int(stream))))

    def synchronize(self):
        _lib.check(self.lib.cdb_synchronize(self.handle))

    # -- data ------------------------------------------------------------------------------------------------------
    def upload(self, agents):
        if _check_agents(agents) != self.model:
            raise InvalidType('agents dtype does not match the device model')
        _lib.check(self.lib.cdb_upload_agents_aos(self.handle, _ptr(agents), len(agents), agents.dtype.itemsize))
        self.n = len(agents)

    def upload_raw(self, ptr, n):
        """Upload from a raw host pointer (e.g. a pinned torch tensor holding packed records)."""
        _lib.check(self.lib.cdb_upload_agents_aos(self.handle, C.c_void_p(int(ptr)), int(n), self.itemsize))
        self.n = int(n)

    def upload_fields(self, agents, mask):
        """Refresh only the fields in ``mask`` of the agents uploaded before (same array, same order)."""
        if _check_agents(agents) != self.model:
            raise InvalidType('agents dtype does not match the device model')
        _lib.check(self.lib.cdb_upload_agents_fields(self.handle, _ptr(agents), len(agents), agents.dtype.itemsize, int(mask)))

    def upload_fields_raw(self, ptr, n, mask):
        _lib.check(self.lib.cdb_upload_agents_fields(self.handle, C.c_void_p(int(ptr)), int(n), self.itemsize, int(mask)))

    def host_register(self, agents):
        """Pin + map the host array so that field-masked transfers run as zero-copy kernels (see include/crowd_b200.h)."""
        _lib.check(self.lib.cdb_host_register(self.handle, _ptr(agents), len(agents), agents.dtype.itemsize))

    def host_unregister(self, agents):
        _lib.check(self.lib.cdb_host_unregister(self.handle, _ptr(agents)))

    def transfer_stats(self, reset=False):
        """-> (h2d_bytes, d2h_bytes) moved over PCIe by upload / download calls so far."""
        a, b = C.c_int64(), C.c_int64()
        _lib.check(self.lib.cdb_transfer_stats(self.handle, C.byref(a), C.byref(b), 1 if reset else 0))
        return a.value, b.value

    def download(self, agents, mask=_lib.F_ALL_MUTABLE):
        if _check_agents(agents) != self.model:
            raise InvalidType('agents dtype does not match the device model')
        _lib.check(self.lib.cdb_download_agents_aos(self.handle, _ptr(agents), len(agents), agents.dtype.itemsize, mask))

    def download_raw(self, ptr, n, mask=_lib.F_WHOLE_RECORD):
        _lib.check(self.lib.cdb_download_agents_aos(self.handle, C.c_void_p(int(ptr)), int(n), self.itemsize, mask))

    def set_obstacles(self, obstacles):
        seg = as_obstacles(obstacles)
        _lib.check(self.lib.cdb_set_obstacles(self.handle, _ptr(seg) if len(seg) else None, len(seg)))

    def build_navigation_field(self, target, target_segments, obstacle_segments, bounds, step, radius=0.5, strength=0.3,
                               want_maps=False):
        """Field.navigation_to_target (reference simulation/field.py:155-164) computed on the device from line-segment
        geometry and installed as the navigation field of ``target``: eikonal distance map around the obstacles buffered by
        ``radius``, normalised gradient, fill of the buffer zone, blend away from the walls (``strength``).
        ``bounds`` = (minx, miny, maxx, maxy) of the domain, ``step`` the grid spacing (quickest_path.meshgrid).
        -> mgrid, or (mgrid, distance_map, (U, V)) as the reference returns when ``want_maps`` (host copies; a 2000 m room at
        step 0.1 is 3 x 3.2 GB -- leave it off and nothing of that size ever exists on the host)."""
        from .synthetic import MeshGrid
        mg = MeshGrid(float(step), *[float(b) for b in bounds])
        ny, nx = mg.shape
        tseg = np.ascontiguousarray(np.asarray(target_segments, dtype=np.float64).reshape(-1, 4))
        oseg = as_obstacles(obstacle_segments) if obstacle_segments is not None else np.zeros((0, 4))
        oseg = np.ascontiguousarray(oseg, dtype=np.float64)
        dmap = np.empty((ny, nx)) if want_maps else None
        U = np.empty((ny, nx)) if want_maps else None
        V = np.empty((ny, nx)) if want_maps else None
        rounds = C.c_int64()
        _lib.check(self.lib.cdb_build_navigation_field(
            self.handle, int(target), _ptr(tseg), len(tseg), _ptr(oseg) if len(oseg) else None, len(oseg), ny, nx, float(bounds[0]),
            float(bounds[1]), float(step), float(radius), float(strength), _ptr(dmap) if want_maps else None,
            _ptr(U) if want_maps else None, _ptr(V) if want_maps else None, C.byref(rounds)))
        self.last_field_rounds = rounds.value
        return (mg, dmap, (U, V)) if want_maps else mg

    def set_navigation_field(self, target, mgrid, direction_map):
        """(mgrid, (U, V)) as returned by Field.navigation_to_target (reference field.py:155-164)."""
        U = np.ascontiguousarray(np.asarray(direction_map[0]), dtype=np.float64)
        V = np.ascontiguousarray(np.asarray(direction_map[1]), dtype=np.float64)
        if U.shape != V.shape or U.ndim != 2:
            raise InvalidValue('direction map must be two (ny, nx) arrays')
        minx, miny = mgrid.bounds[0], mgrid.bounds[1]
        _lib.check(self.lib.cdb_set_navigation_field(self.handle, int(target), _ptr(U), _ptr(V), U.shape[0], U.shape[1],
                                                     float(minx), float(miny), float(mgrid.step)))

    def clear_navigation(self):
        _lib.check(self.lib.cdb_clear_navigation(self.handle))

    # -- nodes -----------------------------------------------------------------------------------------------------
    def reset(self):
        _lib.check(self.lib.cdb_reset(self.handle))

    def set_seed(self, seed):
        _lib.check(self.lib.cdb_set_seed(self.handle, int(seed) & (2 ** 64 - 1)))

    def fluctuation(self):
        _lib.check(self.lib.cdb_fluctuation(self.handle))

    def navigation(self):
        _lib.check(self.lib.cdb_navigation(self.handle))

    def orientation(self):
        _lib.check(self.lib.cdb_orientation(self.handle))

    def adjust(self):
        _lib.check(self.lib.cdb_adjust(self.handle))

    def agent_agent(self, cell_size):
        _lib.check(self.lib.cdb_agent_agent(self.handle, float(cell_size)))

    def agent_obstacle(self):
        _lib.check(self.lib.cdb_agent_obstacle(self.handle))

    def integrate(self, dt_min, dt_max):
        dt = C.c_double()
        _lib.check(self.lib.cdb_integrate(self.handle, float(dt_min), float(dt_max), C.byref(dt)))
        return dt.value

    def step(self, n_steps=1, flags=_lib.STEP_ALL, cell_size=3.6, dt_min=0.01, dt_max=0.01, want_dt=True):
        """n_steps fused iterations of the selected nodes on the device; returns the dt of every step (or None)."""
        dts = np.zeros(int(n_steps), dtype=np.float64) if want_dt else None
        _lib.check(self.lib.cdb_step(self.handle, int(flags), float(cell_size), float(dt_min), float(dt_max),
                                     int(n_steps), _ptr(dts) if want_dt and n_steps else None))
        return dts

    # -- asynchronous host-visible state ---------------------------------------------------------------------------------
    def set_deferred_sync(self, enable=True):
        _lib.check(self.lib.cdb_set_deferred_sync(self.handle, 1 if enable else 0))

    def sync_count(self):
        """Blocking host synchronisations the library has performed for this sim so far."""
        return int(self.lib.cdb_sync_count(self.handle))

    def snapshot_begin(self):
        """Queue a copy of the whole packed records (current device state) into a pinned host buffer; returns the slot."""
        slot = C.c_int64()
        _lib.check(self.lib.cdb_snapshot_begin(self.handle, C.byref(slot)))
        return slot.value

    def snapshot_wait(self, slot, dtype):
        """-> structured array VIEW of the slot's pinned buffer (valid until the slot is reused two snapshots later)."""
        ptr, n = C.c_void_p(), C.c_int64()
        _lib.check(self.lib.cdb_snapshot_wait(self.handle, int(slot), C.byref(ptr), C.byref(n)))
        if n.value == 0:
            return np.zeros(0, dtype=dtype)
        buf = (C.c_uint8 * (n.value * dtype.itemsize)).from_address(ptr.value)
        return np.frombuffer(buf, dtype=dtype, count=n.value)

    def scalars_begin(self):
        slot = C.c_int64()
        _lib.check(self.lib.cdb_scalars_begin(self.handle, C.byref(slot)))
        return slot.value

    def scalars_wait(self, slot, n_targets=0):
        """-> (dt, time_tot, inside_domain_changes, target_counts) as they were when scalars_begin(slot) was queued."""
        dt, tt, ch = C.c_double(), C.c_double(), C.c_int64()
        counts = np.zeros(int(n_targets), dtype=np.int64)
        _lib.check(self.lib.cdb_scalars_wait(self.handle, int(slot), C.byref(dt), C.byref(tt), C.byref(ch),
                                             _ptr(counts) if n_targets else None, int(n_targets)))
        return dt.value, tt.value, ch.value, counts

    def set_rebuild_policy(self, skin_fraction=0.10, max_interval=16, min_agents=16384):
        """Resident-order steps (include/crowd_b200.h): rebuild the block list at most every ``max_interval`` steps, on search
        cells ``1 + skin_fraction`` times wider; ``max_interval=1`` rebuilds at every step."""
        _lib.check(self.lib.cdb_set_rebuild_policy(self.handle, float(skin_fraction), int(max_interval), int(min_agents)))

    def rebuild_stats(self):
        """-> dict(rebuilds, kept, stale, interval)"""
        v = [C.c_int64() for _ in range(4)]
        _lib.check(self.lib.cdb_get_rebuild_stats(self.handle, *[C.byref(x) for x in v]))
        return dict(zip(('rebuilds', 'kept', 'stale', 'interval'), (x.value for x in v)))

    def set_small_crowd_max(self, max_agents):
        """Crowds up to this size are stepped by one thread block, all steps of a call in one launch (0: never)."""
        _lib.check(self.lib.cdb_set_small_crowd_max(self.handle, int(max_agents)))

    def set_graphs(self, enable):
        _lib.check(self.lib.cdb_set_graphs(self.handle, 1 if enable else 0))

    def time(self):
        t, it = C.c_double(), C.c_int64()
        _lib.check(self.lib.cdb_get_time(self.handle, C.byref(t), C.byref(it)))
        return t.value, it.value

    # -- instrumentation ---------------------------------------------------------------------------------------------
    def set_variant(self, variant):
        _lib.check(self.lib.cdb_set_variant(self.handle, int(variant)))

    def launch_count(self):
        return int(self.lib.cdb_launch_count(self.handle))

    def profile(self, enable=True):
        _lib.check(self.lib.cdb_profile_enable(self.handle, 1 if enable else 0))

    def profile_read(self):
        """-> (ms_blocklist_and_pre, ms_agent_agent, ms_post, steps) summed since the last read."""
        ms = (C.c_double * 3)()
        steps = C.c_int64()
        _lib.check(self.lib.cdb_profile_read(self.handle, ms, C.byref(steps)))
        return ms[0], ms[1], ms[2], steps.value

    def profile_read_phases(self):
        """-> (ms_blocklist_and_pre, ms_pair_sweep, ms_pair_eval, ms_step_kernel, ms_post, steps) summed since the last read."""
        ms = (C.c_double * 5)()
        steps = C.c_int64()
        _lib.check(self.lib.cdb_profile_read_phases(self.handle, ms, C.byref(steps)))
        return ms[0], ms[1], ms[2], ms[3], ms[4], steps.value

    def set_pair_capacity(self, pairs):
        """Fix the capacity of the pair list of kernel variant 3 (0 = automatic); a test hook for the repeat path."""
        _lib.check(self.lib.cdb_set_pair_capacity(self.handle, int(pairs)))

    def set_search_refinement(self, refinement):
        """0 = automatic (search on cell_size / 2 where valid), 1 = always search on the cell_size lattice."""
        _lib.check(self.lib.cdb_set_search_refinement(self.handle, int(refinement)))

    def ext_max(self):
        """Bound on the radius / body extent of the uploaded agents (the refined search needs 3 + 2 ext_max < cell_size)."""
        v = C.c_double()
        _lib.check(self.lib.cdb_get_ext_max(self.handle, C.byref(v)))
        return v.value

    def pair_stats(self):
        """-> (capacity, pairs listed by the last step, steps repeated after an overflow)."""
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        _lib.check(self.lib.cdb_get_pair_stats(self.handle, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    # -- block list exports ------------------------------------------------------------------------------------------
    def build_block_list(self, cell_size):
        _lib.check(self.lib.cdb_build_block_list(self.handle, float(cell_size)))

    def grid(self):
        g = (C.c_int64 * 4)()
        _lib.check(self.lib.cdb_get_grid(self.handle, g))
        return tuple(int(x) for x in g)

    def cell_ids(self):
        out = np.empty(self.n, dtype=np.int64)
        _lib.check(self.lib.cdb_get_cell_ids(self.handle, _ptr(out), self.n))
        return out

    def cell_tables(self):
        """-> (points_indices, cells_count, cells_offset, grid_shape) like cell_lists.add_to_cells."""
        g = self.grid()
        nc = g[2] * g[3]
        pi = np.empty(self.n, dtype=np.int64)
        cc = np.empty(nc, dtype=np.int64)
        co = np.empty(nc, dtype=np.int64)
        _lib.check(self.lib.cdb_get_cell_tables(self.handle, _ptr(pi), self.n, _ptr(cc), _ptr(co), nc))
        return pi, cc, co, np.array(g[2:], dtype=np.int64)

    def neighbor_pairs(self):
        cnt = C.c_int64()
        _lib.check(self.lib.cdb_get_neighbor_pairs(self.handle, None, 0, C.byref(cnt)))
        out = np.empty((cnt.value, 2), dtype=np.int64)
        if cnt.value:
            _lib.check(self.lib.cdb_get_neighbor_pairs(self.handle, _ptr(out), cnt.value, C.byref(cnt)))
        return out

    # -- collective motion ---------------------------------------------------------------------------
    def set_states(self, agents, target=True):
        """Send the States fields the collective-motion nodes read (agents.py:33-60) next to the uploaded records."""
        n = len(agents)
        f = {k: np.ascontiguousarray(agents[k]) for k in ('is_leader', 'is_follower', 'index_leader', 'familiar_exit')}
        tg = np.ascontiguousarray(agents['target']) if target else None
        _lib.check(self.lib.cdb_set_states(self.handle, _ptr(tg) if target and n else None,
                                           _ptr(f['is_leader'].view(np.uint8)) if n else None,
                                           _ptr(f['is_follower'].view(np.uint8)) if n else None,
                                           _ptr(f['index_leader']) if n else None, _ptr(f['familiar_exit']) if n else None, n))

    def get_states(self, agents):
        """Write back target / is_follower / index_leader (the States fields these nodes mutate)."""
        n = len(agents)
        if n == 0:
            return
        tg, fo, il = np.empty(n, np.int64), np.empty(n, np.uint8), np.empty(n, np.int64)
        _lib.check(self.lib.cdb_get_states(self.handle, _ptr(tg), _ptr(fo), _ptr(il), n))
        agents['target'] = tg
        agents['is_follower'] = fo.astype(bool)
        agents['index_leader'] = il

    def exit_detection(self, center_door, detection_range, apply=False):
        doors = np.ascontiguousarray(center_door, dtype=np.float64).reshape(-1, 2)
        _lib.check(self.lib.cdb_exit_detection(self.handle, _ptr(doors) if len(doors) else None, len(doors),
                                               float(detection_range), 1 if apply else 0))

    def exit_detection_result(self):
        det, has = np.empty(self.n, np.int64), np.empty(self.n, np.uint8)
        _lib.check(self.lib.cdb_get_exit_detection(self.handle, _ptr(det) if self.n else None,
                                                   _ptr(has) if self.n else None, self.n))
        return det, has.astype(bool)

    def nearest_neighbors(self, sight, size_nearest_other):
        out = np.full((self.n, int(size_nearest_other)), -1, dtype=np.int64)
        _lib.check(self.lib.cdb_nearest_neighbors(self.handle, float(sight), int(size_nearest_other),
                                                  _ptr(out) if out.size else None))
        return out

    def leader_follower(self, sight, phi=0.45 * np.pi, weight_position_leader=0.40):
        _lib.check(self.lib.cdb_leader_follower(self.handle, float(sight), float(phi), float(weight_position_leader)))

    def leader_follower_with_herding(self, sight, size_nearest_other, phi=0.45 * np.pi, weight_position_herding=0.15,
                                     weight_position_leader=0.40, weight_direction_leader=0.65):
        _lib.check(self.lib.cdb_leader_follower_with_herding(
            self.handle, float(sight), int(size_nearest_other), float(phi), float(weight_position_herding),
            float(weight_position_leader), float(weight_direction_leader)))

    def direction(self):
        out = np.zeros((self.n, 2), dtype=np.float64)
        _lib.check(self.lib.cdb_get_direction(self.handle, _ptr(out) if self.n else None, self.n))
        return out

    # -- host-visible state nodes ------------------------------------------------------------------------
    def set_polygons(self, which, polygons):
        """polygons: list of (nv, 2) vertex arrays (a repeated closing vertex, as in shapely's exterior, is dropped)."""
        polys = []
        for v in polygons:
            v = np.ascontiguousarray(v, dtype=np.float64).reshape(-1, 2)
            if len(v) > 1 and (v[0] == v[-1]).all():
                v = v[:-1]
            polys.append(v)
        offsets = np.zeros(len(polys) + 1, dtype=np.int64)
        offsets[1:] = np.cumsum([len(v) for v in polys])
        xy = np.ascontiguousarray(np.concatenate(polys)) if polys else np.zeros((0, 2))
        _lib.check(self.lib.cdb_set_polygons(self.handle, int(which), _ptr(xy) if len(xy) else None,
                                             _ptr(offsets) if polys else None, len(polys)))

    def set_active(self, active):
        a = np.ascontiguousarray(active).view(np.uint8)
        _lib.check(self.lib.cdb_set_active(self.handle, _ptr(a) if len(a) else None, len(a)))

    def get_active(self, n=None):
        """``n``: size of the flag array -- the uploaded agents, or the whole crowd in strip mode (cdb_strip_set_global_agents)"""
        n = self.n if n is None else int(n)
        out = np.zeros(n, dtype=np.uint8)
        _lib.check(self.lib.cdb_get_active(self.handle, _ptr(out) if n else None, n))
        return out.astype(bool)

    def inside_domain(self, want_count=True):
        c = C.c_int64(0)
        _lib.check(self.lib.cdb_inside_domain(self.handle, C.byref(c) if want_count else None))
        return c.value if want_count else None

    def target_reached(self, n_polygons, want_counts=True):
        counts = np.zeros(int(n_polygons), dtype=np.int64)
        _lib.check(self.lib.cdb_target_reached(self.handle, _ptr(counts) if want_counts and n_polygons else None, int(n_polygons)))
        return counts if want_counts else None

    def target_reached_by(self, n_polygons, n=None):
        n = self.n if n is None else int(n)
        out = np.zeros((int(n_polygons), n), dtype=np.uint8)
        _lib.check(self.lib.cdb_get_target_reached(self.handle, _ptr(out) if out.size else None, int(n_polygons), n))
        return out.astype(bool)

    def set_lattice(self, ix_min, iy_min, nx, ny):
        _lib.check(self.lib.cdb_set_lattice(self.handle, int(ix_min), int(iy_min), int(nx), int(ny)))

    def clear_lattice(self):
        _lib.check(self.lib.cdb_clear_lattice(self.handle))


_CACHE = {}


def device_agents_for(agents, device=0):
    """Cached DeviceAgents for (device, model) -- used by the strict functional API."""
    model = _check_agents(agents)
    key = (device, model)
    da = _CACHE.get(key)
    if da is None:
        da = _CACHE[key] = DeviceAgents(model, capacity=len(agents), device=device)
    return da


def host_round_trips(crowds, n_updates, flags=_lib.STEP_ALL, cell_size=3.6, dt_min=0.01, dt_max=0.01):
    """Step independent crowds through the HOST boundary concurrently: a replica study (the reference runs its replicas as
    independent processes, ``MultiAgentProcess``, ``crowddynamics/simulation/multiagent.py:58-101``) on one device.

    ``crowds`` is a list of ``(DeviceAgents, host_pointer, n_agents)`` -- every crowd on its own device handle (hence its own
    CUDA stream), its packed host records pinned.  Every update of every crowd is the strict round trip ``upload whole
    records -> one fused step -> download whole records``; the crowds run on one host thread each (the C ABI calls release
    the GIL), so the upload of one crowd overlaps the download of another -- PCIe is full duplex, a single crowd can only use
    one direction at a time -- and the kernels of one hide behind the copies of the others.  Results are those of running
    the crowds one after the other (nothing is shared between the handles)."""
    import threading
    errors = []

    def work(dev, ptr, n):
        try:
            for _ in range(int(n_updates)):
                dev.upload_raw(ptr, n)
                dev.step(1, flags, cell_size, dt_min, dt_max, want_dt=False)
                dev.download_raw(ptr, n)
        except BaseException as exc:      # surfaced in the calling thread
            errors.append(exc)

    threads = [threading.Thread(target=work, args=c) for c in crowds]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
