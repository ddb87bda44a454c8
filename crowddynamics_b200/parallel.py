"""Spatial strip decomposition of one crowd over the GPUs of a node (no reference counterpart).

One process per GPU (``torch.distributed``, NCCL over NVLink).  The cell lattice is global and anchored at multiples of
the cell size, so every rank computes identical cell coordinates; rank g owns a contiguous range of cell *columns*.
Per step and per neighbour one halo message (the first / last owned cell column as packed neighbour records + per-cell
counts) and one migrant message (agents whose column left the strip, with their full state) are exchanged with
``batch_isend_irecv``; when dt_min != dt_max the two maxima of ``adaptive_timestep`` are MAX-all-reduced.  Forces are
computed in gather form on the owner of each agent, so no force ever travels back.

The message buffers are ordinary torch tensors; the CUDA library packs / unpacks them through raw device pointers
(``include/crowd_b200.h``: cdb_strip_begin / cdb_strip_finish / cdb_strip_absorb).
"""
import ctypes as C
import math

import numpy as np

from . import _lib
from .structures import MODEL_CIRCULAR, MODEL_THREE_CIRCLE, model_of


def partition_columns(ix_min, nx, world):
    """Split cell columns [ix_min, ix_min + nx) into `world` contiguous strips of (almost) equal width.
    Returns world + 1 boundaries."""
    if nx < world:
        raise ValueError('fewer cell columns (%d) than ranks (%d)' % (nx, world))
    return [ix_min + (nx * g) // world for g in range(world + 1)]


def lattice_of(positions, cell_size, pad=1):
    """Global lattice (ix_min, iy_min, nx, ny) covering all positions, padded by `pad` cells."""
    c = np.floor(np.asarray(positions) / cell_size).astype(np.int64)
    lo, hi = c.min(axis=0) - pad, c.max(axis=0) + pad
    return int(lo[0]), int(lo[1]), int(hi[0] - lo[0] + 1), int(hi[1] - lo[1] + 1)


def owner_of_columns(cols, bounds):
    """Rank owning each cell column (columns outside the lattice belong to the first / last rank)."""
    return np.clip(np.searchsorted(np.asarray(bounds[1:-1]), cols, side='right'), 0, len(bounds) - 2)


class CudaStripDevice:
    """Adapter: DeviceAgents + the strip entry points of the C ABI, taking torch tensors as message buffers."""

    def __init__(self, model, capacity, device_index, stream=None):
        from .engine import DeviceAgents
        self.dev = DeviceAgents(model, capacity=capacity, device=device_index, stream=stream)
        self.lib = self.dev.lib
        self.handle = self.dev.handle

    @staticmethod
    def _p(t):
        return C.c_void_p(t.data_ptr()) if t is not None else None

    def upload(self, agents, ids):
        self.dev.upload(agents)
        ids = np.ascontiguousarray(ids, dtype=np.int64)
        _lib.check(self.lib.cdb_set_agent_ids(self.handle, C.c_void_p(ids.ctypes.data), len(ids)))

    def ext_max(self):
        return self.dev.ext_max()

    def set_search_refinement(self, refinement):
        self.dev.set_search_refinement(refinement)

    def set_rebuild_policy(self, skin_fraction, max_interval, min_agents=0):
        self.dev.set_rebuild_policy(skin_fraction, max_interval, min_agents)

    def set_kind(self, kind):
        _lib.check(self.lib.cdb_strip_set_kind(self.handle, int(kind)))

    def drift(self):
        """-> (largest displacement of the last step, drift bound since the last rebuild, its limit); one synchronisation"""
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        _lib.check(self.lib.cdb_strip_drift(self.handle, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def rebuild_stats(self):
        return self.dev.rebuild_stats()

    # -- host-visible state nodes in strip mode: flag arrays indexed by GLOBAL agent id --------------------------------------
    def set_global_agents(self, n_global):
        self.n_global = int(n_global)
        _lib.check(self.lib.cdb_strip_set_global_agents(self.handle, int(n_global)))

    def set_polygons(self, which, polygons):
        self.dev.set_polygons(which, polygons)

    def set_active(self, active):
        self.dev.set_active(active)

    def get_active(self):
        return self.dev.get_active(self.n_global)

    def inside_domain(self):
        return self.dev.inside_domain()

    def target_reached(self, n_polygons):
        return self.dev.target_reached(n_polygons)

    def target_reached_by(self, n_polygons):
        return self.dev.target_reached_by(n_polygons, self.n_global)

    def set_strip(self, ix_min, iy_min, nx_owned, ny, has_left, has_right, halo_cap, mig_cap):
        _lib.check(self.lib.cdb_set_strip(self.handle, ix_min, iy_min, nx_owned, ny, int(has_left), int(has_right),
                                          halo_cap, mig_cap))
        return int(self.lib.cdb_halo_buffer_doubles(self.handle)), int(self.lib.cdb_migrant_buffer_doubles(self.handle))

    def set_obstacles(self, obstacles):
        self.dev.set_obstacles(obstacles)

    def set_navigation_field(self, target, mgrid, direction_map):
        self.dev.set_navigation_field(target, mgrid, direction_map)

    # -- one-sided exchange over peer memory (NVLink) ---------------------------------------------------------------------
    def exchange_alloc(self):
        _lib.check(self.lib.cdb_strip_exchange_alloc(self.handle))

    def exchange_handles(self):
        buf = C.create_string_buffer(int(self.lib.cdb_strip_exchange_handle_bytes()))
        _lib.check(self.lib.cdb_strip_exchange_handles(self.handle, buf))
        return buf.raw

    def connect_ipc(self, left, right):
        lb = C.create_string_buffer(left, len(left)) if left is not None else None
        rb = C.create_string_buffer(right, len(right)) if right is not None else None
        _lib.check(self.lib.cdb_strip_exchange_connect_ipc(self.handle, lb, rb))

    def connect_local(self, left, right):
        _lib.check(self.lib.cdb_strip_exchange_connect_local(self.handle, left.handle if left is not None else None,
                                                             right.handle if right is not None else None))

    def begin_direct(self, flags, cell_size, send_halo=True):
        _lib.check(self.lib.cdb_strip_begin_direct(self.handle, flags, cell_size, 1 if send_halo else 0))

    def finish_direct(self, flags, dt_min, dt_max, recv_halo=True):
        _lib.check(self.lib.cdb_strip_finish_direct(self.handle, flags, dt_min, dt_max, 1 if recv_halo else 0))

    def absorb_direct(self, exact=False):
        n = C.c_int64()
        _lib.check(self.lib.cdb_strip_absorb_direct(self.handle, C.byref(n) if exact else None))
        return n.value if exact else None

    def begin(self, flags, cell_size, halo_left_out, halo_right_out):
        _lib.check(self.lib.cdb_strip_begin(self.handle, flags, cell_size, self._p(halo_left_out), self._p(halo_right_out)))

    def export_vmax(self, buf):
        _lib.check(self.lib.cdb_strip_export_vmax(self.handle, self._p(buf)))

    def import_vmax(self, buf):
        _lib.check(self.lib.cdb_strip_import_vmax(self.handle, self._p(buf)))

    def finish(self, flags, dt_min, dt_max, halo_left_in, halo_right_in, mig_left_out, mig_right_out):
        _lib.check(self.lib.cdb_strip_finish(self.handle, flags, dt_min, dt_max, self._p(halo_left_in),
                                             self._p(halo_right_in), self._p(mig_left_out), self._p(mig_right_out)))

    def absorb(self, mig_left_in, mig_right_in, exact=False):
        """Appends the received migrants; returns None: the exact count stays on the device (see count()) unless ``exact``
        (one host synchronisation; required after bursts such as the initial settle())."""
        n = C.c_int64()
        _lib.check(self.lib.cdb_strip_absorb(self.handle, self._p(mig_left_in), self._p(mig_right_in), C.byref(n) if exact else None))
        return n.value if exact else None

    def count(self):
        n = C.c_int64()
        _lib.check(self.lib.cdb_strip_count(self.handle, C.byref(n)))
        return n.value

    def export_agents(self, dtype):
        cnt = C.c_int64()
        cap = int(self.lib.cdb_num_agents(self.handle))
        out = np.zeros(max(cap, 1), dtype=dtype)
        ids = np.zeros(max(cap, 1), dtype=np.int64)
        _lib.check(self.lib.cdb_export_agents(self.handle, C.c_void_p(out.ctypes.data), C.c_void_p(ids.ctypes.data), cap,
                                              C.byref(cnt)))
        return out[:cnt.value], ids[:cnt.value]

    def time(self):
        return self.dev.time()

    # instrumentation passthrough (bench.py)
    def profile(self, enable=True):
        self.dev.profile(enable)

    def profile_read(self):
        return self.dev.profile_read()

    def profile_read_phases(self):
        return self.dev.profile_read_phases()

    def launch_count(self):
        return self.dev.launch_count()


class StripSimulation:
    """One rank's strip of a crowd that is decomposed along x."""

    def __init__(self, dev, rank, world, bounds, lattice, cell_size, halo_cap, mig_cap, tensor_device, n_owned,
                 dist=None, flags=_lib.STEP_ALL, dt_min=0.01, dt_max=0.01, model=None, ext_max=None, skin=0.0, max_interval=16):
        """``bounds`` / ``lattice`` are in columns of ``cell_size * (1 + skin)``.  ``skin`` > 0: block lists are kept for up to
        ``max_interval`` steps (cdb_strip_set_kind); every rank must pass the same values."""
        import torch
        self.torch = torch
        self.dist = dist
        self.dev = dev
        self.rank, self.world = rank, world
        self.bounds = list(bounds)
        self.cell_size = float(cell_size)
        self.flags, self.dt_min, self.dt_max = flags, dt_min, dt_max
        self.left = rank - 1 if rank > 0 else None
        self.right = rank + 1 if rank < world - 1 else None
        ix_min, iy_min, nx, ny = lattice
        # Search lattice: every rank must bin on the same one, and the choice mirrors what a single device does on its own
        # (cells of cell_size / 2 whenever no pair can interact beyond cell_size, which needs the largest radius / body
        # extent of the WHOLE crowd), so that the strips reproduce the single-device run bit for bit.
        self.refinement = 1
        if model is not None and hasattr(dev, 'set_search_refinement'):
            if ext_max is None:
                ext_max = dev.ext_max()
                if dist is not None and world > 1:
                    tmax = torch.tensor([ext_max], dtype=torch.float64, device=tensor_device)
                    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
                    ext_max = float(tmax.item())
            if (3.0 + 2.0 * ext_max) * (1.0 + 1e-9) < self.cell_size:
                self.refinement = 2
        if hasattr(dev, 'set_search_refinement'):
            dev.set_search_refinement(self.refinement)
        # kept block lists: the same interval on every rank, sized from the all-reduced displacement (adapt_interval)
        self.skin = float(skin) if hasattr(dev, 'set_kind') else 0.0
        self.max_interval = int(max_interval)
        self.interval = 1
        self._since = 0               # steps issued on the current block list
        self._steps = 0
        self._force_rebuild = True    # the first step, and the first one after settle(), rebuild
        if self.skin > 0.0:
            if ext_max is None or not (3.0 + 2.0 * ext_max) * (1.0 + 1e-9) < self.cell_size:
                raise ValueError('kept block lists need 3 + 2 max R < cell_size (the pair set must not depend on the lattice)')
            dev.set_rebuild_policy(self.skin, max(2, self.max_interval))
        halo_doubles, mig_doubles = dev.set_strip(self.bounds[rank], iy_min, self.bounds[rank + 1] - self.bounds[rank], ny,
                                                  self.left is not None, self.right is not None, halo_cap, mig_cap)

        def buf(n):
            return torch.zeros(n, dtype=torch.float64, device=tensor_device)
        self.halo_out = {s: buf(halo_doubles) for s in ('l', 'r')}
        self.halo_in = {s: buf(halo_doubles) for s in ('l', 'r')}
        self.mig_out = {s: buf(mig_doubles) for s in ('l', 'r')}
        self.mig_in = {s: buf(mig_doubles) for s in ('l', 'r')}
        self.vmax = buf(4)           # {max |v|, max v0, NaN flags of the two}: see k_vmax_export
        self._n_owned = n_owned
        self.direct = False          # True: halo / migrants travel by one-sided writes into the neighbours' buffers

    def connect_direct(self):
        """One process per GPU: replace the two NCCL send/recv rounds of a step by one-sided writes over NVLink peer memory
        (CUDA IPC).  Collective: every rank must call it.  Falls back to NCCL (returns False) where IPC is not possible."""
        ok = 1
        handles = None
        try:
            self.dev.exchange_alloc()
            handles = self.dev.exchange_handles()
        except Exception:
            ok = 0
        gathered = [None] * self.world
        self.dist.all_gather_object(gathered, (ok, handles))
        if not all(g[0] for g in gathered):
            return False
        try:
            self.dev.connect_ipc(gathered[self.left][1] if self.left is not None else None,
                                 gathered[self.right][1] if self.right is not None else None)
        except Exception:
            ok = 0
        flags = [None] * self.world
        self.dist.all_gather_object(flags, ok)
        self.direct = all(flags)
        return self.direct

    def n_owned(self):
        """Agents currently owned by this rank (asks the device when the last step did not report it)."""
        if self._n_owned is None:
            self._n_owned = self.dev.count()
        return self._n_owned

    # -- neighbour exchange ---------------------------------------------------------------------------------------------
    def _exchange(self, out, inn):
        """Send out['l'] to the left neighbour and out['r'] to the right one; receive into inn['l'] / inn['r']."""
        if self.world == 1 or self.direct:
            return
        dist = self.dist
        ops = []
        if self.left is not None:
            ops.append(dist.P2POp(dist.isend, out['l'], self.left))
            ops.append(dist.P2POp(dist.irecv, inn['l'], self.left))
        if self.right is not None:
            ops.append(dist.P2POp(dist.isend, out['r'], self.right))
            ops.append(dist.P2POp(dist.irecv, inn['r'], self.right))
        for req in dist.batch_isend_irecv(ops):
            req.wait()

    # -- the phases of one step (driven by step() over torch.distributed, or by LocalGroup in one process) ---------------
    def phase_begin(self, flags=None, send_halo=True):
        L, R = self.left is not None, self.right is not None
        flags = self.flags if flags is None else flags
        if self.direct:
            return self.dev.begin_direct(flags, self.cell_size, send_halo)
        self.dev.begin(flags, self.cell_size, self.halo_out['l'] if L and send_halo else None, self.halo_out['r'] if R and send_halo else None)

    def phase_finish(self, flags=None, use_halo=True):
        L, R = self.left is not None, self.right is not None
        flags = self.flags if flags is None else flags
        if self.direct:
            return self.dev.finish_direct(flags, self.dt_min, self.dt_max, use_halo)
        self.dev.finish(flags, self.dt_min, self.dt_max,
                        self.halo_in['l'] if L and use_halo else None, self.halo_in['r'] if R and use_halo else None,
                        self.mig_out['l'] if L else None, self.mig_out['r'] if R else None)

    def phase_absorb(self, exact=False):
        L, R = self.left is not None, self.right is not None
        if self.direct:
            self._n_owned = self.dev.absorb_direct(exact)
            return
        self._n_owned = self.dev.absorb(self.mig_in['l'] if L else None, self.mig_in['r'] if R else None, exact)

    @property
    def adaptive(self):
        return self.dt_min != self.dt_max and bool(self.flags & _lib.STEP_INTEGRATOR)

    PHASES = ('block_list_and_halo_pack', 'halo_exchange', 'halo_unpack_pairs_finish_migrant_pack', 'migrant_exchange', 'absorb')

    def profile_phases(self, enable=True):
        """CUDA-event timing of the five phases of step() on the current stream (bench.py, N > 1)."""
        self._phase_events = [] if enable else None

    def phase_ms(self):
        """-> {phase: mean ms per step} over the steps recorded since profile_phases(True)."""
        ev = getattr(self, '_phase_events', None) or []
        self.torch.cuda.synchronize()
        out = {}
        for k, name in enumerate(self.PHASES):
            out[name] = float(np.mean([e[k].elapsed_time(e[k + 1]) for e in ev])) if ev else 0.0
        return out

    def _mark(self, marks):
        if marks is not None:
            e = self.torch.cuda.Event(enable_timing=True)
            e.record()
            marks.append(e)

    # -- kept block lists: which kind of step comes next (cdb_strip_set_kind); identical on every rank ---------------------
    def plan_step(self):
        """-> (kind, migrate): 0 rebuild + migrants (no kept lists at all), 1 rebuild, 2 kept, 3 kept + migrants, 4 rebuild + migrants"""
        if self.skin <= 0.0:
            return 0, True
        rebuild = self._force_rebuild or self._since == 0
        last = self._since + 1 >= self.interval          # the next step rebuilds
        kind = (4 if last else 1) if rebuild else (3 if last else 2)
        self.dev.set_kind(kind)
        return kind, last

    def end_step(self, migrate):
        self._force_rebuild = False
        self._since = 0 if migrate else self._since + 1
        self._steps += 1

    def adapt_due(self):
        return self.skin > 0.0 and (self._steps == 2 or self._steps % 32 == 0)

    def adapt_interval(self, disp_max, limit):
        """Common rebuild interval from the largest per-step displacement over ALL ranks (margin 1.5; the device refuses to
        sweep a stale list, which in strip mode is an error rather than a repeat)."""
        if disp_max > 0.0 and np.isfinite(disp_max) and limit > 0.0:
            self.interval = int(max(1, min(self.max_interval, np.floor(limit / (1.5 * disp_max)))))
        else:
            self.interval = 1

    def step(self, n_steps=1):
        rec = getattr(self, '_phase_events', None)
        for _ in range(n_steps):
            marks = [] if rec is not None and len(rec) < 2048 else None
            self._mark(marks)
            kind, migrate = self.plan_step()
            self.phase_begin()
            self._mark(marks)
            self._exchange(self.halo_out, self.halo_in)
            if self.adaptive and self.world > 1:
                self.dev.export_vmax(self.vmax)
                self.dist.all_reduce(self.vmax, op=self.dist.ReduceOp.MAX)
                self.dev.import_vmax(self.vmax)
            self._mark(marks)
            self.phase_finish()
            self._mark(marks)
            if migrate:
                self._exchange(self.mig_out, self.mig_in)
            self._mark(marks)
            if migrate:
                self.phase_absorb()
            self._mark(marks)
            self.end_step(migrate)
            if marks is not None:
                rec.append(marks)
            if self.adapt_due():
                last, _, limit = self.dev.drift()
                if self.dist is not None and self.world > 1:
                    t = self.torch.tensor([last], dtype=self.torch.float64, device=self.vmax.device)
                    self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
                    last = float(t.item())
                self.adapt_interval(last, limit)

    # -- InsideDomain / TargetReached on strips (simulation/logic.py:343-387) ---------------------------------------------
    def set_domain(self, polygon, n_global, active=None):
        """``polygon``: (nv, 2) vertices of the domain; ``n_global``: agents of the whole crowd; ``active``: the initial
        States.active flags by global id (default: all active).  Every rank calls it with the same arguments."""
        self._n_global = int(n_global)
        self.dev.set_global_agents(n_global)
        self.dev.set_polygons(_lib.POLY_DOMAIN, [np.asarray(polygon, dtype=np.float64)])
        self.dev.set_active(np.ones(n_global, dtype=bool) if active is None else np.ascontiguousarray(active, dtype=bool))

    def set_targets(self, polygons, n_global):
        if getattr(self, '_n_global', 0) != int(n_global):
            self._n_global = int(n_global)
            self.dev.set_global_agents(n_global)
        self._n_targets = len(polygons)
        self.dev.set_polygons(_lib.POLY_TARGETS, [np.asarray(p, dtype=np.float64) for p in polygons])

    def _sum_over_ranks(self, values):
        values = np.asarray(values, dtype=np.int64)
        if self.dist is None or self.world == 1:
            return values
        t = self.torch.tensor(values, dtype=self.torch.int64, device=self.vmax.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return t.cpu().numpy()

    def inside_domain(self):
        """InsideDomain.update on the owned agents -> number of ``active`` flags that changed in the WHOLE crowd"""
        return int(self._sum_over_ranks([self.dev.inside_domain()])[0])

    def target_reached(self):
        """TargetReached.update on the owned agents -> per polygon, agents of the whole crowd that have ever been inside it"""
        return self._sum_over_ranks(self.dev.target_reached(self._n_targets))

    def owned_flags(self, dtype):
        """-> (ids, active, reached_by[np, len(ids)]) of the agents this rank owns now (authoritative entries only)"""
        _, ids = self.export(dtype)
        active = self.dev.get_active()[ids] if getattr(self, '_n_global', 0) else None
        reached = self.dev.target_reached_by(self._n_targets)[:, ids] if getattr(self, '_n_targets', 0) else None
        return ids, active, reached

    def settle(self):
        """Move agents that were generated outside this rank's columns to their owner (one hop), without stepping."""
        if self.skin > 0.0:
            self.dev.set_kind(0)
        self._force_rebuild, self._since = True, 0
        self.phase_begin(flags=0, send_halo=False)
        self.phase_finish(flags=0, use_halo=False)
        self._exchange(self.mig_out, self.mig_in)
        self.phase_absorb(exact=True)

    def export(self, dtype):
        return self.dev.export_agents(dtype)

    # -- construction helpers ---------------------------------------------------------------------------------------------
    @classmethod
    def from_global(cls, agents, obstacles, fields, cell_size, rank, world, device_index=0, dist=None, lattice=None,
                    flags=_lib.STEP_ALL, dt_min=0.01, dt_max=0.01, make_device=None, tensor_device=None, slack=1.5,
                    skin=0.0, max_interval=16):
        """Every rank holds the same global `agents` array and keeps the agents of its own cell columns
        (global id = row index).  Used by the tests: the union over ranks must reproduce the single-GPU result."""
        model = model_of(agents)
        bin_size = cell_size * (1.0 + skin)            # columns of the (possibly widened) strip lattice
        lattice = lattice or lattice_of(agents['position'], bin_size)
        ix_min, iy_min, nx, ny = lattice
        bounds = partition_columns(ix_min, nx, world)
        cols = np.floor(agents['position'][:, 0] / bin_size).astype(np.int64)
        mine = owner_of_columns(cols, bounds) == rank
        ids = np.nonzero(mine)[0].astype(np.int64)
        local = np.ascontiguousarray(agents[mine])
        per_col = max(1, int(np.bincount(np.clip(cols - ix_min, 0, nx - 1), minlength=nx).max()))
        halo_cap = int(slack * per_col) + 64
        mig_cap = halo_cap
        capacity = int(slack * len(local)) + 2 * mig_cap + 1024
        if make_device is None:
            import torch
            dev = CudaStripDevice(model, capacity, device_index, stream=torch.cuda.current_stream().cuda_stream)
            tensor_device = torch.device('cuda', device_index)
        else:
            dev = make_device(model, capacity)
        dev.upload(local, ids)
        dev.set_obstacles(obstacles)
        for t, (mg, uv) in enumerate(fields or ()):
            dev.set_navigation_field(t, mg, uv)
        ext = None                                      # the whole crowd is known to every rank here (k_ext_max's formula)
        if model == MODEL_CIRCULAR and len(agents):
            ext = float(np.max(agents['radius']))
        elif len(agents):
            p = agents['position']
            arm = np.maximum(np.hypot(*(agents['position_ls'] - p).T), np.hypot(*(agents['position_rs'] - p).T))
            arm = np.maximum(arm, np.abs(agents['r_ts']) * (1.0 + 1e-9))
            ext = float(np.max(np.maximum(agents['r_t'], arm + agents['r_s'])))
        return cls(dev, rank, world, bounds, lattice, cell_size, halo_cap, mig_cap, tensor_device, len(local), dist=dist,
                   flags=flags, dt_min=dt_min, dt_max=dt_max, model=model, ext_max=ext, skin=skin, max_interval=max_interval)

    @classmethod
    def synthetic(cls, model, n_per_rank, density, rank, world, device_index, seed=0, cell_size=3.6, dist=None,
                  dt_min=0.01, dt_max=0.01, skin=0.0, max_interval=16):
        """Weak-scaling benchmark crowd: every rank generates its own n_per_rank agents in its own square of a
        world x 1 row of rooms without inner walls (one walled rectangle); strips are aligned to cell columns, so agents a
        rank generated beyond its last column are handed to the neighbour before the first step."""
        import torch
        from . import synthetic as S
        m = int(math.ceil(math.sqrt(n_per_rank)))
        side = m / math.sqrt(density)
        agents, _, _ = S.uniform_crowd(n_per_rank, model, density=density, seed=seed, origin=(rank * side, 0.0))
        obstacles = S.walls_of_box(0.0, 0.0, world * side, side)
        ix_min, iy_min = -1, -1
        bin_size = cell_size * (1.0 + skin)
        nx = int(math.floor(world * side / bin_size)) + 3
        ny = int(math.floor(side / bin_size)) + 3
        # column boundaries follow the rooms: rank g owns the columns whose left edge lies in [g * side, (g + 1) * side)
        bounds = [ix_min] + [int(math.ceil(g * side / bin_size)) for g in range(1, world)] + [ix_min + nx]
        mid = MODEL_CIRCULAR if model == 'circular' else MODEL_THREE_CIRCLE
        per_col = int(1.3 * density * bin_size * side) + 256
        halo_cap, mig_cap = per_col, per_col
        capacity = int(1.05 * n_per_rank) + 4 * mig_cap
        dev = CudaStripDevice(mid, capacity, device_index, stream=torch.cuda.current_stream().cuda_stream)
        ids = np.arange(n_per_rank, dtype=np.int64) + rank * n_per_rank
        dev.upload(agents, ids)
        dev.set_obstacles(obstacles)
        x0 = rank * side
        mg, uv = S.direction_field(1.0, (x0 - 8.0, 0.0, x0 + side + 8.0, side), 'exit', point=(world * side, side / 2))
        dev.set_navigation_field(0, mg, uv)
        sim = cls(dev, rank, world, bounds, (ix_min, iy_min, nx, ny), cell_size, halo_cap, mig_cap,
                  torch.device('cuda', device_index), n_per_rank, dist=dist, dt_min=dt_min, dt_max=dt_max, model=mid,
                  skin=skin, max_interval=max_interval)
        sim.settle()
        return sim


def strong_scaling_strip(model, n_total, density, rank, world, device_index, geometry='room', seed=0, cell_size=3.6, dist=None,
                         dt_min=0.01, dt_max=0.01, skin=0.0, max_interval=16):
    """Strong-scaling benchmark crowd (BASELINE configs 4 and 5): ONE square room of n_total agents split into `world` strips
    of (almost) equal width; every rank generates the lattice columns of its own strip (``synthetic.uniform_slab``).
    geometry: 'room' = four walls (config 5), 'room_exit' = a door in the right wall + exit hall, 11 wall segments (config 4).
    Static exit direction field (step 1 m) over the rank's x range."""
    import torch
    from . import synthetic as S
    m = int(math.ceil(math.sqrt(n_total)))
    side = m / math.sqrt(density)
    ix_min, iy_min = -1, -1
    bin_size = cell_size * (1.0 + skin)
    nx = int(math.floor((side + (5.0 if geometry == 'room_exit' else 0.0)) / bin_size)) + 3
    ny = int(math.floor(side / bin_size)) + 3
    inner = int(math.floor(side / bin_size)) + 1           # columns that hold agents initially
    bounds = [ix_min] + [int(round(inner * g / world)) for g in range(1, world)] + [ix_min + nx]
    x_lo = -np.inf if rank == 0 else bounds[rank] * bin_size
    x_hi = np.inf if rank == world - 1 else bounds[rank + 1] * bin_size
    agents, ids, _ = S.uniform_slab(n_total, model, density=density, seed=seed, x_lo=x_lo, x_hi=x_hi)
    if geometry == 'room_exit':
        obstacles = S.room_exit_walls(side)
        agents['target'] = 0
    else:
        obstacles = S.walls_of_box(0.0, 0.0, side, side)
    mid = MODEL_CIRCULAR if model == 'circular' else MODEL_THREE_CIRCLE
    per_col = int(1.3 * density * bin_size * side) + 256
    halo_cap, mig_cap = per_col, per_col
    capacity = int(1.05 * len(agents)) + 4 * mig_cap
    dev = CudaStripDevice(mid, capacity, device_index, stream=torch.cuda.current_stream().cuda_stream)
    dev.upload(agents, ids)
    dev.set_obstacles(obstacles)
    fx0 = max(0.0, (bounds[rank] if rank else 0) * bin_size - 8.0)
    fx1 = min(side + 5.0, (bounds[rank + 1] if rank < world - 1 else nx) * bin_size + 8.0)
    mg, uv = S.direction_field(1.0, (fx0, 0.0, fx1, side), 'exit', point=(side + 0.15, side / 2))
    dev.set_navigation_field(0, mg, uv)
    sim = StripSimulation(dev, rank, world, bounds, (ix_min, iy_min, nx, ny), cell_size, halo_cap, mig_cap,
                          torch.device('cuda', device_index), len(agents), dist=dist, dt_min=dt_min, dt_max=dt_max, model=mid,
                          skin=skin, max_interval=max_interval)
    sim.settle()
    return sim


class LocalGroup:
    """All strips of a decomposition inside ONE process (message exchange by tensor copies instead of send/recv).
    Functionally identical to one process per GPU; used to validate the strip kernels on a single device."""

    def __init__(self, sims, direct=False):
        self.sims = list(sims)
        if direct:
            # the one-sided exchange inside one process: every strip writes straight into its neighbours' receive buffers
            for s in self.sims:
                s.dev.exchange_alloc()
            for g, s in enumerate(self.sims):
                s.dev.connect_local(self.sims[g - 1].dev if s.left is not None else None,
                                    self.sims[g + 1].dev if s.right is not None else None)
                s.direct = True

    def _exchange(self, out_name, in_name):
        if self.sims[0].direct:
            return
        for g, s in enumerate(self.sims):
            if s.right is not None:
                r = self.sims[g + 1]
                getattr(r, in_name)['l'].copy_(getattr(s, out_name)['r'])
                getattr(s, in_name)['r'].copy_(getattr(r, out_name)['l'])

    def settle(self):
        for s in self.sims:
            if s.skin > 0.0:
                s.dev.set_kind(0)
            s._force_rebuild, s._since = True, 0
            s.phase_begin(flags=0, send_halo=False)
        for s in self.sims:
            s.phase_finish(flags=0, use_halo=False)
        self._exchange('mig_out', 'mig_in')
        for s in self.sims:
            s.phase_absorb(exact=True)

    def step(self, n_steps=1):
        torch = self.sims[0].torch
        for _ in range(n_steps):
            plans = [s.plan_step() for s in self.sims]
            assert len(set(plans)) == 1                # every strip issues the same kind of step
            migrate = plans[0][1]
            for s in self.sims:
                s.phase_begin()
            self._exchange('halo_out', 'halo_in')
            if self.sims[0].adaptive and len(self.sims) > 1:
                for s in self.sims:
                    s.dev.export_vmax(s.vmax)
                # plain maximum, like the all_reduce(MAX) of the multi-process path: NaN (np.max in the reference's
                # adaptive_timestep propagates it) travels as a flag, see k_vmax_export
                m = torch.stack([s.vmax for s in self.sims]).max(0).values
                for s in self.sims:
                    s.vmax.copy_(m)
                    s.dev.import_vmax(s.vmax)
            for s in self.sims:
                s.phase_finish()
            if migrate:
                self._exchange('mig_out', 'mig_in')
                for s in self.sims:
                    s.phase_absorb()
            for s in self.sims:
                s.end_step(migrate)
            if self.sims[0].adapt_due():
                drifts = [s.dev.drift() for s in self.sims]
                for s in self.sims:
                    s.adapt_interval(max(d[0] for d in drifts), min(d[2] for d in drifts))

    def export(self, dtype):
        """-> (agents, ids) of the whole crowd ordered by global id."""
        parts = [s.export(dtype) for s in self.sims]
        agents = np.concatenate([p[0] for p in parts])
        ids = np.concatenate([p[1] for p in parts])
        order = np.argsort(ids, kind='stable')
        return agents[order], ids[order]
