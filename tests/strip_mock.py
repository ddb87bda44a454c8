"""TEST INFRASTRUCTURE: a numpy/oracle stand-in for CudaStripDevice, so that the host-side strip protocol
(crowddynamics_b200.parallel: partitioning, neighbour schedule, message plumbing, dt all-reduce) can be exercised on CPU
with the gloo backend.  Messages are torch CPU float64 tensors: header {count, -, -, -} then `count` records of W doubles
(the packed agent record padded to a multiple of 8 bytes + the global id)."""
import numpy as np

from crowddynamics_b200 import _lib
from oracle import crowd_oracle as O

HEADER = 4


class NumpyStripDevice:
    def __init__(self, model, capacity):
        self.model = model
        self.agents = None
        self.ids = None
        self.obstacles = np.zeros(0)
        self.fields = []

    # -- records <-> messages ---------------------------------------------------------------------------------------------
    def _w(self):
        return (self.agents.dtype.itemsize + 7) // 8 + 1

    def _pack(self, buf, sel):
        w = self._w()
        n = int(sel.sum())
        assert HEADER + n * w <= buf.numel(), 'message capacity exceeded'
        out = buf.numpy()
        out[0] = n
        raw = np.zeros((n, (w - 1) * 8), dtype=np.uint8)
        raw[:, :self.agents.dtype.itemsize] = np.ascontiguousarray(self.agents[sel]).view(np.uint8).reshape(n, self.agents.dtype.itemsize)
        body = out[HEADER:HEADER + n * w].reshape(n, w)
        body[:, :w - 1] = raw.view(np.float64).reshape(n, w - 1)
        body[:, w - 1] = self.ids[sel]

    def _unpack(self, buf):
        w = self._w()
        arr = buf.numpy()
        n = int(arr[0])
        body = arr[HEADER:HEADER + n * w].reshape(n, w)
        raw = np.ascontiguousarray(body[:, :w - 1]).view(np.uint8).reshape(n, (w - 1) * 8)[:, :self.agents.dtype.itemsize]
        recs = np.ascontiguousarray(raw).view(self.agents.dtype).reshape(-1).copy()
        return recs, body[:, w - 1].astype(np.int64)

    # -- the CudaStripDevice surface -----------------------------------------------------------------------------------------
    def upload(self, agents, ids):
        self.agents = agents.copy()
        self.ids = np.asarray(ids, dtype=np.int64).copy()

    def set_strip(self, ix_min, iy_min, nx_owned, ny, has_left, has_right, halo_cap, mig_cap):
        self.col_lo, self.col_hi = ix_min, ix_min + nx_owned - 1
        self.has_left, self.has_right = has_left, has_right
        w = self._w()
        return HEADER + halo_cap * w, HEADER + mig_cap * w

    def set_obstacles(self, obstacles):
        self.obstacles = obstacles

    def set_navigation_field(self, target, mgrid, direction_map):
        assert target == len(self.fields)
        self.fields.append((mgrid, direction_map))

    def _cols(self):
        return np.floor(self.agents['position'][:, 0] / self.cell_size).astype(np.int64)

    def begin(self, flags, cell_size, halo_left_out, halo_right_out):
        self.cell_size = cell_size
        cols = self._cols()
        if halo_left_out is not None:
            self._pack(halo_left_out, cols == self.col_lo)
        if halo_right_out is not None:
            self._pack(halo_right_out, cols == self.col_hi)
        v = np.hypot(*self.agents['velocity'].T) if len(self.agents) else np.zeros(0)
        self._vmax = np.array([v.max() if len(v) else 0.0,
                               self.agents['target_velocity'].max() if len(v) else -np.inf])

    def export_vmax(self, buf):
        # {max |v|, max v0, NaN flags}: NaN travels as a flag because all_reduce(MAX) need not propagate it (k_vmax_export)
        v = self._vmax
        buf.numpy()[:] = [0.0 if np.isnan(v[0]) else v[0], -np.inf if np.isnan(v[1]) else v[1], float(np.isnan(v[0])), float(np.isnan(v[1]))]

    def import_vmax(self, buf):
        b = buf.numpy()
        self._vmax = np.array([np.nan if b[2] > 0 else b[0], np.nan if b[3] > 0 else b[1]])

    def finish(self, flags, dt_min, dt_max, halo_left_in, halo_right_in, mig_left_out, mig_right_out):
        ghosts = [self._unpack(b) for b in (halo_left_in, halo_right_in) if b is not None]
        n_own = len(self.agents)
        allrec = np.concatenate([self.agents] + [g[0] for g in ghosts])
        allid = np.concatenate([self.ids] + [g[1] for g in ghosts])
        order = np.argsort(allid, kind='stable')          # index order == global id order, like the single-domain run
        work = np.ascontiguousarray(allrec[order])
        if flags & _lib.STEP_NAVIGATION and self.fields:
            O.navigation(work, self.fields)
        if flags & _lib.STEP_ORIENTATION:
            O.orientation(work)
        if flags & _lib.STEP_ADJUSTING:
            O.adjusting(work)
        if flags & _lib.STEP_AGENT_AGENT:
            O.agent_agent_block_list(work, self.cell_size)
        if flags & _lib.STEP_AGENT_OBSTACLE:
            O.agent_obstacle(work, self.obstacles)
        back = np.empty_like(work)
        back[order] = work
        own = np.ascontiguousarray(back[:n_own])
        if flags & _lib.STEP_INTEGRATOR:
            v_max, v0_max = self._vmax
            if v_max == 0.0:
                dt = dt_max
            else:
                dt = 1.1 * v0_max * dt_max / v_max
                dt = dt_max if dt > dt_max else (dt_min if dt < dt_min else dt)
            if len(own):
                O.velocity_verlet_integrator(own, dt, dt)
            self.last_dt = dt
        if flags & _lib.STEP_RESET:
            O.reset(own)
        self.agents = own
        cols = self._cols()
        go_l = (cols < self.col_lo) & bool(self.has_left)
        go_r = (cols > self.col_hi) & bool(self.has_right)
        if mig_left_out is not None:
            self._pack(mig_left_out, go_l)
        if mig_right_out is not None:
            self._pack(mig_right_out, go_r)
        keep = ~(go_l | go_r)
        self.agents, self.ids = np.ascontiguousarray(self.agents[keep]), self.ids[keep]

    def absorb(self, mig_left_in, mig_right_in, exact=False):
        for b in (mig_left_in, mig_right_in):
            if b is not None:
                recs, ids = self._unpack(b)
                self.agents = np.concatenate((self.agents, recs))
                self.ids = np.concatenate((self.ids, ids))
        return len(self.agents)

    def export_agents(self, dtype):
        return self.agents.copy(), self.ids.copy()
