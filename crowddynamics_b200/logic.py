"""Drop-in logic nodes for the replaced sub-tree of ``MultiAgentSimulation`` (reference simulation/logic.py).

Same class names, constructor ``Node(simulation, **params)``, ``update()`` without arguments, composition with ``<<``
and lookup ``tree['Name']`` as the reference (simulation/base.py:27-84, logic.py:31-54).  The reference's base classes
need traitlets + anytree, which are not installed in the build image; when they are importable (a real crowddynamics
install) our nodes subclass the reference ``LogicNode`` so they can be mixed with its other nodes, otherwise a small
duck-typed tree (``LogicNodeBase`` below) provides the same surface.

State handling: all nodes of one simulation share a ``DeviceState`` that mirrors
``simulation.agents.array`` on the GPU.

* ``mode='strict'``   -- each ``update()`` uploads the DIRTY fields of the host array (everything the first time; afterwards
  only what host-side nodes declared with ``mark_dirty`` / ``HostNode(writes=...)``), runs its kernel(s) and downloads the
  fields the node writes, field-granular over PCIe: the reference's contract (the host array is always coherent; host-side
  nodes such as ``InsideDomain`` / ``SaveSimulationData`` can sit anywhere in the tree).  ``upload='always'`` re-sends the
  whole array before every node, for trees that mix in host code which does not declare its writes.
* ``mode='resident'`` -- the array is uploaded once; nodes only launch kernels; ``DeviceState.sync_host()`` refreshes
  the host array on demand (call it before any host-side node reads agents).  ``FusedStep`` runs the whole replaced
  sub-tree for many iterations in one C-ABI call.
"""
import numpy as np

from . import _lib
from .engine import DeviceAgents
from .exceptions import InvalidValue, CrowdDynamicsException
from .structures import is_model, model_of, as_obstacles, MODEL_THREE_CIRCLE

try:  # pragma: no cover - reference package not installable here
    from crowddynamics.simulation.logic import LogicNode as _RefLogicNode
    from crowddynamics.core.geometry import geom_to_linear_obstacles as _geom_to_linear_obstacles
    _HAVE_REFERENCE = True
except Exception:  # noqa
    _RefLogicNode = None
    _geom_to_linear_obstacles = None
    _HAVE_REFERENCE = False


class LogicNodeBase:
    """Minimal stand-in for reference simulation/base.py:27-84 (anytree NodeMixin + traitlets ``name``), used only where the
    reference package (hence its real ``LogicNode``) is not importable.  It restates that interface on purpose --
    ``inject_before / inject_after / add_children``, ``<<`` composition, ``tree['Name']`` lookup and its ``KeyError`` text
    are the reference's, method for method, so that node trees written for the reference run unchanged; there is no logic
    of its own here."""

    def __init__(self, *args, **kwargs):
        self.name = kwargs.pop('name', self.__class__.__name__)
        self._parent = None
        self.children = ()

    # -- tree ------------------------------------------------------------------------------------------------------
    @property
    def parent(self):
        return self._parent

    @parent.setter
    def parent(self, node):
        if self._parent is not None:
            self._parent.children = tuple(c for c in self._parent.children if c is not self)
        self._parent = node
        if node is not None:
            node.children = node.children + (self,)

    @property
    def root(self):
        node = self
        while node._parent is not None:
            node = node._parent
        return node

    def inject_before(self, node):
        parent = self.parent
        self.parent = node
        node.parent = parent

    def inject_after(self, node):
        for child in self.children:
            child.parent = node
        node.parent = self

    def add_children(self, node):
        node.parent = self
        return self

    def __lshift__(self, other):
        if isinstance(other, LogicNodeBase) or hasattr(other, 'update'):
            self.add_children(other)
        else:
            for _other in other:
                self.add_children(_other)
        return self

    def __repr__(self):
        return self.name

    def __getitem__(self, item):
        for node in pre_order_iter(self.root):
            if node.name == item:
                return node
        raise KeyError('Key: "{}" not in the tree.'.format(item))

    def update(self):
        raise NotImplementedError


def pre_order_iter(node):
    yield node
    for c in node.children:
        yield from pre_order_iter(c)


def post_order_iter(node):
    for c in node.children:
        yield from post_order_iter(c)
    yield node


class _Params:
    """Poor man's traits: class-level defaults overridable by constructor keywords (reference nodes use traitlets)."""
    _params = {}

    def _init_params(self, kwargs):
        for key, default in self._params.items():
            setattr(self, key, kwargs.pop(key, default))


class DeviceState:
    """GPU mirror of ``simulation.agents.array`` shared by all nodes of one simulation.

    Field-granular bookkeeping (masks of ``_lib.F_*`` bits over the mutable fields):

    * ``host_dirty`` -- fields the HOST changed since the device last saw them.  The first node uploads the whole array
      (constants included; they then stay on the device, keyed on the identity of the host array); afterwards ``begin()``
      uploads only ``host_dirty``.  Host-side nodes declare their writes with ``mark_dirty(mask)`` (see ``HostNode``);
      ``upload='always'`` restores the conservative behaviour (whole array before every node) for trees that mix in host
      code which does not declare anything.
    * ``dev_ahead`` -- fields the DEVICE changed that the host array has not received yet.  Strict mode downloads what a
      node wrote right away (``end(mask)``), resident mode accumulates it until ``sync_host()``.
    """

    def __init__(self, simulation, mode='strict', device=0, upload='dirty'):
        assert mode in ('strict', 'resident') and upload in ('dirty', 'always')
        self.simulation = simulation
        self.mode = mode
        self.upload_policy = upload
        self.device = device
        self.dev = None
        self._host_id = None
        self._obstacles_key = None
        self._nav_key = None
        self.host_dirty = 0
        self.dev_ahead = 0
        self.dirty_states = False   # device is ahead of the States fields target / is_follower / index_leader
        self.dirty_active = False   # ... and of States.active
        self._active_id = None
        self._states_id = None
        self._registered = None
        self.pending_scalars = set()   # deferred nodes waiting for this update's scalars (ScalarsSync)

    @property
    def dirty_host(self):
        """device is ahead of the host array (some mutable field has not been downloaded yet)"""
        return bool(self.dev_ahead)

    @dirty_host.setter
    def dirty_host(self, value):
        self.dev_ahead = _lib.F_ALL_MUTABLE if value else 0

    @classmethod
    def of(cls, simulation, mode=None, device=0, upload=None):
        st = getattr(simulation, '_b200_state', None)
        if st is None:
            st = cls(simulation, mode or 'strict', device, upload or 'dirty')
            try:
                simulation._b200_state = st
            except Exception:  # traitlets HasTraits accept new attributes; be defensive anyway
                object.__setattr__(simulation, '_b200_state', st)
        else:
            if mode is not None and mode != st.mode:
                st.sync_host()
                st.mode = mode
            if upload is not None:
                st.upload_policy = upload
        return st

    @property
    def agents(self):
        return self.simulation.agents.array

    def _ensure(self):
        agents = self.agents
        model = model_of(agents)      # raises InvalidType for an unknown dtype, like interactions.py:204-205
        if self.dev is None or self.dev.model != model:
            self.dev = DeviceAgents(model, capacity=len(agents), device=self.device)
            # nothing the old device held is on the new one: walls, fields, seed, States, flags are all sent again
            self._host_id = None
            self._obstacles_key = None
            self._nav_key = None
            self._states_id = None
            self._active_id = None
            self._registered = None
            self._seed = None
            self.host_dirty = 0
            self.dev_ahead = 0
            self.dirty_states = self.dirty_active = False
        return agents

    def begin(self):
        """Make the device state current before a node runs."""
        agents = self._ensure()
        key = (id(agents), len(agents))
        if self._host_id != key or (self.mode == 'strict' and self.upload_policy == 'always'):
            self.dev.upload(agents)
            if self._host_id != key and self.upload_policy == 'dirty' and len(agents) and hasattr(self.dev, 'host_register'):
                # pinned + mapped: field-masked transfers then run as zero-copy kernels over PCIe
                try:
                    self.dev.host_register(agents)
                    self._registered = key
                except CrowdDynamicsException:
                    self._registered = None      # e.g. memory that cannot be pinned: the bounce-buffer path still works
            self._host_id = key
            self.host_dirty = 0
            self.dev_ahead = 0
        elif self.host_dirty:
            self.dev.upload_fields(agents, self.host_dirty)
            self.host_dirty = 0
        return self.dev

    def begin_states(self):
        """``begin()`` for the collective-motion nodes, which also read the States fields (agents.py:33-60)."""
        dev = self.begin()
        agents = self.agents
        if (self.mode == 'strict' and self.upload_policy == 'always') or self._states_id != self._host_id:
            dev.set_states(agents, target=False)      # target travels with the records
            self._states_id = self._host_id
        return dev

    def end(self, mask, states=False):
        """Publish what a node wrote (``states``: also target / is_follower / index_leader)."""
        if self.mode == 'strict':
            if mask:
                self.dev.download(self.agents, mask)
            if states:
                self.dev.get_states(self.agents)
        else:
            self.dev_ahead |= int(mask)
            self.dirty_states = self.dirty_states or states

    def sync_host(self, mask=_lib.F_ALL_MUTABLE):
        """Resident mode: refresh the host array from the device (before host-side nodes / IO read it)."""
        pending = self.dev_ahead & int(mask)
        if self.dev is not None and pending:
            self.dev.download(self.agents, pending)
            self.dev_ahead &= ~pending
        if self.dev is not None and self.dirty_states:
            self.dev.get_states(self.agents)
            self.dirty_states = False
        if self.dev is not None and self.dirty_active:
            self.agents['active'] = self.dev.get_active()
            self.dirty_active = False

    def mark_dirty(self, mask=_lib.F_ALL_MUTABLE):
        """A host-side node wrote the mutable fields in ``mask`` of the host array: the next GPU node uploads them (and only
        them).  The host's values win over device-side changes of the same fields that were never downloaded."""
        mask = int(mask) & _lib.F_ALL_MUTABLE
        self.host_dirty |= mask
        self.dev_ahead &= ~mask

    def invalidate(self, mask=None):
        """The host array was modified by host code.  ``mask``: the mutable fields that were edited (same as
        ``mark_dirty``); ``None``: anything may have changed, constants and States fields included -- the whole array is
        uploaded again at the next node.  That discards whatever the device has not published yet, so it is refused while the
        device is ahead: call ``sync_host()`` BEFORE editing the array (downloading now would overwrite the edits)."""
        if mask is not None:
            return self.mark_dirty(mask)
        if self.dev_ahead or self.dirty_states or self.dirty_active:
            raise CrowdDynamicsException(
                'invalidate(): the device holds results that were never downloaded (fields 0x%x%s%s); call sync_host() '
                'before editing the host array, or pass the mask of the edited fields'
                % (self.dev_ahead, ', States' if self.dirty_states else '', ', active' if self.dirty_active else ''))
        self._host_id = None
        self._states_id = None
        self._active_id = None
        self.host_dirty = 0

    # -- geometry / fields -------------------------------------------------------------------------------------------
    def ensure_obstacles(self):
        field = getattr(self.simulation, 'field', None)
        obstacles = getattr(field, 'obstacles', None) if field is not None else None
        if obstacles is None:
            seg = np.zeros((0, 4))
        elif isinstance(obstacles, np.ndarray):
            seg = as_obstacles(obstacles)
        elif _geom_to_linear_obstacles is not None:      # shapely geometry, as logic.py:126-129
            seg = as_obstacles(_geom_to_linear_obstacles(obstacles))
        else:
            raise TypeError('field.obstacles must be an obstacle_type_linear array (shapely is not available)')
        key = seg.tobytes()
        if key != self._obstacles_key:
            self.dev.set_obstacles(seg)
            self._obstacles_key = key

    def ensure_navigation(self, step, radius, strength):
        field = self.simulation.field
        n_targets = len(field.targets)
        key = (id(field), n_targets, step, radius, strength)
        if key != self._nav_key:
            self.dev.clear_navigation()
            geometry = getattr(field, 'navigation_geometry', None)
            for target in range(n_targets):
                if geometry is not None:
                    # line-segment geometry: the field is BUILT on the device (eikonal solver + gradient + wall blend,
                    # csrc/field_kernels.cuh) -- what Field.navigation_to_target does on the host with skfmm (field.py:155-164)
                    tseg, oseg, bounds = geometry(target)
                    self.dev.build_navigation_field(target, tseg, oseg, bounds, step, radius, strength)
                else:
                    mgrid, distance_map, direction_map = field.navigation_to_target(target, step, radius, strength)
                    self.dev.set_navigation_field(target, mgrid, direction_map)
            self._nav_key = key


_Base = _RefLogicNode if _HAVE_REFERENCE else LogicNodeBase


class LogicNode(_Base, _Params):
    """Base of the GPU nodes: ``Node(simulation, mode='strict'|'resident', device=0, **params)``."""

    def __init__(self, simulation, *args, **kwargs):
        mode = kwargs.pop('mode', None)
        device = kwargs.pop('device', 0)
        upload = kwargs.pop('upload', None)
        self._init_params(kwargs)
        if _HAVE_REFERENCE:
            super().__init__(simulation, *args, **kwargs)
        else:
            super().__init__(*args, **kwargs)
            self.simulation = simulation
        self.state = DeviceState.of(simulation, mode, device, upload)

    def update(self):
        raise NotImplementedError


class HostNode(LogicNode):
    """Wraps a host-side node of the reference (anything with ``update()``) in a tree of GPU nodes: the host array is made
    current before it runs (``reads``: the mutable fields it looks at; resident mode only downloads those that are pending)
    and the fields it ``writes`` are sent to the device before the next GPU node -- the "upload dirty fields" half of strict
    mode.  Example: ``HostNode(sim, node=reference_fluctuation, reads=0, writes=F_FORCE | F_TORQUE)``."""
    _params = dict(node=None, reads=_lib.F_ALL_MUTABLE, writes=0, name=None)

    def __init__(self, simulation, *args, **kwargs):
        super().__init__(simulation, *args, **kwargs)
        if self.node is None or not hasattr(self.node, 'update'):
            raise InvalidValue('HostNode needs node=<object with update()>')

    def update(self):
        if self.reads:
            self.state.sync_host(self.reads)
        self.node.update()
        if self.writes:
            self.state.mark_dirty(self.writes)


class Reset(LogicNode):
    """logic.py:59-64"""

    def update(self):
        dev = self.state.begin()
        dev.reset()
        self.state.end(_lib.F_FORCE | _lib.F_TORQUE)


class Integrator(LogicNode):
    """logic.py:67-75"""
    _params = dict(dt_min=0.01, dt_max=0.01)

    def update(self):
        dev = self.state.begin()
        dt = dev.integrate(self.dt_min, self.dt_max)
        self.state.end(_lib.F_POSITION | _lib.F_VELOCITY | _lib.F_FORCE_PREV | _lib.F_SHOULDERS | _lib.F_ORIENTATION |
                       _lib.F_ANGULAR_VELOCITY | _lib.F_TORQUE_PREV)
        self.simulation.data['dt'] = dt
        self.simulation.data['time_tot'] += dt


class Fluctuation(LogicNode):
    """logic.py:78-86 -- stochastic force / torque.  The reference draws from numpy's unseeded global RNG; this node uses a
    counter-based Philox generator on the device (``seed`` param), so parity with the reference is distributional:
    |xi| / (m * std_rand_force) ~ TruncNormal[0, 3], direction ~ U(0, 2 pi), eta / (I * std_rand_torque) ~ TruncNormal[-3, 3]."""
    _params = dict(seed=None)

    def update(self):
        dev = self.state.begin()
        if self.seed is not None and getattr(self.state, '_seed', None) != self.seed:
            dev.set_seed(self.seed)
            self.state._seed = self.seed
        dev.fluctuation()
        self.state.end(_lib.F_FORCE | _lib.F_TORQUE)


class Adjusting(LogicNode):
    """logic.py:89-94"""

    def update(self):
        dev = self.state.begin()
        dev.adjust()
        self.state.end(_lib.F_FORCE | _lib.F_TORQUE)


class AgentAgentInteractions(LogicNode):
    """logic.py:97-119.  As in the reference, ``sight_soc`` / ``f_soc_max`` only size the cells; the kernels use the
    module constants SIGTH_SOC = 3.0 (interactions.py:45) and F_SOC_MAX = 2e3 (power_law.py:52)."""
    _params = dict(sight_soc=3.0, max_agent_radius=0.3, f_soc_max=2e3, cell_size=None)

    def __init__(self, simulation, *args, **kwargs):
        super().__init__(simulation, *args, **kwargs)
        if self.cell_size is None:
            self.cell_size = self.sight_soc + 2 * self.max_agent_radius

    def update(self):
        dev = self.state.begin()
        dev.agent_agent(self.cell_size)
        self.state.end(_lib.F_FORCE | _lib.F_TORQUE)


class AgentObstacleInteractions(LogicNode):
    """logic.py:122-130"""

    def update(self):
        dev = self.state.begin()
        self.state.ensure_obstacles()
        dev.agent_obstacle()
        self.state.end(_lib.F_FORCE | _lib.F_TORQUE)


class Navigation(LogicNode):
    """logic.py:135-165 (sampling of the static direction field; the field itself is built by the host ``Field``)."""
    _params = dict(step=0.1, radius=0.5, strength=0.3)

    def update(self):
        dev = self.state.begin()
        self.state.ensure_navigation(self.step, self.radius, self.strength)
        dev.navigation()
        self.state.end(_lib.F_TARGET_DIRECTION)


class Orientation(LogicNode):
    """logic.py:258-261"""

    def update(self):
        if is_model(self.simulation.agents.array, 'three_circle'):
            dev = self.state.begin()
            dev.orientation()
            self.state.end(_lib.F_TARGET_ORIENTATION)


class ExitDetection(LogicNode):
    """logic.py:237-256 -- herding agents detect an exit within ``detection_range`` that is in their line of sight.
    ``center_door``: (n_doors, 2) door centres; default: the mean of every ``field.targets`` geometry, as in the reference."""
    _params = dict(detection_range=20.0, center_door=None)

    def _doors(self):
        if self.center_door is not None:
            return np.asarray(self.center_door, dtype=np.float64).reshape(-1, 2)
        doors = np.asarray([np.mean(np.asarray(target, dtype=np.float64), axis=0) for target in self.simulation.field.targets])
        if doors.ndim != 2 or doors.shape[1] != 2:
            raise InvalidValue('ExitDetection needs door centres: field.targets must be geometries (coordinate arrays), '
                               'or pass center_door=(n_doors, 2)')
        return doors

    def update(self):
        dev = self.state.begin_states()
        self.state.ensure_obstacles()
        dev.exit_detection(self._doors(), self.detection_range, apply=True)
        self.state.end(0, states=True)


class LeaderFollower(LogicNode):
    """logic.py:168-182"""
    _params = dict(sight=20.0)

    def update(self):
        dev = self.state.begin_states()
        self.state.ensure_obstacles()
        dev.leader_follower(self.sight)
        self.state.end(_lib.F_TARGET_DIRECTION, states=True)


class LeaderFollowerWithHerding(LogicNode):
    """logic.py:185-221"""
    _params = dict(sight_follower=10.0, size_nearest_other=5)

    def update(self):
        dev = self.state.begin_states()
        self.state.ensure_obstacles()
        dev.leader_follower_with_herding(self.sight_follower, self.size_nearest_other)
        self.state.end(_lib.F_TARGET_DIRECTION, states=True)


def _exterior(geom):
    """Vertices of a polygon given as an (nv, 2) array or as a shapely Polygon (``np.asarray(geom.exterior)``, logic.py:349)."""
    if geom is None:
        raise InvalidValue('a polygon is required (field.domain is None?)')
    ext = getattr(geom, 'exterior', geom)
    coords = getattr(ext, 'coords', ext)
    v = np.asarray(coords, dtype=np.float64)
    if v.ndim != 2 or v.shape[1] != 2 or len(v) < 3:
        raise InvalidValue('a polygon is (nv >= 3, 2) vertices, got shape %s' % (v.shape,))
    return v


class InsideDomain(LogicNode):
    """logic.py:343-357 -- sets agents not inside the domain inactive.  ``domain``: (nv, 2) vertices or a shapely Polygon;
    default ``simulation.field.domain``."""
    _params = dict(domain=None, deferred=False)

    def __init__(self, simulation, *args, **kwargs):
        super().__init__(simulation, *args, **kwargs)
        self.simulation.data['inactive'] = 0
        self._vertices = _exterior(self.domain if self.domain is not None else self.simulation.field.domain)
        self._sent_to = None

    def update(self):
        dev = self.state.begin()
        if self._sent_to is not dev:
            dev.set_polygons(_lib.POLY_DOMAIN, [self._vertices])
            self._sent_to = dev
        agents = self.state.agents
        if (self.state.mode == 'strict' and self.state.upload_policy == 'always') or self.state._active_id != self.state._host_id:
            dev.set_active(agents['active'])
            self.state._active_id = self.state._host_id
        if self.deferred and self.state.mode == 'resident':
            # no read-back now: the change count of THIS update is collected by the tree's scalar snapshot one update later
            dev.inside_domain(want_count=False)
            self.state.pending_scalars.add(self)
            self.state.dirty_active = True
            return
        self.simulation.data['inactive'] += dev.inside_domain()
        if self.state.mode == 'strict':
            agents['active'] = dev.get_active()
        else:
            self.state.dirty_active = True

    def collect(self, scalars):
        """deferred mode: ``scalars`` = (dt, time_tot, inside changes, target counts) of the update before"""
        self.simulation.data['inactive'] += int(scalars[2])


class TargetReached(LogicNode):
    """logic.py:360-387 -- counts, per polygon target, the agents that have been inside it at any update so far
    (``simulation.data['target_<i>']``).  ``polygons``: list of (nv, 2) vertex arrays / shapely Polygons, or None for an index
    to be skipped; default ``simulation.field.targets`` (non-polygon targets are skipped, as in the reference)."""
    prefix = 'target_{index}'
    _params = dict(polygons=None, deferred=False)

    def __init__(self, simulation, *args, **kwargs):
        super().__init__(simulation, *args, **kwargs)
        targets = self.polygons if self.polygons is not None else self.simulation.field.targets
        self.names, self._polys = [], []
        for i, target in enumerate(targets):
            if target is None or not (hasattr(target, 'exterior') or isinstance(target, (np.ndarray, list, tuple))):
                continue
            try:
                vertices = _exterior(target)
            except (InvalidValue, ValueError, TypeError):
                continue                       # a line / point target: "we can only measure polygon targets" (logic.py:373)
            name = self.prefix.format(index=i)
            self.names.append(name)
            self._polys.append(vertices)
            self.simulation.data[name] = 0
        self._sent_to = None

    @property
    def reached_by(self):
        """list over the measured targets of bool arrays (the reference's ``reached_by``)."""
        return list(self.state.dev.target_reached_by(len(self._polys))) if self._sent_to is not None else []

    def update(self):
        dev = self.state.begin()
        if self._sent_to is not dev:
            dev.set_polygons(_lib.POLY_TARGETS, self._polys)
            self._sent_to = dev
        if self.deferred and self.state.mode == 'resident':
            dev.target_reached(len(self._polys), want_counts=False)
            self.state.pending_scalars.add(self)
            return
        for name, count in zip(self.names, dev.target_reached(len(self._polys))):
            self.simulation.data[name] = int(count)

    def collect(self, scalars):
        for name, count in zip(self.names, scalars[3]):
            self.simulation.data[name] = int(count)


class FusedStep(LogicNode):
    """The whole replaced sub-tree in one C-ABI call per ``update()``: navigation -> orientation -> adjusting ->
    agent-agent -> agent-obstacle -> integrator -> reset (post-order of examples/simulations.py:123-136), resident on
    the device.  ``steps_per_update`` > 1 advances several iterations per call."""
    _params = dict(dt_min=0.01, dt_max=0.01, cell_size=3.6, step=0.1, radius=0.5, strength=0.3, steps_per_update=1,
                   navigation=True, fluctuation=False, seed=None, sync_every_update=False, deferred=False)

    def __init__(self, simulation, *args, **kwargs):
        kwargs.setdefault('mode', 'resident')
        super().__init__(simulation, *args, **kwargs)

    def update(self):
        dev = self.state.begin()
        self.state.ensure_obstacles()
        flags = _lib.STEP_ALL
        field = getattr(self.simulation, 'field', None)
        if self.navigation and field is not None and getattr(field, 'targets', None) is not None and len(field.targets):
            self.state.ensure_navigation(self.step, self.radius, self.strength)
        else:
            flags &= ~_lib.STEP_NAVIGATION
        if self.fluctuation:
            flags |= _lib.STEP_FLUCTUATION
            if self.seed is not None and getattr(self.state, '_seed', None) != self.seed:
                dev.set_seed(self.seed)
                self.state._seed = self.seed
        if self.deferred and self.state.mode == 'resident':
            # nothing is read back here: simulation.data['dt'] / ['time_tot'] follow one update later (collect), the
            # library does not wait for its own bookkeeping either (cdb_set_deferred_sync)
            if not getattr(self.state, '_deferred_on', False):
                dev.set_deferred_sync(True)
                self.state._deferred_on = True
            dev.step(self.steps_per_update, flags, self.cell_size, self.dt_min, self.dt_max, want_dt=False)
            self.state.dev_ahead |= _lib.F_ALL_MUTABLE
            self.state.pending_scalars.add(self)
            return
        dts = dev.step(self.steps_per_update, flags, self.cell_size, self.dt_min, self.dt_max)
        self.state.dev_ahead |= _lib.F_ALL_MUTABLE
        if len(dts):
            self.simulation.data['dt'] = float(dts[-1])
            self.simulation.data['time_tot'] += float(dts.sum())
        if self.sync_every_update or self.state.mode == 'strict':
            self.state.sync_host()


    def collect(self, scalars):
        self.simulation.data['dt'] = float(scalars[0])
        self.simulation.data['time_tot'] = float(scalars[1])


class ScalarsSync(LogicNode):
    """Last node of a resident tree whose nodes run ``deferred``: queues ONE small asynchronous copy of this update's
    scalars (dt, time_tot, InsideDomain's change count, TargetReached's counts) and hands the copy queued by the PREVIOUS
    update -- complete by now -- to the nodes that asked for it.  ``simulation.data`` therefore trails the device by one
    update and no node ever waits; ``flush()`` (or ``DeviceState.sync_host()``) settles the last update."""

    def update(self):
        st = self.state
        if st.dev is None:
            return
        prev = getattr(st, '_scalar_slot', None)
        if prev is not None:
            self._deliver(prev)
        st._scalar_slot = (st.dev.scalars_begin(), tuple(st.pending_scalars))
        st.pending_scalars.clear()

    def _deliver(self, prev):
        slot, nodes = prev
        n_targets = max([len(getattr(n, '_polys', ())) for n in nodes] + [0])
        scalars = self.state.dev.scalars_wait(slot, n_targets)
        for node in nodes:
            node.collect(scalars)

    def flush(self):
        prev = getattr(self.state, '_scalar_slot', None)
        if prev is not None:
            self._deliver(prev)
            self.state._scalar_slot = None


class SaveSimulationData(LogicNode):
    """logic.py:266-337 (+ io.py:19-45 ``save_npy``): every update the agents array goes into a buffer, and when
    ``save_condition(simulation)`` holds the buffer is stacked and written to ``<directory>/agents_<index>.npy``.

    The reference appends ``simulation.agents.array`` itself (a reference to the one live array, so a dumped file holds
    copies of the LAST state -- evidently a bug); the evident intent, one record array per update, is what is stored here.
    Resident mode: the array is not downloaded synchronously.  Each update queues an asynchronous snapshot of the whole
    packed records (``cdb_snapshot_begin``: device-side copy, then D2H on a side stream into one of two pinned buffers) and
    collects the snapshot queued one update earlier, which has arrived by then; only an update that dumps waits for its own
    snapshot.  Strict mode (host array always coherent) copies the host array like the reference."""
    _params = dict(save_condition=None, base_directory='.', save_directory='simulation', basename='agents')

    def __init__(self, simulation, *args, **kwargs):
        super().__init__(simulation, *args, **kwargs)
        import os
        self.full_path = os.path.join(os.path.abspath(self.base_directory), self.save_directory)
        os.makedirs(self.full_path, exist_ok=True)
        self.buffer = []
        self.index = 0
        self._slot = None
        self.files = []

    def _collect(self):
        if self._slot is not None:
            view = self.state.dev.snapshot_wait(self._slot, self.state.agents.dtype)
            self.buffer.append(np.array(view))         # out of the pinned slot before it is reused
            self._slot = None

    def update(self):
        import os
        save = bool(self.save_condition(self.simulation)) if self.save_condition is not None else False
        if self.state.mode == 'resident' and self.state.dev is not None:
            self._collect()
            self._slot = self.state.dev.snapshot_begin()
            if save:
                self._collect()
        else:
            self.buffer.append(self.state.agents.copy())
        if save and self.buffer:
            path = os.path.join(self.full_path, '%s_%d.npy' % (self.basename, self.index))
            np.save(path, np.vstack(self.buffer))
            self.files.append(path)
            self.buffer.clear()
            self.index += 1

    def flush(self):
        """collect the snapshot still in flight (end of the run)"""
        if self.state.mode == 'resident' and self.state.dev is not None:
            self._collect()


class MultiAgentSimulation:
    """Minimal host with the surface the nodes use (reference simulation/multiagent.py:23-55): ``agents.array``,
    ``field`` (``obstacles``, ``targets``, ``navigation_to_target``), ``logic``, ``data`` and ``update()`` =
    post-order traversal of the logic tree."""

    class _Agents:
        def __init__(self, array):
            self.array = array

    class ArrayField:
        """Field given directly as arrays: obstacle segments + per-target (mgrid, (U, V))."""

        def __init__(self, obstacles=None, fields=(), domain=None):
            self.obstacles = obstacles
            self.targets = list(range(len(fields)))
            self._fields = list(fields)
            self.domain = domain          # (nv, 2) vertices of the domain polygon (InsideDomain)

        def navigation_to_target(self, target, step, radius, strength):
            mgrid, direction_map = self._fields[target]
            return mgrid, None, direction_map

    class GeometryField:
        """Field given as geometry, like the reference's ``Field`` (simulation/field.py): obstacle segments, one list of
        target line segments per target (doors), the bounds of the domain.  The navigation fields are built on the device."""

        def __init__(self, obstacles, targets, bounds, domain=None):
            self.obstacles = obstacles
            self.targets = [np.asarray(t, dtype=np.float64).reshape(-1, 4) for t in targets]
            self.bounds = tuple(float(b) for b in bounds)
            self.domain = domain

        def navigation_geometry(self, target):
            return self.targets[target], self.obstacles, self.bounds

    def __init__(self, agents, obstacles=None, fields=(), logic=None, domain=None, field=None):
        self.agents = self._Agents(agents)
        self.field = field if field is not None else self.ArrayField(obstacles, fields, domain)
        self.logic = logic
        self.data = {'iterations': 0, 'time_tot': 0.0, 'dt': 0.0}

    def update(self):
        for node in post_order_iter(self.logic.root):
            node.update()
        self.data['iterations'] += 1


def hallway_logic(simulation, mode='strict', dt_min=0.01, dt_max=0.01, cell_size=None, fluctuation=False, seed=None):
    """The replaced part of the Hallway / RoomWithOneExit logic tree (examples/simulations.py:123-136); InsideDomain stays a
    host-side node of the reference, Fluctuation is optional (stochastic)."""
    aa = dict(cell_size=cell_size) if cell_size else {}
    fl = (Fluctuation(simulation, seed=seed),) if fluctuation else ()
    return Reset(simulation, mode=mode) << (
        Integrator(simulation, dt_min=dt_min, dt_max=dt_max) << fl + (
            Adjusting(simulation) << (Navigation(simulation), Orientation(simulation)),
            AgentAgentInteractions(simulation, **aa),
            AgentObstacleInteractions(simulation)))
