"""TEST INFRASTRUCTURE ONLY -- loads the *unmodified* reference hot-path modules from /root/reference.

Used in the build container (where /root/reference exists) to
  * pin the C restatement in ``oracle/crowd_oracle.c`` against the reference's own numba code, and
  * generate the golden vectors committed under ``tests/golden/`` (see ``tests/golden/generate.py``).

It never runs on the GPU box (the reference tree is not there) and is never imported by the
product package.  No reference source is copied: modules are imported from where they lie, and the
three functions that live in modules we cannot import as a whole (``shoulders`` in
simulation/agents.py, ``getdefault``/``is_inside`` in core/steering/navigation.py, ``meshgrid`` in
core/steering/quickest_path.py -- their modules import traitlets/shapely/skfmm, which are absent)
are pulled out of the reference files at run time with ``ast`` and exec'd unchanged.

Shims:
  1. stub packages whose ``__path__`` points into /root/reference (skips crowddynamics/__init__.py -> versioneer)
  2. ``numba.generated_jit`` (removed from numba >= 0.59; only ``vector2D.unit_vector`` uses it)
  3. ``crowddynamics.exceptions`` (the real one imports traitlets)
  4. ``crowddynamics.simulation.agents`` exposing the two structured dtypes (hand-built from
     agents.py:447-457 + traits.py:158-219, itemsize asserted 228/316), ``is_model`` and ``shoulders``
  5. ``cell_lists`` -- third-party, unpinned, not installed: restated below per the in-tree spec
     crowddynamics/core/block_list.py:28-52 ("parity unpinned" at that boundary).
  6. ``numba.typing.typeof`` (moved to numba.core.typing in numba >= 0.49; imported by sensory_region.py,
     collective_motion.py and evacuation.py for ``typeof(obstacle_type_linear)``)
"""
import ast
import importlib
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get('CROWD_REFERENCE_ROOT', '/root/reference')


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'crowddynamics', 'core'))


# ---------------------------------------------------------------------------------------------
# dtypes (data contract a1/a2 of SURVEY section 8)
# ---------------------------------------------------------------------------------------------
_STATES = [('active', np.bool_), ('target_reached', np.bool_), ('target', np.int64),
           ('is_leader', np.bool_), ('is_follower', np.bool_), ('index_leader', np.int64),
           ('familiar_exit', np.int64)]
_BODY = [(n, np.float64) for n in ('radius', 'r_t', 'r_s', 'r_ts', 'mass', 'inertia_rot',
                                   'target_velocity', 'target_angular_velocity')]
_TRANS = [(n, np.float64, (2,)) for n in ('position', 'velocity', 'target_direction', 'force', 'force_prev')] + \
         [(n, np.float64) for n in ('tau_adj', 'k_soc', 'tau_0', 'mu', 'kappa', 'damping', 'std_rand_force')]
_ROT = [(n, np.float64) for n in ('orientation', 'angular_velocity', 'target_orientation', 'torque',
                                  'torque_prev', 'tau_rot', 'std_rand_torque')]
_THREE = [('position_ls', np.float64, (2,)), ('position_rs', np.float64, (2,))]

agent_type_circular = np.dtype(_STATES + _BODY + _TRANS)
agent_type_three_circle = np.dtype(_THREE + _STATES + _BODY + _TRANS + _ROT)
assert agent_type_circular.itemsize == 228 and agent_type_three_circle.itemsize == 316
obstacle_type_linear = np.dtype([('p0', np.float64, (2,)), ('p1', np.float64, (2,))])


def _extract(path, names):
    """Return the unmodified source of top-level definitions ``names`` from a reference file."""
    with open(path) as f:
        src = f.read()
    tree = ast.parse(src)
    out = []
    for node in tree.body:
        name = getattr(node, 'name', None)
        if name is None and isinstance(node, ast.Assign) and isinstance(node.targets[0], ast.Name):
            name = node.targets[0].id
        if name in names:
            out.append(ast.get_source_segment(src, node) if not getattr(node, 'decorator_list', None)
                       else '\n'.join(src.splitlines()[node.decorator_list[0].lineno - 1:node.end_lineno]))
    assert len(out) == len(names), (path, names)
    return '\n\n'.join(out)


# ---------------------------------------------------------------------------------------------
# cell_lists restatement (spec: crowddynamics/core/block_list.py:28-52; call sites interactions.py:191-205)
# ---------------------------------------------------------------------------------------------
def _make_cell_lists():
    import numba
    from numba import f8, i8

    mod = types.ModuleType('cell_lists')

    @numba.jit(nopython=True, nogil=True)
    def add_to_cells(points, cell_size):
        n = points.shape[0]
        ix = np.empty(n, dtype=np.int64)
        iy = np.empty(n, dtype=np.int64)
        if n == 0:
            z = np.zeros(0, dtype=np.int64)
            return z, z.copy(), z.copy(), np.zeros(2, dtype=np.int64)
        for k in range(n):
            ix[k] = np.int64(np.floor(points[k, 0] / cell_size))
            iy[k] = np.int64(np.floor(points[k, 1] / cell_size))
        x_min = ix.min(); y_min = iy.min()
        nx = ix.max() - x_min + 1
        ny = iy.max() - y_min + 1
        grid_shape = np.array((nx, ny), dtype=np.int64)
        ncell = nx * ny
        flat = (ix - x_min) * ny + (iy - y_min)
        cells_count = np.zeros(ncell, dtype=np.int64)
        for k in range(n):
            cells_count[flat[k]] += 1
        cells_offset = np.zeros(ncell, dtype=np.int64)
        acc = 0
        for c in range(ncell):
            cells_offset[c] = acc
            acc += cells_count[c]
        fill = np.zeros(ncell, dtype=np.int64)
        points_indices = np.empty(n, dtype=np.int64)
        for k in range(n):          # stable: ascending agent index inside a cell
            c = flat[k]
            points_indices[cells_offset[c] + fill[c]] = k
            fill[c] += 1
        return points_indices, cells_count, cells_offset, grid_shape

    def neighboring_cells(grid_shape):
        """Forward half stencil, flattened (C, 4): for cell (x, y) the neighbours
        (x, y+1), (x+1, y-1), (x+1, y), (x+1, y+1); -1 where outside the grid."""
        nx, ny = int(grid_shape[0]), int(grid_shape[1])
        out = -np.ones((nx * ny, 4), dtype=np.int64)
        if nx * ny == 0:
            return out.reshape(-1)
        x, y = np.divmod(np.arange(nx * ny, dtype=np.int64), ny)
        for s, (dx, dy) in enumerate(((0, 1), (1, -1), (1, 0), (1, 1))):
            ok = (x + dx < nx) & (y + dy >= 0) & (y + dy < ny)
            out[ok, s] = ((x + dx) * ny + (y + dy))[ok]
        return out.reshape(-1)

    @numba.jit(nopython=True, nogil=True)
    def iter_nearest_neighbors(cell_indices, neigh_cells, points_indices, cells_count, cells_offset):
        for c in cell_indices:
            n_c = cells_count[c]
            o_c = cells_offset[c]
            for a in range(n_c):          # same cell: i < j
                i = points_indices[o_c + a]
                for b in range(a + 1, n_c):
                    yield i, points_indices[o_c + b]
            for s in range(4):            # forward neighbours: i in c, j in the neighbour
                d = neigh_cells[4 * c + s]
                if d < 0:
                    continue
                n_d = cells_count[d]
                o_d = cells_offset[d]
                for a in range(n_c):
                    i = points_indices[o_c + a]
                    for b in range(n_d):
                        yield i, points_indices[o_d + b]

    mod.add_to_cells = add_to_cells
    mod.neighboring_cells = neighboring_cells
    mod.iter_nearest_neighbors = iter_nearest_neighbors
    return mod


_LOADED = None


def load():
    """Import the reference hot-path modules; returns a namespace with the reference callables."""
    global _LOADED
    if _LOADED is not None:
        return _LOADED
    if not available():
        raise RuntimeError('reference tree not present at %s' % REFERENCE_ROOT)
    # The reference kernels are compiled with cache=True and inline *our* cell_lists restatement; key the numba
    # cache on this file's content so that a stale cache built against another stub can never be picked up.
    import hashlib
    import tempfile
    with open(os.path.abspath(__file__), 'rb') as f:
        tag = hashlib.sha1(f.read()).hexdigest()[:12]
    os.environ['NUMBA_CACHE_DIR'] = os.path.join(tempfile.gettempdir(), 'crowd_b200_refharness_' + tag)
    import numba
    from numba.core import config as _nb_config
    _nb_config.CACHE_DIR = os.environ['NUMBA_CACHE_DIR']
    from numba.extending import overload

    root = os.path.join(REFERENCE_ROOT, 'crowddynamics')

    # (2) numba.generated_jit shim: build an @overload-ed function from the type-dispatching body.
    if not hasattr(numba, 'generated_jit'):
        def generated_jit(*dargs, **dkwargs):
            def deco(gen):
                def stub(*a):
                    raise NotImplementedError
                stub.__name__ = gen.__name__
                overload(stub)(gen)

                @numba.jit(nopython=True, nogil=True)
                def entry(x):
                    return stub(x)
                return entry
            return deco
        numba.generated_jit = generated_jit

    # (6) numba.typing.typeof
    if 'numba.typing' not in sys.modules:
        typing_pkg = types.ModuleType('numba.typing')
        typing_pkg.__path__ = []
        typeof_mod = types.ModuleType('numba.typing.typeof')
        typeof_mod.typeof = numba.typeof
        typing_pkg.typeof = typeof_mod
        sys.modules['numba.typing'] = typing_pkg
        sys.modules['numba.typing.typeof'] = typeof_mod

    # (1) stub packages
    for name, sub in (('crowddynamics', ''), ('crowddynamics.core', 'core'),
                      ('crowddynamics.core.motion', 'core/motion'),
                      ('crowddynamics.core.steering', 'core/steering'),
                      ('crowddynamics.simulation', 'simulation')):
        pkg = types.ModuleType(name)
        pkg.__path__ = [os.path.join(root, sub)]
        sys.modules[name] = pkg

    # (3) exceptions
    exc = types.ModuleType('crowddynamics.exceptions')

    class CrowdDynamicsException(Exception):
        pass

    class InvalidType(CrowdDynamicsException, TypeError):
        pass

    class InvalidValue(CrowdDynamicsException, ValueError):
        pass
    exc.CrowdDynamicsException = CrowdDynamicsException
    exc.InvalidType = InvalidType
    exc.InvalidValue = InvalidValue
    sys.modules['crowddynamics.exceptions'] = exc

    # (5) cell_lists
    sys.modules['cell_lists'] = _make_cell_lists()

    # (4) agents stub; ``shoulders`` is exec'd from the reference file after vector2D is importable
    agents_mod = types.ModuleType('crowddynamics.simulation.agents')
    agents_mod.agent_type_circular = agent_type_circular
    agents_mod.agent_type_three_circle = agent_type_three_circle
    agents_mod.AgentModelToType = {'circular': agent_type_circular, 'three_circle': agent_type_three_circle}
    agents_mod.NO_TARGET = -1
    agents_mod.NO_LEADER = -1
    sys.modules['crowddynamics.simulation.agents'] = agents_mod
    vector2D = importlib.import_module('crowddynamics.core.vector2D')
    ns = {'numba': numba, 'np': np, 'void': numba.void, 'typeof': numba.typeof,
          'agent_type_three_circle': agent_type_three_circle,
          'AgentModelToType': agents_mod.AgentModelToType,
          'rotate270': vector2D.rotate270, 'unit_vector': vector2D.unit_vector}
    src = _extract(os.path.join(root, 'simulation', 'agents.py'), ['is_model', 'shoulders'])
    exec(compile(src.replace('cache=True', 'cache=False'), 'reference:simulation/agents.py', 'exec'), ns)
    agents_mod.is_model = ns['is_model']
    agents_mod.shoulders = ns['shoulders']

    R = types.SimpleNamespace()
    R.vector2D = vector2D
    R.structures = importlib.import_module('crowddynamics.core.structures')
    R.distance = importlib.import_module('crowddynamics.core.distance')
    R.contact = importlib.import_module('crowddynamics.core.motion.contact')
    R.power_law = importlib.import_module('crowddynamics.core.motion.power_law')
    R.adjusting = importlib.import_module('crowddynamics.core.motion.adjusting')
    R.integrator = importlib.import_module('crowddynamics.core.integrator')
    R.interactions = importlib.import_module('crowddynamics.core.interactions')
    R.orientation = importlib.import_module('crowddynamics.core.steering.orientation')
    R.geom2D = importlib.import_module('crowddynamics.core.geom2D')
    R.sensory_region = importlib.import_module('crowddynamics.core.sensory_region')
    R.collective_motion = importlib.import_module('crowddynamics.core.steering.collective_motion')
    R.evacuation = importlib.import_module('crowddynamics.core.evacuation')
    R.agents = agents_mod
    R.cell_lists = sys.modules['cell_lists']
    R.exceptions = exc

    # navigation sampling: getdefault/is_inside (navigation.py:60-78) and meshgrid (quickest_path.py:23-51)
    from typing import NamedTuple, Callable
    nav_ns = {'numba': numba, 'np': np, 'f8': numba.f8, 'i8': numba.i8}
    src = _extract(os.path.join(root, 'core', 'steering', 'navigation.py'), ['is_inside', 'getdefault'])
    exec(compile(src.replace('cache=True', 'cache=False'), 'reference:core/steering/navigation.py', 'exec'), nav_ns)
    R.getdefault = nav_ns['getdefault']
    mg_ns = {'np': np, 'NamedTuple': NamedTuple, 'Callable': Callable}
    src = _extract(os.path.join(root, 'core', 'steering', 'quickest_path.py'), ['MeshGrid', 'meshgrid'])
    exec(compile(src, 'reference:core/steering/quickest_path.py', 'exec'), mg_ns)
    R.meshgrid = mg_ns['meshgrid']
    _LOADED = R
    return R


# ---------------------------------------------------------------------------------------------
# The reference logic nodes, restated as the plain function calls their update() bodies make
# (simulation/logic.py:59-165,258-261) -- the node classes themselves need traitlets/anytree.
# ---------------------------------------------------------------------------------------------
def node_reset(R, agents):                      # logic.py:59-64
    agents['force'] = 0
    if R.agents.is_model(agents, 'three_circle'):
        agents['torque'] = 0


def node_navigation(R, agents, fields):         # logic.py:149-165
    """fields: list over targets of (mgrid, (U, V))."""
    for target in range(len(fields)):
        has_target = agents['target'] == target
        if not has_target.size:
            continue
        mgrid, direction_map = fields[target]
        indices = np.fliplr(mgrid.indicer(agents[has_target]['position']))
        new_direction = R.getdefault(np.ascontiguousarray(indices), direction_map,
                                     np.ascontiguousarray(agents[has_target]['target_direction']))
        agents['target_direction'][has_target] = new_direction


def node_orientation(R, agents):                # logic.py:258-261
    if R.agents.is_model(agents, 'three_circle'):
        R.orientation.orient_towards_target_direction(agents)


def node_adjusting(R, agents):                  # logic.py:89-94
    R.adjusting.force_adjust_agents(agents)
    if R.agents.is_model(agents, 'three_circle'):
        R.adjusting.torque_adjust_agents(agents)


def node_agent_agent(R, agents, cell_size):     # logic.py:118-119
    R.interactions.agent_agent_block_list(agents, cell_size)


def node_agent_obstacle(R, agents, obstacles):  # logic.py:122-130
    R.interactions.agent_obstacle(agents, obstacles)


def node_integrator(R, agents, dt_min, dt_max):  # logic.py:71-75
    return R.integrator.velocity_verlet_integrator(agents, dt_min, dt_max)


def step(R, agents, obstacles, fields, cell_size, dt_min, dt_max):
    """One MultiAgentSimulation.update() in the Hallway post-order, Fluctuation/InsideDomain omitted."""
    if fields:
        node_navigation(R, agents, fields)
    node_orientation(R, agents)
    node_adjusting(R, agents)
    node_agent_agent(R, agents, cell_size)
    node_agent_obstacle(R, agents, obstacles)
    dt = node_integrator(R, agents, dt_min, dt_max)
    node_reset(R, agents)
    return dt
