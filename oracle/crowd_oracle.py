"""TEST INFRASTRUCTURE ONLY -- ctypes binding of the C oracle (oracle/crowd_oracle.c -> oracle/liboracle.so).

Function names and argument meaning follow the reference (core/interactions.py, core/integrator.py, ...); every
function mutates the structured ``agents`` array in place, like the reference's numba kernels.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class InvalidType(TypeError):
    pass


def build(force=False):
    so = os.path.join(_HERE, 'liboracle.so')
    src = os.path.join(_HERE, 'crowd_oracle.c')
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(['make', '-C', _HERE, '-s', '-B'])
    return so


class _Cells(C.Structure):
    _fields_ = [('n', C.c_int64), ('ncell', C.c_int64), ('grid', C.c_int64 * 4),
                ('cell_of_agent', C.POINTER(C.c_int64)), ('points_indices', C.POINTER(C.c_int64)),
                ('cells_count', C.POINTER(C.c_int64)), ('cells_offset', C.POINTER(C.c_int64))]


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        i64, f64, vp = C.c_int64, C.c_double, C.c_void_p
        L.oracle_wrap_to_pi.restype = f64
        L.oracle_wrap_to_pi.argtypes = [f64]
        L.oracle_force_social_circular.argtypes = [vp, i64, i64, i64, vp, vp]
        L.oracle_force_social_three_circle.argtypes = [vp, i64, i64, vp, vp]
        L.oracle_add_to_cells.argtypes = [vp, i64, i64, f64, C.POINTER(_Cells)]
        L.oracle_free_cells.argtypes = [C.POINTER(_Cells)]
        L.oracle_neighbor_pairs.restype = i64
        L.oracle_neighbor_pairs.argtypes = [vp, i64, i64, f64, vp, vp, i64]
        L.oracle_agent_agent_block_list.argtypes = [vp, i64, i64, f64]
        L.oracle_agent_agent_brute.argtypes = [vp, i64, i64]
        L.oracle_agent_obstacle.argtypes = [vp, i64, i64, vp, i64]
        L.oracle_adjusting.argtypes = [vp, i64, i64]
        L.oracle_orientation.argtypes = [vp, i64, i64]
        L.oracle_navigation.argtypes = [vp, i64, i64, i64, vp, vp, i64, i64, f64, f64, f64]
        L.oracle_adaptive_timestep.restype = f64
        L.oracle_adaptive_timestep.argtypes = [vp, i64, i64, f64, f64]
        L.oracle_velocity_verlet_integrator.argtypes = [vp, i64, i64, f64, f64, C.POINTER(f64)]
        L.oracle_reset.argtypes = [vp, i64, i64]
        L.oracle_step.argtypes = [vp, i64, i64, vp, i64, i64, vp, vp, i64, i64, f64, f64, f64, f64, f64, f64,
                                  C.POINTER(f64)]
        u8p = vp
        L.oracle_line_intersect.argtypes = [vp, vp, vp, vp]
        L.oracle_is_obstacle_between_points.argtypes = [vp, vp, vp, i64]
        L.oracle_exit_detection.argtypes = [vp, i64, vp, i64, i64, vp, i64, f64, vp, u8p]
        L.oracle_herding_relationship.argtypes = [vp, vp, vp, vp, f64]
        L.oracle_find_nearest_neighbors.argtypes = [vp, i64, i64, f64, i64, vp, i64, vp]
        L.oracle_herding_interaction.argtypes = [vp, i64, i64, u8p, vp, i64, f64, f64, vp, u8p]
        L.oracle_leader_follower_interaction_brute.argtypes = [vp, i64, i64, f64, f64, vp, i64, f64, vp, u8p]
        L.oracle_leader_follower_interaction.argtypes = [vp, i64, i64, vp, i64, f64, f64, f64, vp]
        L.oracle_leader_follower_with_herding_interaction.argtypes = [vp, i64, i64, vp, i64, f64, i64, f64, f64, f64, f64, vp]
        L.oracle_point_in_polygon.argtypes = [vp, i64, f64, f64]
        L.oracle_inside_domain.restype = i64
        L.oracle_inside_domain.argtypes = [vp, i64, i64, vp, i64]
        L.oracle_target_reached.restype = i64
        L.oracle_target_reached.argtypes = [vp, i64, i64, vp, i64, vp]
        _LIB = L
    return _LIB


def _a(agents):
    if not (isinstance(agents, np.ndarray) and agents.flags.c_contiguous and agents.dtype.fields is not None):
        raise InvalidType('agents must be a C-contiguous structured array')
    return agents.ctypes.data, len(agents), agents.dtype.itemsize


def _check(rc):
    if rc:
        raise InvalidType('unknown agent dtype')


def wrap_to_pi(x):
    return lib().oracle_wrap_to_pi(float(x))


def force_social_circular(agents, i, j):
    p, n, sz = _a(agents)
    fi, fj = np.zeros(2), np.zeros(2)
    lib().oracle_force_social_circular(p, sz, i, j, fi.ctypes.data, fj.ctypes.data)
    return fi, fj


def force_social_three_circle(agents, i, j):
    p, n, sz = _a(agents)
    assert sz == 316
    fi, fj = np.zeros(2), np.zeros(2)
    lib().oracle_force_social_three_circle(p, i, j, fi.ctypes.data, fj.ctypes.data)
    return fi, fj


def add_to_cells(agents, cell_size):
    """-> dict(cell_of_agent, points_indices, cells_count, cells_offset, grid=(ix_min, iy_min, nx, ny))."""
    p, n, sz = _a(agents)
    cl = _Cells()
    _check(lib().oracle_add_to_cells(p, n, sz, cell_size, C.byref(cl)))
    nc = cl.ncell

    def arr(ptr, m):
        return np.ctypeslib.as_array(ptr, shape=(m,)).copy() if m else np.zeros(0, dtype=np.int64)
    out = dict(cell_of_agent=arr(cl.cell_of_agent, n), points_indices=arr(cl.points_indices, n),
               cells_count=arr(cl.cells_count, nc), cells_offset=arr(cl.cells_offset, nc),
               grid=tuple(int(g) for g in cl.grid))
    lib().oracle_free_cells(C.byref(cl))
    return out


def neighbor_pairs(agents, cell_size):
    """All candidate (i, j) pairs in generator order, shape (P, 2)."""
    p, n, sz = _a(agents)
    cnt = lib().oracle_neighbor_pairs(p, n, sz, cell_size, None, None, 0)
    if cnt < 0:
        raise InvalidType('unknown agent dtype')
    oi, oj = np.empty(cnt, dtype=np.int64), np.empty(cnt, dtype=np.int64)
    lib().oracle_neighbor_pairs(p, n, sz, cell_size, oi.ctypes.data, oj.ctypes.data, cnt)
    return np.stack((oi, oj), axis=1)


def agent_agent_block_list(agents, cell_size):
    p, n, sz = _a(agents)
    _check(lib().oracle_agent_agent_block_list(p, n, sz, cell_size))


def agent_agent_brute(agents):
    p, n, sz = _a(agents)
    _check(lib().oracle_agent_agent_brute(p, n, sz))


def agent_obstacle(agents, obstacles):
    p, n, sz = _a(agents)
    obstacles = np.ascontiguousarray(obstacles)
    _check(lib().oracle_agent_obstacle(p, n, sz, obstacles.ctypes.data, len(obstacles)))


def adjusting(agents):
    _check(lib().oracle_adjusting(*_a(agents)))


def orientation(agents):
    _check(lib().oracle_orientation(*_a(agents)))


def navigation(agents, fields):
    """fields: list over targets of (mgrid, (U, V)) -- the reference's navigation_to_target layout."""
    p, n, sz = _a(agents)
    for target, (mg, (U, V)) in enumerate(fields):
        U = np.ascontiguousarray(U, dtype=np.float64)
        V = np.ascontiguousarray(V, dtype=np.float64)
        ny, nx = U.shape
        _check(lib().oracle_navigation(p, n, sz, target, U.ctypes.data, V.ctypes.data, ny, nx,
                                       mg.bounds[0], mg.bounds[1], mg.step))


def adaptive_timestep(agents, dt_min, dt_max):
    p, n, sz = _a(agents)
    return lib().oracle_adaptive_timestep(p, n, sz, dt_min, dt_max)


def velocity_verlet_integrator(agents, dt_min, dt_max):
    p, n, sz = _a(agents)
    dt = C.c_double()
    _check(lib().oracle_velocity_verlet_integrator(p, n, sz, dt_min, dt_max, C.byref(dt)))
    return dt.value


def reset(agents):
    _check(lib().oracle_reset(*_a(agents)))


def step(agents, obstacles, fields, cell_size, dt_min, dt_max):
    """One update() of the replaced sub-tree; returns dt."""
    if fields:
        navigation(agents, fields)
    orientation(agents)
    adjusting(agents)
    agent_agent_block_list(agents, cell_size)
    agent_obstacle(agents, obstacles)
    dt = velocity_verlet_integrator(agents, dt_min, dt_max)
    reset(agents)
    return dt


# ---- exit detection, herding, leader-follower -----------------------------------------------------
def _obs(obstacles):
    obs = np.ascontiguousarray(obstacles)
    return obs, obs.ctypes.data, len(obs)


def _v2(x):
    return np.ascontiguousarray(x, dtype=np.float64)


def line_intersect(x0, x1, y0, y1):
    x0, x1, y0, y1 = _v2(x0), _v2(x1), _v2(y0), _v2(y1)
    return bool(lib().oracle_line_intersect(x0.ctypes.data, x1.ctypes.data, y0.ctypes.data, y1.ctypes.data))


def is_obstacle_between_points(p0, p1, obstacles):
    p0, p1 = _v2(p0), _v2(p1)
    obs, po, no = _obs(obstacles)
    return bool(lib().oracle_is_obstacle_between_points(p0.ctypes.data, p1.ctypes.data, po, no))


def exit_detection(center_door, agents, obstacles, detection_range):
    """core/evacuation.py:137-174 on agents['position'] -> (detected_exit int64[n], has_detected bool[n])."""
    p, n, sz = _a(agents)
    doors = _v2(center_door).reshape(-1, 2)
    obs, po, no = _obs(obstacles)
    detected = np.empty(n, dtype=np.int64)
    has = np.zeros(n, dtype=np.uint8)
    _check(lib().oracle_exit_detection(doors.ctypes.data, len(doors), p, n, sz, po, no, float(detection_range),
                                       detected.ctypes.data, has.ctypes.data))
    return detected, has.astype(bool)


def herding_relationship(x1, x2, v1, v2, phi=np.pi / 2):
    x1, x2, v1, v2 = _v2(x1), _v2(x2), _v2(v1), _v2(v2)
    r = lib().oracle_herding_relationship(x1.ctypes.data, x2.ctypes.data, v1.ctypes.data, v2.ctypes.data, float(phi))
    return bool(r & 1), bool(r & 2)


def find_nearest_neighbors(agents, sight, size_nearest_other, obstacles):
    p, n, sz = _a(agents)
    obs, po, no = _obs(obstacles)
    neighbors = np.full((n, size_nearest_other), -1, dtype=np.int64)
    _check(lib().oracle_find_nearest_neighbors(p, n, sz, float(sight), int(size_nearest_other), po, no, neighbors.ctypes.data))
    return neighbors


def herding_interaction(agents, is_herding, neighbors, weight_position, phi):
    p, n, sz = _a(agents)
    ish = np.ascontiguousarray(is_herding, dtype=np.uint8)
    nb = np.ascontiguousarray(neighbors, dtype=np.int64)
    out = np.zeros((n, 2))
    has = np.zeros(n, dtype=np.uint8)
    _check(lib().oracle_herding_interaction(p, n, sz, ish.ctypes.data, nb.ctypes.data, nb.shape[1] if nb.ndim == 2 else 0,
                                            float(weight_position), float(phi), out.ctypes.data, has.ctypes.data))
    return out, has.astype(bool)


def leader_follower_interaction_brute(agents, weight_position, phi, obstacles, sight):
    p, n, sz = _a(agents)
    obs, po, no = _obs(obstacles)
    out = np.zeros((n, 2))
    has = np.zeros(n, dtype=np.uint8)
    _check(lib().oracle_leader_follower_interaction_brute(p, n, sz, float(weight_position), float(phi), po, no, float(sight),
                                                          out.ctypes.data, has.ctypes.data))
    return out, has.astype(bool)


def leader_follower_interaction(agents, obstacles, sight, phi=0.45 * np.pi, weight_position_leader=0.40):
    p, n, sz = _a(agents)
    obs, po, no = _obs(obstacles)
    out = np.zeros((n, 2))
    _check(lib().oracle_leader_follower_interaction(p, n, sz, po, no, float(sight), float(phi), float(weight_position_leader),
                                                    out.ctypes.data))
    return out


def leader_follower_with_herding_interaction(agents, obstacles, sight, size_nearest_other, phi=0.45 * np.pi,
                                             weight_position_herding=0.15, weight_position_leader=0.40,
                                             weight_direction_leader=0.65):
    p, n, sz = _a(agents)
    obs, po, no = _obs(obstacles)
    out = np.zeros((n, 2))
    _check(lib().oracle_leader_follower_with_herding_interaction(
        p, n, sz, po, no, float(sight), int(size_nearest_other), float(phi), float(weight_position_herding),
        float(weight_position_leader), float(weight_direction_leader), out.ctypes.data))
    return out


# ---- InsideDomain / TargetReached ----------------------------------------------------------------------
def _poly(vertices):
    v = np.ascontiguousarray(vertices, dtype=np.float64).reshape(-1, 2)
    if len(v) > 1 and (v[0] == v[-1]).all():      # a closed ring (shapely exterior): the closing vertex adds nothing
        v = np.ascontiguousarray(v[:-1])
    return v


def point_in_polygon(vertices, x, y):
    v = _poly(vertices)
    return bool(lib().oracle_point_in_polygon(v.ctypes.data, len(v), float(x), float(y)))


def inside_domain(agents, vertices):
    """InsideDomain.update (logic.py:351-357) -> number of agents whose ``active`` flag changed."""
    p, n, sz = _a(agents)
    v = _poly(vertices)
    r = lib().oracle_inside_domain(p, n, sz, v.ctypes.data, len(v))
    _check(r < 0)
    return int(r)


def target_reached(agents, vertices, reached_by):
    """TargetReached.update (logic.py:383-387) for one polygon; ``reached_by``: bool/uint8 array updated in place."""
    p, n, sz = _a(agents)
    v = _poly(vertices)
    rb = reached_by.view(np.uint8)
    assert rb.flags.c_contiguous and len(rb) == n
    r = lib().oracle_target_reached(p, n, sz, v.ctypes.data, len(v), rb.ctypes.data)
    _check(r < 0)
    return int(r)
