"""Data contract of the hot path: the structured dtypes the logic nodes operate on.

* ``agent_type_circular`` / ``agent_type_three_circle``: reference simulation/agents.py:447-457, generated there by
  traits.class_to_struct_dtype (traits.py:200-219) walking the MRO ``Circular -> States -> Body -> TranslationalMotion``
  (``BodyType`` excluded, agents.py:307); ThreeCircle puts its own ``position_ls/position_rs`` first and
  ``RotationalMotion`` last.  ``np.dtype([...])`` without ``align=True`` packs: itemsize 228 / 316 bytes.
* ``obstacle_type_linear``: reference core/structures.py:6-9.
"""
import numpy as np

from .exceptions import InvalidType

_STATES = [('active', np.bool_), ('target_reached', np.bool_), ('target', np.int64),
           ('is_leader', np.bool_), ('is_follower', np.bool_), ('index_leader', np.int64),
           ('familiar_exit', np.int64)]
_BODY = [(n, np.float64) for n in ('radius', 'r_t', 'r_s', 'r_ts', 'mass', 'inertia_rot',
                                   'target_velocity', 'target_angular_velocity')]
_TRANSLATIONAL = [(n, np.float64, (2,)) for n in ('position', 'velocity', 'target_direction', 'force', 'force_prev')] \
    + [(n, np.float64) for n in ('tau_adj', 'k_soc', 'tau_0', 'mu', 'kappa', 'damping', 'std_rand_force')]
_ROTATIONAL = [(n, np.float64) for n in ('orientation', 'angular_velocity', 'target_orientation', 'torque',
                                         'torque_prev', 'tau_rot', 'std_rand_torque')]
_THREE_CIRCLE = [('position_ls', np.float64, (2,)), ('position_rs', np.float64, (2,))]

agent_type_circular = np.dtype(_STATES + _BODY + _TRANSLATIONAL)
agent_type_three_circle = np.dtype(_THREE_CIRCLE + _STATES + _BODY + _TRANSLATIONAL + _ROTATIONAL)
obstacle_type_linear = np.dtype([('p0', np.float64, (2,)), ('p1', np.float64, (2,))])

assert agent_type_circular.itemsize == 228
assert agent_type_three_circle.itemsize == 316
assert obstacle_type_linear.itemsize == 32

AgentModelToType = {'circular': agent_type_circular, 'three_circle': agent_type_three_circle}
MODEL_CIRCULAR = 0
MODEL_THREE_CIRCLE = 1
NO_TARGET = -1
NO_LEADER = -1


def is_model(agents, model):
    """Same test as reference simulation/agents.py:460-470 (dtype identity by hash)."""
    return hash(agents.dtype) == hash(AgentModelToType[model])


def model_of(agents):
    """MODEL_* id of a structured agents array; raises InvalidType like interactions.py:204-205."""
    if not isinstance(agents, np.ndarray) or agents.dtype.fields is None:
        raise InvalidType('agents must be a structured numpy array')
    if is_model(agents, 'circular'):
        return MODEL_CIRCULAR
    if is_model(agents, 'three_circle'):
        return MODEL_THREE_CIRCLE
    raise InvalidType('unknown agent dtype (itemsize %d)' % agents.dtype.itemsize)


def as_obstacles(obstacles):
    """Accept a structured ``obstacle_type_linear`` array, an (W, 2, 2)/(W, 4) float array or None -> (W, 4) f64."""
    if obstacles is None:
        return np.zeros((0, 4), dtype=np.float64)
    obstacles = np.asarray(obstacles)
    if obstacles.dtype.fields is not None:
        if obstacles.dtype.itemsize != 32:
            raise InvalidType('obstacles must have dtype obstacle_type_linear')
        return np.ascontiguousarray(obstacles).view(np.float64).reshape(-1, 4)
    return np.ascontiguousarray(obstacles, dtype=np.float64).reshape(-1, 4)
