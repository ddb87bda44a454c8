"""Drop-in for reference core/integrator.py:209-256."""
from .. import _lib
from ..engine import device_agents_for

_MASK = (_lib.F_POSITION | _lib.F_VELOCITY | _lib.F_FORCE_PREV | _lib.F_SHOULDERS | _lib.F_ORIENTATION |
         _lib.F_ANGULAR_VELOCITY | _lib.F_TORQUE_PREV)


def velocity_verlet_integrator(agents, dt_min, dt_max):
    """Adaptive-dt velocity Verlet (+ rotational Verlet and shoulders for three_circle); returns dt."""
    dev = device_agents_for(agents)
    dev.upload(agents)
    dt = dev.integrate(dt_min, dt_max)
    dev.download(agents, _MASK)
    return dt
