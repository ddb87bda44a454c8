"""Drop-in for reference core/motion/adjusting.py:101-121."""
from ... import _lib
from ...engine import device_agents_for
from ...structures import is_model


def adjust_agents(agents):
    """force_adjust_agents + (three_circle) torque_adjust_agents, i.e. Adjusting.update (logic.py:89-94)."""
    dev = device_agents_for(agents)
    dev.upload(agents)
    dev.adjust()
    dev.download(agents, _lib.F_FORCE | _lib.F_TORQUE)


def force_adjust_agents(agents):
    dev = device_agents_for(agents)
    dev.upload(agents)
    dev.adjust()
    dev.download(agents, _lib.F_FORCE)


def torque_adjust_agents(agents):
    if not is_model(agents, 'three_circle'):
        raise TypeError('torque_adjust_agents needs three_circle agents')
    dev = device_agents_for(agents)
    dev.upload(agents)
    dev.adjust()
    dev.download(agents, _lib.F_TORQUE)
