"""Drop-in for reference core/steering/orientation.py:17-21."""
from ... import _lib
from ...engine import device_agents_for


def orient_towards_target_direction(agents):
    dev = device_agents_for(agents)
    dev.upload(agents)
    dev.orientation()
    dev.download(agents, _lib.F_TARGET_ORIENTATION)
