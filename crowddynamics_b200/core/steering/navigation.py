"""Drop-in for the sampling part of Navigation.update (reference logic.py:149-165, navigation.py:60-78)."""
from ... import _lib
from ...engine import device_agents_for


def navigate(agents, fields):
    """fields: list over targets of (mgrid, (U, V)); writes agents['target_direction'] for agents with that target."""
    dev = device_agents_for(agents)
    dev.upload(agents)
    dev.clear_navigation()
    for target, (mgrid, direction_map) in enumerate(fields):
        dev.set_navigation_field(target, mgrid, direction_map)
    dev.navigation()
    dev.download(agents, _lib.F_TARGET_DIRECTION)
