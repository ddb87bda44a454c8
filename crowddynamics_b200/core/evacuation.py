"""Drop-in for reference core/evacuation.py:137-174."""
import numpy as np

from ..engine import device_agents_for
from ..structures import agent_type_circular


def exit_detection(center_door, position, obstacles, detection_range):
    """evacuation.py:137-174: -> (detected_exit int64[n], has_detected bool[n]); ``position`` may be an (n, 2) array (as in
    the reference) or a structured agents array."""
    if isinstance(position, np.ndarray) and position.dtype.fields is not None:
        agents = position
    else:
        position = np.asarray(position, dtype=np.float64).reshape(-1, 2)
        agents = np.zeros(len(position), dtype=agent_type_circular)
        agents['position'] = position
    dev = device_agents_for(agents)
    dev.upload(agents)
    dev.set_obstacles(obstacles)
    dev.exit_detection(center_door, detection_range, apply=False)
    return dev.exit_detection_result()
