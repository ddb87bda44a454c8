"""Drop-in for reference core/steering/collective_motion.py:69-289 (function-level API, strict mode: the host array is
uploaded, the kernels run, the States fields they mutate are written back)."""
import numpy as np

from ...engine import device_agents_for

MISSING_NEIGHBOR = -1


def _prepare(agents, obstacles):
    dev = device_agents_for(agents)
    dev.upload(agents)
    dev.set_states(agents, target=False)
    dev.set_obstacles(obstacles)
    return dev


def find_nearest_neighbors(agents, sight, size_nearest_other, obstacles):
    """collective_motion.py:69-110 over the block list of :262-267 (cell_size = sight) -> neighbors (n, k), -1 = missing.
    (The reference takes the position array and the block-list tables; the tables are built on the device here.)"""
    return _prepare(agents, obstacles).nearest_neighbors(sight, size_nearest_other)


def leader_follower_interaction(agents, obstacles, sight, phi=0.45 * np.pi, weight_position_leader=0.40):
    """collective_motion.py:229-243 -> direction (n, 2); mutates agents['target'] / agents['index_leader']."""
    dev = _prepare(agents, obstacles)
    dev.leader_follower(sight, phi, weight_position_leader)
    dev.get_states(agents)
    return dev.direction()


def leader_follower_with_herding_interaction(agents, obstacles, sight, size_nearest_other, phi=0.45 * np.pi,
                                             weight_position_herding=0.15, weight_position_leader=0.40,
                                             weight_direction_leader=0.65):
    """collective_motion.py:246-289 -> direction (n, 2); mutates agents['target'] / agents['index_leader']."""
    dev = _prepare(agents, obstacles)
    dev.leader_follower_with_herding(sight, size_nearest_other, phi, weight_position_herding, weight_position_leader,
                                     weight_direction_leader)
    dev.get_states(agents)
    return dev.direction()
