"""Drop-in for reference core/interactions.py:191-214 (higher level API)."""
from .. import _lib
from ..engine import device_agents_for


def agent_agent_block_list(agents, cell_size):
    """interactions.py:191-205: block-list neighbour search + agent-agent social/contact forces (and torques)."""
    dev = device_agents_for(agents)
    dev.upload(agents)
    dev.agent_agent(cell_size)
    dev.download(agents, _lib.F_FORCE | _lib.F_TORQUE)


def agent_obstacle(agents, obstacles):
    """interactions.py:208-214: agent - linear obstacle contact forces (and torques)."""
    dev = device_agents_for(agents)
    dev.upload(agents)
    dev.set_obstacles(obstacles)
    dev.agent_obstacle()
    dev.download(agents, _lib.F_FORCE | _lib.F_TORQUE)


def block_list(agents, cell_size):
    """cell_lists.add_to_cells(agents['position'], cell_size) as called at interactions.py:192-193:
    -> (points_indices, cells_count, cells_offset, grid_shape)."""
    dev = device_agents_for(agents)
    dev.upload(agents)
    dev.build_block_list(cell_size)
    return dev.cell_tables()


def neighbor_pairs(agents, cell_size):
    """All (i, j) the reference's iter_nearest_neighbors loop visits (interactions.py:152-155), as an (P, 2) array
    (order unspecified)."""
    dev = device_agents_for(agents)
    dev.upload(agents)
    dev.build_block_list(cell_size)
    return dev.neighbor_pairs()
