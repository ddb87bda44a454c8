"""The drop-in logic nodes (crowddynamics_b200.logic) driving a MultiAgentSimulation-like host, against the oracle."""
import numpy as np
import pytest

from conftest import load_golden, rel_err_fields, vec_rel_err
from crowddynamics_b200 import synthetic as S, logic as L
from oracle import crowd_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('model', ['circular', 'three_circle'])
def test_strict_tree_matches_oracle_step_by_step(model):
    agents, obstacles, fields = S.hallway(seed=1, model=model)
    ref = agents.copy()
    sim = L.MultiAgentSimulation(agents, obstacles, fields)
    sim.logic = L.hallway_logic(sim, mode='strict', dt_min=0.001, dt_max=0.01)
    assert [n.name for n in L.post_order_iter(sim.logic.root)] == [
        'Navigation', 'Orientation', 'Adjusting', 'AgentAgentInteractions', 'AgentObstacleInteractions', 'Integrator',
        'Reset']
    assert sim.logic['Integrator'].dt_max == 0.01 and sim.logic['AgentAgentInteractions'].cell_size == 3.6
    t = 0.0
    for it in range(20):
        sim.update()
        t += O.step(ref, obstacles, fields, 3.6, 0.001, 0.01)
        assert np.abs(agents['position'] - ref['position']).max() <= 1e-9
    assert sim.data['iterations'] == 20
    assert abs(sim.data['time_tot'] - t) <= 1e-12
    assert (agents['force'] == 0).all()


def test_hallway_reference_trajectory_resident():
    """BASELINE config 1: 200 updates of the Hallway, fused + resident, against the reference's own trajectory."""
    g = load_golden('hallway.npz')
    agents, obstacles, fields = S.hallway(seed=0)
    sim = L.MultiAgentSimulation(agents, obstacles, fields)
    sim.logic = L.FusedStep(sim, dt_min=0.01, dt_max=0.01, steps_per_update=50)
    traj = [agents['position'].copy()]
    for _ in range(4):
        sim.update()
        sim.logic.state.sync_host()
        traj.append(agents['position'].copy())
    assert np.abs(np.stack(traj) - g['positions']).max() <= 1e-6
    assert abs(sim.data['time_tot'] - 2.0) < 1e-9
    # reference examples/tests/test_validation.py:13-38 flavour: the crowd actually walks
    assert np.abs(agents['position'][:, 0] - traj[0][:, 0]).mean() > 1.0


def test_resident_nodes_and_host_sync():
    agents, obstacles, fields = S.hallway(seed=2, model='three_circle')
    ref = agents.copy()
    sim = L.MultiAgentSimulation(agents, obstacles, fields)
    sim.logic = L.hallway_logic(sim, mode='resident')
    for _ in range(10):
        sim.update()
        O.step(ref, obstacles, fields, 3.6, 0.01, 0.01)
    assert (agents['position'] != ref['position']).any()      # host not refreshed yet
    sim.logic.state.sync_host()
    assert np.abs(agents['position'] - ref['position']).max() <= 1e-9
    assert np.abs(agents['orientation'] - ref['orientation']).max() <= 1e-8
    # a host-side node edits the array -> invalidate -> next node re-uploads
    agents['velocity'] = 0.0
    ref['velocity'] = 0.0
    sim.logic.state.invalidate()
    sim.update(); O.step(ref, obstacles, fields, 3.6, 0.01, 0.01)
    sim.logic.state.sync_host()
    assert np.abs(agents['position'] - ref['position']).max() <= 1e-9


@pytest.mark.parametrize('model', ['circular', 'three_circle'])
def test_room_evacuation_style_config(model):
    """BASELINE config 4 flavour at test size: room with a door, 11 wall segments, static exit field, all target 0;
    10 fused resident updates against the oracle.  (At 2 agents/m^2 with stiff contacts the dynamics amplify a 1e-16
    difference roughly 4x per step -- 1e-6 m after 20 steps -- so the horizon is kept short.)"""
    agents, obstacles, fields, side = S.room_with_exit(4000, model, density=2.0, seed=6)
    assert len(obstacles) == 11
    ref = agents.copy()
    sim = L.MultiAgentSimulation(agents, obstacles, fields)
    sim.logic = L.FusedStep(sim, dt_min=0.001, dt_max=0.01, steps_per_update=5, step=0.5)
    t = 0.0
    for _ in range(2):
        sim.update()
        for _ in range(5):
            t += O.step(ref, obstacles, fields, 3.6, 0.001, 0.01)
    sim.logic.state.sync_host()
    assert abs(sim.data['time_tot'] - t) <= 1e-12
    assert np.abs(agents['position'] - ref['position']).max() <= 1e-8
    assert np.abs(agents['velocity'] - ref['velocity']).max() <= 1e-6
    # the crowd heads for the door
    d0 = np.hypot(side - ref['position'][:, 0], side / 2 - ref['position'][:, 1]).mean()
    assert np.hypot(side - agents['position'][:, 0], side / 2 - agents['position'][:, 1]).mean() <= d0 + 1e-9


def _validation_agents(model, positions, orientations):
    """Attributes of reference examples/validation.py:26-40,76-90."""
    from crowddynamics_b200.structures import agent_type_circular, agent_type_three_circle
    a = np.zeros(len(positions), dtype=agent_type_circular if model == 'circular' else agent_type_three_circle)
    S.fill_adult_bodies(a, np.random.default_rng(0), omega0=0.0)
    a['radius'] = 0.255; a['r_t'] = 0.5882 * 0.255; a['r_s'] = 0.3725 * 0.255; a['r_ts'] = 0.6275 * 0.255
    a['mass'] = 73.5; a['target_velocity'] = 1.0
    a['inertia_rot'] = 4.0 * np.pi * (73.5 / 80.0) * (0.255 / 0.27) ** 2
    a['position'] = positions
    e = np.stack((np.cos(orientations), np.sin(orientations)), 1)
    a['velocity'] = 1.0 * e
    a['target_direction'] = e
    if model != 'circular':
        a['orientation'] = orientations
        a['target_orientation'] = orientations
        S.set_shoulders(a)
    return a


@pytest.mark.parametrize('model', ['circular', 'three_circle'])
def test_reference_validation_movement(model):
    """reference examples/tests/test_validation.py:13-24 (TestMovement): one agent, Reset << Integrator << Adjusting <<
    Orientation, 1000 iterations of dt = 0.01 at v0 = 1 => it has walked (at least about) 10 m."""
    agents = _validation_agents(model, [(0.0, 0.0)], np.array([0.0]))
    ref = agents.copy()
    sim = L.MultiAgentSimulation(agents, None, [])
    sim.logic = L.Reset(sim) << (L.Integrator(sim) << (L.Adjusting(sim) << L.Orientation(sim)))
    for _ in range(1000):
        sim.update()
        O.orientation(ref); O.adjusting(ref); O.velocity_verlet_integrator(ref, 0.01, 0.01); O.reset(ref)
    dist = np.hypot(*(agents['position'][0]))
    assert dist >= 10.0 or np.isclose(dist, 10.0)
    assert np.abs(agents['position'] - ref['position']).max() <= 1e-12
    assert abs(sim.data['time_tot'] - 10.0) < 1e-9


@pytest.mark.parametrize('model', ['circular', 'three_circle'])
def test_reference_validation_agent_interaction(model):
    """reference examples/tests/test_validation.py:27-38 (TestAgentInteraction): two agents walking head-on from (0, 0) and
    (5, 0) with Fluctuation, Adjusting << Orientation and AgentAgentInteractions; the reference asserts that agent 0 still
    covers >= 8 m in 1000 iterations, which depends on its unseeded random fluctuation breaking the head-on symmetry (the
    noise-free oracle gives 6.4 m circular / 7.8 m three-circle).  Here: clear progress, and the agents got past each other."""
    agents = _validation_agents(model, [(0.0, 0.0), (5.0, 0.0)], np.array([0.0, np.pi]))
    sim = L.MultiAgentSimulation(agents, None, [])
    sim.logic = L.Reset(sim, mode='resident') << (L.Integrator(sim) << (
        L.Fluctuation(sim, seed=3), L.Adjusting(sim) << L.Orientation(sim), L.AgentAgentInteractions(sim)))
    start = agents['position'].copy()
    for _ in range(1000):
        sim.update()
    sim.logic.state.sync_host()
    dist = np.hypot(*(agents['position'][0] - start[0]))
    assert dist >= 6.0
    assert agents['position'][0, 0] > agents['position'][1, 0]                 # they passed each other
    assert np.isfinite(agents['position']).all()
