"""SURVEY 8(f) rank 3 on the GPU: InsideDomain / TargetReached through the C ABI and the logic nodes, bit-exact against the
oracle's restatement of matplotlib's contains_points rule (see tests/test_oracle_domain.py for what pins that)."""
import numpy as np
import pytest

from crowddynamics_b200 import synthetic as S, logic as L, _lib
from crowddynamics_b200.engine import DeviceAgents
from crowddynamics_b200.exceptions import CrowdDynamicsException, InvalidValue
from crowddynamics_b200.structures import model_of
from oracle import crowd_oracle as O

pytestmark = pytest.mark.gpu


def _star(cx, cy, r_out, r_in, n=9):
    a = np.linspace(0, 2 * np.pi, 2 * n, endpoint=False)
    r = np.where(np.arange(2 * n) % 2 == 0, r_out, r_in)
    return np.stack((cx + r * np.cos(a), cy + r * np.sin(a)), axis=1)


@pytest.mark.parametrize('model', ['circular', 'three_circle'])
def test_kernels_against_oracle(model):
    agents, obstacles, side = S.uniform_crowd(30000, model, density=1.0, seed=4)
    agents['active'] = np.random.default_rng(0).random(len(agents)) < 0.5
    ref = agents.copy()
    domain = _star(side / 2, side / 2, 0.55 * side, 0.3 * side)
    goals = [_star(side * 0.3, side * 0.3, 0.2 * side, 0.1 * side, 5), np.array([(0, 0), (side / 3, 0), (0, side / 3)]),
             np.array([(side, side), (side + 5, side), (side + 5, side + 5)])]          # outside the room: only reached after the jump
    dev = DeviceAgents(model_of(agents), capacity=len(agents))
    dev.upload(agents)
    dev.set_polygons(_lib.POLY_DOMAIN, [np.vstack((domain, domain[:1]))])             # closed ring, like shapely's exterior
    dev.set_polygons(_lib.POLY_TARGETS, goals)
    dev.set_active(agents['active'])
    reached = [np.zeros(len(agents), dtype=bool) for _ in goals]
    for it in range(4):
        changed = dev.inside_domain()
        assert changed == O.inside_domain(ref, domain)
        assert (dev.get_active() == ref['active']).all()
        counts = dev.target_reached(len(goals))
        assert list(counts) == [O.target_reached(ref, g, r) for g, r in zip(goals, reached)]
        assert (dev.target_reached_by(len(goals)) == np.stack(reached)).all()
        # move: fused steps on the device (re-sorts the planes), the oracle beside it
        dev.step(3, _lib.STEP_ALL & ~_lib.STEP_NAVIGATION, 3.6, 0.01, 0.01)
        for _ in range(3):
            O.step(ref, obstacles[:0], [], 3.6, 0.01, 0.01)
        if it == 1:                                                                    # and a big jump for a part of the crowd
            out = agents.copy(); dev.download(out)
            out['position'][::7] += 0.4 * side
            ref['position'] = out['position']                                          # (also re-synchronises the two trajectories)
            if model != 'circular':
                S.set_shoulders(out)
                ref['position_ls'], ref['position_rs'] = out['position_ls'], out['position_rs']
            act = dev.get_active()
            dev.upload(out); dev.set_active(act)
    assert counts[0] > 0 and counts[1] > 0


def test_nodes_strict_and_resident():
    for mode in ('strict', 'resident'):
        agents, obstacles, side = S.uniform_crowd(2000, 'circular', density=1.0, seed=6)
        agents['active'] = True
        ref = agents.copy()
        domain = np.array([(1.0, 1.0), (side - 1.0, 1.0), (side - 1.0, side - 1.0), (1.0, side - 1.0)])
        goal = _star(side / 2, side / 2, 6.0, 3.0)
        sim = L.MultiAgentSimulation(agents, obstacles, (), domain=domain)
        sim.logic = L.Reset(sim, mode=mode) << (L.InsideDomain(sim), L.TargetReached(sim, polygons=[None, goal]),
                                                L.Integrator(sim) << (L.Adjusting(sim), L.AgentAgentInteractions(sim)))
        assert sim.data['inactive'] == 0 and sim.data['target_1'] == 0 and 'target_0' not in sim.data
        inactive, reached = 0, np.zeros(len(ref), dtype=bool)
        for _ in range(5):
            sim.update()
            inactive += O.inside_domain(ref, domain)
            count = O.target_reached(ref, goal, reached)
            O.adjusting(ref); O.agent_agent_block_list(ref, 3.6); O.velocity_verlet_integrator(ref, 0.01, 0.01); O.reset(ref)
            assert sim.data['inactive'] == inactive and sim.data['target_1'] == count
            if mode == 'strict':
                assert (agents['active'] == ref['active']).all()
        sim.logic.state.sync_host()
        assert (agents['active'] == ref['active']).all() and 0 < ref['active'].sum() < len(ref)
        assert (sim.logic['TargetReached'].reached_by[0] == reached).all()


def test_errors():
    agents, _, side = S.uniform_crowd(100, 'circular', density=1.0, seed=1)
    dev = DeviceAgents(model_of(agents), capacity=len(agents))
    dev.upload(agents)
    with pytest.raises(CrowdDynamicsException):
        dev.inside_domain()                                   # no domain polygon yet
    dev.set_polygons(_lib.POLY_DOMAIN, [np.array([(0, 0), (1, 0), (1, 1)])])
    with pytest.raises(CrowdDynamicsException):
        dev.inside_domain()                                   # no active flags yet
    with pytest.raises(InvalidValue):
        dev.set_active(np.zeros(5, dtype=bool))               # wrong length
    with pytest.raises(InvalidValue):
        dev.set_polygons(_lib.POLY_DOMAIN, [np.zeros((3, 2)), np.zeros((3, 2))])
    with pytest.raises(InvalidValue):
        dev.target_reached(2)                                 # no target polygons set
    dev.set_active(np.ones(100, dtype=bool))
    assert dev.inside_domain() == 100 - sum(O.point_in_polygon([(0, 0), (1, 0), (1, 1)], *p) for p in agents['position'])
    assert list(dev.target_reached(0)) == []


# ---- a fully resident tree never waits for the device (SaveSimulationData, deferred counters) ----------------------------------
def test_resident_tree_with_io_nodes_never_blocks(tmp_path):
    """FusedStep + InsideDomain + TargetReached + SaveSimulationData, all resident and deferred: in steady state an update
    performs ZERO blocking host synchronisations (counted inside the library), the saved chunks equal the oracle trajectory
    update by update, and the counters equal the oracle's once flushed."""
    from crowddynamics_b200 import logic as L
    agents, obstacles, fields = S.hallway(seed=4)
    domain = np.array([(1.0, -1.0), (39.0, -1.0), (39.0, 6.0), (1.0, 6.0)])       # agents walk out of it at both ends
    goal = np.array([(30.0, -1.0), (41.0, -1.0), (41.0, 6.0), (30.0, 6.0)])
    ref = agents.copy()
    sim = L.MultiAgentSimulation(agents, obstacles, fields, domain=domain)
    updates, every = 60, 25
    step = L.FusedStep(sim, step=0.1, deferred=True)
    inside = L.InsideDomain(sim, deferred=True)
    reached = L.TargetReached(sim, polygons=[goal], deferred=True)
    saver = L.SaveSimulationData(sim, save_condition=lambda s: (s.data['iterations'] + 1) % every == 0, base_directory=str(tmp_path),
                                 save_directory='run')
    scal = L.ScalarsSync(sim)
    sim.logic = scal << (saver << (reached << (inside << step)))
    assert [n.name for n in L.post_order_iter(sim.logic.root)] == ['FusedStep', 'InsideDomain', 'TargetReached', 'SaveSimulationData', 'ScalarsSync']
    syncs, traj, inactive_ref, reached_ref = [], [], 0, np.zeros(len(ref), dtype=bool)
    act = ref['active'].copy()
    for it in range(updates):
        before = sim.logic.state.dev.sync_count() if sim.logic.state.dev is not None else None
        sim.update()
        if before is not None:
            syncs.append(sim.logic.state.dev.sync_count() - before)
        O.step(ref, obstacles, fields, 3.6, 0.01, 0.01)
        inactive_ref += O.inside_domain(ref, domain)
        O.target_reached(ref, goal, reached_ref)
        traj.append(ref.copy())
    dumping = [(it + 1) % every == 0 for it in range(1, updates)]
    steady = [s for s, d in zip(syncs, dumping) if not d][5:]
    assert steady and max(steady) == 0, syncs                     # nothing waits outside the dumping updates
    saver.flush(); scal.flush()
    assert sim.data['inactive'] == inactive_ref and sim.data['target_0'] == int(reached_ref.sum())
    assert abs(sim.data['time_tot'] - 0.01 * updates) < 1e-12
    assert len(saver.files) == updates // every
    saved = np.concatenate([np.load(f) for f in saver.files] + ([np.vstack(saver.buffer)] if saver.buffer else []))
    assert saved.shape == (updates, len(ref))
    for it in range(updates):
        assert np.abs(saved[it]['position'] - traj[it]['position']).max() <= 1e-9, it
        assert (saved[it]['active'] == traj[it]['active']).all(), it
        assert (saved[it]['radius'] == ref['radius']).all()
