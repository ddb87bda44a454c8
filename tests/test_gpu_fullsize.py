"""Size-independent properties at the BASELINE sizes (1 M agents), where the serial oracle would take minutes:
block-list invariants, the reference's known-answer property (equal velocities => no social force) on a whole crowd,
two independent GPU implementations against each other, run-to-run determinism, relabelling invariance, and a sampled
comparison against the oracle (the oracle evaluates a window of the big crowd cut out with a halo)."""
import numpy as np
import pytest

from conftest import vec_rel_err
from crowddynamics_b200 import _lib, synthetic as S
from crowddynamics_b200.engine import DeviceAgents
from crowddynamics_b200.structures import MODEL_CIRCULAR, MODEL_THREE_CIRCLE
from oracle import crowd_oracle as O

pytestmark = pytest.mark.gpu
N = 1000000
CELL = 3.6
PRE = _lib.STEP_ALL & ~(_lib.STEP_INTEGRATOR | _lib.STEP_RESET | _lib.STEP_NAVIGATION)


def _mid(model):
    return MODEL_CIRCULAR if model == 'circular' else MODEL_THREE_CIRCLE


@pytest.fixture(scope='module', params=['circular', 'three_circle'])
def crowd(request):
    agents, obstacles, side = S.uniform_crowd(N, request.param, density=1.0, seed=0)
    return request.param, agents, obstacles, side


def test_block_list_invariants_1m(crowd):
    model, agents, obstacles, side = crowd
    dev = DeviceAgents(_mid(model))
    dev.upload(agents)
    dev.build_block_list(CELL)
    pi, cc, co, gs = dev.cell_tables()
    cells = dev.cell_ids()
    dev.close()
    assert cc.sum() == N and (co == np.concatenate(([0], np.cumsum(cc)[:-1]))).all()
    assert (np.sort(pi) == np.arange(N)).all()                       # a permutation of the agents
    assert (np.diff(cells[pi]) >= 0).all()                           # grouped by cell, cells ascending
    same = np.diff(cells[pi]) == 0
    assert (np.diff(pi)[same] > 0).all()                             # ascending agent index inside a cell
    ix = np.floor(agents['position'][:, 0] / CELL).astype(np.int64)
    iy = np.floor(agents['position'][:, 1] / CELL).astype(np.int64)
    assert (cells == (ix - ix.min()) * gs[1] + (iy - iy.min())).all()  # cell = floor(p / c), bit exact
    assert tuple(gs) == (ix.max() - ix.min() + 1, iy.max() - iy.min() + 1)


def test_equal_velocities_mean_no_social_force_1m(crowd):
    """reference core/motion/tests/test_power_law_benchmark.py:13-35, on a whole non-overlapping crowd."""
    model, agents, obstacles, side = crowd
    a = agents.copy()
    a['velocity'] = (0.7, -0.3)
    dev = DeviceAgents(_mid(model))
    dev.upload(a)
    dev.agent_agent(CELL)
    dev.download(a)
    dev.close()
    assert (a['force'] == 0).all()
    if model == 'three_circle':
        assert (a['torque'] == 0).all()


def test_two_gpu_implementations_agree_and_are_deterministic_1m(crowd):
    model, agents, obstacles, side = crowd
    out = []
    for variant in (3, 3, 1, 2):
        dev = DeviceAgents(_mid(model))
        dev.set_variant(variant)
        dev.upload(agents)
        dev.set_obstacles(obstacles)
        dev.step(1, PRE, CELL, 0.01, 0.01, want_dt=False)
        f = agents.copy()
        dev.download(f)
        dev.close()
        out.append(f)
    assert (out[0]['force'] == out[1]['force']).all()                 # bit-reproducible (whatever order the atomics resolved in)
    assert vec_rel_err(out[0]['force'], out[2]['force']) <= 1e-10     # once-per-pair pipeline vs one-phase kernels
    assert (out[0]['force'] == out[3]['force']).all()                 # ... and bit-identical to the both-sides fused kernel
    if model == 'three_circle':
        assert (out[0]['torque'] == out[1]['torque']).all()
        assert vec_rel_err(out[0]['torque'], out[2]['torque']) <= 1e-10
        assert (out[0]['torque'] == out[3]['torque']).all()


def test_window_against_oracle_1m(crowd):
    """Per-agent forces of the agents in a 40 m x 40 m window of the 1 M crowd against the oracle run on the window plus a
    4 m halo (everything an inner agent can interact with)."""
    model, agents, obstacles, side = crowd
    dev = DeviceAgents(_mid(model))
    dev.upload(agents)
    dev.set_obstacles(obstacles)
    dev.step(1, PRE, CELL, 0.01, 0.01, want_dt=False)
    got = agents.copy()
    dev.download(got)
    dev.close()
    for (x0, y0) in ((0.0, 0.0), (side / 2 - 20.0, side / 2 - 20.0), (side - 40.0, side - 40.0)):
        p = agents['position']
        halo = (p[:, 0] >= x0 - 4.0) & (p[:, 0] < x0 + 44.0) & (p[:, 1] >= y0 - 4.0) & (p[:, 1] < y0 + 44.0)
        sub = np.ascontiguousarray(agents[halo])          # index order preserved: same pair orientation as the full crowd
        inner = (sub['position'][:, 0] >= x0) & (sub['position'][:, 0] < x0 + 40.0) & \
                (sub['position'][:, 1] >= y0) & (sub['position'][:, 1] < y0 + 40.0)
        O.adjusting(sub)
        # anchor the oracle's lattice like the full crowd's so that the (cell, index) pair orientation matches
        O.agent_agent_block_list(sub, CELL)
        O.agent_obstacle(sub, obstacles)
        ref = sub[inner]
        mine = got[halo][inner]
        assert inner.sum() > 1000
        assert vec_rel_err(mine['force'], ref['force']) <= 1e-9
        if model == 'three_circle':
            assert vec_rel_err(mine['torque'], ref['torque']) <= 1e-9


def test_every_agent_against_oracle_1m(crowd):
    """ALL 1 M agents: forces / torques after the force nodes of one step, and the state after the integrator, against the
    serial C oracle on the whole crowd (~10 s of CPU; the headline workload of bench.py itself, not a window of it)."""
    model, agents, obstacles, side = crowd
    fields = [S.direction_field(1.0, (0.0, 0.0, side, side), 'exit', point=(side, side / 2))]
    ref = agents.copy()
    O.navigation(ref, fields); O.orientation(ref); O.adjusting(ref)
    O.agent_agent_block_list(ref, CELL); O.agent_obstacle(ref, obstacles)
    dev = DeviceAgents(_mid(model))
    dev.upload(agents)
    dev.set_obstacles(obstacles)
    dev.set_navigation_field(0, *fields[0])
    dev.step(1, _lib.STEP_ALL & ~(_lib.STEP_INTEGRATOR | _lib.STEP_RESET), CELL, 0.001, 0.01, want_dt=False)
    got = agents.copy()
    dev.download(got)
    assert vec_rel_err(got['force'], ref['force']) <= 1e-9
    assert (got['target_direction'] == ref['target_direction']).all()
    if model == 'three_circle':
        assert vec_rel_err(got['torque'], ref['torque']) <= 1e-9
        assert (got['target_orientation'] == ref['target_orientation']).all() or \
            np.abs(got['target_orientation'] - ref['target_orientation']).max() <= 1e-15
    interacting = (np.abs(ref['force'] - agents['force']).sum(axis=1) > 0).mean()
    assert interacting > 0.9                                   # the comparison is about real pair forces
    dt = dev.step(1, _lib.STEP_INTEGRATOR | _lib.STEP_RESET, CELL, 0.001, 0.01)[0]
    dt_ref = O.velocity_verlet_integrator(ref, 0.001, 0.01)
    O.reset(ref)
    dev.download(got)
    dev.close()
    assert abs(dt - dt_ref) <= 1e-15
    assert np.abs(got['position'] - ref['position']).max() <= 1e-12 and np.abs(got['velocity'] - ref['velocity']).max() <= 1e-11
    if model == 'three_circle':
        assert np.abs(got['orientation'] - ref['orientation']).max() <= 1e-11
        assert np.abs(got['position_ls'] - ref['position_ls']).max() <= 1e-11


def test_relabelling_invariance_circular():
    """Shuffling the agent order leaves every agent's force unchanged up to summation order (circular pairs are exactly
    swap symmetric, so pair orientation does not matter)."""
    agents, obstacles, side = S.uniform_crowd(200000, 'circular', density=1.5, seed=4, overlap_fraction=0.02)
    perm = np.random.default_rng(0).permutation(len(agents))
    res = []
    for arr in (agents, np.ascontiguousarray(agents[perm])):
        dev = DeviceAgents(MODEL_CIRCULAR)
        dev.upload(arr)
        dev.agent_agent(CELL)
        out = arr.copy()
        dev.download(out)
        dev.close()
        res.append(out)
    assert vec_rel_err(res[1]['force'], res[0]['force'][perm], floor=1e-9) <= 1e-10


def test_thousand_steps_stay_finite_and_inside_the_room():
    """Config 2 flavour: 1000 fused steps of a walled room (smaller crowd to keep the test short)."""
    agents, obstacles, side = S.uniform_crowd(100000, 'circular', density=1.0, seed=2)
    dev = DeviceAgents(MODEL_CIRCULAR)
    dev.upload(agents)
    dev.set_obstacles(obstacles)
    dts = dev.step(1000, _lib.STEP_ALL & ~_lib.STEP_NAVIGATION, CELL, 0.001, 0.01)
    out = agents.copy()
    dev.download(out)
    t, it = dev.time()
    dev.close()
    assert it == 1000 and abs(t - dts.sum()) < 1e-9 and ((dts >= 0.001) & (dts <= 0.01)).all()
    assert np.isfinite(out['position']).all() and np.isfinite(out['velocity']).all()
    assert (out['position'] > -0.5).all() and (out['position'] < side + 0.5).all()      # walls hold the crowd
    assert (out['force'] == 0).all()


def test_million_three_circle_agents_500_steps_with_fluctuation():
    """Config 3 flavour as a soak: 1 M three-circle agents, adaptive dt, stochastic node on, 500 fused steps (graph replay
    on the sim's own stream): state stays finite and inside the walls, time bookkeeping adds up, agents keep their identity."""
    agents, obstacles, side = S.uniform_crowd(N, 'three_circle', density=1.0, seed=9)
    agents['std_rand_force'] = 0.1
    agents['std_rand_torque'] = 0.1
    dev = DeviceAgents(MODEL_THREE_CIRCLE)
    dev.upload(agents)
    dev.set_obstacles(obstacles)
    dev.set_navigation_field(0, *S.direction_field(2.0, (0, 0, side, side), 'exit', point=(side, side / 2)))
    dts = dev.step(500, _lib.STEP_ALL | _lib.STEP_FLUCTUATION, CELL, 0.001, 0.01)
    out = agents.copy()
    dev.download(out)
    t, it = dev.time()
    dev.close()
    assert it == 500 and abs(t - dts.sum()) < 1e-9 and ((dts >= 0.001) & (dts <= 0.01)).all()
    for f in ('position', 'velocity', 'orientation', 'angular_velocity', 'position_ls', 'position_rs'):
        assert np.isfinite(out[f]).all(), f
    assert (out['position'] > -0.5).all() and (out['position'] < side + 0.5).all()
    assert (np.abs(out['orientation']) <= np.pi).all()
    assert (out['mass'] == agents['mass']).all() and (out['radius'] == agents['radius']).all()     # records kept their rows
    moved = np.hypot(*(out['position'] - agents['position']).T)
    assert 0.5 < moved.mean() < 6.0                     # ~ 4 s of walking at ~ 1 m/s against a crowd
