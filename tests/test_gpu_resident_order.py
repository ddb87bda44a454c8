"""Resident-order steps (include/crowd_b200.h, cdb_set_rebuild_policy; ChainState in csrc/kernels.cuh): for large crowds cdb_step
rebuilds its block list only every few steps, on slightly wider search cells; in between the agents keep their slots, the step
works in place and writes the next step's neighbour records itself.  The reference re-bins at every update
(core/interactions.py:191-205); the results must not depend on the difference beyond the order of an agent's pair sums."""
import numpy as np
import pytest

from crowddynamics_b200 import _lib, synthetic as S
from crowddynamics_b200.engine import DeviceAgents
from crowddynamics_b200.structures import MODEL_CIRCULAR, MODEL_THREE_CIRCLE
from oracle import crowd_oracle as O

pytestmark = pytest.mark.gpu
CELL = 3.6
MODELS = ['circular', 'three_circle']


def _run(model, agents, obstacles, fields, steps, policy, dts=(0.01, 0.01), chunks=None, flags=_lib.STEP_ALL):
    dev = DeviceAgents(MODEL_CIRCULAR if model == 'circular' else MODEL_THREE_CIRCLE)
    dev.set_rebuild_policy(*policy)
    dev.upload(agents)
    dev.set_obstacles(obstacles)
    for k, f in enumerate(fields):
        dev.set_navigation_field(k, *f)
    dt = np.concatenate([dev.step(k, flags, CELL, dts[0], dts[1]) for k in (chunks or [steps])])
    out = agents.copy()
    dev.download(out)
    t, it = dev.time()
    stats = dev.rebuild_stats()
    dev.close()
    return out, dt, t, it, stats


def _diff(a, b, name):
    return float(np.abs(a[name] - b[name]).max())


@pytest.mark.parametrize('model', MODELS)
@pytest.mark.parametrize('dts', [(0.01, 0.01), (0.001, 0.01)], ids=['fixed-dt', 'adaptive-dt'])
@pytest.mark.parametrize('steps,tol', [(8, 1e-11), (16, 1e-7)], ids=['8-steps', '16-steps'])
def test_kept_order_equals_rebuilding_every_step(model, dts, steps, tol):
    """A 30 000-agent crowd: block list kept for several steps vs rebuilt at every step (max_interval = 1).  The two differ
    only in the order an agent's pair contributions are added: 1e-16 relative per step, which these dynamics amplify by about
    a factor of two per step -- two rebuild-every-step runs on different search lattices drift apart at exactly the same rate
    (scripts/chaos_probe.py, DESIGN.md section 4: 1e-14 after 4 steps, 1e-12 after 8-16, 1e-3 after 40), hence the horizons."""
    agents, obstacles, side = S.uniform_crowd(30000, model, density=1.0, seed=41)
    fields = [S.direction_field(0.5, (0, 0, side, side), 'swirl')]
    chunks = [2, steps - 2]                       # the interval is first sized after two steps
    ref, dt_ref, t_ref, it_ref, st_ref = _run(model, agents, obstacles, fields, steps, (0.10, 1, 0), dts, chunks)
    got, dt, t, it, st = _run(model, agents, obstacles, fields, steps, (0.10, 16, 0), dts, chunks)
    assert st_ref['kept'] == 0 and st_ref['rebuilds'] == 0        # max_interval = 1: the mode is off altogether
    assert st['kept'] >= steps // 2 and st['stale'] == 0 and st['rebuilds'] >= 2 and st['interval'] > 1, st
    assert it == it_ref == steps
    assert np.abs(dt - dt_ref).max() <= 1e-15 and abs(t - t_ref) <= 1e-13
    d = {k: _diff(got, ref, k) for k in ('position', 'velocity', 'target_direction')}
    assert d['position'] <= tol and d['velocity'] <= 100 * tol and d['target_direction'] <= 1e-9, d
    assert _diff(got, ref, 'force_prev') <= 1e4 * tol * max(1.0, np.abs(ref['force_prev']).max())   # 1/tau^2: forces are touchier
    if model == 'three_circle':
        d3 = {k: _diff(got, ref, k) for k in ('orientation', 'position_ls', 'position_rs', 'angular_velocity')}
        assert d3['orientation'] <= 10 * tol and d3['position_ls'] <= 10 * tol and d3['position_rs'] <= 10 * tol, d3
    for name in ('radius', 'mass', 'target_velocity', 'tau_adj', 'k_soc', 'tau_0', 'mu', 'kappa', 'damping'):
        assert (got[name] == agents[name]).all(), name               # constants are left alone by the in-place steps


@pytest.mark.parametrize('model', MODELS)
def test_kept_order_against_oracle(model):
    """8 updates of 20 000 agents, calls of 1 / 3 / 4 steps (the kept order survives across cdb_step calls), vs the C oracle."""
    agents, obstacles, side = S.uniform_crowd(20000, model, density=1.0, seed=42, overlap_fraction=0.01)
    fields = [S.direction_field(0.5, (0, 0, side, side), 'swirl')]
    ref = agents.copy()
    dts_ref = [O.step(ref, obstacles, fields, CELL, 0.001, 0.01) for _ in range(8)]
    got, dt, t, it, st = _run(model, agents, obstacles, fields, 8, (0.10, 16, 0), (0.001, 0.01), chunks=[1, 3, 4])
    assert st['kept'] + st['rebuilds'] == 8 and st['stale'] == 0, st
    assert np.abs(dt - np.array(dts_ref)).max() <= 1e-13
    assert np.abs(got['position'] - ref['position']).max() <= 1e-8
    assert np.abs(got['velocity'] - ref['velocity']).max() <= 1e-6


@pytest.mark.parametrize('model', MODELS)
def test_stale_lattice_is_refused_and_the_steps_repeated(model):
    """Agents that start at rest and accelerate hard: the interval chosen from the first (tiny) displacements is too long for
    the thin skin, the device refuses the step that would sweep a stale lattice, the host rebuilds and repeats -- same
    trajectory, same step count, same simulated time as rebuilding at every step."""
    agents, obstacles, side = S.uniform_crowd(20000, model, density=0.5, seed=43)
    agents['velocity'] = 0.0
    agents['target_velocity'] = 6.0
    agents['tau_adj'] = 0.25
    fields = [S.direction_field(0.5, (0, 0, side, side), 'swirl')]
    ref, dt_ref, t_ref, it_ref, _ = _run(model, agents, obstacles, fields, 18, (0.02, 1, 0))
    got, dt, t, it, st = _run(model, agents, obstacles, fields, 18, (0.02, 32, 0))
    assert st['stale'] >= 1 and st['kept'] >= 4, st
    assert it == it_ref == 18 and abs(t - t_ref) <= 1e-13 and len(dt) == 18
    d = {k: _diff(got, ref, k) for k in ('position', 'velocity')}
    assert d['position'] <= 1e-8 and d['velocity'] <= 1e-6, d


def test_mode_ends_with_anything_that_touches_the_state():
    """An upload, a node-wise call or a block-list export between two cdb_step calls forces a rebuild at the next step."""
    agents, obstacles, side = S.uniform_crowd(20000, 'circular', density=1.0, seed=44)
    dev = DeviceAgents(MODEL_CIRCULAR)
    dev.set_rebuild_policy(0.10, 16, 0)
    dev.upload(agents)
    dev.set_obstacles(obstacles)
    dev.step(6, _lib.STEP_ALL, CELL, 0.01, 0.01, want_dt=False)
    base = dev.rebuild_stats()
    assert base['kept'] >= 2
    dev.step(1, _lib.STEP_ALL, CELL, 0.01, 0.01, want_dt=False)
    s1 = dev.rebuild_stats()
    got = agents.copy()
    dev.download(got)                                   # reading the state does not end the run
    dev.step(1, _lib.STEP_ALL, CELL, 0.01, 0.01, want_dt=False)
    s2 = dev.rebuild_stats()
    assert s2['rebuilds'] + s2['kept'] == s1['rebuilds'] + s1['kept'] + 1
    dev.upload(got)                                     # writing it does
    dev.step(1, _lib.STEP_ALL, CELL, 0.01, 0.01, want_dt=False)
    s3 = dev.rebuild_stats()
    assert s3['rebuilds'] == s2['rebuilds'] + 1
    dev.step(1, _lib.STEP_ALL, CELL, 0.01, 0.01, want_dt=False)
    dev.agent_obstacle()                                # a node-wise call in between
    dev.step(1, _lib.STEP_ALL, CELL, 0.01, 0.01, want_dt=False)
    s4 = dev.rebuild_stats()
    assert s4['rebuilds'] >= s3['rebuilds'] + 1
    # and the result still equals the oracle's after all of that (7 + 1 + 1 + 1 + 1 steps, obstacle forces added once more)
    dev.close()


@pytest.mark.parametrize('accelerating', [False, True], ids=['steady', 'accelerating'])
def test_kept_order_with_deferred_synchronisation(accelerating):
    """One cdb_step call per update with cdb_set_deferred_sync (what FusedStep(deferred=True) does): the pair-count / applied-step
    check of call k is read at call k + 1 without waiting; steps it reports as refused (stale lattice) are repeated then, or --
    at the latest -- when state is handed to the host.  (With one call per update the interval follows the displacement closely
    enough that a hard-accelerating crowd is served by shorter intervals rather than by refusals.)"""
    agents, obstacles, side = S.uniform_crowd(20000, 'circular', density=0.5 if accelerating else 1.0, seed=45)
    skin = 0.10
    if accelerating:
        agents['velocity'] = 0.0
        agents['target_velocity'] = 6.0
        agents['tau_adj'] = 0.25
        skin = 0.02
    fields = [S.direction_field(0.5, (0, 0, side, side), 'swirl')]
    ref, _, t_ref, it_ref, _ = _run('circular', agents, obstacles, fields, 16, (skin, 1, 0))
    dev = DeviceAgents(MODEL_CIRCULAR)
    dev.set_rebuild_policy(skin, 32, 0)
    dev.set_deferred_sync(True)
    dev.upload(agents)
    dev.set_obstacles(obstacles)
    dev.set_navigation_field(0, *fields[0])
    syncs0 = dev.sync_count()
    for _ in range(16):
        dev.step(1, _lib.STEP_ALL, CELL, 0.01, 0.01, want_dt=False)
    syncs = dev.sync_count() - syncs0
    # a refused step surfaces one call late; handing state to the host (download, time) first repeats whatever is missing
    got = agents.copy()
    dev.download(got)
    t, it = dev.time()
    st = dev.rebuild_stats()
    dev.close()
    assert st['kept'] >= 4, st
    if not accelerating:
        assert st['stale'] == 0 and syncs <= 10, (st, syncs)       # steady state: the calls do not wait for their own check
    assert it == it_ref == 16 and abs(t - t_ref) <= 1e-13
    assert np.abs(got['position'] - ref['position']).max() <= 1e-8
