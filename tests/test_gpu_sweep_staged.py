"""k_sweep_staged (csrc/pair_kernels.cuh): the classification sweep with its candidates staged in shared memory by TMA bulk
copies (`cp.async.bulk` + mbarrier) must list exactly the pairs the plain sweep lists -- the per-agent sums are added in
ascending partner order whatever the order of the list, so whole trajectories have to be BIT-identical between the two.
The kernel is selected per sim at creation time (environment variable CROWD_B200_SWEEP = plain | staged)."""
import numpy as np
import pytest

from crowddynamics_b200 import _lib, synthetic as S
from crowddynamics_b200.engine import DeviceAgents
from crowddynamics_b200.structures import MODEL_CIRCULAR, MODEL_THREE_CIRCLE
from oracle import crowd_oracle as O

pytestmark = pytest.mark.gpu
CELL = 3.6
MODELS = ['circular', 'three_circle']
MODES = ['staged']
FIELDS = ['position', 'velocity', 'force', 'force_prev', 'target_direction']
FIELDS3 = FIELDS + ['orientation', 'angular_velocity', 'torque', 'torque_prev', 'position_ls', 'position_rs']


def _run(monkeypatch, mode, model, agents, obstacles, fields, chunks, policy, flags=_lib.STEP_ALL, dts=(0.001, 0.01)):
    monkeypatch.setenv('CROWD_B200_SWEEP', mode)
    dev = DeviceAgents(MODEL_CIRCULAR if model == 'circular' else MODEL_THREE_CIRCLE)
    dev.set_small_crowd_max(0)                    # always the general pipeline
    dev.set_rebuild_policy(*policy)
    dev.upload(agents)
    dev.set_obstacles(obstacles)
    for k, f in enumerate(fields):
        dev.set_navigation_field(k, *f)
    dt = np.concatenate([np.atleast_1d(dev.step(k, flags, CELL, dts[0], dts[1])) for k in chunks])
    out = agents.copy()
    dev.download(out)
    stats = dev.rebuild_stats()
    dev.close()
    return out, dt, stats


def _same(a, b, model):
    for name in (FIELDS3 if model == 'three_circle' else FIELDS):
        assert np.array_equal(a[name], b[name]), name


@pytest.mark.parametrize('mode', MODES)
@pytest.mark.parametrize('model', MODELS)
@pytest.mark.parametrize('n,density', [(1, 1.0), (2, 1.0), (129, 1.0), (5000, 0.125), (30000, 1.0), (30000, 2.5)],
                         ids=['n1', 'n2', 'n129', 'sparse', 'rho1', 'rho2.5'])
def test_staged_equals_plain_rebuilding_every_step(monkeypatch, mode, model, n, density):
    agents, obstacles, side = S.uniform_crowd(n, model, density=density, seed=7 + n, overlap_fraction=0.02)
    fields = [S.direction_field(0.5, (0, 0, side, side), 'swirl')]
    ref, dt_ref, _ = _run(monkeypatch, 'plain', model, agents, obstacles, fields, [1, 3], (0.10, 1, 0))
    got, dt, _ = _run(monkeypatch, mode, model, agents, obstacles, fields, [1, 3], (0.10, 1, 0))
    assert np.array_equal(dt, dt_ref)
    _same(got, ref, model)


@pytest.mark.parametrize('mode', MODES)
@pytest.mark.parametrize('model', MODELS)
def test_staged_equals_plain_on_kept_block_lists(monkeypatch, mode, model):
    """resident-order steps: the staged sweep reads the records k_finish wrote in place on the step before"""
    agents, obstacles, side = S.uniform_crowd(40000, model, density=1.0, seed=51)
    fields = [S.direction_field(0.5, (0, 0, side, side), 'swirl')]
    ref, dt_ref, st_ref = _run(monkeypatch, 'plain', model, agents, obstacles, fields, [2, 10], (0.10, 16, 0))
    got, dt, st = _run(monkeypatch, mode, model, agents, obstacles, fields, [2, 10], (0.10, 16, 0))
    assert st == st_ref and st['kept'] >= 6 and st['stale'] == 0, (st, st_ref)
    assert np.array_equal(dt, dt_ref)
    _same(got, ref, model)


@pytest.mark.parametrize('model', MODELS)
def test_staged_with_density_jumps_falls_back_per_column(monkeypatch, model):
    """A sparse crowd next to a 30x denser one: CTAs at the border find hulls beyond the staging capacity in the forward
    columns and sweep those from global memory; cell_size lattice (reach 1) and the finer one (reach 2)."""
    rng = np.random.default_rng(3)
    a, obstacles, side = S.uniform_crowd(6000, model, density=0.1, seed=11)
    b, _, side_b = S.uniform_crowd(24000, model, density=3.0, seed=12, overlap_fraction=0.05)
    b = b.copy()
    for name in ('position',) + (('position_ls', 'position_rs') if model == 'three_circle' else ()):
        b[name][:, 0] += side + 0.3                  # the dense block starts where the sparse one ends
    agents = np.concatenate([a, b])
    agents = agents[rng.permutation(len(agents))]
    for refinement in (0, 1):
        outs = []
        for mode in ('plain', 'staged'):
            monkeypatch.setenv('CROWD_B200_SWEEP', mode)
            dev = DeviceAgents(MODEL_CIRCULAR if model == 'circular' else MODEL_THREE_CIRCLE)
            dev.set_search_refinement(refinement)
            dev.set_rebuild_policy(0.10, 1, 0)
            dev.upload(agents)
            dev.step(2, _lib.STEP_ALL & ~_lib.STEP_NAVIGATION & ~_lib.STEP_AGENT_OBSTACLE, CELL, 0.001, 0.01)
            out = agents.copy()
            dev.download(out)
            dev.close()
            outs.append(out)
        _same(outs[1], outs[0], model)


@pytest.mark.parametrize('mode', MODES)
@pytest.mark.parametrize('model', MODELS)
def test_staged_single_step_against_oracle(monkeypatch, mode, model):
    agents, obstacles, side = S.uniform_crowd(8000, model, density=1.2, seed=23, overlap_fraction=0.03)
    ref = agents.copy()
    O.agent_agent_block_list(ref, CELL)
    got, _, _ = _run(monkeypatch, mode, model, agents, obstacles, [], [1], (0.10, 1, 0), flags=_lib.STEP_AGENT_AGENT)
    scale = np.abs(ref['force']).max()
    assert np.abs(got['force'] - ref['force']).max() <= 1e-9 * scale
    if model == 'three_circle':
        assert np.abs(got['torque'] - ref['torque']).max() <= 1e-9 * max(1.0, np.abs(ref['torque']).max())
