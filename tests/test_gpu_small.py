"""Small crowds (<= 256 agents; the size of the reference's example simulations, examples/simulations.py:74-163) are advanced by
ONE thread block that keeps the crowd in shared memory and runs all steps of a cdb_step call in one launch
(csrc/small_kernel.cuh).  It must give what the general pipeline and the oracle give."""
import numpy as np
import pytest

from conftest import vec_rel_err
from crowddynamics_b200 import _lib, synthetic as S
from crowddynamics_b200.engine import DeviceAgents
from crowddynamics_b200.structures import MODEL_CIRCULAR, MODEL_THREE_CIRCLE
from oracle import crowd_oracle as O

pytestmark = pytest.mark.gpu
CELL = 3.6
MODELS = ['circular', 'three_circle']


def _run(model, agents, obstacles, fields, chunks, small, flags=_lib.STEP_ALL, dts=(0.001, 0.01), seed=None):
    dev = DeviceAgents(MODEL_CIRCULAR if model == 'circular' else MODEL_THREE_CIRCLE)
    dev.set_small_crowd_max(256 if small else 0)
    dev.upload(agents)
    dev.set_obstacles(obstacles)
    for k, f in enumerate(fields):
        dev.set_navigation_field(k, *f)
    if seed is not None:
        dev.set_seed(seed)
    l0 = dev.launch_count()
    dt = np.concatenate([dev.step(k, flags, CELL, dts[0], dts[1]) for k in chunks])
    launches = dev.launch_count() - l0
    out = agents.copy()
    dev.download(out)
    t, it = dev.time()
    dev.close()
    return out, dt, t, it, launches


@pytest.mark.parametrize('model', MODELS)
@pytest.mark.parametrize('n,density', [(1, 1.0), (2, 1.0), (7, 2.0), (50, 1.0), (200, 0.5), (256, 3.0)])
def test_small_kernel_equals_general_path(model, n, density):
    agents, obstacles, side = S.uniform_crowd(n, model, density=density, seed=70 + n, overlap_fraction=0.05 if n > 20 else 0.0)
    agents['std_rand_force'] = 0.1
    fields = [S.direction_field(0.25, (0, 0, side, side), 'swirl')]
    flags = _lib.STEP_ALL | _lib.STEP_FLUCTUATION
    ref, dt_ref, t_ref, it_ref, l_ref = _run(model, agents, obstacles, fields, [1, 5], False, flags, seed=5)
    got, dt, t, it, l = _run(model, agents, obstacles, fields, [1, 5], True, flags, seed=5)
    assert l == 2 and l_ref > 10 * l                           # one launch per cdb_step call
    assert it == it_ref == 6 and np.abs(dt - dt_ref).max() <= 1e-15 and abs(t - t_ref) <= 1e-14
    for name, tol in (('position', 1e-10), ('velocity', 1e-8), ('target_direction', 0.0), ('force_prev', None)):
        d = float(np.abs(got[name] - ref[name]).max())
        assert d <= (tol if tol is not None else 1e-8 * max(1.0, np.abs(ref[name]).max())), (name, d)
    assert (got['force'] == 0).all()                            # Reset ran
    if model == 'three_circle':
        assert np.abs(got['orientation'] - ref['orientation']).max() <= 1e-9
        assert np.abs(got['position_ls'] - ref['position_ls']).max() <= 1e-9


@pytest.mark.parametrize('model', MODELS)
def test_single_step_forces_against_oracle(model):
    agents, obstacles, side = S.uniform_crowd(180, model, density=1.5, seed=81, overlap_fraction=0.05)
    fields = [S.direction_field(0.5, (0, 0, side, side), 'swirl')]
    ref = agents.copy()
    O.navigation(ref, fields); O.orientation(ref); O.adjusting(ref)
    O.agent_agent_block_list(ref, CELL); O.agent_obstacle(ref, obstacles)
    got, _, _, _, launches = _run(model, agents, obstacles, fields, [1], True, _lib.STEP_ALL & ~(_lib.STEP_INTEGRATOR | _lib.STEP_RESET))
    assert launches == 1
    assert vec_rel_err(got['force'], ref['force']) <= 1e-9
    if model == 'three_circle':
        assert vec_rel_err(got['torque'], ref['torque']) <= 1e-9
    assert (got['position'] == agents['position']).all()


@pytest.mark.parametrize('model', MODELS)
def test_trajectory_against_oracle_and_long_calls(model):
    """12 adaptive updates vs the oracle; then a 700-step call (several passes over the dt ring) returns 700 dts that sum to
    the simulated time."""
    agents, obstacles, side = S.uniform_crowd(120, model, density=1.0, seed=82, overlap_fraction=0.02)
    fields = [S.direction_field(0.5, (0, 0, side, side), 'swirl')]
    ref = agents.copy()
    dts_ref = [O.step(ref, obstacles, fields, CELL, 0.001, 0.01) for _ in range(12)]
    got, dt, t, it, _ = _run(model, agents, obstacles, fields, [12], True)
    assert np.abs(dt - np.array(dts_ref)).max() <= 1e-13
    assert np.abs(got['position'] - ref['position']).max() <= 1e-7
    got, dt, t, it, launches = _run(model, agents, obstacles, fields, [700], True)
    assert it == 700 and len(dt) == 700 and abs(dt.sum() - t) <= 1e-9 and (dt > 0).all() and launches <= 8


def test_general_path_is_kept_where_the_pair_set_depends_on_the_lattice():
    """cell_size 3.0 < 3 + 2 R: agents in range of each other may sit in non-adjacent cells, which the reference then does NOT
    pair -- the all-pairs kernel would; the general path has to run (many launches), and equals the oracle."""
    agents, obstacles, side = S.uniform_crowd(100, 'circular', density=1.0, seed=83)
    ref = agents.copy()
    O.agent_agent_block_list(ref, 3.0)
    dev = DeviceAgents(MODEL_CIRCULAR)
    dev.upload(agents)
    l0 = dev.launch_count()
    dev.step(1, _lib.STEP_AGENT_AGENT, 3.0, 0.01, 0.01, want_dt=False)
    assert dev.launch_count() - l0 > 5
    got = agents.copy()
    dev.download(got)
    dev.close()
    assert vec_rel_err(got['force'], ref['force']) <= 1e-9
