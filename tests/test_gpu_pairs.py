"""The once-per-pair pipeline (kernel variant 3: forward-half-stencil sweep -> pair list -> one evaluation per pair ->
ordered per-agent sums, csrc/pair_kernels.cuh) against the oracle and against the both-sides fused kernel (variant 2), and
its "a step whose pairs do not fit is not applied, the host grows the list and repeats it" protocol."""
import numpy as np
import pytest

from conftest import rel_err_fields, vec_rel_err
from crowddynamics_b200 import _lib, synthetic as S
from crowddynamics_b200.engine import DeviceAgents
from crowddynamics_b200.structures import MODEL_CIRCULAR, MODEL_THREE_CIRCLE
from oracle import crowd_oracle as O

pytestmark = pytest.mark.gpu
CELL = 3.6
MODELS = ['circular', 'three_circle']


def _mid(model):
    return MODEL_CIRCULAR if model == 'circular' else MODEL_THREE_CIRCLE


def _run(model, agents, obstacles, fields, variant, steps, pair_cap=0, chunks=None, flags=_lib.STEP_ALL):
    dev = DeviceAgents(_mid(model))
    dev.set_variant(variant)
    dev.upload(agents)
    dev.set_obstacles(obstacles)
    for k, f in enumerate(fields):
        dev.set_navigation_field(k, *f)
    if pair_cap:
        dev.set_pair_capacity(pair_cap)
    dts = [dev.step(k, flags, CELL, 0.001, 0.01) for k in (chunks or [steps])]
    out = agents.copy()
    dev.download(out)
    t, it = dev.time()
    stats = dev.pair_stats() if variant == 3 else None
    dev.close()
    return np.concatenate(dts), out, t, it, stats


@pytest.mark.parametrize('model', MODELS)
@pytest.mark.parametrize('density', [0.3, 1.0, 3.0])
def test_single_step_forces_against_oracle(model, density):
    agents, obstacles, side = S.uniform_crowd(6000, model, density=density, seed=31, overlap_fraction=0.05 if density > 1 else 0.01)
    fields = [S.direction_field(0.5, (0, 0, side, side), 'swirl')]
    ref = agents.copy()
    O.navigation(ref, fields); O.orientation(ref); O.adjusting(ref)
    O.agent_agent_block_list(ref, CELL); O.agent_obstacle(ref, obstacles)
    _, got, _, _, stats = _run(model, agents, obstacles, fields, 3, 1, flags=_lib.STEP_ALL & ~(_lib.STEP_INTEGRATOR | _lib.STEP_RESET))
    assert vec_rel_err(got['force'], ref['force']) <= 1e-9
    if model == 'three_circle':
        assert vec_rel_err(got['torque'], ref['torque']) <= 1e-9
    assert stats[1] > 0 and stats[2] == 0          # pairs were listed, nothing had to be repeated


@pytest.mark.parametrize('model', MODELS)
def test_trajectory_bit_identical_to_both_sides_kernel(model):
    agents, obstacles, side = S.uniform_crowd(8000, model, density=1.5, seed=32, overlap_fraction=0.03)
    agents['std_rand_force'] = 0.1
    fields = [S.direction_field(0.5, (0, 0, side, side), 'swirl')]
    flags = _lib.STEP_ALL | _lib.STEP_FLUCTUATION
    a = _run(model, agents, obstacles, fields, 3, 0, chunks=[1, 6, 70], flags=flags)
    b = _run(model, agents, obstacles, fields, 2, 0, chunks=[1, 6, 70], flags=flags)
    assert (a[0] == b[0]).all() and a[2] == b[2] and a[3] == b[3] == 77
    assert rel_err_fields(a[1], b[1])[0] == 0


@pytest.mark.parametrize('model', MODELS)
def test_overflowing_pair_list_repeats_the_steps(model):
    """A pair list far too small for the crowd: the overflowing steps are not applied on the device, the host grows the list
    and repeats them -- the trajectory, the dt log, the time and the iteration count are the ones of an ample list."""
    agents, obstacles, side = S.uniform_crowd(5000, model, density=2.0, seed=33, overlap_fraction=0.03)
    agents['std_rand_force'] = 0.1
    fields = [S.direction_field(0.5, (0, 0, side, side), 'exit', point=(side, side / 2))]
    flags = _lib.STEP_ALL | _lib.STEP_FLUCTUATION
    ample = _run(model, agents, obstacles, fields, 3, 0, chunks=[3, 1, 40], flags=flags)
    tight = _run(model, agents, obstacles, fields, 3, 0, pair_cap=64, chunks=[3, 1, 40], flags=flags)
    assert ample[4][2] == 0 and tight[4][2] >= 1 and tight[4][0] >= tight[4][1]
    assert (ample[0] == tight[0]).all() and ample[2] == tight[2] and ample[3] == tight[3] == 44
    assert rel_err_fields(ample[1], tight[1])[0] == 0


@pytest.mark.parametrize('model', MODELS)
def test_overflow_in_the_node_wise_entry_point(model):
    agents, obstacles, side = S.uniform_crowd(3000, model, density=2.0, seed=34, overlap_fraction=0.05)
    ref = agents.copy()
    O.agent_agent_block_list(ref, CELL)
    dev = DeviceAgents(_mid(model))
    dev.upload(agents)
    dev.set_pair_capacity(16)
    dev.agent_agent(CELL)
    got = agents.copy()
    dev.download(got)
    cap, found, repeats = dev.pair_stats()
    dev.close()
    assert repeats >= 1 and cap >= found > 16
    assert vec_rel_err(got['force'], ref['force']) <= 1e-9
    if model == 'three_circle':
        assert vec_rel_err(got['torque'], ref['torque']) <= 1e-9


@pytest.mark.parametrize('model', MODELS)
def test_dense_jam_grows_the_list_on_its_own(model):
    """55 agents / m^2: ~100x the pairs per agent of an ordinary crowd -- the automatic capacity overflows once and is grown."""
    agents, obstacles, _ = S.random_crowd(2000, model, half_width=3.0, seed=35)
    ref = agents.copy()
    O.adjusting(ref); O.agent_agent_block_list(ref, CELL); O.agent_obstacle(ref, obstacles)
    dev = DeviceAgents(_mid(model))
    dev.upload(agents)
    dev.set_obstacles(obstacles)
    dev.step(1, _lib.STEP_ADJUSTING | _lib.STEP_AGENT_AGENT | _lib.STEP_AGENT_OBSTACLE, CELL, 0.01, 0.01, want_dt=False)
    got = agents.copy()
    dev.download(got)
    cap, found, repeats = dev.pair_stats()
    dev.close()
    assert cap >= found
    assert vec_rel_err(got['force'], ref['force']) <= 1e-9
    if model == 'three_circle':
        assert vec_rel_err(got['torque'], ref['torque']) <= 1e-9


# ---- search lattice refinement (cells of cell_size / 2, reach 2) -----------------------------------------------------------
def _forces(model, agents, obstacles, cell, refinement, variant=3):
    dev = DeviceAgents(_mid(model))
    dev.set_variant(variant)
    dev.set_search_refinement(refinement)
    dev.upload(agents)
    dev.set_obstacles(obstacles)
    dev.step(1, _lib.STEP_ADJUSTING | _lib.STEP_ORIENTATION | _lib.STEP_AGENT_AGENT | _lib.STEP_AGENT_OBSTACLE, cell, 0.01, 0.01,
             want_dt=False)
    out = agents.copy()
    dev.download(out)
    # the exports keep describing the cell_size lattice, whatever the search used
    dev.build_block_list(cell)
    grid = dev.grid()
    dev.close()
    return out, grid


@pytest.mark.parametrize('model', MODELS)
@pytest.mark.parametrize('cell', [3.6, 5.0])
def test_refined_search_lattice_finds_the_same_pairs(model, cell):
    agents, obstacles, side = S.uniform_crowd(7000, model, density=1.3, seed=41, overlap_fraction=0.03)
    ref = agents.copy()
    O.orientation(ref); O.adjusting(ref); O.agent_agent_block_list(ref, cell); O.agent_obstacle(ref, obstacles)
    coarse, g1 = _forces(model, agents, obstacles, cell, 1)
    fine, g2 = _forces(model, agents, obstacles, cell, 0)
    both_sides, _ = _forces(model, agents, obstacles, cell, 0, variant=2)
    assert g1 == g2
    for got in (coarse, fine):
        assert vec_rel_err(got['force'], ref['force']) <= 1e-9
        if model == 'three_circle':
            assert vec_rel_err(got['torque'], ref['torque']) <= 1e-9
    # same pair set, only the summation order differs between the two lattices
    assert vec_rel_err(fine['force'], coarse['force']) <= 1e-12
    assert rel_err_fields(fine, both_sides)[0] == 0


@pytest.mark.parametrize('model', MODELS)
def test_refinement_is_refused_when_pairs_can_interact_beyond_a_cell(model):
    """Agents twice as wide (3 + 2 R > cell_size) or a small cell_size: pairs of non-adjacent cells can be closer than the
    interaction range, the reference's block list does not see them, and neither may we -- the search must stay on the
    cell_size lattice."""
    agents, obstacles, side = S.uniform_crowd(3000, model, density=0.5, seed=42)
    for f in ('radius', 'r_t', 'r_s', 'r_ts'):
        if f in agents.dtype.names:
            agents[f] *= 2.2
    if model == 'three_circle':
        S.set_shoulders(agents)
    for cell in (3.6, 2.0):
        ref = agents.copy()
        O.adjusting(ref); O.agent_agent_block_list(ref, cell); O.agent_obstacle(ref, obstacles)
        brute = agents.copy()
        O.adjusting(brute); O.agent_agent_brute(brute); O.agent_obstacle(brute, obstacles)
        got, _ = _forces(model, agents, obstacles, cell, 0)
        assert vec_rel_err(got['force'], ref['force']) <= 1e-9
        if cell == 2.0:
            assert vec_rel_err(ref['force'], brute['force']) > 1e-6     # the block list really misses pairs here
