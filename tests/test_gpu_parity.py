"""Parity of the CUDA path (through the C ABI / the drop-in Python boundary) against the CPU oracle and the golden
vectors of the reference.  Needs a GPU: run with -m gpu on the B200 box.

Bars (BASELINE.json north_star):
  * block-list cell assignment, tables and neighbour (candidate-pair) sets: bit exact;
  * single-step per-agent force / torque: within 1e-9 relative (fp64);
  * short-horizon trajectories: positions within 1e-7 m after 5 adaptive steps / 1e-6 m after 200 Hallway steps
    (the dynamics are chaotic; differences start at the 1e-16 level from summation order and libm).
"""
import numpy as np
import pytest

from conftest import load_golden, from_raw, rel_err_fields, vec_rel_err
from crowddynamics_b200 import _lib, synthetic as S
from crowddynamics_b200.engine import DeviceAgents
from crowddynamics_b200.exceptions import InvalidType, InvalidValue
from crowddynamics_b200.structures import (agent_type_circular, agent_type_three_circle, obstacle_type_linear,
                                           MODEL_CIRCULAR, MODEL_THREE_CIRCLE)
from crowddynamics_b200.core import interactions as gi, integrator as gint
from crowddynamics_b200.core.motion import adjusting as gadj
from crowddynamics_b200.core.steering import navigation as gnav, orientation as gori
from oracle import crowd_oracle as O

pytestmark = pytest.mark.gpu

DT = {'circular': agent_type_circular, 'three_circle': agent_type_three_circle}
FORCE_TOL = 1e-9
CELL = 3.6


def _fields(g):
    mg = S.MeshGrid(float(g['field_step']), *g['field_bounds'])
    return [(mg, (g['U'][t], g['V'][t])) for t in range(len(g['U']))]


def _obs(g):
    return np.ascontiguousarray(g['obstacles']).view(obstacle_type_linear).reshape(-1)


def _assert_forces(a, ref, model, tol=FORCE_TOL):
    assert vec_rel_err(a['force'], ref['force']) <= tol
    if model == 'three_circle':
        assert vec_rel_err(a['torque'], ref['torque']) <= tol


# ---- block list: bit exact --------------------------------------------------------------------------------------------
def _crowds():
    yield S.uniform_crowd(3000, 'circular', density=1.0, seed=0)[0]
    yield S.uniform_crowd(1500, 'three_circle', density=2.0, seed=1, overlap_fraction=0.05)[0]
    yield S.random_crowd(1000, 'circular', seed=2)[0]                      # negative coordinates, 0.125 agents/m^2
    yield S.uniform_crowd(700, 'circular', density=0.05, seed=3, origin=(-40.0, 13.0))[0]   # mostly empty cells
    a = S.random_crowd(64, 'circular', half_width=1.0, seed=4)[0]          # everything in <= 4 cells
    yield a
    b = S.random_crowd(50, 'three_circle', seed=5)[0]
    b['position'][:, 1] = 0.5                                            # a single row of cells
    yield b


@pytest.mark.parametrize('k', range(6))
def test_block_list_bit_exact(k):
    a = list(_crowds())[k]
    cl = O.add_to_cells(a, CELL)
    pi, cc, co, gs = gi.block_list(a, CELL)
    assert tuple(gs) == tuple(cl['grid'][2:])
    assert (pi == cl['points_indices']).all()
    assert (cc == cl['cells_count']).all()
    assert (co == cl['cells_offset']).all()
    dev = DeviceAgents(MODEL_CIRCULAR if a.dtype.itemsize == 228 else MODEL_THREE_CIRCLE)
    dev.upload(a)
    dev.build_block_list(CELL)
    assert dev.grid() == cl['grid']
    assert (dev.cell_ids() == cl['cell_of_agent']).all()
    pairs = dev.neighbor_pairs()
    ref = O.neighbor_pairs(a, CELL)
    assert pairs.shape == ref.shape
    key = lambda p: p[np.lexsort((p[:, 1], p[:, 0]))]
    assert (key(pairs) == key(ref)).all()          # same ordered (i, j) pairs, i.e. same sets and same orientation
    dev.close()


@pytest.mark.parametrize('model', ['circular', 'three_circle'])
@pytest.mark.parametrize('n', [0, 1, 2])
def test_tiny_crowds(model, n):
    """reference core/tests/test_interactions.py:42-55: block list with 0-2 agents must simply work."""
    a, obs, _ = S.random_crowd(n, model, half_width=0.5, seed=n)
    ref = a.copy()
    O.agent_agent_block_list(ref, CELL); O.agent_obstacle(ref, obs)
    gi.agent_agent_block_list(a, CELL); gi.agent_obstacle(a, obs)
    _assert_forces(a, ref, model)
    d_ref = O.velocity_verlet_integrator(ref, 0.001, 0.01)
    d = gint.velocity_verlet_integrator(a, 0.001, 0.01)
    assert d == d_ref
    assert rel_err_fields(a, ref)[0] <= 1e-12


# ---- single-step forces -----------------------------------------------------------------------------------------------
@pytest.mark.parametrize('model', ['circular', 'three_circle'])
@pytest.mark.parametrize('seed,density,overlap', [(0, 1.0, 0.0), (1, 1.0, 0.05), (2, 0.125, 0.0), (3, 3.0, 0.1)])
def test_agent_agent_forces(model, seed, density, overlap):
    n = 4000 if model == 'circular' else 2000
    a, _, _ = S.uniform_crowd(n, model, density=density, seed=seed, overlap_fraction=overlap)
    a['force'] = np.random.default_rng(seed).normal(0, 50, (n, 2))     # pre-existing force must be kept and added to
    ref = a.copy()
    O.agent_agent_block_list(ref, CELL)
    gi.agent_agent_block_list(a, CELL)
    _assert_forces(a, ref, model)
    untouched = [f for f in a.dtype.names if f not in ('force', 'torque')]
    assert rel_err_fields(a, ref, untouched)[0] == 0


@pytest.mark.parametrize('model', ['circular', 'three_circle'])
def test_reference_benchmark_workload(model):
    """The reference's own benchmark crowd (core/tests/test_interactions_benchmark.py:10-33), N = 1000."""
    a, _, _ = S.random_crowd(1000, model, seed=42)
    ref = a.copy()
    O.agent_agent_block_list(ref, CELL)
    gi.agent_agent_block_list(a, CELL)
    _assert_forces(a, ref, model)


@pytest.mark.parametrize('model', ['circular', 'three_circle'])
def test_agent_obstacle_forces(model):
    a, obs, side = S.uniform_crowd(3000, model, density=1.0, seed=5)
    # push a band of agents into the walls so that h < 0 happens, plus a degenerate and a far segment
    a['position'][:200, 0] = np.random.default_rng(1).uniform(-0.1, 0.2, 200)
    if model != 'circular':
        S.set_shoulders(a)
    extra = np.zeros(3, dtype=obstacle_type_linear)
    extra[0]['p0'] = extra[0]['p1'] = (1.0, 1.0)
    extra[1]['p0'], extra[1]['p1'] = (side / 2, side / 3), (side / 2 + 4.0, side / 3 + 1.0)
    extra[2]['p0'], extra[2]['p1'] = (1e6, 1e6), (1e6 + 1, 1e6)
    obs = np.concatenate((obs, extra))
    ref = a.copy()
    O.agent_obstacle(ref, obs)
    gi.agent_obstacle(a, obs)
    assert (np.abs(ref['force']).sum(1) > 0).sum() > 50
    _assert_forces(a, ref, model, tol=1e-12)


# ---- golden vectors of the reference, node by node ------------------------------------------------------------------------
@pytest.mark.parametrize('name', ['step_%s.npz', 'step_sparse_%s.npz'])
@pytest.mark.parametrize('model', ['circular', 'three_circle'])
def test_nodes_against_reference_golden(name, model):
    g = load_golden(name % model)
    dt = DT[model]
    fields, obs, cell = _fields(g), _obs(g), float(g['cell_size'])
    a = from_raw(g['initial'], dt)
    gnav.navigate(a, fields)
    assert rel_err_fields(a, from_raw(g['after_navigation'], dt))[0] == 0          # pure gather: bit exact
    gori.orient_towards_target_direction(a) if model == 'three_circle' else None
    assert rel_err_fields(a, from_raw(g['after_orientation'], dt))[0] <= 1e-14
    a = from_raw(g['after_orientation'], dt)
    gadj.adjust_agents(a)
    ref = from_raw(g['after_adjusting'], dt)
    _assert_forces(a, ref, model, tol=1e-13)
    a = ref.copy()
    pi, cc, co, gs = gi.block_list(a, cell)
    assert (pi == g['points_indices']).all() and (cc == g['cells_count']).all() and (co == g['cells_offset']).all()
    assert tuple(gs) == tuple(g['grid_shape'])
    gi.agent_agent_block_list(a, cell)
    ref = from_raw(g['after_agent_agent'], dt)
    _assert_forces(a, ref, model)
    a = ref.copy()
    gi.agent_obstacle(a, obs)
    ref = from_raw(g['after_agent_obstacle'], dt)
    _assert_forces(a, ref, model, tol=1e-12)
    a = ref.copy()
    d = gint.velocity_verlet_integrator(a, float(g['dt_min']), float(g['dt_max']))
    assert abs(d - g['dts'][0]) <= 1e-15 * abs(g['dts'][0])
    ref = from_raw(g['after_integrator'], dt)
    worst, f = rel_err_fields(a, ref)
    assert worst <= 1e-12, (worst, f)


@pytest.mark.parametrize('model', ['circular', 'three_circle'])
def test_five_steps_against_reference_golden(model):
    """Short-horizon trajectory: 5 adaptive-dt steps of the whole replaced sub-tree, resident on the device."""
    g = load_golden('step_%s.npz' % model)
    dt = DT[model]
    a = from_raw(g['initial'], dt)
    dev = DeviceAgents(MODEL_CIRCULAR if model == 'circular' else MODEL_THREE_CIRCLE)
    dev.upload(a)
    dev.set_obstacles(_obs(g))
    for t, (mg, uv) in enumerate(_fields(g)):
        dev.set_navigation_field(t, mg, uv)
    dts = dev.step(5, _lib.STEP_ALL, float(g['cell_size']), float(g['dt_min']), float(g['dt_max']))
    dev.download(a)
    dev.close()
    ref = from_raw(g['after_5_steps'], dt)
    np.testing.assert_allclose(dts, g['dts'], rtol=1e-9, atol=0)
    assert np.abs(a['position'] - ref['position']).max() <= 1e-7
    assert np.abs(a['velocity'] - ref['velocity']).max() <= 1e-5
    assert (a['force'] == 0).all()


def test_known_answers_on_gpu():
    """reference core/motion/tests/test_power_law_benchmark.py:13-63 through the GPU pair kernels."""
    g = load_golden('known_answers.npz')
    for model in ('circular', 'three_circle'):
        a = from_raw(g['%s_not_colliding_agents' % model], DT[model])
        gi.agent_agent_block_list(a, CELL)
        assert (a['force'] == 0).all()
        a = from_raw(g['%s_colliding_agents' % model], DT[model])
        gi.agent_agent_block_list(a, CELL)
        assert vec_rel_err(a['force'], g['%s_colliding_force' % model]) <= FORCE_TOL
        assert np.hypot(*a['force'][0]) > 0 and np.hypot(*a['force'][1]) > 0


def test_pair_vectors_on_gpu():
    """Golden pair interactions of the reference (incl. coincident centres, zero relative velocity, overlaps)."""
    g = load_golden('pairs.npz')
    for model in ('circular', 'three_circle'):
        a = from_raw(g[model + '_agents'], DT[model])
        ref = from_raw(g[model + '_after_interaction'], DT[model])
        # evaluate pair (2k, 2k+1) in isolation: move pair k far away from every other pair (exact: power of two shift
        # would still change rounding, so instead run each pair as its own two-agent crowd)
        out = a.copy()
        dev = DeviceAgents(MODEL_CIRCULAR if model == 'circular' else MODEL_THREE_CIRCLE)
        for k in range(len(a) // 2):
            two = a[2 * k:2 * k + 2].copy()
            dev.upload(two); dev.agent_agent(CELL); dev.download(two)
            out[2 * k:2 * k + 2] = two
        dev.close()
        if model == 'three_circle':
            # the golden pairs were evaluated as (i, j) = (2k, 2k+1); the block list orients a pair by
            # (cell_x, cell_y, index), so only pairs whose first agent is in the not-larger cell are comparable
            cells = [O.add_to_cells(a[2 * k:2 * k + 2].copy(), CELL)['cell_of_agent'] for k in range(len(a) // 2)]
            keep = np.repeat([c[0] <= c[1] for c in cells], 2)
            assert keep.sum() >= len(a) // 2
            out, ref = out[keep], ref[keep]
        _assert_forces(out, ref, model)


# ---- integrator / adaptive dt ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('model', ['circular', 'three_circle'])
def test_integrator_and_adaptive_dt(model):
    rng = np.random.default_rng(9)
    for trial in range(6):
        a, _, _ = S.uniform_crowd(2500, model, density=1.0, seed=trial)
        a['force'] = rng.normal(0, 300, (len(a), 2)); a['force_prev'] = rng.normal(0, 300, (len(a), 2))
        if model != 'circular':
            a['torque'] = rng.normal(0, 30, len(a)); a['torque_prev'] = rng.normal(0, 30, len(a))
            a['orientation'] = rng.uniform(-np.pi, np.pi, len(a)) * (1 if trial else 0.9999)
        a['velocity'] *= (0.0, 0.3, 1.0, 3.0, 10.0, 100.0)[trial]
        lo, hi = 0.001, 0.01
        ref = a.copy()
        d_ref = O.velocity_verlet_integrator(ref, lo, hi)
        d = gint.velocity_verlet_integrator(a, lo, hi)
        assert lo <= d <= hi                                  # reference core/tests/test_integrator.py:8-53
        assert abs(d - d_ref) <= 1e-15 * d_ref
        worst, f = rel_err_fields(a, ref)
        assert worst <= 1e-12, (worst, f)


# ---- error behaviour -------------------------------------------------------------------------------------------------------------
def test_invalid_dtype_raises_invalid_type():
    bad = np.zeros(4, dtype=obstacle_type_linear)
    with pytest.raises(InvalidType):
        gi.agent_agent_block_list(bad, CELL)
    with pytest.raises(InvalidType):
        gi.agent_obstacle(np.zeros(3), None)
    dev = DeviceAgents(MODEL_CIRCULAR)
    with pytest.raises(InvalidType):
        dev.upload(np.zeros(2, dtype=agent_type_three_circle))
    a = S.random_crowd(5, 'circular')[0]
    dev.upload(a)
    with pytest.raises(InvalidValue):
        dev.agent_agent(0.0)
    a['position'][2] = np.nan
    dev.upload(a)
    with pytest.raises(InvalidValue):
        dev.agent_agent(CELL)
    dev.close()


# ---- fused two-phase kernel vs the one-phase kernels (independent implementations of the same arithmetic) ----------------
@pytest.mark.parametrize('model', ['circular', 'three_circle'])
@pytest.mark.parametrize('density', [0.125, 1.0, 4.0])
def test_fused_kernel_matches_one_phase_kernel(model, density):
    n = 20000
    a, obs, side = S.uniform_crowd(n, model, density=density, seed=7, overlap_fraction=0.03 if density > 1 else 0.0)
    mid = MODEL_CIRCULAR if model == 'circular' else MODEL_THREE_CIRCLE
    out = {}
    for variant in (1, 2, 3):
        dev = DeviceAgents(mid)
        dev.set_variant(variant)
        # same search lattice for all three: resident-order steps (wider cells, variant 3 only) would change the ORDER in which
        # an agent's pair contributions are added, and the bit-for-bit claim below is about that order
        dev.set_rebuild_policy(0.10, 1)
        dev.upload(a)
        dev.set_obstacles(obs)
        dev.set_navigation_field(0, *S.direction_field(0.5, (0, 0, side, side), 'swirl'))
        dev.step(1, _lib.STEP_ALL & ~(_lib.STEP_INTEGRATOR | _lib.STEP_RESET), CELL, 0.001, 0.01, want_dt=False)
        f = a.copy(); dev.download(f)
        dts = dev.step(3, _lib.STEP_ALL, CELL, 0.001, 0.01)
        g = a.copy(); dev.download(g)
        out[variant] = (f, g, dts)
        dev.close()
    _assert_forces(out[2][0], out[1][0], model, tol=1e-10)
    np.testing.assert_allclose(out[1][2], out[2][2], rtol=1e-9, atol=0)
    assert np.abs(out[1][1]['position'] - out[2][1]['position']).max() <= 1e-9
    # the once-per-pair pipeline (variant 3, the default) adds the same per-pair numbers in the same order as variant 2
    assert rel_err_fields(out[3][0], out[2][0])[0] == 0
    assert (out[3][2] == out[2][2]).all()
    assert rel_err_fields(out[3][1], out[2][1])[0] == 0
    # and the one-phase path against the oracle on the same crowd (single step forces)
    ref = a.copy()
    O.navigation(ref, [S.direction_field(0.5, (0, 0, side, side), 'swirl')]); O.orientation(ref); O.adjusting(ref)
    O.agent_agent_block_list(ref, CELL); O.agent_obstacle(ref, obs)
    _assert_forces(out[2][0], ref, model)


# ---- CUDA-graph replay of step pairs is only a launch optimisation ------------------------------------------------------------
@pytest.mark.parametrize('model', ['circular', 'three_circle'])
def test_graph_replay_is_bit_identical_to_plain_launches(model):
    agents, obstacles, side = S.uniform_crowd(5000, model, density=1.0, seed=8, overlap_fraction=0.02)
    agents['std_rand_force'] = 0.2
    fields = [S.direction_field(0.5, (0, 0, side, side), 'swirl')]
    out = {}
    for graphs in (True, False):
        dev = DeviceAgents(MODEL_CIRCULAR if model == 'circular' else MODEL_THREE_CIRCLE)
        dev.set_graphs(graphs)
        dev.upload(agents)
        dev.set_obstacles(obstacles)
        dev.set_navigation_field(0, *fields[0])
        dts = [dev.step(n, _lib.STEP_ALL | _lib.STEP_FLUCTUATION, CELL, 0.001, 0.01) for n in (1, 7, 2, 131, 1030)]   # odd / even, > lattice window, > dt ring
        launches = dev.launch_count()
        a = agents.copy()
        dev.download(a)
        t, it = dev.time()
        dev.close()
        out[graphs] = (np.concatenate(dts), a, t, it, launches)
    assert out[True][3] == out[False][3] == 1171
    assert (out[True][0] == out[False][0]).all() and out[True][2] == out[False][2]
    assert rel_err_fields(out[True][1], out[False][1])[0] == 0
    assert out[True][4] == out[False][4]                # same kernels launched, just batched


def test_legacy_default_stream_is_not_captured():
    """bench.py runs the sim on torch's current stream, which is the legacy default stream: it cannot be captured into a CUDA
    graph, so cdb_step must fall back to plain launches there (and give the same result)."""
    import torch
    agents, obstacles, side = S.uniform_crowd(3000, 'circular', density=1.0, seed=9)
    out = []
    for use_default in (True, False):
        dev = DeviceAgents(MODEL_CIRCULAR)
        if use_default:
            dev.set_stream(torch.cuda.current_stream().cuda_stream)
        dev.upload(agents); dev.set_obstacles(obstacles)
        dts = dev.step(9, _lib.STEP_ALL & ~_lib.STEP_NAVIGATION, CELL, 0.001, 0.01)
        a = agents.copy(); dev.download(a); dev.close()
        out.append((dts, a))
    assert (out[0][0] == out[1][0]).all() and rel_err_fields(out[0][1], out[1][1])[0] == 0
