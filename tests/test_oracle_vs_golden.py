"""The C oracle (oracle/crowd_oracle.c) against golden vectors produced by the reference's own numba code
(tests/golden/generate.py) and against the reference's known-answer tests.  CPU only.

Bar: bit-exact wherever the arithmetic is IEEE basic operations; 1e-13 relative is allowed for results that pass
through libm (hypot/exp/sin/cos/atan2), whose last-bit behaviour may differ between hosts."""
import numpy as np
import pytest

from conftest import load_golden, from_raw, rel_err_fields
from crowddynamics_b200.structures import agent_type_circular, agent_type_three_circle, obstacle_type_linear
from crowddynamics_b200 import synthetic as S
from oracle import crowd_oracle as O

DT = {'circular': agent_type_circular, 'three_circle': agent_type_three_circle}
TOL = 1e-13


def _fields(g):
    mg = S.MeshGrid(float(g['field_step']), *g['field_bounds'])
    assert mg.shape == g['U'].shape[1:]
    return [(mg, (g['U'][t], g['V'][t])) for t in range(len(g['U']))]


def _obs(g):
    return np.ascontiguousarray(g['obstacles']).view(obstacle_type_linear).reshape(-1)


@pytest.mark.parametrize('name', ['step_%s.npz', 'step_sparse_%s.npz'])
@pytest.mark.parametrize('model', ['circular', 'three_circle'])
def test_nodes_against_reference(name, model):
    g = load_golden(name % model)
    dt = DT[model]
    a = from_raw(g['initial'], dt)
    fields, obs, cell = _fields(g), _obs(g), float(g['cell_size'])
    O.navigation(a, fields)
    assert rel_err_fields(a, from_raw(g['after_navigation'], dt))[0] == 0
    O.orientation(a)
    assert rel_err_fields(a, from_raw(g['after_orientation'], dt))[0] <= TOL
    O.adjusting(a)
    assert rel_err_fields(a, from_raw(g['after_adjusting'], dt))[0] <= TOL
    # block list tables: bit exact
    before = a.copy()
    cl = O.add_to_cells(before, cell)
    assert (cl['points_indices'] == g['points_indices']).all()
    assert (cl['cells_count'] == g['cells_count']).all()
    assert (cl['cells_offset'] == g['cells_offset']).all()
    assert tuple(cl['grid'][2:]) == tuple(g['grid_shape'])
    O.agent_agent_block_list(a, cell)
    assert rel_err_fields(a, from_raw(g['after_agent_agent'], dt))[0] <= TOL
    O.agent_obstacle(a, obs)
    assert rel_err_fields(a, from_raw(g['after_agent_obstacle'], dt))[0] <= TOL
    d = O.velocity_verlet_integrator(a, float(g['dt_min']), float(g['dt_max']))
    assert d == g['dts'][0]
    assert rel_err_fields(a, from_raw(g['after_integrator'], dt))[0] <= TOL
    O.reset(a)
    assert rel_err_fields(a, from_raw(g['after_reset'], dt))[0] <= TOL


@pytest.mark.parametrize('model', ['circular', 'three_circle'])
def test_five_steps_against_reference(model):
    g = load_golden('step_%s.npz' % model)
    dt = DT[model]
    a = from_raw(g['initial'], dt)
    fields, obs = _fields(g), _obs(g)
    dts = [O.step(a, obs, fields, float(g['cell_size']), float(g['dt_min']), float(g['dt_max'])) for _ in range(5)]
    np.testing.assert_allclose(dts, g['dts'], rtol=TOL, atol=0)
    worst, f = rel_err_fields(a, from_raw(g['after_5_steps'], dt))
    assert worst <= 1e-11, (worst, f)   # five chaotic steps amplify a last-bit libm difference, if there is one


@pytest.mark.parametrize('model', ['circular', 'three_circle'])
def test_pair_vectors(model):
    g = load_golden('pairs.npz')
    dt = DT[model]
    a = from_raw(g[model + '_agents'], dt)
    n_pairs = len(a) // 2
    force = O.force_social_circular if model == 'circular' else O.force_social_three_circle
    for k in range(n_pairs):
        fi, fj = force(a, 2 * k, 2 * k + 1)
        np.testing.assert_allclose(fi, g[model + '_social_i'][k], rtol=TOL, atol=0)
        np.testing.assert_allclose(fj, g[model + '_social_j'][k], rtol=TOL, atol=0)
    # the full pair interaction (gate, social, contact, torque): pairs (2k, 2k+1) are far apart from other pairs'
    # members only by construction of the test, so apply them one by one through a two-agent block list
    b = a.copy()
    for k in range(n_pairs):
        two = b[2 * k:2 * k + 2].copy()
        O.agent_agent_brute(two)
        b[2 * k:2 * k + 2] = two
    assert rel_err_fields(b, from_raw(g[model + '_after_interaction'], dt))[0] <= TOL


@pytest.mark.parametrize('model', ['circular', 'three_circle'])
def test_wall_vectors(model):
    g = load_golden('pairs.npz')
    dt = DT[model]
    a = from_raw(g[model + '_wall_agents'], dt)
    segs = g['wall_segments']
    for k in range(len(a)):
        one = a[k:k + 1].copy()
        obs = np.ascontiguousarray(segs[k:k + 1]).view(obstacle_type_linear).reshape(-1)
        O.agent_obstacle(one, obs)
        a[k] = one[0]
    assert rel_err_fields(a, from_raw(g[model + '_wall_after'], dt))[0] <= TOL


def test_wrap_to_pi():
    g = load_golden('pairs.npz')
    out = np.array([O.wrap_to_pi(x) for x in g['wrap_in']])
    assert (out == g['wrap_out']).all()
    assert (np.abs(out) <= np.pi).all()            # reference core/tests/test_vector2D.py:9-18
    assert O.wrap_to_pi(np.pi) == np.pi and O.wrap_to_pi(-np.pi) == -np.pi
    assert O.wrap_to_pi(3 * np.pi) == np.pi and O.wrap_to_pi(-3 * np.pi) == -np.pi


@pytest.mark.parametrize('model', ['circular', 'three_circle'])
def test_reference_known_answers(model):
    """reference core/motion/tests/test_power_law_benchmark.py:13-35 (zero when not colliding), :41-63 (non-zero)."""
    g = load_golden('known_answers.npz')
    dt = DT[model]
    force = O.force_social_circular if model == 'circular' else O.force_social_three_circle
    a = from_raw(g['%s_not_colliding_agents' % model], dt)
    fi, fj = force(a, 0, 1)
    assert np.hypot(*fi) == 0 and np.hypot(*fj) == 0
    a = from_raw(g['%s_colliding_agents' % model], dt)
    fi, fj = force(a, 0, 1)
    assert np.hypot(*fi) > 0 and np.hypot(*fj) > 0
    np.testing.assert_allclose(np.stack((fi, fj)), g['%s_colliding_force' % model], rtol=TOL, atol=0)


def test_hallway_trajectory():
    """BASELINE config 1 (Hallway, 50 circular agents): 200 updates of the reference vs the oracle."""
    g = load_golden('hallway.npz')
    agents, obstacles, fields = S.hallway(seed=0)
    assert (np.ascontiguousarray(agents).view(np.uint8).reshape(len(agents), -1) == g['initial']).all()
    a = agents.copy()
    traj = [a['position'].copy()]
    for k in range(int(g['steps'])):
        O.step(a, obstacles, fields, 3.6, 0.01, 0.01)
        if (k + 1) % 50 == 0:
            traj.append(a['position'].copy())
    np.testing.assert_allclose(np.stack(traj), g['positions'], rtol=0, atol=1e-9)


def test_block_list_loses_no_pair():
    """Block list + gate == brute force over all pairs (same pair kernels), up to summation order."""
    for model in ('circular', 'three_circle'):
        a, _, _ = S.uniform_crowd(400, model, density=1.5, seed=11, overlap_fraction=0.05)
        b = a.copy()
        O.agent_agent_block_list(a, 3.6)
        O.agent_agent_brute(b)
        if model == 'circular':
            assert rel_err_fields(a, b, ['force'])[0] < 1e-9
        # three_circle: the brute-force (i<j) orientation differs from the block list's for some pairs -> only
        # check that exactly the same agents feel a force
        assert ((np.abs(a['force']).sum(1) > 0) == (np.abs(b['force']).sum(1) > 0)).all()


def test_reference_property_tests():
    """Properties the reference's own tests assert: h >= -r_tot (core/tests/test_distance.py:23) is implied by
    hypot >= 0; dt_min <= dt <= dt_max (core/tests/test_integrator.py:8-53); N in {0, 1, 2} runs
    (core/tests/test_interactions.py:42-55)."""
    rng = np.random.default_rng(5)
    for model in ('circular', 'three_circle'):
        for n in (0, 1, 2):
            a, obs, _ = S.random_crowd(n, model, seed=n)
            O.agent_agent_block_list(a, 3.6)
            O.agent_obstacle(a, obs)
        for _ in range(20):
            a, _, _ = S.random_crowd(50, model, seed=int(rng.integers(1 << 30)))
            lo = float(rng.uniform(1e-4, 1e-2)); hi = lo + float(rng.uniform(0, 1e-1))
            a['velocity'] *= rng.uniform(0, 10)
            dt = O.velocity_verlet_integrator(a, lo, hi)
            assert lo <= dt <= hi
    with pytest.raises(O.InvalidType):
        O.agent_agent_block_list(np.zeros(3, dtype=obstacle_type_linear), 3.6)
