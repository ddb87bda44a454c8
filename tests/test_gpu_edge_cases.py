"""Adversarial inputs through the CUDA path against the oracle: the edge cases the reference's own tests poke at
(core/tests/test_interactions.py:42-79, test_distance.py, test_vector2D.py) plus the ones the GPU design could get wrong
(list flushes at extreme density, clamped lattices, huge / negative coordinates, degenerate geometry)."""
import numpy as np
import pytest

from conftest import vec_rel_err, rel_err_fields
from crowddynamics_b200 import _lib, synthetic as S
from crowddynamics_b200.engine import DeviceAgents
from crowddynamics_b200.structures import MODEL_CIRCULAR, MODEL_THREE_CIRCLE, obstacle_type_linear
from oracle import crowd_oracle as O

pytestmark = pytest.mark.gpu
CELL = 3.6
MODELS = ['circular', 'three_circle']


def _mid(model):
    return MODEL_CIRCULAR if model == 'circular' else MODEL_THREE_CIRCLE


def _one_step_forces(agents, obstacles, variant=3):
    dev = DeviceAgents(_mid('circular' if agents.dtype.itemsize == 228 else 'three_circle'))
    dev.set_variant(variant)
    dev.upload(agents)
    dev.set_obstacles(obstacles)
    dev.step(1, _lib.STEP_ADJUSTING | _lib.STEP_ORIENTATION | _lib.STEP_AGENT_AGENT | _lib.STEP_AGENT_OBSTACLE, CELL, 0.01, 0.01,
             want_dt=False)
    out = agents.copy()
    dev.download(out)
    dev.close()
    return out


def _oracle_forces(agents, obstacles):
    ref = agents.copy()
    O.orientation(ref); O.adjusting(ref); O.agent_agent_block_list(ref, CELL); O.agent_obstacle(ref, obstacles)
    return ref


def _check(agents, obstacles, tol=1e-9):
    ref = _oracle_forces(agents, obstacles)
    for variant in (3, 2, 1):
        got = _one_step_forces(agents, obstacles, variant)
        assert vec_rel_err(got['force'], ref['force']) <= tol, variant
        if 'torque' in agents.dtype.names:
            assert vec_rel_err(got['torque'], ref['torque']) <= tol, variant


@pytest.mark.parametrize('model', MODELS)
def test_extreme_density_many_list_flushes(model):
    """2000 agents in a 6 m box (55 /m^2, everybody overlaps somebody): hundreds of survivors per agent, so the per-lane lists
    overflow and are flushed many times."""
    agents, obstacles, _ = S.random_crowd(2000, model, half_width=3.0, seed=1)
    _check(agents, obstacles)


@pytest.mark.parametrize('model', MODELS)
def test_coincident_agents_and_equal_velocities(model):
    agents, obstacles, _ = S.random_crowd(400, model, half_width=6.0, seed=2)
    agents['position'][1] = agents['position'][0]            # d == 0: zero normal (distance.py:42-43)
    agents['position'][3] = agents['position'][2]
    agents['velocity'][3] = agents['velocity'][2]            # and a == 0 as well
    agents['velocity'][10:60] = (0.3, -0.2)                  # a == 0 for many pairs
    agents['velocity'][60:80] = 0.0
    if model != 'circular':
        S.set_shoulders(agents)
    _check(agents, obstacles)


@pytest.mark.parametrize('model', MODELS)
def test_far_away_negative_and_clustered_coordinates(model):
    """Two clusters 2 km apart at negative / large coordinates: the block list spans ~3e5 cells of which a handful are used."""
    a1, _, _ = S.random_crowd(300, model, half_width=8.0, seed=3)
    a2, _, _ = S.random_crowd(300, model, half_width=8.0, seed=4)
    a1['position'] += (-1500.0, -700.0)
    a2['position'] += (600.0, 950.0)
    agents = np.concatenate((a1, a2))
    if model != 'circular':
        S.set_shoulders(agents)
    obstacles = S.walls_of_box(-1510.0, -710.0, -1490.0, -690.0)
    _check(agents, obstacles)
    # fused resident steps on the same crowd (padded, re-used lattice)
    dev = DeviceAgents(_mid(model))
    dev.upload(agents); dev.set_obstacles(obstacles)
    dev.step(5, _lib.STEP_ALL & ~_lib.STEP_NAVIGATION, CELL, 0.01, 0.01, want_dt=False)
    got = agents.copy(); dev.download(got); dev.close()
    ref = agents.copy()
    for _ in range(5):
        O.step(ref, obstacles, [], CELL, 0.01, 0.01)
    assert np.abs(got['position'] - ref['position']).max() <= 1e-8


@pytest.mark.parametrize('model', MODELS)
def test_agents_outside_a_fixed_lattice_are_binned_into_border_cells(model):
    """cdb_set_lattice smaller than the crowd: agents outside are clamped into the border cells; forces must not change."""
    agents, obstacles, side = S.uniform_crowd(3000, model, density=1.0, seed=5)
    ref = _oracle_forces(agents, obstacles)
    dev = DeviceAgents(_mid(model))
    dev.upload(agents)
    dev.set_obstacles(obstacles)
    dev.set_lattice(3, 2, 6, 7)          # covers only the middle of the 55 m room
    dev.step(1, _lib.STEP_ADJUSTING | _lib.STEP_ORIENTATION | _lib.STEP_AGENT_AGENT | _lib.STEP_AGENT_OBSTACLE, CELL, 0.01, 0.01,
             want_dt=False)
    got = agents.copy(); dev.download(got); dev.close()
    assert vec_rel_err(got['force'], ref['force']) <= 1e-9
    if model != 'circular':
        assert vec_rel_err(got['torque'], ref['torque']) <= 1e-9


@pytest.mark.parametrize('model', MODELS)
def test_degenerate_and_touching_walls(model):
    """reference core/tests/test_interactions.py:60-79: arbitrary segments incl. degenerate ones; centre exactly on the line
    (np.sign(0) = 0 -> zero normal, distance.py:140-142); agents overlapping end caps."""
    agents, _, _ = S.random_crowd(200, model, half_width=3.0, seed=6)
    obs = np.zeros(6, dtype=obstacle_type_linear)
    obs[0]['p0'] = obs[0]['p1'] = (0.5, 0.5)                           # degenerate
    obs[1]['p0'], obs[1]['p1'] = (-3.0, 0.0), (3.0, 0.0)
    obs[2]['p0'], obs[2]['p1'] = (0.0, -3.0), (0.0, 3.0)
    obs[3]['p0'], obs[3]['p1'] = (1.0, 1.0), (1.0 + 1e-9, 1.0)         # almost degenerate
    obs[4]['p0'], obs[4]['p1'] = (-2.0, -2.0), (2.0, 2.0)
    obs[5]['p0'], obs[5]['p1'] = (1e7, 1e7), (1e7 + 5, 1e7)           # far away
    agents['position'][0] = (1.25, 0.0)                                # on the line of wall 1
    agents['position'][1] = (3.1, 0.05)                                # beyond the end cap of wall 1
    agents['position'][2] = (0.0, 0.0)                                 # on three walls at once
    if model != 'circular':
        S.set_shoulders(agents)
    ref = agents.copy(); O.agent_obstacle(ref, obs)
    dev = DeviceAgents(_mid(model))
    dev.upload(agents); dev.set_obstacles(obs); dev.agent_obstacle()
    got = agents.copy(); dev.download(got)
    dev.step(1, _lib.STEP_AGENT_OBSTACLE, CELL, 0.01, 0.01, want_dt=False)      # the fused kernel's wall path, added on top
    twice = agents.copy(); dev.download(twice); dev.close()
    assert vec_rel_err(got['force'], ref['force']) <= 1e-12
    assert vec_rel_err(twice['force'], 2 * ref['force']) <= 1e-12
    if model != 'circular':
        assert vec_rel_err(got['torque'], ref['torque']) <= 1e-12


def test_orientation_wrapping_extremes():
    """wrap_to_pi with Python-modulo semantics at +-pi and for large angles (reference core/tests/test_vector2D.py:9-18),
    through the rotational Verlet of the integrator."""
    agents, _, _ = S.random_crowd(64, 'three_circle', half_width=50.0, seed=7)
    phi = np.concatenate((np.pi * np.arange(-8, 8), np.linspace(-40.0, 40.0, 32), np.full(16, np.pi)))
    agents['orientation'] = phi
    agents['angular_velocity'] = 0.0
    agents['torque'] = 0.0
    agents['torque_prev'] = 0.0
    agents['force'] = 0.0
    agents['force_prev'] = 0.0
    ref = agents.copy()
    O.velocity_verlet_integrator(ref, 0.01, 0.01)
    dev = DeviceAgents(MODEL_THREE_CIRCLE)
    dev.upload(agents); dev.integrate(0.01, 0.01)
    got = agents.copy(); dev.download(got); dev.close()
    assert (got['orientation'] == ref['orientation']).all()         # fmod is exact on both sides
    assert (np.abs(got['orientation']) <= np.pi).all()
    assert rel_err_fields(got, ref, ['position_ls', 'position_rs'])[0] <= 1e-14
