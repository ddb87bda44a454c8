"""Randomised differential test: 48 seeded scenarios (both agent models; crowd size, density, overlaps, number / placement of
wall segments, direction fields, targets, dt range all drawn at random), three fused steps each, CUDA path vs oracle."""
import numpy as np
import pytest

from conftest import vec_rel_err
from crowddynamics_b200 import _lib, synthetic as S
from crowddynamics_b200.engine import DeviceAgents
from crowddynamics_b200.structures import MODEL_CIRCULAR, MODEL_THREE_CIRCLE, obstacle_type_linear
from oracle import crowd_oracle as O

pytestmark = pytest.mark.gpu


def _scenario(seed):
    rng = np.random.default_rng(1000 + seed)
    model = ('circular', 'three_circle')[seed % 2]
    n = int(rng.integers(3, 2500))
    density = float(np.exp(rng.uniform(np.log(0.05), np.log(6.0))))
    overlap = float(rng.choice([0.0, 0.02, 0.2]))
    if rng.random() < 0.3:
        agents, _, side = S.random_crowd(n, model, half_width=float(rng.uniform(2.0, 30.0)), seed=seed)
        side *= 1.0
        origin = -side / 2
    else:
        origin = float(rng.uniform(-50, 50))
        agents, _, side = S.uniform_crowd(n, model, density=density, seed=seed, origin=(origin, origin), overlap_fraction=overlap)
    agents['velocity'] *= rng.uniform(0.0, 3.0)
    agents['target'] = rng.integers(-1, 3, n)
    w = int(rng.integers(0, 13))
    obs = np.zeros(w, dtype=obstacle_type_linear)
    obs['p0'] = rng.uniform(origin, origin + side, (w, 2))
    obs['p1'] = obs['p0'] + rng.uniform(-side / 2, side / 2, (w, 2))
    if w and rng.random() < 0.3:
        obs['p1'][0] = obs['p0'][0]
    step = float(rng.choice([0.25, 0.5, 1.0]))
    b = (origin - rng.uniform(0, 3), origin - rng.uniform(0, 3), origin + side * rng.uniform(0.5, 1.1), origin + side * rng.uniform(0.5, 1.1))
    fields = [S.direction_field(step, b, kind, point=(origin + side / 2, origin)) for kind in ('swirl', 'exit', 'x-')]
    dt_max = float(rng.choice([0.01, 0.02]))
    dt_min = dt_max if rng.random() < 0.5 else dt_max / 10
    return model, agents, obs, fields, dt_min, dt_max


@pytest.mark.parametrize('seed', range(48))
def test_random_scenario(seed):
    model, agents, obs, fields, dt_min, dt_max = _scenario(seed)
    dev = DeviceAgents(MODEL_CIRCULAR if model == 'circular' else MODEL_THREE_CIRCLE)
    dev.upload(agents)
    dev.set_obstacles(obs)
    for t, (mg, uv) in enumerate(fields):
        dev.set_navigation_field(t, mg, uv)
    # forces of the first step (everything except integrator / reset)
    dev.step(1, _lib.STEP_ALL & ~(_lib.STEP_INTEGRATOR | _lib.STEP_RESET), 3.6, dt_min, dt_max, want_dt=False)
    got = agents.copy()
    dev.download(got)
    ref = agents.copy()
    O.navigation(ref, fields); O.orientation(ref); O.adjusting(ref); O.agent_agent_block_list(ref, 3.6); O.agent_obstacle(ref, obs)
    assert vec_rel_err(got['force'], ref['force']) <= 1e-9
    if model == 'three_circle':
        assert vec_rel_err(got['torque'], ref['torque'], floor=1e-9) <= 1e-9
    assert (got['target_direction'] == ref['target_direction']).all()
    # then three whole steps from the original state
    dev.upload(agents)
    dts = dev.step(3, _lib.STEP_ALL, 3.6, dt_min, dt_max)
    dev.download(got)
    dev.close()
    ref = agents.copy()
    rdts = [O.step(ref, obs, fields, 3.6, dt_min, dt_max) for _ in range(3)]
    np.testing.assert_allclose(dts, rdts, rtol=1e-9)
    scale = 1.0 + np.abs(ref['position']).max()
    assert np.abs(got['position'] - ref['position']).max() <= 1e-9 * scale
    assert np.abs(got['velocity'] - ref['velocity']).max() <= 1e-6 * (1.0 + np.abs(ref['velocity']).max())
