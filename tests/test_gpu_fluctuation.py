"""Fluctuation node (reference core/motion/fluctuation.py:14-62, logic.py:78-86) on the GPU.  The reference draws from
numpy's unseeded global RNG, so the parity bar is distributional: Kolmogorov-Smirnov against scipy's truncated normal (the
very distribution the reference samples, core/rand.py:8-33) and the uniform direction, plus the structural properties."""
import numpy as np
import pytest
import scipy.stats as st

from crowddynamics_b200 import _lib, synthetic as S, logic as L
from crowddynamics_b200.engine import DeviceAgents
from crowddynamics_b200.structures import MODEL_CIRCULAR, MODEL_THREE_CIRCLE

pytestmark = pytest.mark.gpu
N = 200000


def _draw(model, seed, calls=1, scale_f=0.1, scale_t=0.1):
    agents, _, _ = S.uniform_crowd(N, model, density=1.0, seed=1)
    agents['std_rand_force'] = scale_f
    if model != 'circular':
        agents['std_rand_torque'] = scale_t
    dev = DeviceAgents(MODEL_CIRCULAR if model == 'circular' else MODEL_THREE_CIRCLE)
    dev.set_seed(seed)
    dev.upload(agents)
    out = []
    for _ in range(calls):
        dev.reset()
        dev.fluctuation()
        a = agents.copy()
        dev.download(a)
        out.append(a)
    dev.close()
    return agents, out


@pytest.mark.parametrize('model', ['circular', 'three_circle'])
def test_distributions_match_the_reference_sampler(model):
    agents, (a,) = _draw(model, seed=7)
    mag = np.hypot(*a['force'].T) / (agents['mass'] * agents['std_rand_force'])
    ang = np.arctan2(a['force'][:, 1], a['force'][:, 0]) % (2 * np.pi)
    assert mag.min() >= 0 and mag.max() <= 3.0
    assert st.kstest(mag, st.truncnorm(0.0, 3.0).cdf).pvalue > 1e-3        # xi ~ truncnorm(0, 3, scale=std_rand_force)
    assert st.kstest(ang, st.uniform(0, 2 * np.pi).cdf).pvalue > 1e-3       # phi ~ U(0, 2 pi)
    assert abs(np.corrcoef(mag, ang)[0, 1]) < 0.01
    if model == 'three_circle':
        eta = a['torque'] / (agents['inertia_rot'] * agents['std_rand_torque'])
        assert np.abs(eta).max() <= 3.0
        assert st.kstest(eta, st.truncnorm(-3.0, 3.0).cdf).pvalue > 1e-3
        assert abs(np.corrcoef(eta, mag)[0, 1]) < 0.01
    else:
        assert 'torque' not in a.dtype.names


def test_streams_are_reproducible_and_independent():
    _, (a1, a2) = _draw('circular', seed=11, calls=2)
    _, (b1, b2) = _draw('circular', seed=11, calls=2)
    _, (c1,) = _draw('circular', seed=12)
    assert (a1['force'] == b1['force']).all() and (a2['force'] == b2['force']).all()     # same seed, same draws
    assert (a1['force'] != a2['force']).any() and (a1['force'] != c1['force']).any()
    assert abs(np.corrcoef(a1['force'][:, 0], a2['force'][:, 0])[0, 1]) < 0.01            # successive calls independent
    assert abs(np.corrcoef(a1['force'][:, 0], c1['force'][:, 0])[0, 1]) < 0.01            # seeds independent
    assert abs(np.corrcoef(a1['force'][:-1, 0], a1['force'][1:, 0])[0, 1]) < 0.01         # agents independent


def test_zero_scale_gives_no_fluctuation_and_force_is_added():
    agents, _, _ = S.uniform_crowd(1000, 'three_circle', density=1.0, seed=2)
    agents['std_rand_force'] = 0.0
    agents['std_rand_torque'] = 0.0
    agents['force'] = 3.0
    agents['torque'] = -2.0
    dev = DeviceAgents(MODEL_THREE_CIRCLE)
    dev.upload(agents)
    dev.fluctuation()
    a = agents.copy()
    dev.download(a)
    assert (a['force'] == 3.0).all() and (a['torque'] == -2.0).all()
    agents['std_rand_force'] = 0.5
    dev.upload(agents)
    dev.fluctuation()
    dev.download(a)
    dev.close()
    assert (a['force'] != 3.0).any() and np.abs(a['force'] - 3.0).max() <= 3 * 0.5 * agents['mass'].max() + 1e-9


def test_fused_step_with_fluctuation_is_decomposition_invariant():
    """Philox is keyed by (seed, step) and counted by the global agent id: strips reproduce the single-device run bit for bit
    even with the stochastic node switched on."""
    import torch
    from crowddynamics_b200.parallel import StripSimulation, LocalGroup
    agents, obstacles, side = S.uniform_crowd(6000, 'three_circle', density=1.0, seed=3)
    agents['std_rand_force'] = 0.3
    agents['std_rand_torque'] = 0.3
    flags = _lib.STEP_ALL | _lib.STEP_FLUCTUATION
    dev = DeviceAgents(MODEL_THREE_CIRCLE)
    dev.upload(agents); dev.set_obstacles(obstacles)
    dev.step(10, flags, 3.6, 0.01, 0.01, want_dt=False)
    ref = agents.copy(); dev.download(ref); dev.close()
    plain = DeviceAgents(MODEL_THREE_CIRCLE)
    plain.upload(agents); plain.set_obstacles(obstacles)
    plain.step(10, _lib.STEP_ALL, 3.6, 0.01, 0.01, want_dt=False)
    det = agents.copy(); plain.download(det); plain.close()
    assert np.abs(ref['position'] - det['position']).max() > 1e-6          # the noise does something
    sims = [StripSimulation.from_global(agents, obstacles, [], 3.6, r, 2, device_index=0, flags=flags) for r in range(2)]
    group = LocalGroup(sims)
    group.step(10)
    torch.cuda.synchronize()
    got, ids = group.export(agents.dtype)
    assert (got['position'] == ref['position']).all() and (got['orientation'] == ref['orientation']).all()


def test_fluctuation_node_in_the_logic_tree():
    agents, obstacles, fields = S.hallway(seed=0)
    sim = L.MultiAgentSimulation(agents, obstacles, fields)
    sim.logic = L.hallway_logic(sim, mode='strict', fluctuation=True, seed=5)
    names = [n.name for n in L.post_order_iter(sim.logic.root)]
    assert names[0] == 'Fluctuation' and names[-1] == 'Reset'
    start = agents['position'].copy()
    for _ in range(50):
        sim.update()
    assert np.isfinite(agents['position']).all() and np.abs(agents['position'] - start).max() > 0.1
