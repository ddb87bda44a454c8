"""Where the reference tree is present (the build container), run the reference's OWN numba code live against the C oracle on
fresh seeded crowds -- beyond the committed golden vectors.  Skipped on machines without /root/reference (the GPU box)."""
import numpy as np
import pytest

from crowddynamics_b200 import synthetic as S
from oracle import ref_harness as H
from oracle import crowd_oracle as O

pytestmark = pytest.mark.skipif(not H.available(), reason='reference tree not present')


@pytest.fixture(scope='module')
def R():
    return H.load()


@pytest.mark.parametrize('model', ['circular', 'three_circle'])
@pytest.mark.parametrize('seed,density,overlap', [(21, 0.125, 0.0), (22, 1.0, 0.05), (23, 3.0, 0.1)])
def test_three_steps_bit_identical(R, model, seed, density, overlap):
    agents, obstacles, side = S.uniform_crowd(700, model, density=density, seed=seed, overlap_fraction=overlap)
    agents['target'] = np.random.default_rng(seed).integers(-1, 2, len(agents))
    fields = [S.direction_field(0.5, (0, 0, side * 0.9, side), 'swirl'), S.direction_field(0.5, (0, 0, side, side * 0.8), 'exit')]
    ra, oa = agents.copy(), agents.copy()
    for _ in range(3):
        dt_r = H.step(R, ra, obstacles, fields, 3.6, 0.001, 0.01)
        dt_o = O.step(oa, obstacles, fields, 3.6, 0.001, 0.01)
        assert dt_r == dt_o
        for f in ra.dtype.names:
            x, y = ra[f], oa[f]
            assert ((x == y) | (np.isnan(x) & np.isnan(y))).all(), f


def test_reference_benchmark_workload_bit_identical(R):
    """core/tests/test_interactions_benchmark.py:10-33 (uniform random positions, overlaps allowed)."""
    for model in ('circular', 'three_circle'):
        agents, _, _ = S.random_crowd(500, model, seed=31)
        ra, oa = agents.copy(), agents.copy()
        H.node_agent_agent(R, ra, 3.6)
        O.agent_agent_block_list(oa, 3.6)
        assert (ra['force'] == oa['force']).all()
        if model == 'three_circle':
            assert (ra['torque'] == oa['torque']).all()


def test_pair_orientation_matters_only_for_three_circle(R):
    """swapping (i, j) is an exact negation for circular agents, but not for three-circle ones -- the
    reason the GPU kernels reproduce the block list's (cell, index) pair orientation."""
    for model, fn in (('circular', R.interactions.interaction_agent_agent_circular),
                      ('three_circle', R.interactions.interaction_agent_agent_three_circle)):
        agents, _, _ = S.random_crowd(400, model, half_width=6.0, seed=41)
        differing = 0
        for k in range(0, 400, 2):
            a, b = agents[k:k + 2].copy(), agents[k:k + 2][::-1].copy()
            fn(0, 1, a)
            fn(0, 1, b)
            same = (a['force'] == b['force'][::-1]).all()
            if model == 'three_circle':
                same = same and (a['torque'] == b['torque'][::-1]).all()
            differing += not same
        assert (differing == 0) if model == 'circular' else (differing > 0)
