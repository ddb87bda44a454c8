"""Strip decomposition on the GPU.  LocalGroup drives all strips of a decomposition from one process on one device (the
message exchange is a tensor copy instead of NCCL send/recv; everything else -- halo / migrant packing, ghost columns,
pair orientation across strip borders -- is the production CUDA code).  A multi-process NCCL run needs one GPU per rank
and is exercised by `bench.py --gpus N` / tests marked multi_gpu."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from crowddynamics_b200 import _lib, synthetic as S
from crowddynamics_b200.engine import DeviceAgents
from crowddynamics_b200.parallel import StripSimulation, LocalGroup
from crowddynamics_b200.structures import MODEL_CIRCULAR, MODEL_THREE_CIRCLE

pytestmark = pytest.mark.gpu
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))


def _single(agents, obstacles, fields, steps, dt_min, dt_max):
    dev = DeviceAgents(MODEL_CIRCULAR if agents.dtype.itemsize == 228 else MODEL_THREE_CIRCLE)
    dev.upload(agents)
    dev.set_obstacles(obstacles)
    for t, (mg, uv) in enumerate(fields):
        dev.set_navigation_field(t, mg, uv)
    dev.step(steps, _lib.STEP_ALL, 3.6, dt_min, dt_max, want_dt=False)
    out = agents.copy()
    dev.download(out)
    dev.close()
    return out


@pytest.mark.parametrize('model', ['circular', 'three_circle'])
@pytest.mark.parametrize('world', [2, 3])
@pytest.mark.parametrize('dts', [(0.01, 0.01), (0.001, 0.01)])
def test_strips_reproduce_single_device(model, world, dts):
    import torch
    agents, obstacles, side = S.uniform_crowd(6000, model, density=1.0, seed=3, overlap_fraction=0.02)
    agents['velocity'] *= 3.0                  # agents cross strip borders within the horizon
    fields = [S.direction_field(0.5, (0, 0, side, side), 'swirl')]
    steps = 15
    ref = _single(agents, obstacles, fields, steps, *dts)
    sims = [StripSimulation.from_global(agents, obstacles, fields, 3.6, r, world, device_index=0, dt_min=dts[0], dt_max=dts[1])
            for r in range(world)]
    n0 = [s.n_owned() for s in sims]
    group = LocalGroup(sims)
    group.step(steps)
    torch.cuda.synchronize()
    got, ids = group.export(agents.dtype)
    assert (ids == np.arange(len(agents))).all()          # every agent owned by exactly one strip
    assert sum(abs(s.n_owned() - a) for s, a in zip(sims, n0)) > 0      # migration actually happened
    # same candidates in the same order on whichever strip owns an agent -> identical arithmetic
    assert np.abs(got['position'] - ref['position']).max() <= 1e-12
    assert np.abs(got['velocity'] - ref['velocity']).max() <= 1e-10
    if model == 'three_circle':
        assert np.abs(got['orientation'] - ref['orientation']).max() <= 1e-10


def test_settle_moves_misplaced_agents():
    import torch
    agents, obstacles, side = S.uniform_crowd(3000, 'circular', density=1.0, seed=1)
    sims = [StripSimulation.from_global(agents, obstacles, [], 3.6, r, 2, device_index=0) for r in range(2)]
    # hand rank 0 a few agents that belong to rank 1: shift them one column to the right after upload is not possible from
    # outside, so instead build rank 0 from a shifted copy
    shifted = agents.copy()
    shifted['position'][:, 0] += 3.6
    sims[0] = StripSimulation.from_global(shifted, obstacles, [], 3.6, 0, 2, device_index=0,
                                          lattice=(sims[1].bounds[0], -1, sims[1].bounds[-1] - sims[1].bounds[0], 20))
    group = LocalGroup([sims[0], sims[1]])
    before = [s.n_owned() for s in group.sims]
    group.settle()
    torch.cuda.synchronize()
    after = [s.n_owned() for s in group.sims]
    assert sum(after) == sum(before)


@pytest.mark.skipif(True, reason='needs >= 2 GPUs; run manually: torchrun --nproc-per-node 2 tests/run_strips_nccl.py')
def test_placeholder_multi_gpu():
    pass
