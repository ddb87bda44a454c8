"""Host-side strip protocol on CPU: two gloo ranks, a numpy/oracle device stand-in (tests/strip_mock.py), against the
single-domain oracle.  No GPU involved: this covers partitioning, the neighbour schedule and the message plumbing of
crowddynamics_b200.parallel."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(__file__))

from crowddynamics_b200 import synthetic as S, _lib  # noqa: E402
from crowddynamics_b200.parallel import StripSimulation, partition_columns, owner_of_columns, lattice_of  # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, model, steps, dt_min, dt_max, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from strip_mock import NumpyStripDevice
    agents, obstacles, side = S.uniform_crowd(600, model, density=1.0, seed=5, overlap_fraction=0.02)
    agents['velocity'] *= 3.0          # make agents cross strip borders within a few steps
    fields = [S.direction_field(0.5, (0, 0, side, side), 'swirl')]
    sim = StripSimulation.from_global(agents, obstacles, fields, 3.6, rank, world, dist=dist, dt_min=dt_min, dt_max=dt_max,
                                      make_device=lambda m, cap: NumpyStripDevice(m, cap), tensor_device=torch.device('cpu'))
    n0 = sim.n_owned()
    sim.step(steps)
    rec, ids = sim.export(agents.dtype)
    np.savez(os.path.join(out_dir, 'rank%d.npz' % rank), rec=rec.view(np.uint8).reshape(len(rec), -1), ids=ids, n0=n0,
             n1=sim.n_owned())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('model', ['circular', 'three_circle'])
@pytest.mark.parametrize('dts', [(0.01, 0.01), (0.001, 0.01)])
def test_two_rank_strips_match_single_domain(tmp_path, model, dts):
    from oracle import crowd_oracle as O
    world, steps = 2, 12
    port = _free_port()
    mp.spawn(_worker, args=(world, port, model, steps, dts[0], dts[1], str(tmp_path)), nprocs=world, join=True)
    agents, obstacles, side = S.uniform_crowd(600, model, density=1.0, seed=5, overlap_fraction=0.02)
    agents['velocity'] *= 3.0
    fields = [S.direction_field(0.5, (0, 0, side, side), 'swirl')]
    ref = agents.copy()
    for _ in range(steps):
        O.step(ref, obstacles, fields, 3.6, dts[0], dts[1])
    got = np.zeros_like(ref)
    seen = np.zeros(len(ref), dtype=int)
    moved = 0
    for r in range(world):
        d = np.load(os.path.join(str(tmp_path), 'rank%d.npz' % r))
        rec = np.ascontiguousarray(d['rec']).view(ref.dtype).reshape(-1)
        got[d['ids']] = rec
        seen[d['ids']] += 1
        moved += abs(int(d['n1']) - int(d['n0']))
    assert (seen == 1).all()                 # every agent owned by exactly one rank
    assert np.abs(got['position'] - ref['position']).max() <= 1e-9
    assert np.abs(got['velocity'] - ref['velocity']).max() <= 1e-7


def _nan_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from strip_mock import NumpyStripDevice
    agents, obstacles, side = S.uniform_crowd(400, 'circular', density=1.0, seed=2)
    agents['velocity'][np.argmin(agents['position'][:, 0])] = np.nan          # owned by rank 0 only
    sim = StripSimulation.from_global(agents, obstacles, [], 3.6, rank, world, dist=dist, dt_min=0.001, dt_max=0.01,
                                      make_device=lambda m, cap: NumpyStripDevice(m, cap), tensor_device=torch.device('cpu'))
    sim.phase_begin()
    sim.dev.export_vmax(sim.vmax)
    local = sim.vmax.numpy().copy()
    dist.all_reduce(sim.vmax, op=dist.ReduceOp.MAX)
    sim.dev.import_vmax(sim.vmax)
    np.savez(os.path.join(out_dir, 'nan%d.npz' % rank), local=local, merged=sim.dev._vmax)
    dist.barrier()
    dist.destroy_process_group()


def test_nan_velocity_reaches_every_rank(tmp_path):
    """np.max in the reference's adaptive_timestep propagates NaN (integrator.py:25-60); all_reduce(MAX) does not promise to.
    The exchanged vector carries NaN as a flag (k_vmax_export / strip_mock.export_vmax), so both ranks end up with NaN."""
    port = _free_port()
    mp.spawn(_nan_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    d0, d1 = (np.load(os.path.join(str(tmp_path), 'nan%d.npz' % r)) for r in range(2))
    assert d0['local'][2] == 1.0 and d1['local'][2] == 0.0 and not np.isnan(d0['local']).any()
    assert np.isnan(d0['merged'][0]) and np.isnan(d1['merged'][0])
    assert d0['merged'][1] == d1['merged'][1] and np.isfinite(d0['merged'][1])


def test_partition_helpers():
    b = partition_columns(-3, 20, 4)
    assert b[0] == -3 and b[-1] == 17 and all(b[i] < b[i + 1] for i in range(4))
    cols = np.array([-10, -3, 1, 2, 6, 7, 16, 40])
    own = owner_of_columns(cols, b)
    assert own[0] == 0 and own[-1] == 3 and (np.diff(own) >= 0).all()
    for c, o in zip(cols[1:-1], own[1:-1]):
        assert b[o] <= c < b[o + 1]
    with pytest.raises(ValueError):
        partition_columns(0, 2, 3)
    pos = np.array([[0.1, 0.2], [10.0, -4.0]])
    ix, iy, nx, ny = lattice_of(pos, 3.6, pad=1)
    assert ix == -1 and iy == -3 and nx == 5 and ny == 5
