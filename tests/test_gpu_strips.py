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
@pytest.mark.parametrize('direct', [False, True], ids=['messages', 'peer-memory'])
def test_strips_reproduce_single_device(model, world, dts, direct):
    import torch
    agents, obstacles, side = S.uniform_crowd(6000, model, density=1.0, seed=3, overlap_fraction=0.02)
    agents['velocity'] *= 6.0                  # agents cross strip borders within the horizon
    fields = [S.direction_field(0.5, (0, 0, side, side), 'swirl')]
    steps = 30
    ref = _single(agents, obstacles, fields, steps, *dts)
    sims = [StripSimulation.from_global(agents, obstacles, fields, 3.6, r, world, device_index=0, dt_min=dts[0], dt_max=dts[1])
            for r in range(world)]
    group = LocalGroup(sims, direct=direct)      # direct: one-sided writes into the neighbours' buffers + sequence flags
    owned0 = [set(s.export(agents.dtype)[1].tolist()) for s in sims]
    group.step(steps)
    torch.cuda.synchronize()
    got, ids = group.export(agents.dtype)
    assert (ids == np.arange(len(agents))).all()          # every agent owned by exactly one strip
    owned1 = [set(s.export(agents.dtype)[1].tolist()) for s in sims]
    assert any(a != b for a, b in zip(owned0, owned1))       # migration actually happened
    # same candidates in the same order on whichever strip owns an agent -> identical arithmetic
    assert np.abs(got['position'] - ref['position']).max() <= 1e-12
    assert np.abs(got['velocity'] - ref['velocity']).max() <= 1e-10
    if model == 'three_circle':
        assert np.abs(got['orientation'] - ref['orientation']).max() <= 1e-10


@pytest.mark.parametrize('model', ['circular', 'three_circle'])
@pytest.mark.parametrize('world', [2, 3])
@pytest.mark.parametrize('dts', [(0.01, 0.01), (0.001, 0.01)], ids=['fixed-dt', 'adaptive-dt'])
@pytest.mark.parametrize('direct', [False, True], ids=['messages', 'peer-memory'])
def test_strips_with_kept_block_lists(model, world, dts, direct):
    """Resident-order steps in strip mode (cdb_strip_set_kind): block lists kept for several steps on widened cells, halo
    records only on kept steps, migrants on the step before a rebuild.  Against one device with the same policy -- the two
    rebuild on different schedules, so they agree up to summation order (14 steps: bar 1e-8 m)."""
    import torch
    agents, obstacles, side = S.uniform_crowd(30000, model, density=1.0, seed=6)
    agents['velocity'] *= 3.0
    # make sure some agents cross every strip border within the horizon, whatever dt the adaptive rule picks
    from crowddynamics_b200.parallel import lattice_of, partition_columns
    bin_size = 3.6 * 1.10
    lat = lattice_of(agents['position'], bin_size)
    for col in partition_columns(lat[0], lat[2], world)[1:-1]:
        near = np.abs(agents['position'][:, 0] - col * bin_size) < 0.25
        agents['velocity'][near] = np.where(agents['position'][near, :1] < col * bin_size, 1.0, -1.0) * np.array([[4.0, 0.0]])
    fields = [S.direction_field(0.5, (0, 0, side, side), 'swirl')]
    steps = 14
    dev = DeviceAgents(MODEL_CIRCULAR if model == 'circular' else MODEL_THREE_CIRCLE)
    dev.set_rebuild_policy(0.10, 16, 0)
    dev.upload(agents)
    dev.set_obstacles(obstacles)
    dev.set_navigation_field(0, *fields[0])
    dev.step(2, _lib.STEP_ALL, 3.6, *dts, want_dt=False)
    dev.step(steps - 2, _lib.STEP_ALL, 3.6, *dts, want_dt=False)
    ref = agents.copy()
    dev.download(ref)
    assert dev.rebuild_stats()['kept'] > 0
    dev.close()
    sims = [StripSimulation.from_global(agents, obstacles, fields, 3.6, r, world, device_index=0, dt_min=dts[0], dt_max=dts[1],
                                        skin=0.10, max_interval=16) for r in range(world)]
    group = LocalGroup(sims, direct=direct)
    owned0 = [set(s.export(agents.dtype)[1].tolist()) for s in sims]
    group.step(steps)
    torch.cuda.synchronize()
    stats = [s.dev.rebuild_stats() for s in sims]
    assert all(st['kept'] >= 3 and st['rebuilds'] >= 2 for st in stats), stats
    assert sims[0].interval > 1
    got, ids = group.export(agents.dtype)
    assert (ids == np.arange(len(agents))).all()          # every agent owned by exactly one strip
    owned1 = [set(s.export(agents.dtype)[1].tolist()) for s in sims]
    if dts[0] == dts[1]:       # (adaptive dt caps the displacement at 1.4 cm per step: too little to cross within the horizon)
        assert any(a != b for a, b in zip(owned0, owned1))   # migration happened (on the steps before a rebuild)
    d = {k: float(np.abs(got[k] - ref[k]).max()) for k in ('position', 'velocity')}
    assert d['position'] <= 1e-8 and d['velocity'] <= 1e-6, d
    if model == 'three_circle':
        assert np.abs(got['orientation'] - ref['orientation']).max() <= 1e-7


@pytest.mark.parametrize('model', ['circular', 'three_circle'])
@pytest.mark.parametrize('world', [2, 3])
@pytest.mark.parametrize('skin', [0.0, 0.10], ids=['rebuild-every-step', 'kept-block-lists'])
def test_domain_nodes_on_strips(model, world, skin):
    """InsideDomain / TargetReached (simulation/logic.py:343-387) with the crowd split into strips: the flags live in arrays
    indexed by global agent id, migrating agents take theirs along, counts are summed over the strips.  Against the oracle
    applied to the single-device trajectory, update by update."""
    import torch
    from oracle import crowd_oracle as O
    agents, obstacles, side = S.uniform_crowd(20000 if skin else 6000, model, density=1.0, seed=9)
    agents['velocity'] *= 6.0 if not skin else 2.5
    n = len(agents)
    c = side / 2
    domain = np.array([(0.2 * side, 0.15 * side), (0.85 * side, 0.1 * side), (0.9 * side, 0.8 * side), (c, 0.95 * side), (0.1 * side, 0.7 * side)])
    goals = [np.array([(c - 9, c - 9), (c + 9, c - 9), (c + 9, c + 9), (c - 9, c + 9)]),
             np.array([(0.0, 0.0), (0.3 * side, 0.0), (0.0, 0.3 * side)])]
    fields = [S.direction_field(0.5, (0, 0, side, side), 'swirl')]
    sims = [StripSimulation.from_global(agents, obstacles, fields, 3.6, r, world, device_index=0, skin=skin) for r in range(world)]
    group = LocalGroup(sims)
    for s in sims:
        s.set_domain(domain, n, agents['active'])
        s.set_targets(goals, n)
    ref_dev = DeviceAgents(MODEL_CIRCULAR if model == 'circular' else MODEL_THREE_CIRCLE)
    ref_dev.set_rebuild_policy(0.10, 16 if skin else 1, 0)
    ref_dev.upload(agents)
    ref_dev.set_obstacles(obstacles)
    ref_dev.set_navigation_field(0, *fields[0])
    ref = agents.copy()
    reached_ref = [np.zeros(n, dtype=bool) for _ in goals]
    inactive_ref = inactive = 0
    updates = 10
    for it in range(updates):
        group.step(1)
        ref_dev.step(1, _lib.STEP_ALL, 3.6, 0.01, 0.01, want_dt=False)
        ref_dev.download(ref)
        inactive += sum(s.dev.inside_domain() for s in sims)              # what StripSimulation.inside_domain() all-reduces
        counts = np.sum([s.dev.target_reached(len(goals)) for s in sims], axis=0)
        inactive_ref += O.inside_domain(ref, domain)
        for g, r in zip(goals, reached_ref):
            O.target_reached(ref, g, r)
        assert inactive == inactive_ref, it
        assert counts.tolist() == [int(r.sum()) for r in reached_ref], it
    torch.cuda.synchronize()
    assert inactive_ref > 0 and all(r.any() for r in reached_ref)
    # every agent's flags, read from the strip that owns it now, equal the oracle's
    seen = np.zeros(n, dtype=int)
    moved = 0
    for r, s in enumerate(sims):
        ids, active, reached = s.owned_flags(agents.dtype)
        seen[ids] += 1
        assert (active == ref['active'][ids]).all()
        for p in range(len(goals)):
            assert (reached[p] == reached_ref[p][ids]).all()
    assert (seen == 1).all()
    ref_dev.close()


def test_settle_moves_misplaced_agents():
    """Set-up path of the weak-scaling benchmark: ranks generate agents by coordinate, strips are aligned to cell columns,
    settle() hands the misplaced ones to their owner."""
    import torch
    from crowddynamics_b200.parallel import CudaStripDevice, lattice_of, partition_columns, owner_of_columns
    agents, obstacles, side = S.uniform_crowd(5000, 'three_circle', density=1.0, seed=1)
    lattice = lattice_of(agents['position'], 3.6)
    bounds = partition_columns(lattice[0], lattice[2], 2)
    split_x = (bounds[1] + 0.45) * 3.6                     # not on a column edge: some agents land on the wrong rank
    sims = []
    for r in range(2):
        mine = (agents['position'][:, 0] < split_x) == (r == 0)
        dev = CudaStripDevice(MODEL_THREE_CIRCLE, 8000, 0, stream=torch.cuda.current_stream().cuda_stream)
        dev.upload(np.ascontiguousarray(agents[mine]), np.nonzero(mine)[0])
        dev.set_obstacles(obstacles)
        sims.append(StripSimulation(dev, r, 2, bounds, lattice, 3.6, 600, 600, torch.device('cuda', 0), int(mine.sum())))
    group = LocalGroup(sims)
    group.settle()
    torch.cuda.synchronize()
    cols = np.floor(agents['position'][:, 0] / 3.6).astype(np.int64)
    owner = owner_of_columns(cols, bounds)
    for r, sim in enumerate(sims):
        rec, ids = sim.export(agents.dtype)
        assert sorted(ids.tolist()) == np.nonzero(owner == r)[0].tolist()
        assert (rec['position'] == agents['position'][ids]).all() and (rec['mass'] == agents['mass'][ids]).all()
    # and the settled strips step like the single device
    ref = _single(agents, obstacles, [], 5, 0.01, 0.01)
    group.step(5)
    got, ids = group.export(agents.dtype)
    assert np.abs(got['position'] - ref['position']).max() <= 1e-12


def _gpu_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_gpu_count() < 2, reason='needs >= 2 GPUs (one process per GPU over NCCL)')
@pytest.mark.parametrize('exchange', ['nccl', 'peer'])
def test_nccl_strips_one_process_per_gpu(exchange):
    """torchrun with one rank per visible GPU (at most 4): halo + migrant exchange over NCCL send / recv, or by one-sided
    writes over NVLink peer memory; the gathered result must equal the single-GPU trajectory bit for bit
    (tests/run_strips_nccl.py)."""
    n = min(_gpu_count(), 4)
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(n), '--master-addr', '127.0.0.1',
           '--master-port', str(port), os.path.join(ROOT, 'tests', 'run_strips_nccl.py')]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, STRIP_EXCHANGE=exchange))
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert out.stdout.count(' OK') == 4 and 'FAIL' not in out.stdout
