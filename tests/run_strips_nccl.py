"""Multi-process NCCL check of the strip decomposition (one GPU per rank):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/run_strips_nccl.py
Every rank steps its strip; rank 0 gathers all agents and compares with a single-GPU run of the same crowd."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
sys.path.insert(0, ROOT)
from crowddynamics_b200 import _lib, synthetic as S  # noqa: E402
from crowddynamics_b200.engine import DeviceAgents  # noqa: E402
from oracle import crowd_oracle as O  # noqa: E402
from crowddynamics_b200.parallel import StripSimulation  # noqa: E402
from crowddynamics_b200.structures import MODEL_CIRCULAR, MODEL_THREE_CIRCLE  # noqa: E402


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    ok = True
    exchange = os.environ.get('STRIP_EXCHANGE', 'nccl')      # 'peer': one-sided writes over NVLink peer memory (CUDA IPC)
    for model in ('circular', 'three_circle'):
        for dts in ((0.01, 0.01), (0.001, 0.01)):
            agents, obstacles, side = S.uniform_crowd(40000, model, density=1.0, seed=3, overlap_fraction=0.02)
            agents['velocity'] *= 3.0
            fields = [S.direction_field(0.5, (0, 0, side, side), 'swirl')]
            sim = StripSimulation.from_global(agents, obstacles, fields, 3.6, rank, world, device_index=local, dist=dist,
                                              dt_min=dts[0], dt_max=dts[1])
            if exchange == 'peer':
                assert sim.connect_direct(), 'CUDA IPC not available'
            # InsideDomain / TargetReached over the strips: flags by global id, counts all-reduced over the ranks
            domain = np.array([(0.2 * side, 0.15 * side), (0.85 * side, 0.1 * side), (0.9 * side, 0.8 * side), (0.1 * side, 0.7 * side)])
            goal = np.array([(0.4 * side, 0.4 * side), (0.6 * side, 0.4 * side), (0.6 * side, 0.6 * side), (0.4 * side, 0.6 * side)])
            sim.set_domain(domain, len(agents), agents['active'])
            sim.set_targets([goal], len(agents))
            inactive, reached = 0, None
            for _ in range(4):
                sim.step(5)
                inactive += sim.inside_domain()
                reached = sim.target_reached()
            torch.cuda.synchronize()
            dist.barrier()
            rec, ids = sim.export(agents.dtype)
            gathered = [None] * world
            dist.all_gather_object(gathered, (rec.view(np.uint8).reshape(len(rec), -1), ids))
            if rank == 0:
                got = np.zeros_like(agents)
                seen = np.zeros(len(agents), dtype=int)
                for raw, i in gathered:
                    got[i] = np.ascontiguousarray(raw).view(agents.dtype).reshape(-1)
                    seen[i] += 1
                dev = DeviceAgents(MODEL_CIRCULAR if model == 'circular' else MODEL_THREE_CIRCLE, device=local)
                dev.set_rebuild_policy(0.10, 1)       # rebuild at every step, like the strips: same summation order
                dev.upload(agents); dev.set_obstacles(obstacles); dev.set_navigation_field(0, *fields[0])
                ref = agents.copy()
                inactive_ref, reached_ref = 0, np.zeros(len(agents), dtype=bool)
                for _ in range(4):
                    dev.step(5, _lib.STEP_ALL, 3.6, dts[0], dts[1], want_dt=False)
                    dev.download(ref)
                    inactive_ref += O.inside_domain(ref, domain)
                    O.target_reached(ref, goal, reached_ref)
                dev.close()
                err = np.abs(got['position'] - ref['position']).max()
                good = (seen == 1).all() and err <= 1e-12 and inactive == inactive_ref and int(reached[0]) == int(reached_ref.sum())
                ok &= bool(good)
                print('strips %s world=%d %-12s dt=%s: owners ok=%s max |dx|=%.3e %s' % (
                    exchange, world, model, dts, (seen == 1).all(), err, 'OK' if good else 'FAIL'), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not ok:
        sys.exit(1)


if __name__ == '__main__':
    main()
