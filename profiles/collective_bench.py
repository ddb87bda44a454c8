"""Timing of the collective-motion nodes on one GPU, with the C oracle (the CPU port of the reference
functions) timed beside them on a smaller crowd.  Usage: python profiles/collective_bench.py [--agents N] > out.json
Host wall clock around call + cdb_synchronize (the calls are single kernels / short kernel sequences); inputs resident."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
from crowddynamics_b200 import synthetic as S  # noqa: E402
from crowddynamics_b200.engine import DeviceAgents  # noqa: E402
from crowddynamics_b200.structures import model_of  # noqa: E402


def timed(fn, sync, reps=5, warm=2, prep=None):
    for _ in range(warm):
        if prep:
            prep()
        fn()
    sync()
    ts = []
    for _ in range(reps):
        if prep:
            prep()
            sync()
        t = time.perf_counter()
        fn()
        sync()
        ts.append(time.perf_counter() - t)
    return 1e3 * float(np.median(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--agents', type=int, default=1_000_000)
    ap.add_argument('--cpu-agents', type=int, default=20_000)
    ap.add_argument('--density', type=float, default=1.0)
    ap.add_argument('--model', default='circular')
    args = ap.parse_args()
    n = args.agents
    agents, obstacles, doors, side = S.leader_follower_crowd(n, args.model, density=args.density, seed=1, n_leaders=max(1, n // 500))
    dev = DeviceAgents(model_of(agents), capacity=n)
    dev.upload(agents)
    dev.set_obstacles(obstacles)
    out = {'agents': n, 'model': args.model, 'density': args.density, 'leaders': int(agents['is_leader'].sum()),
           'obstacles': len(obstacles), 'doors': len(doors), 'gpu_ms': {}, 'cpu_port': {}}

    def reset_states():
        dev.set_states(agents)

    reset_states()
    out['gpu_ms']['exit_detection'] = timed(lambda: dev.exit_detection(doors, 20.0, apply=False), dev.synchronize)
    out['gpu_ms']['leader_follower_with_herding(sight=10,k=5)'] = timed(
        lambda: dev.leader_follower_with_herding(10.0, 5), dev.synchronize, prep=reset_states)
    out['gpu_ms']['leader_follower(sight=20)'] = timed(lambda: dev.leader_follower(20.0), dev.synchronize, prep=reset_states)
    out['gpu_ms']['nearest_neighbors(sight=10,k=5) incl. D2H of the table'] = timed(lambda: dev.nearest_neighbors(10.0, 5), dev.synchronize)
    out['gpu_ms']['set_states (pageable H2D of the five States arrays, not in the lines above)'] = timed(reset_states, dev.synchronize)
    out['gpu_agents_per_s'] = {k: n / (v * 1e-3) for k, v in out['gpu_ms'].items()}

    from oracle import crowd_oracle as O
    m = args.cpu_agents
    a, obs, drs, _ = S.leader_follower_crowd(m, args.model, density=args.density, seed=1, n_leaders=max(1, m // 500))
    for name, fn in (('exit_detection', lambda x: O.exit_detection(drs, x, obs, 20.0)),
                     ('leader_follower_with_herding(sight=10,k=5)', lambda x: O.leader_follower_with_herding_interaction(x, obs, 10.0, 5)),
                     ('leader_follower(sight=20)', lambda x: O.leader_follower_interaction(x, obs, 20.0))):
        b = a.copy()
        t = time.perf_counter()
        fn(b)
        dt = time.perf_counter() - t
        out['cpu_port'][name] = {'agents': m, 'ms': 1e3 * dt, 'agents_per_s': m / dt, 'cores': 1}
    print(json.dumps(out))


if __name__ == '__main__':
    main()
