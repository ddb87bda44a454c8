#!/usr/bin/env python
"""bench.py -- agent-steps/sec of the replaced crowddynamics sub-tree on B200 (BASELINE.json metric).

One "step" = one MultiAgentSimulation.update() of the replaced nodes over the whole crowd:
navigation sample -> orientation -> adjusting -> block list + agent-agent -> agent-obstacle -> adaptive-dt
velocity Verlet -> reset.  Workload at N=1: synthetic 1M ThreeCircle agents (BASELINE.json configs[2], the
configuration the metric is quoted on), uniform density 1 agent/m^2 in a walled square room, static direction field.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--model three_circle|circular] [--agents M] [--density R]
    python bench.py --impl reference ...      # the CPU arm: the oracle port of the reference on the host cores

Under torchrun (N > 1) every rank owns one strip of a domain N times as wide (weak scaling), exchanging halo agents and
migrants with its strip neighbours over NCCL each step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALGO_BYTES = {'circular': 184 + 16, 'three_circle': 312 + 16}   # state bytes per agent-step + (U,V) sample
CELL = 3.6


def profile_counters(model, n_agents, density):
    """Per-kernel counters of ONE launch from the committed ncu --set full summary (profiles/ncu_full_<model>_<tag>.txt,
    the newest tag) -- only valid for the workload that was profiled (1 M agents, 1 /m^2).
    -> {'source': path, 'kernels': {base kernel name: {metric: value}}}"""
    if n_agents != 1000000 or abs(density - 1.0) > 1e-12:
        return None
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, 'profiles', 'ncu_full_%s_r*.txt' % model)))     # tags sort by round
    if not files:
        return None
    out = {'source': os.path.relpath(files[-1], ROOT), 'kernels': {}}
    unit_scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}
    cur = None
    with open(files[-1]) as f:
        for line in f:
            if line.startswith('kernel:'):
                name = line.split(':', 1)[1].strip()
                name = name.replace('void ', '').split('<')[0].split('(')[0].strip()
                cur = out['kernels'].setdefault(name, {})
                continue
            c = line.split()
            if cur is None or len(c) < 2:
                continue
            try:
                if c[0] in ('dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__time_duration.sum'):
                    cur[c[0]] = float(c[-1].replace(',', '')) * unit_scale.get(c[1], 1.0)
                elif c[0] in ('smsp__inst_executed_pipe_fp64.sum', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active'):
                    cur[c[0]] = float(c[-1].replace(',', ''))
            except ValueError:
                pass
    return out


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index=0):
        self.rows = []
        self.proc = None
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu_index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def mark(self):
        return time.time()

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        rows = [r for (ts, r) in self.rows if (t0 is None or ts >= t0) and (t1 is None or ts <= t1 + 0.12)]
        if not rows:
            rows = [r for (_, r) in self.rows[-2:]]
        for r in rows:
            c = [x.strip() for x in r.split(',')]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); smax.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(names, c[5:9]):
                if v.lower() == 'active':
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(smax) if smax else None,
                'samples': len(sm), 'reasons': sorted(reasons)}


def make_crowd(model, n, density, seed, origin=(0.0, 0.0)):
    from crowddynamics_b200 import synthetic as S
    agents, obstacles, side = S.uniform_crowd(n, model, density=density, seed=seed, origin=origin)
    return agents, obstacles, side


def make_field(side, origin=(0.0, 0.0), step=1.0):
    """Static direction field towards an exit in the middle of the right wall (synthetic, (ny, nx) [iy, ix] layout).
    step = 1 m keeps the two maps at 2 x 8 MB for the 1000 m room (the reference default 0.1 m would be 2 x 800 MB)."""
    from crowddynamics_b200 import synthetic as S
    x0, y0 = origin
    return S.direction_field(step, (x0, y0, x0 + side, y0 + side), 'exit', point=(x0 + side, y0 + side / 2))


# ---------------------------------------------------------------------------------------------------------------------
def cpu_port_rate(model, n_sample, steps, density, threads=1):
    """agent-steps/s of the C oracle port (serial algorithm, like the reference) on `threads` independent replicas."""
    from oracle import crowd_oracle as O
    from crowddynamics_b200 import synthetic as S
    O.lib()
    crowds = []
    for t in range(threads):
        a, obs, side = S.uniform_crowd(n_sample, model, density=density, seed=100 + t)
        crowds.append((a, obs, [make_field(side)]))

    def work(k):
        a, obs, fields = crowds[k]
        for _ in range(steps):
            O.step(a, obs, fields, CELL, 0.01, 0.01)
    for k in range(threads):       # warm-up: one step each
        a, obs, fields = crowds[k]
        O.step(a, obs, fields, CELL, 0.01, 0.01)
    t0 = time.perf_counter()
    if threads == 1:
        work(0)
    else:
        th = [threading.Thread(target=work, args=(k,)) for k in range(threads)]
        [x.start() for x in th]
        [x.join() for x in th]
    dt = time.perf_counter() - t0
    return threads * n_sample * steps / dt, dt


def run_reference(args, rank, world):
    """--impl reference: the CPU arm.  The reference is Python + numba and /root/reference does not exist on the GPU box,
    so this times the oracle port (bit-identical to the numba code on the golden vectors) on all host threads: the
    algorithm is serial, so the threads run independent replicas of a bounded sample of the same workload."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_sample = args.cpu_agents or (50000 if args.model == 'three_circle' else 200000)
    # warm-up steps are run inside cpu_port_rate (1 per replica); `steps` timed steps per replica
    steps = max(1, min(args.steps, 5))
    rate, secs = cpu_port_rate(args.model, n_sample, steps, args.density, threads=cores)
    sample = ('%d independent replicas of %d %s agents (density %.3g /m^2) x %d steps, C port of the numba reference (bit-identical to it; '
              'the numba original is 8 - 10x slower than this port: profiles/reference_numba_vs_port_container_cpu.json)' % (
                  cores, n_sample, args.model, args.density, steps))
    line = {
        'impl': 'reference', 'metric': 'agent-steps/sec', 'value': rate, 'unit': 'agent-steps/s', 'n_gpus': args.gpus,
        'steps': steps, 'warmup': 1, 'ms_per_step': 1e3 * secs / steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': workload_config(args, args.agents, args.gpus),
        'cpu_baseline': {'value': rate, 'unit': 'agent-steps/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': rate, 'unit': 'agent-steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line), flush=True)


def run_hallway(args):
    """BASELINE config 1: examples.simulations.Hallway, 50 Circular agents, 1000 update() iterations -- the launch-bound end of
    the path.  Ways through the boundary: (a) one resident FusedStep node per update(), dt read back every update like the
    reference's Integrator node does (a host synchronisation per update), and (a') the same with deferred scalars (no
    synchronisation); (b) the same 1000 iterations as ONE cdb_step call -- one launch of the single-block small-crowd kernel
    (csrc/small_kernel.cuh) -- and (b') on the general pipeline (pairs of steps replayed as a CUDA graph); (c) the seven
    strict nodes per update.  Beside them the serial C port of the reference on one host core (the numba original is ~10x
    slower than the port)."""
    import torch
    from crowddynamics_b200 import _lib, logic as L, synthetic as S
    from crowddynamics_b200.engine import DeviceAgents
    from crowddynamics_b200.structures import MODEL_CIRCULAR
    from oracle import crowd_oracle as O
    updates = 1000

    def fused_nodes(deferred=False):
        agents, obstacles, fields = S.hallway(seed=0)
        sim = L.MultiAgentSimulation(agents, obstacles, fields)
        step = L.FusedStep(sim, step=0.1, deferred=deferred)
        sim.logic = (L.ScalarsSync(sim) << step) if deferred else step
        sim.update()
        t0 = time.perf_counter()
        for _ in range(updates):
            sim.update()
        if deferred:
            sim.logic.flush()
        sim.logic.state.sync_host()
        return time.perf_counter() - t0, agents

    def one_call(small=True):
        agents, obstacles, fields = S.hallway(seed=0)
        dev = DeviceAgents(MODEL_CIRCULAR)
        dev.set_small_crowd_max(256 if small else 0)
        dev.upload(agents); dev.set_obstacles(obstacles)
        for t, f in enumerate(fields):
            dev.set_navigation_field(t, *f)
        dev.step(2, _lib.STEP_ALL, CELL, 0.01, 0.01, want_dt=False)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        dev.step(updates, _lib.STEP_ALL, CELL, 0.01, 0.01, want_dt=False)
        dev.download(agents)
        secs = time.perf_counter() - t0
        dev.close()
        return secs, agents

    def strict_nodes():
        agents, obstacles, fields = S.hallway(seed=0)
        sim = L.MultiAgentSimulation(agents, obstacles, fields)
        sim.logic = L.hallway_logic(sim, mode='strict')
        sim.update()
        t0 = time.perf_counter()
        for _ in range(updates):
            sim.update()
        return time.perf_counter() - t0, agents

    def cpu_port():
        agents, obstacles, fields = S.hallway(seed=0)
        O.lib()
        O.step(agents, obstacles, fields, CELL, 0.01, 0.01)
        t0 = time.perf_counter()
        for _ in range(updates):
            O.step(agents, obstacles, fields, CELL, 0.01, 0.01)
        return time.perf_counter() - t0, agents

    out = {}
    for name, fn in (('fused_node_per_update', fused_nodes), ('fused_node_per_update_deferred_scalars', lambda: fused_nodes(True)),
                     ('one_cdb_step_call', one_call), ('one_cdb_step_call_general_pipeline_cuda_graphs', lambda: one_call(False)),
                     ('strict_seven_nodes_per_update', strict_nodes), ('cpu_port_1_core', cpu_port)):
        secs, agents = fn()
        out[name] = {'seconds': secs, 'updates_per_s': updates / secs, 'agent_steps_per_s': 50 * updates / secs,
                     'us_per_update': 1e6 * secs / updates}
    best = out['one_cdb_step_call']
    line = {'metric': 'agent-steps/sec', 'value': best['agent_steps_per_s'], 'unit': 'agent-steps/s', 'n_gpus': 1, 'steps': updates,
            'warmup': 2, 'ms_per_step': best['us_per_update'] * 1e-3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': 'BASELINE config 1: Hallway, 50 Circular agents, 40 x 5 m, 2 walls, 2 targets (constant +-x fields '
                                   'instead of the FMM field), dt 0.01, 1000 updates; latency bound on a GPU: one thread block holds the crowd'},
            'hallway': out,
            'cpu_baseline': {'value': out['cpu_port_1_core']['agent_steps_per_s'], 'unit': 'agent-steps/s', 'cores': 1, 'kind': 'port',
                             'sample': 'the whole workload: 1000 updates of the 50-agent Hallway, serial C port of the reference'}}
    print(json.dumps(line), flush=True)


def strip_parity_check(model, rank, world, local_rank, dist):
    """N > 1: before the timed run, every rank steps its strip of a small crowd over NCCL (halo + migrants + adaptive-dt
    all-reduce); rank 0 gathers the agents and compares them with a single-device run of the same crowd.
    -> {'status': 'ok' | 'mismatch', ...} on rank 0, None elsewhere.  The strips must reproduce the single device bit for bit
    (same pairs, same arithmetic, same summation order)."""
    import torch
    from crowddynamics_b200 import _lib, synthetic as S
    from crowddynamics_b200.engine import DeviceAgents
    from crowddynamics_b200.parallel import StripSimulation
    from crowddynamics_b200.structures import MODEL_CIRCULAR, MODEL_THREE_CIRCLE
    steps, n = 12, 40000
    agents, obstacles, side = S.uniform_crowd(n, model, density=1.0, seed=3, overlap_fraction=0.02)
    agents['velocity'] *= 3.0                           # plenty of migrants across the strip borders
    fields = [S.direction_field(0.5, (0, 0, side, side), 'swirl')]
    sim = StripSimulation.from_global(agents, obstacles, fields, CELL, rank, world, device_index=local_rank, dist=dist,
                                      dt_min=0.001, dt_max=0.01)
    sim.step(steps)
    torch.cuda.synchronize()
    rec, ids = sim.export(agents.dtype)
    gathered = [None] * world
    dist.all_gather_object(gathered, (rec.view(np.uint8).reshape(len(rec), -1), ids))
    if rank != 0:
        return None
    got = np.zeros_like(agents)
    seen = np.zeros(len(agents), dtype=int)
    for raw, i in gathered:
        got[i] = np.ascontiguousarray(raw).view(agents.dtype).reshape(-1)
        seen[i] += 1
    dev = DeviceAgents(MODEL_CIRCULAR if model == 'circular' else MODEL_THREE_CIRCLE, device=local_rank)
    dev.set_rebuild_policy(0.10, 1)      # block list rebuilt at every step, like these strips: same summation order, bit for bit
    dev.upload(agents); dev.set_obstacles(obstacles); dev.set_navigation_field(0, *fields[0])
    dev.step(steps, _lib.STEP_ALL, CELL, 0.001, 0.01, want_dt=False)
    ref = agents.copy(); dev.download(ref); dev.close()
    err = float(np.abs(got['position'] - ref['position']).max())
    migrated = int(sum(len(i) for _, i in gathered)) == n and bool((seen == 1).all())
    ok = migrated and err == 0.0
    return {'status': 'ok' if ok else 'mismatch', 'agents': n, 'steps': steps, 'ranks': world, 'every_agent_owned_once': migrated,
            'max_abs_position_diff_vs_single_device': err, 'adaptive_dt': True}


def strip_kept_check(model, rank, world, local_rank, dist):
    """N > 1: the strips with KEPT block lists (resident-order steps) against one device with the same policy.  The rebuild
    schedules differ (each adapts on its own), so the two agree up to summation order: 1e-16 per step, doubled per step by the
    dynamics -- 10 steps, bar 1e-9 m."""
    import torch
    from crowddynamics_b200 import _lib, synthetic as S
    from crowddynamics_b200.engine import DeviceAgents
    from crowddynamics_b200.parallel import StripSimulation
    from crowddynamics_b200.structures import MODEL_CIRCULAR, MODEL_THREE_CIRCLE
    steps, n = 10, 60000
    agents, obstacles, side = S.uniform_crowd(n, model, density=1.0, seed=5)
    agents['velocity'] *= 1.5
    fields = [S.direction_field(0.5, (0, 0, side, side), 'swirl')]
    sim = StripSimulation.from_global(agents, obstacles, fields, CELL, rank, world, device_index=local_rank, dist=dist,
                                      skin=0.10, max_interval=16)
    sim.step(steps)
    torch.cuda.synchronize()
    stats = sim.dev.rebuild_stats()
    rec, ids = sim.export(agents.dtype)
    gathered = [None] * world
    dist.all_gather_object(gathered, (rec.view(np.uint8).reshape(len(rec), -1), ids))
    if rank != 0:
        return None
    got = np.zeros_like(agents)
    seen = np.zeros(len(agents), dtype=int)
    for raw, i in gathered:
        got[i] = np.ascontiguousarray(raw).view(agents.dtype).reshape(-1)
        seen[i] += 1
    dev = DeviceAgents(MODEL_CIRCULAR if model == 'circular' else MODEL_THREE_CIRCLE, device=local_rank)
    dev.set_rebuild_policy(0.10, 16, 0)
    dev.upload(agents); dev.set_obstacles(obstacles); dev.set_navigation_field(0, *fields[0])
    dev.step(2, _lib.STEP_ALL, CELL, 0.01, 0.01, want_dt=False)
    dev.step(steps - 2, _lib.STEP_ALL, CELL, 0.01, 0.01, want_dt=False)
    ref = agents.copy(); dev.download(ref); dev.close()
    err = float(np.abs(got['position'] - ref['position']).max())
    owned = bool((seen == 1).all())
    return {'status': 'ok' if owned and err <= 1e-9 and stats['kept'] > 0 else 'mismatch', 'agents': n, 'steps': steps,
            'kept_steps_rank0': stats['kept'], 'rebuilds_rank0': stats['rebuilds'], 'interval': sim.interval,
            'every_agent_owned_once': owned, 'max_abs_position_diff_vs_single_device': err, 'bar': 1e-9}


def workload_config(args, n_per_gpu, world):
    geom = ('a room with a door (11 wall segments)' if getattr(args, 'workload', 'room') == 'room_exit'
            else 'a walled square room, 4 wall segments')
    strong = world > 1 and getattr(args, 'scaling', 'weak') == 'strong'
    per = n_per_gpu // world if strong else n_per_gpu
    field = ('exit navigation field built on the device from the geometry (step %g m)' % args.field_step
             if getattr(args, 'field_step', 0) > 0 and args.workload == 'room_exit' else 'static exit direction field (step 1 m)')
    return {'workload': 'synthetic %d %s agents %s, uniform %.3g agents/m^2 in %s, %s, cell 3.6 m, dt_min=dt_max=0.01'
                        % (n_per_gpu, args.model, 'in total (one room, %d strips)' % world if strong else 'per GPU', args.density, geom, field),
            'agents_per_gpu': per, 'agents_total': per * world, 'agent_model': args.model, 'density': args.density,
            'parallelism': 'strips%d' % world if world > 1 else 'single',
            'l2_policy': 'inputs larger than L2 (%d MB of SoA state streamed per step)'
                         % (per * (36 if args.model == 'three_circle' else 20) * 8 // 2 ** 20)}


# ---------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=300)
    ap.add_argument('--warmup', type=int, default=10)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--model', default='three_circle', choices=['circular', 'three_circle'])
    ap.add_argument('--agents', type=int, default=1000000, help='agents per GPU')
    ap.add_argument('--density', type=float, default=1.0)
    ap.add_argument('--workload', default='room', choices=['room', 'room_exit', 'hallway'],
                    help="room: walled square (configs 2/3/5); room_exit: room with a door, 11 wall segments, exit field (config 4); "
                         "hallway: BASELINE config 1 (50 agents x 1000 updates, launch-bound)")
    ap.add_argument('--cpu-agents', type=int, default=0, help='agents in the CPU baseline sample')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-fp64-peak', action='store_true', help='skip the DFMA peak micro-benchmark (used under ncu)')
    ap.add_argument('--e2e-steps', type=int, default=3)
    ap.add_argument('--e2e-crowds', type=int, default=4,
                    help='independent crowds in flight in the e2e leg (1: a single crowd, copies and step strictly serial)')
    ap.add_argument('--field-step', type=float, default=0.0,
                    help='room_exit: build the exit navigation field ON THE DEVICE from the wall / door geometry at this grid '
                         'step (reference default 0.1) instead of uploading a synthetic host field')
    ap.add_argument('--scaling', default='weak', choices=['weak', 'strong'],
                    help='N > 1: weak = --agents per GPU (a row of rooms); strong = --agents in total, ONE room split into strips '
                         '(BASELINE configs 4 / 5)')
    ap.add_argument('--exchange', default='peer', choices=['peer', 'nccl'],
                    help='N > 1: halo / migrant exchange by one-sided writes over NVLink peer memory (CUDA IPC; falls back to '
                         'NCCL send/recv when IPC is unavailable) or by NCCL send/recv')
    ap.add_argument('--refinement', type=int, default=0, help='search lattice: 0 automatic (cell_size / 2 where valid), 1 cell_size')
    ap.add_argument('--rebuild-max', type=int, default=16,
                    help='rebuild the block list at most every this many steps (resident-order steps, include/crowd_b200.h); '
                         '1 = rebuild at every step like the reference')
    ap.add_argument('--rebuild-min-agents', type=int, default=16384, help='crowds below this size rebuild at every step (and replay CUDA graphs)')
    ap.add_argument('--skin', type=float, default=0.10, help='widening of the search cells that the kept block list relies on')
    ap.add_argument('--variant', type=int, default=3, help='agent-agent kernel variant (3 once-per-pair, 2 both-sides fused kernel)')
    args = ap.parse_args()

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist
    from crowddynamics_b200 import _lib
    from crowddynamics_b200.engine import DeviceAgents
    from crowddynamics_b200.structures import MODEL_CIRCULAR, MODEL_THREE_CIRCLE

    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm')
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    _lib.load()
    if args.workload == 'hallway':
        return run_hallway(args)

    n = args.agents
    mid = MODEL_CIRCULAR if args.model == 'circular' else MODEL_THREE_CIRCLE
    itemsize = 228 if args.model == 'circular' else 316

    strip_parity = strip_kept = None
    field_build = None
    if world > 1:
        from crowddynamics_b200.parallel import StripSimulation
        strip_parity = strip_parity_check(args.model, rank, world, local_rank, dist)
        strip_kept = strip_kept_check(args.model, rank, world, local_rank, dist) if args.rebuild_max > 1 else None
        if args.scaling == 'strong':
            from crowddynamics_b200.parallel import strong_scaling_strip
            sim = strong_scaling_strip(args.model, n, args.density, rank, world, local_rank,
                                       geometry='room_exit' if args.workload == 'room_exit' else 'room', dist=dist,
                                       skin=args.skin if args.rebuild_max > 1 else 0.0, max_interval=args.rebuild_max)
        else:
            sim = StripSimulation.synthetic(args.model, n, args.density, rank, world, local_rank, seed=rank, dist=dist,
                                            skin=args.skin if args.rebuild_max > 1 else 0.0, max_interval=args.rebuild_max)
        exchange = 'nccl send/recv'
        if args.exchange == 'peer' and sim.connect_direct():
            exchange = 'one-sided writes over NVLink peer memory (CUDA IPC) + sequence flags'
        step_fn, dev = sim.step, sim.dev
        n_local = sim.n_owned
    else:
        if args.workload == 'room_exit':
            from crowddynamics_b200 import synthetic as S
            agents, obstacles, fields, side = S.room_with_exit(n, args.model, density=args.density, seed=0, step=1.0)
            field = fields[0]
        else:
            agents, obstacles, side = make_crowd(args.model, n, args.density, seed=0)
            field = make_field(side)
        dev = DeviceAgents(mid, capacity=n, device=local_rank)
        dev.set_stream(torch.cuda.current_stream().cuda_stream)
        dev.set_variant(args.variant)
        dev.set_search_refinement(args.refinement)
        dev.set_rebuild_policy(args.skin, args.rebuild_max, args.rebuild_min_agents)
        dev.set_obstacles(obstacles)
        field_build = None
        if args.workload == 'room_exit' and args.field_step > 0:
            # Field.navigation_to_target on the device (cdb_build_navigation_field): nothing of grid size on the host
            door = [(side + 5.0, side / 2 - 2.6, side + 5.0, side / 2 + 2.6)]
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            mg = dev.build_navigation_field(0, door, obstacles, (0.0, 0.0, side + 5.0, side), args.field_step, 0.5, 0.3)
            torch.cuda.synchronize()
            field_build = {'seconds': time.perf_counter() - t0, 'grid': list(mg.shape), 'cells': int(mg.shape[0] * mg.shape[1]),
                           'step': args.field_step, 'relaxation_rounds': dev.last_field_rounds,
                           'what': 'eikonal distance to the exit around the walls buffered by 0.5 m (fast iterative method), '
                                   'normalised gradient, buffer fill, wall blend; two eikonal solves'}
        else:
            dev.set_navigation_field(0, *field)
        # pinned host image of simulation.agents.array (packed records), the e2e leg copies it every step
        host = torch.empty(n * itemsize, dtype=torch.uint8).pin_memory()
        host.numpy()[:] = agents.view(np.uint8).reshape(-1)
        dev.upload_raw(host.data_ptr(), n)

        def step_fn(k):
            dev.step(k, _lib.STEP_ALL, CELL, 0.01, 0.01, want_dt=False)
        n_local = lambda: n   # noqa: E731

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    step_fn(args.warmup)
    sync()
    if world > 1:
        sim.profile_phases(True)
    dev.profile(True)
    launches0 = dev.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    t_begin = time.time()
    ev0.record()
    step_fn(args.steps)
    ev1.record()
    sync()
    clocks = sampler.stop(t_begin, time.time()) if rank == 0 else None
    ms = ev0.elapsed_time(ev1)
    launches = dev.launch_count() - launches0
    rebuild_stats = dev.rebuild_stats()
    if world > 1:
        rebuild_stats['interval'] = sim.interval      # agreed across ranks (all-reduced displacement)
    ph = dev.profile_read_phases()       # pre + block list, pair sweep, pair evaluation, step kernel, post, steps
    prof = (ph[0], ph[1] + ph[2] + ph[3], ph[4], ph[5])
    dev.profile(False)
    exchange_ms = sim.phase_ms() if world > 1 else None
    agents_total = n_local() if callable(n_local) else n_local
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        c = torch.tensor([float(agents_total), float(launches)], dtype=torch.float64, device='cuda')
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        agents_total, launches = int(c[0].item()), int(c[1].item())
    value = agents_total * args.steps / (ms * 1e-3)

    # ---- end to end through the public boundary with HOST buffers (rank-local, aggregated like `value`) -------------
    # `e2e`: the full round trip of simulation.agents.array -- H2D of the whole packed records, one step, D2H of the whole
    # records.  PCIe-bound by construction: profiles/pcie_probe_r2.txt measures 55.6 / 52.6 GB/s for whole-record DMA and
    # 18 - 26 GB/s for any strided subset of the 316-byte records, so moving only the 152 mutable bytes per direction is NOT
    # faster than moving the record; what the field masks buy is shown by the other legs (`e2e_variants`), where the bytes a
    # real node tree needs per update are a fraction of the record.
    e2e = None
    e2e_variants = None
    if world == 1:
        e2e_steps = max(1, args.e2e_steps)
        for _ in range(1):    # warm-up
            dev.upload_raw(host.data_ptr(), n); step_fn(1); dev.download_raw(host.data_ptr(), n)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            dev.upload_raw(host.data_ptr(), n)          # H2D of simulation.agents.array (pinned) + AoS -> SoA
            step_fn(1)
            dev.download_raw(host.data_ptr(), n)        # SoA -> AoS + D2H of the whole records
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        e2e_single = {'value': n * e2e_steps / e2e_s, 'unit': 'agent-steps/s', 'h2d_bytes_per_step': n * itemsize,
                      'd2h_bytes_per_step': n * itemsize, 'steps': e2e_steps,
                      'path': 'cdb_upload_agents_aos -> cdb_step(1) -> cdb_download_agents_aos on a pinned host array, ONE crowd: '
                              'upload, step and download strictly one after the other (PCIe used in one direction at a time)'}
        e2e = e2e_single
        if args.e2e_crowds > 1 and field_build is None:
            # The same round trip for `e2e_crowds` independent crowds of the SAME workload in flight (a replica study; the
            # reference arm likewise runs independent replicas on the host cores): one device handle + CUDA stream + host
            # thread per crowd (crowddynamics_b200.engine.host_round_trips), so the upload of one crowd overlaps the download
            # of another -- PCIe is full duplex -- and the kernels hide behind the copies.  Every step of every crowd still
            # moves its whole records both ways inside the timed region.
            from crowddynamics_b200.engine import host_round_trips
            crowds, keep = [], []
            for _k in range(args.e2e_crowds):
                d = DeviceAgents(mid, capacity=n, device=local_rank)       # own non-blocking stream
                d.set_variant(args.variant)
                d.set_search_refinement(args.refinement)
                d.set_rebuild_policy(args.skin, args.rebuild_max, args.rebuild_min_agents)
                d.set_obstacles(obstacles)
                d.set_navigation_field(0, *field)
                h = torch.empty(n * itemsize, dtype=torch.uint8).pin_memory()
                h.numpy()[:] = agents.view(np.uint8).reshape(-1)
                crowds.append((d, h.data_ptr(), n))
                keep.append(h)
            host_round_trips(crowds, 1, _lib.STEP_ALL, CELL, 0.01, 0.01)          # warm-up
            torch.cuda.synchronize()
            per_crowd = max(4, e2e_steps)
            t0 = time.perf_counter()
            host_round_trips(crowds, per_crowd, _lib.STEP_ALL, CELL, 0.01, 0.01)
            torch.cuda.synchronize()
            secs = time.perf_counter() - t0
            e2e = {'value': n * per_crowd * len(crowds) / secs, 'unit': 'agent-steps/s', 'h2d_bytes_per_step': n * itemsize,
                   'd2h_bytes_per_step': n * itemsize, 'steps': per_crowd * len(crowds), 'crowds_in_flight': len(crowds),
                   'single_crowd_value': e2e_single['value'],
                   'path': '%d independent crowds of the workload in flight, each stepped by cdb_upload_agents_aos -> cdb_step(1) -> '
                           'cdb_download_agents_aos on its own pinned host records (one host thread and CUDA stream per crowd): whole '
                           'records both ways for every step of every crowd; uploads overlap downloads (full-duplex PCIe). '
                           'One crowd alone: single_crowd_value' % len(crowds)}
            for d, _, _ in crowds:
                d.close()
            del crowds, keep

        def leg(name, body, what):
            body()                                       # warm-up
            torch.cuda.synchronize()
            dev.transfer_stats(reset=True)
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                body()
            torch.cuda.synchronize()
            secs = time.perf_counter() - t0
            up, down = dev.transfer_stats()
            return {'value': n * e2e_steps / secs, 'unit': 'agent-steps/s', 'h2d_bytes_per_step': up // e2e_steps,
                    'd2h_bytes_per_step': down // e2e_steps, 'steps': e2e_steps, 'path': what}

        F = _lib
        observer = F.F_POSITION | F.F_VELOCITY | (F.F_ORIENTATION if args.model == 'three_circle' else 0)
        force = F.F_FORCE | (F.F_TORQUE if args.model == 'three_circle' else 0)

        def strict_fused():
            step_fn(1); dev.download_raw(host.data_ptr(), n, F.F_ALL_MUTABLE)

        def strict_fused_with_host_force_node():
            dev.upload_fields_raw(host.data_ptr(), n, force); step_fn(1); dev.download_raw(host.data_ptr(), n, F.F_ALL_MUTABLE)

        def resident_observer():
            step_fn(1); dev.download_raw(host.data_ptr(), n, observer)

        e2e_variants = {
            'single_crowd': e2e_single,
            'strict_fused_step': leg('strict', strict_fused,
                                     'FusedStep in strict mode, steady state: nothing dirty on the host, one step, every '
                                     'mutable field written back into the pinned host records (zero-copy kernel)'),
            'strict_fused_step_host_force_node': leg('strict+host', strict_fused_with_host_force_node,
                                                     'the same with a host-side node that writes force (/ torque) every '
                                                     'update: only those fields are uploaded (cdb_upload_agents_fields)'),
            'resident_with_observer': leg('observer', resident_observer,
                                          'resident step + download of what an observer / SaveSimulationData-style node '
                                          'reads (position, velocity, orientation)'),
        }

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peaks()
    n_rank0 = agents_total // world if world > 1 else n
    roofline = None
    fp64 = None
    if prof[3] > 0:
        # One step = block list + three kernels of the once-per-pair pipeline.  The longest of them, k_finish, streams every
        # agent's state once (the SURVEY 8(d) algorithmic bytes) and is the HBM-bound one: it carries `roofline`.  k_sweep is
        # bound by the fp64 pipe: `roofline_fp64`.  Durations: CUDA events on the sim's stream around each phase.
        steps_p = ph[5]
        phases = {'pre_and_block_list': ph[0] / steps_p, 'pair_sweep': ph[1] / steps_p, 'pair_eval': ph[2] / steps_p,
                  'finish': ph[3] / steps_p, 'post': ph[4] / steps_p}
        k_ms = phases['finish']
        algo_bytes = ALGO_BYTES[args.model] * n_rank0
        achieved = algo_bytes / (k_ms * 1e-3) / 1e9
        pc = profile_counters(args.model, n_rank0, args.density)
        traffic = None
        kf = (pc or {}).get('kernels', {}).get('k_finish', {})
        if 'dram__bytes_read.sum' in kf:
            traffic = kf['dram__bytes_read.sum'] + kf['dram__bytes_write.sum']
        roofline = {'bound': 'hbm', 'kernel': 'k_finish<%s> (per-agent nodes, ordered sum of the pair contributions, walls, '
                                              'integrator, reset; the dominant kernel of the step)' % args.model,
                    'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                    'frac': achieved / peak, 'peak_source': peak_src, 'traffic': traffic,
                    'traffic_source': pc['source'] if pc else None,
                    'kernel_ms': k_ms, 'algorithmic_bytes_per_launch': algo_bytes,
                    'phase_ms_per_step': phases,
                    'whole_step': {'ms': ms / args.steps, 'achieved': algo_bytes / (ms / args.steps * 1e-3) / 1e9,
                                   'frac': algo_bytes / (ms / args.steps * 1e-3) / 1e9 / peak},
                    'note': 'the step is split over an fp64-bound kernel (k_sweep), a latency/divergence-bound one '
                            '(k_pair_eval) and this HBM-streaming one; see roofline_fp64 and DESIGN.md section 4'}
        # fp64 roofline of the sweep: DFMA peak measured live, fp64 warp-instructions per launch from the committed profile
        try:
            if args.no_fp64_peak:
                raise RuntimeError('skipped (--no-fp64-peak)')
            import ctypes as C
            tf = C.c_double()
            _lib.check(_lib.load().cdb_measure_fp64_peak(local_rank, C.byref(tf)))
            fp64 = {'kernel': 'k_sweep<%s> (forward-half-stencil pair classification)' % args.model,
                    'peak_tflops_measured_dfma': tf.value, 'phase_ms': phases['pair_sweep']}
            ks = (pc or {}).get('kernels', {}).get('k_sweep', {})
            key = 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active'
            if key in ks and 'gpu__time_duration.sum' in ks:
                fp64.update({'pipe_frac_ncu': ks[key] / 100.0, 'kernel_ms_ncu': ks['gpu__time_duration.sum'],
                             'source': pc['source']})
        except Exception as exc:   # pragma: no cover
            fp64 = {'note': str(exc)}
    line = {
        'metric': 'agent-steps/sec', 'value': value, 'unit': 'agent-steps/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': args.scaling if world > 1 else 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic', 'config': workload_config(args, n, world),
        'gpu_launches': launches, 'clocks': clocks, 'roofline': roofline, 'roofline_fp64': fp64, 'e2e': e2e,
        'e2e_variants': e2e_variants, 'field_build': field_build if world == 1 else None, 'strip_parity': strip_parity, 'strip_kept_block_lists': strip_kept,
        'block_list_policy': {'skin_fraction': args.skin, 'max_interval': args.rebuild_max, 'since_upload': rebuild_stats},
        'strip_phase_ms_rank0': exchange_ms, 'strip_exchange': exchange if world > 1 else None,
        'hbm_fraction_whole_step': value / world * ALGO_BYTES[args.model] / 1e9 / peak,
    }
    if world == 1 and not args.no_cpu_baseline:
        # the workload itself where one step of it fits the budget (1 M agents: ~8 s per three-circle step on one core)
        n_sample = args.cpu_agents or min(n, 1000000 if args.model == 'three_circle' else 2000000)
        cpu_steps = 2 if n_sample * (8e-6 if args.model == 'three_circle' else 1.5e-6) > 3 else 3
        rate, secs = cpu_port_rate(args.model, n_sample, cpu_steps, args.density, threads=1)
        line['cpu_baseline'] = {'value': rate, 'unit': 'agent-steps/s', 'cores': 1, 'kind': 'port',
                                'sample': '%d %s agents (density %.3g /m^2) x %d steps after 1 warm-up step, serial C '
                                          'port of the numba reference (bit-identical on the golden vectors), %.1f s'
                                          % (n_sample, args.model, args.density, cpu_steps, secs),
                                'host_cores': os.cpu_count()}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
