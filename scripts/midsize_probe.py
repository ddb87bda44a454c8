"""Mid-size crowds (257 .. 16 384 agents): one cdb_step(2000) call on the sim's own stream, (a) CUDA-graph replay of step pairs with the
block list rebuilt at every step, (b) resident-order steps (block list kept, plain launches)."""
import sys, time
import numpy as np
sys.path.insert(0, '.')
import torch
from crowddynamics_b200 import _lib, synthetic as S
from crowddynamics_b200.engine import DeviceAgents
from crowddynamics_b200.structures import MODEL_CIRCULAR, MODEL_THREE_CIRCLE

for model in ('circular', 'three_circle'):
    for n in (300, 1000, 4000, 16000, 60000):
        agents, obstacles, side = S.uniform_crowd(n, model, density=1.0, seed=1)
        fields = [S.direction_field(1.0, (0, 0, side, side), 'exit', point=(side, side / 2))]
        agents['target'] = 0
        row = []
        for name, policy, graphs in (('graph K=1', (0.10, 1, 0), True), ('plain K=1', (0.10, 1, 0), False), ('kept lists', (0.10, 16, 0), True)):
            dev = DeviceAgents(MODEL_CIRCULAR if model == 'circular' else MODEL_THREE_CIRCLE)
            dev.set_rebuild_policy(*policy); dev.set_graphs(graphs)
            dev.upload(agents); dev.set_obstacles(obstacles); dev.set_navigation_field(0, *fields[0])
            dev.step(20, _lib.STEP_ALL, 3.6, 0.01, 0.01, want_dt=False); dev.synchronize()
            t0 = time.perf_counter()
            dev.step(2000, _lib.STEP_ALL, 3.6, 0.01, 0.01, want_dt=False); dev.synchronize()
            us = (time.perf_counter() - t0) / 2000 * 1e6
            row.append('%s %.1f us' % (name, us))
            dev.close()
        print('%-12s n=%-6d  %s' % (model, n, '   '.join(row)), flush=True)
