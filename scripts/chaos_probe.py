"""How fast do two runs that differ only in summation order drift apart?  (a) rebuild-every-step on the cell_size/2 lattice vs
on the cell_size lattice; (b) rebuild-every-step vs resident-order steps."""
import sys
import numpy as np
sys.path.insert(0, '.')
from crowddynamics_b200 import _lib, synthetic as S
from crowddynamics_b200.engine import DeviceAgents
from crowddynamics_b200.structures import MODEL_CIRCULAR, MODEL_THREE_CIRCLE

def run(model, agents, obstacles, fields, policy, refinement, checkpoints):
    dev = DeviceAgents(MODEL_CIRCULAR if model == 'circular' else MODEL_THREE_CIRCLE)
    dev.set_rebuild_policy(*policy)
    dev.set_search_refinement(refinement)
    dev.upload(agents); dev.set_obstacles(obstacles)
    for k, f in enumerate(fields):
        dev.set_navigation_field(k, *f)
    outs = []
    done = 0
    for c in checkpoints:
        dev.step(c - done, _lib.STEP_ALL, 3.6, 0.01, 0.01, want_dt=False); done = c
        o = agents.copy(); dev.download(o); outs.append(o)
    st = dev.rebuild_stats(); dev.close()
    return outs, st

cps = [2, 4, 8, 12, 16, 24, 32, 40]
for model in ('circular', 'three_circle'):
    agents, obstacles, side = S.uniform_crowd(30000, model, density=1.0, seed=41)
    fields = [S.direction_field(0.5, (0, 0, side, side), 'swirl')]
    a, _ = run(model, agents, obstacles, fields, (0.10, 1, 0), 1, cps)
    b, _ = run(model, agents, obstacles, fields, (0.10, 1, 0), 2, cps)
    c, st = run(model, agents, obstacles, fields, (0.10, 16, 0), 1, cps)
    print(model, 'chain stats', st)
    for k, cp in enumerate(cps):
        print('  step %2d  lattice-vs-lattice %.2e   rebuild-vs-kept %.2e   (max speed %.2f)' % (
            cp, np.abs(a[k]['position'] - b[k]['position']).max(), np.abs(a[k]['position'] - c[k]['position']).max(),
            np.hypot(*a[k]['velocity'].T).max()))
