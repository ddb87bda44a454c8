"""Times the reference's OWN numba code (unmodified modules of /root/reference, loaded through oracle/ref_harness.py) next to the
C port on this container's host CPU -- the numba number BASELINE.md section 3 asks for.  It cannot run on the GPU box
(/root/reference does not exist there); the ratio port : numba measured here is what relates the box's cpu_baseline to numba."""
import sys, time, json
import numpy as np
sys.path.insert(0, '.')
from crowddynamics_b200 import synthetic as S
from oracle import ref_harness as H, crowd_oracle as O

R = H.load()
out = {}
for model, n in (('circular', 100000), ('three_circle', 50000)):
    agents, obstacles, side = S.uniform_crowd(n, model, density=1.0, seed=0)
    fields = [S.direction_field(1.0, (0, 0, side, side), 'exit', point=(side, side / 2))]
    agents['target'] = 0
    ra, oa = agents.copy(), agents.copy()
    H.step(R, ra, obstacles, fields, 3.6, 0.01, 0.01)            # JIT warm-up
    O.step(oa, obstacles, fields, 3.6, 0.01, 0.01)
    t0 = time.perf_counter(); steps = 2
    for _ in range(steps):
        H.step(R, ra, obstacles, fields, 3.6, 0.01, 0.01)
    t_ref = (time.perf_counter() - t0) / steps
    t0 = time.perf_counter()
    for _ in range(steps):
        O.step(oa, obstacles, fields, 3.6, 0.01, 0.01)
    t_port = (time.perf_counter() - t0) / steps
    same = all(((ra[f] == oa[f]) | (np.isnan(ra[f]) & np.isnan(oa[f]))).all() for f in ra.dtype.names)
    out[model] = {'agents': n, 'numba_reference_agent_steps_per_s': n / t_ref, 'c_port_agent_steps_per_s': n / t_port,
                  'port_over_numba': t_ref / t_port, 'bit_identical_after_3_steps': bool(same)}
    print(model, out[model], flush=True)
json.dump(out, open('profiles/reference_numba_vs_port_container_cpu.json', 'w'), indent=1)
