"""Generates the golden vectors under tests/golden/ from the REFERENCE's own numba code.

Run in the build container only (needs /root/reference):   python tests/golden/generate.py
The reference hot-path modules are imported unmodified through oracle/ref_harness.py (shims listed there);
`cell_lists` is the harness's restatement ("parity unpinned" at that boundary -- see DESIGN.md).
Outputs are small .npz files; structured agent arrays are stored as raw uint8 rows (n, itemsize).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, '..', '..')))

from oracle import ref_harness as H  # noqa: E402
from crowddynamics_b200 import synthetic as S  # noqa: E402

CELL = 3.6


def raw(a):
    return np.ascontiguousarray(a).view(np.uint8).reshape(len(a), a.dtype.itemsize).copy()


def gen_step_fixture(R, model, n, seed, density, overlap, fstep=0.25):
    agents, obstacles, side = S.uniform_crowd(n, model, density=density, seed=seed, overlap_fraction=overlap)
    # two targets so that the per-target masking of Navigation.update is exercised; a few agents without target
    rng = np.random.default_rng(seed + 100)
    agents['target'] = rng.integers(-1, 2, size=n)
    bounds = (0.0, 0.0, side * 0.8, side * 0.9)      # smaller than the room: some agents fall outside the grid
    fields = [S.direction_field(fstep, bounds, 'swirl'), S.direction_field(fstep, bounds, 'exit', point=(1.0, 2.0))]
    out = dict(initial=raw(agents), obstacles=obstacles.view(np.float64).reshape(-1, 4), cell_size=CELL,
               field_bounds=np.array(bounds), field_step=fstep,
               U=np.stack([f[1][0] for f in fields]), V=np.stack([f[1][1] for f in fields]),
               dt_min=0.001, dt_max=0.01)
    a = agents.copy()
    H.node_navigation(R, a, fields); out['after_navigation'] = raw(a)
    H.node_orientation(R, a); out['after_orientation'] = raw(a)
    H.node_adjusting(R, a); out['after_adjusting'] = raw(a)
    b = a.copy()
    H.node_agent_agent(R, a, CELL); out['after_agent_agent'] = raw(a)
    H.node_agent_obstacle(R, a, obstacles); out['after_agent_obstacle'] = raw(a)
    dt = H.node_integrator(R, a, 0.001, 0.01); out['after_integrator'] = raw(a)
    H.node_reset(R, a); out['after_reset'] = raw(a)
    dts = [dt]
    for _ in range(4):
        dts.append(H.step(R, a, obstacles, fields, CELL, 0.001, 0.01))
    out['after_5_steps'] = raw(a)
    out['dts'] = np.array(dts)
    # block list tables of the (restated) cell_lists on the state the agent-agent node saw
    pi, cc, co, gs = R.cell_lists.add_to_cells(np.ascontiguousarray(b['position']), CELL)
    out.update(points_indices=pi, cells_count=cc, cells_offset=co, grid_shape=gs)
    return out


def gen_pair_fixture(R, seed=7, n_pairs=400):
    """Pair-level vectors straight from the reference kernels: social force (both models), distances, walls."""
    rng = np.random.default_rng(seed)
    out = {}
    for model in ('circular', 'three_circle'):
        a, _, _ = S.random_crowd(2 * n_pairs, model, half_width=1.0, seed=seed)
        # pairs (2k, 2k+1) at separations from overlapping to beyond the sight gate
        sep = rng.uniform(0.05, 4.5, n_pairs)
        ang = rng.uniform(-np.pi, np.pi, n_pairs)
        a['position'][0::2] = rng.uniform(-5, 5, (n_pairs, 2))
        a['position'][1::2] = a['position'][0::2] + np.stack((np.cos(ang), np.sin(ang)), 1) * sep[:, None]
        # a few exactly degenerate cases: coincident centres, zero relative velocity
        a['position'][1] = a['position'][0]
        a['velocity'][3] = a['velocity'][2]
        if model != 'circular':
            S.set_shoulders(a)
        fi = np.zeros((n_pairs, 2)); fj = np.zeros((n_pairs, 2))
        force = R.power_law.force_social_circular if model == 'circular' else R.power_law.force_social_three_circle
        inter = R.interactions.interaction_agent_agent_circular if model == 'circular' else \
            R.interactions.interaction_agent_agent_three_circle
        for k in range(n_pairs):
            fi[k], fj[k] = force(a, 2 * k, 2 * k + 1)
        b = a.copy()
        for k in range(n_pairs):
            inter(2 * k, 2 * k + 1, b)
        out[model + '_agents'] = raw(a)
        out[model + '_social_i'] = fi
        out[model + '_social_j'] = fj
        out[model + '_after_interaction'] = raw(b)
    # walls: one agent x arbitrary segments, incl. degenerate (p0 == p1) and centre exactly on the line
    segs = rng.uniform(-3, 3, (60, 4))
    segs[0] = (1.0, 1.0, 1.0, 1.0)
    segs[1] = (-1.0, 0.0, 1.0, 0.0)
    for model in ('circular', 'three_circle'):
        a, _, _ = S.random_crowd(60, model, half_width=2.5, seed=seed + 1)
        a['position'][1] = (0.25, 0.0)       # on the line: np.sign(0) = 0 -> zero normal
        if model != 'circular':
            S.set_shoulders(a)
        obs = np.zeros(60, dtype=H.obstacle_type_linear)
        obs['p0'] = segs[:, :2]; obs['p1'] = segs[:, 2:]
        b = a.copy()
        fn = R.interactions.interaction_agent_circular_obstacle if model == 'circular' else \
            R.interactions.interaction_agent_three_circle_obstacle
        for k in range(60):
            fn(k, k, b, obs)
        out[model + '_wall_agents'] = raw(a)
        out[model + '_wall_after'] = raw(b)
    out['wall_segments'] = segs
    xs = np.concatenate((rng.uniform(-50, 50, 200), np.pi * np.arange(-7, 8), [0.0, -0.0, 2 * np.pi, -2 * np.pi]))
    out['wrap_in'] = xs
    out['wrap_out'] = R.vector2D.wrap_to_pi(xs)
    return out


def gen_known_answers(R):
    """The reference's two known-answer cases (core/motion/tests/test_power_law_benchmark.py:13-63), with the
    'adult' body means (conf/body_types.cfg) instead of its random draw."""
    out = {}
    for model, dtype in (('circular', H.agent_type_circular), ('three_circle', H.agent_type_three_circle)):
        for case, v2, phi2 in (('not_colliding', (1.0, 0.0), 0.0), ('colliding', (-1.0, 0.0), np.pi)):
            a = np.zeros(2, dtype=dtype)
            S.fill_adult_bodies(a, np.random.default_rng(0))
            a['radius'] = 0.255; a['r_t'] = 0.5882 * 0.255; a['r_s'] = 0.3725 * 0.255; a['r_ts'] = 0.6275 * 0.255
            a['mass'] = 73.5
            a['position'][1] = (2.0, 0.0)
            a['velocity'][0] = (1.0, 0.0); a['velocity'][1] = v2
            a['target_direction'] = a['velocity']
            if model != 'circular':
                a['orientation'][1] = phi2; a['target_orientation'][1] = phi2
                a['angular_velocity'] = 0.0
                S.set_shoulders(a)
            f = R.power_law.force_social_circular if model == 'circular' else R.power_law.force_social_three_circle
            fi, fj = f(a, 0, 1)
            out['%s_%s_agents' % (model, case)] = raw(a)
            out['%s_%s_force' % (model, case)] = np.stack((fi, fj))
    return out


def gen_hallway(R, steps=200):
    """BASELINE config 1 (Hallway, 50 Circular agents) as arrays; reference trajectory for `steps` updates."""
    agents, obstacles, fields = S.hallway(seed=0)
    a = agents.copy()
    traj = [a['position'].copy()]
    for k in range(steps):
        H.step(R, a, obstacles, fields, CELL, 0.01, 0.01)
        if (k + 1) % 50 == 0:
            traj.append(a['position'].copy())
    return dict(initial=raw(agents), final=raw(a), positions=np.stack(traj), steps=steps,
                obstacles=obstacles.view(np.float64).reshape(-1, 4))


def gen_collective_fixture(R, model, n=500, seed=11):
    """exit detection, k-nearest neighbours, herding, leader-follower (collective_motion.py,
    evacuation.py:137-174) on a leader/follower crowd in a room with inner walls."""
    cm, cl = R.collective_motion, R.cell_lists
    agents, obstacles, doors, side = S.leader_follower_crowd(n, model, density=0.5, seed=seed)
    sight, k, phi = 10.0, 5, 0.45 * np.pi
    pos = np.ascontiguousarray(agents['position'])
    vel = np.ascontiguousarray(agents['velocity'])
    out = dict(initial=raw(agents), obstacles=obstacles.view(np.float64).reshape(-1, 4), center_door=doors,
               sight=sight, size_nearest_other=k, phi=phi, detection_range=20.0)
    pts, cnt, off, shape = cl.add_to_cells(pos, sight)
    nbr = cm.find_nearest_neighbors(pos, sight, k, np.arange(len(cnt)), cl.neighboring_cells(shape), pts, cnt, off, obstacles)
    out['neighbors'] = nbr
    d, h = cm.herding_interaction(agents['is_follower'].copy(), pos, vel, nbr, 0.15, phi)
    out['herding_direction'], out['herding_has_direction'] = d, h
    a = agents.copy()
    out['lfh_direction'] = cm.leader_follower_with_herding_interaction(a, obstacles, sight, k)
    out['lfh_after'] = raw(a)
    a = agents.copy()
    out['lf_direction'] = cm.leader_follower_interaction(a, obstacles, 20.0)
    out['lf_after'] = raw(a)
    tg, has = R.evacuation.exit_detection(np.ascontiguousarray(doors), pos, obstacles, 20.0)
    out['detected_exit'], out['has_detected'] = tg, has
    # herding_relationship known answers (collective_motion.py:25-58), phi = pi / 2 and 0.45 pi
    rng = np.random.default_rng(seed)
    q = rng.normal(size=(200, 8))
    q[:10, 4:6] = 0.0                                 # zero velocity: (False, False)
    rel = np.array([[cm.herding_relationship(r[0:2].copy(), r[2:4].copy(), r[4:6].copy(), r[6:8].copy(), p)
                     for p in (np.pi / 2, phi)] for r in q])
    out['relationship_inputs'], out['relationship'] = q, rel
    return out


def main():
    R = H.load()
    if '--only-collective' in sys.argv:
        for model in ('circular', 'three_circle'):
            np.savez_compressed(os.path.join(HERE, 'collective_%s.npz' % model), **gen_collective_fixture(R, model))
        return
    for model in ('circular', 'three_circle'):
        fx = gen_step_fixture(R, model, n=300, seed=3, density=1.0, overlap=0.03)
        np.savez_compressed(os.path.join(HERE, 'step_%s.npz' % model), **fx)
        fx = gen_step_fixture(R, model, n=120, seed=4, density=0.125, overlap=0.0, fstep=0.75)
        np.savez_compressed(os.path.join(HERE, 'step_sparse_%s.npz' % model), **fx)
    np.savez_compressed(os.path.join(HERE, 'pairs.npz'), **gen_pair_fixture(R))
    np.savez_compressed(os.path.join(HERE, 'known_answers.npz'), **gen_known_answers(R))
    np.savez_compressed(os.path.join(HERE, 'hallway.npz'), **gen_hallway(R))
    for model in ('circular', 'three_circle'):
        np.savez_compressed(os.path.join(HERE, 'collective_%s.npz' % model), **gen_collective_fixture(R, model))
    for f in sorted(os.listdir(HERE)):
        if f.endswith('.npz'):
            print('%-28s %8d bytes' % (f, os.path.getsize(os.path.join(HERE, f))))


if __name__ == '__main__':
    main()
