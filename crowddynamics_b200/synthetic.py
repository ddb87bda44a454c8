"""Synthetic crowds, wall geometry and navigation fields for tests and benchmarks.

Everything is seeded (``np.random.default_rng(seed)``) and built directly as arrays -- the reference's own population
setup (simulation/agents.py:633-680) needs traitlets/configobj/shapely and an unseeded RNG.  Value ranges follow the
'adult' body type (conf/body_types.cfg:1-10, truncated normal at 3 sigma, agents.py:131-134) and the default
constants of TranslationalMotion / RotationalMotion (agents.py:223-252,283-287).
"""
import numpy as np

from .structures import agent_type_circular, agent_type_three_circle, obstacle_type_linear


def _truncnorm(rng, mean, scale, size):
    """Truncated normal on [mean - scale, mean + scale] with sigma = scale / 3 (reference core/rand.py:28)."""
    out = rng.normal(mean, scale / 3.0, size)
    bad = np.abs(out - mean) > scale
    while bad.any():
        out[bad] = rng.normal(mean, scale / 3.0, int(bad.sum()))
        bad = np.abs(out - mean) > scale
    return out


def fill_adult_bodies(agents, rng, omega0=4 * np.pi):
    n = len(agents)
    r = _truncnorm(rng, 0.255, 0.035, n)
    agents['radius'] = r
    agents['r_t'] = 0.5882 * r
    agents['r_s'] = 0.3725 * r
    agents['r_ts'] = 0.6275 * r
    agents['mass'] = _truncnorm(rng, 73.5, 8.0, n)
    agents['target_velocity'] = _truncnorm(rng, 1.25, 0.3, n)
    agents['inertia_rot'] = 4.0 * np.pi * (agents['mass'] / 80.0) * (r / 0.27) ** 2
    agents['target_angular_velocity'] = omega0
    agents['active'] = True
    agents['target'] = 0
    agents['index_leader'] = -1
    agents['familiar_exit'] = -1
    agents['tau_adj'] = 0.5
    agents['k_soc'] = 1.5
    agents['tau_0'] = 3.0
    agents['mu'] = 1.2e5
    agents['kappa'] = 4e4
    agents['damping'] = 500.0
    agents['std_rand_force'] = 0.1
    if 'tau_rot' in agents.dtype.names:
        agents['tau_rot'] = 0.2
        agents['std_rand_torque'] = 0.1


def set_shoulders(agents):
    """position_ls/rs from position + orientation (reference simulation/agents.py:473-486), host-side for setup."""
    phi = agents['orientation']
    off = np.stack((np.sin(phi), -np.cos(phi)), axis=1) * agents['r_ts'][:, None]
    agents['position_ls'] = agents['position'] - off
    agents['position_rs'] = agents['position'] + off


def uniform_crowd(n, model='circular', density=1.0, seed=0, origin=(0.0, 0.0), jitter=None, overlap_fraction=0.0):
    """``n`` agents on a jittered square lattice of pitch 1/sqrt(density) inside a square room.

    Returns (agents, obstacles, side).  The four walls are the room boundary.  ``overlap_fraction`` > 0 moves that
    fraction of agents onto a neighbour (within 0.3 m) so that the contact-force branch (h < 0) is exercised.
    """
    rng = np.random.default_rng(seed)
    dtype = agent_type_circular if model == 'circular' else agent_type_three_circle
    agents = np.zeros(n, dtype=dtype)
    fill_adult_bodies(agents, rng)
    pitch = 1.0 / np.sqrt(density)
    m = int(np.ceil(np.sqrt(n)))
    side = m * pitch
    if jitter is None:
        jitter = max(0.0, 0.5 * (pitch - 0.6))
    k = np.arange(n)
    gx, gy = k // m, k % m
    pos = np.stack(((gx + 0.5) * pitch, (gy + 0.5) * pitch), axis=1)
    pos += rng.uniform(-jitter, jitter, size=(n, 2))
    pos += np.asarray(origin, dtype=np.float64)
    if overlap_fraction > 0 and n > 1:
        movers = rng.choice(n, size=max(1, int(overlap_fraction * n)), replace=False)
        mates = (movers + 1) % n
        ang = rng.uniform(-np.pi, np.pi, len(movers))
        dist = rng.uniform(0.05, 0.45, len(movers))
        pos[movers] = pos[mates] + np.stack((np.cos(ang), np.sin(ang)), axis=1) * dist[:, None]
    agents['position'] = pos
    phi = rng.uniform(-np.pi, np.pi, n)
    speed = rng.uniform(0.0, 1.3, n)
    e = np.stack((np.cos(phi), np.sin(phi)), axis=1)
    agents['velocity'] = e * speed[:, None]
    agents['target_direction'] = e
    if model != 'circular':
        agents['orientation'] = phi
        agents['target_orientation'] = phi
        agents['angular_velocity'] = rng.uniform(-1.0, 1.0, n)
        set_shoulders(agents)
    x0, y0 = origin
    obstacles = walls_of_box(x0, y0, x0 + side, y0 + side)
    return agents, obstacles, side


def uniform_slab(n_total, model='circular', density=1.0, seed=0, x_lo=-np.inf, x_hi=np.inf):
    """The agents of ``uniform_crowd(n_total, ...)``'s lattice whose lattice column centre lies in [x_lo, x_hi) -- what one
    rank of a strong-scaling strip run generates for itself (same lattice and body distributions; its own random stream).
    Returns (agents, ids, side): ids = position of each agent in the full lattice (global agent index)."""
    dtype = agent_type_circular if model == 'circular' else agent_type_three_circle
    pitch = 1.0 / np.sqrt(density)
    m = int(np.ceil(np.sqrt(n_total)))
    side = m * pitch
    centres = (np.arange(m) + 0.5) * pitch
    cols = np.nonzero((centres >= x_lo) & (centres < x_hi))[0]
    ids = (cols[:, None] * m + np.arange(m)[None, :]).reshape(-1)
    ids = ids[ids < n_total]
    n = len(ids)
    rng = np.random.default_rng([seed, int(cols[0]) if len(cols) else 0])
    agents = np.zeros(n, dtype=dtype)
    fill_adult_bodies(agents, rng)
    jitter = max(0.0, 0.5 * (pitch - 0.6))
    gx, gy = ids // m, ids % m
    pos = np.stack(((gx + 0.5) * pitch, (gy + 0.5) * pitch), axis=1) + rng.uniform(-jitter, jitter, size=(n, 2))
    agents['position'] = pos
    phi = rng.uniform(-np.pi, np.pi, n)
    e = np.stack((np.cos(phi), np.sin(phi)), axis=1)
    agents['velocity'] = e * rng.uniform(0.0, 1.3, n)[:, None]
    agents['target_direction'] = e
    if model != 'circular':
        agents['orientation'] = phi
        agents['target_orientation'] = phi
        agents['angular_velocity'] = rng.uniform(-1.0, 1.0, n)
        set_shoulders(agents)
    return agents, ids.astype(np.int64), side


def room_exit_walls(side, door_width=1.2):
    """The 11 wall segments of ``room_with_exit`` for a room of the given side."""
    y_lo, y_hi = side / 2 - door_width / 2, side / 2 + door_width / 2
    t = 0.3
    segs = [((side, 0.0), (0.0, 0.0)), ((0.0, 0.0), (0.0, side)), ((0.0, side), (side, side))]

    def rect(x0, y0, x1, y1):
        c = [(x0, y0), (x1, y0), (x1, y1), (x0, y1)]
        return [(c[k], c[(k + 1) % 4]) for k in range(4)]
    segs += rect(side, 0.0, side + t, y_lo) + rect(side, y_hi, side + t, side)
    obstacles = np.zeros(len(segs), dtype=obstacle_type_linear)
    for k, (p0, p1) in enumerate(segs):
        obstacles[k]['p0'], obstacles[k]['p1'] = p0, p1
    return obstacles


def random_crowd(n, model='circular', half_width=None, seed=0):
    """The reference benchmark's own workload (core/tests/test_interactions_benchmark.py:10-33): positions uniform in
    [-sqrt(2N), sqrt(2N)]^2 (0.125 agents/m^2), overlaps allowed."""
    rng = np.random.default_rng(seed)
    dtype = agent_type_circular if model == 'circular' else agent_type_three_circle
    agents = np.zeros(n, dtype=dtype)
    fill_adult_bodies(agents, rng)
    if half_width is None:
        half_width = np.sqrt(2.0 * max(n, 1))
    agents['position'] = rng.uniform(-half_width, half_width, size=(n, 2))
    phi = rng.uniform(-np.pi, np.pi, n)
    e = np.stack((np.cos(phi), np.sin(phi)), axis=1)
    agents['velocity'] = e * rng.uniform(0.0, 1.3, n)[:, None]
    agents['target_direction'] = e
    if model != 'circular':
        agents['orientation'] = rng.uniform(-np.pi, np.pi, n)
        agents['target_orientation'] = phi
        agents['angular_velocity'] = rng.uniform(-1.0, 1.0, n)
        set_shoulders(agents)
    obstacles = walls_of_box(-half_width, -half_width, half_width, half_width)
    return agents, obstacles, 2 * half_width


def walls_of_box(x0, y0, x1, y1):
    obs = np.zeros(4, dtype=obstacle_type_linear)
    corners = [(x0, y0), (x1, y0), (x1, y1), (x0, y1)]
    for w in range(4):
        obs[w]['p0'] = corners[w]
        obs[w]['p1'] = corners[(w + 1) % 4]
    return obs


class MeshGrid:
    """Minimal stand-in for reference quickest_path.MeshGrid (quickest_path.py:14-51): shape, step, bounds, indicer."""

    def __init__(self, step, minx, miny, maxx, maxy):
        x = np.arange(minx, maxx + step, step=step)
        y = np.arange(miny, maxy + step, step=step)
        self.shape = (len(y), len(x))
        self.step = step
        self.bounds = (minx, miny, maxx, maxy)

    def indicer(self, position):
        shifted = np.asarray(position) - np.array(self.bounds[:2])
        return (shifted / self.step).astype(np.int64)


def direction_field(step, bounds, kind='exit', point=(0.0, 0.0), seed=0):
    """Synthetic static navigation field with the reference layout ``(mgrid, (U, V))``, U/V of shape (ny, nx) indexed
    [iy, ix] (logic.py:159-164).  kind: 'exit' -> unit vectors towards ``point``; 'x+' / 'x-' -> constant;
    'swirl' -> smooth position-dependent unit field (exercises the gather)."""
    mg = MeshGrid(step, *bounds)
    ny, nx = mg.shape
    xs = bounds[0] + step * np.arange(nx)
    ys = bounds[1] + step * np.arange(ny)
    X, Y = np.meshgrid(xs, ys, indexing='xy')
    if kind == 'exit':
        dx, dy = point[0] - X, point[1] - Y
        nrm = np.hypot(dx, dy)
        nrm[nrm == 0] = 1.0
        U, V = dx / nrm, dy / nrm
    elif kind == 'x+':
        U, V = np.ones_like(X), np.zeros_like(X)
    elif kind == 'x-':
        U, V = -np.ones_like(X), np.zeros_like(X)
    elif kind == 'swirl':
        ang = 0.37 * X - 0.23 * Y + 0.05 * X * Y / (1.0 + np.abs(X))
        U, V = np.cos(ang), np.sin(ang)
    else:
        raise ValueError(kind)
    return mg, (np.ascontiguousarray(U), np.ascontiguousarray(V))


def hallway(seed=0, model='circular', size=50, width=40.0, height=5.0, ratio=1.0 / 3.0, step=0.1):
    """BASELINE config 1 rebuilt as arrays (reference examples/simulations.py:74-163, examples/fields.py:20-76):
    two groups walking towards each other in a ``width`` x ``height`` corridor with two wall segments and two targets.
    The FMM navigation field (skfmm, absent) is replaced by constant +x / -x fields -- an approximation stated in
    DESIGN.md; oracle-vs-GPU parity is unaffected because both consume the same arrays."""
    rng = np.random.default_rng(seed)
    dtype = agent_type_circular if model == 'circular' else agent_type_three_circle
    agents = np.zeros(size, dtype=dtype)
    fill_adult_bodies(agents, rng)
    half = size // 2
    pitch = 0.75
    spawn_w = width * ratio

    def place(count, x_lo):
        cols = int(spawn_w // pitch)
        rows = int(height // pitch)
        cells = rng.permutation(cols * rows)[:count]
        cx, cy = cells // rows, cells % rows
        pos = np.stack((x_lo + (cx + 0.5) * pitch, (cy + 0.5) * pitch), axis=1)
        return pos + rng.uniform(-0.05, 0.05, size=pos.shape)
    agents['position'][:half] = place(half, 0.0)
    agents['position'][half:] = place(size - half, width - spawn_w)
    agents['target'][:half] = 1     # walk to the right end
    agents['target'][half:] = 0     # walk to the left end
    agents['target_direction'][:half] = (1.0, 0.0)
    agents['target_direction'][half:] = (-1.0, 0.0)
    if model != 'circular':
        agents['orientation'][:half] = 0.0
        agents['orientation'][half:] = np.pi
        agents['target_orientation'] = agents['orientation']
        set_shoulders(agents)
    obstacles = np.zeros(2, dtype=obstacle_type_linear)
    obstacles[0]['p0'], obstacles[0]['p1'] = (0.0, 0.0), (width, 0.0)
    obstacles[1]['p0'], obstacles[1]['p1'] = (0.0, height), (width, height)
    bounds = (0.0, 0.0, width, height)
    fields = [direction_field(step, bounds, 'x-'), direction_field(step, bounds, 'x+')]
    return agents, obstacles, fields


def room_with_exit(n, model='circular', density=1.0, seed=0, door_width=1.2, hall_length=5.0, step=0.5):
    """BASELINE config 4 flavour (reference examples/simulations.py:166-236, examples/fields.py:175-210): a square room
    whose right wall has a door gap in the middle leading into an exit hall; 11 wall segments (3 room walls as one open
    polyline + two 4-edge rectangles flanking the door, as geom_to_linear_obstacles would emit them); every agent has
    target 0; the direction field points at the door (synthetic stand-in for the FMM field, same (U, V)[iy, ix] layout)."""
    agents, _, side = uniform_crowd(n, model, density=density, seed=seed)
    y_lo, y_hi = side / 2 - door_width / 2, side / 2 + door_width / 2
    t = 0.3                                            # wall thickness of the two door posts
    segs = [((side, 0.0), (0.0, 0.0)), ((0.0, 0.0), (0.0, side)), ((0.0, side), (side, side))]

    def rect(x0, y0, x1, y1):
        c = [(x0, y0), (x1, y0), (x1, y1), (x0, y1)]
        return [(c[k], c[(k + 1) % 4]) for k in range(4)]
    segs += rect(side, 0.0, side + t, y_lo) + rect(side, y_hi, side + t, side)
    obstacles = np.zeros(len(segs), dtype=obstacle_type_linear)
    for k, (p0, p1) in enumerate(segs):
        obstacles[k]['p0'], obstacles[k]['p1'] = p0, p1
    agents['target'] = 0
    bounds = (0.0, 0.0, side + hall_length, side)
    fields = [direction_field(step, bounds, 'exit', point=(side + 0.5 * t, side / 2))]
    e = np.stack((side - agents['position'][:, 0], side / 2 - agents['position'][:, 1]), axis=1)
    e /= np.hypot(*e.T)[:, None]
    agents['target_direction'] = e
    if model != 'circular':
        agents['target_orientation'] = np.arctan2(e[:, 1], e[:, 0])
    return agents, obstacles, fields, side


def leader_follower_crowd(n, model='circular', density=0.5, seed=0, n_leaders=None, n_doors=2, inner_walls=6):
    """Stand-in for the crowds of the reference's examples/collective_motion.py: a walled square room with two door gaps
    (left / right wall) and a few free-standing inner walls that block lines of sight; a small group of leaders
    (``is_leader``, target = one of the exits) among herding followers (``is_follower``, no target, a random familiar
    exit, a quarter of them already following some leader).  Returns (agents, obstacles, center_door, side)."""
    agents, _, side = uniform_crowd(n, model, density=density, seed=seed)
    rng = np.random.default_rng(seed + 1000)
    n_leaders = max(1, n // 100) if n_leaders is None else n_leaders
    leaders = rng.choice(n, size=min(n_leaders, n), replace=False)
    agents['is_follower'] = True
    agents['is_leader'][leaders] = True
    agents['is_follower'][leaders] = False
    agents['target'] = -1
    agents['target'][leaders] = rng.integers(0, n_doors, size=len(leaders))
    agents['familiar_exit'] = rng.integers(0, n_doors, size=n)
    agents['index_leader'] = -1
    followers = np.flatnonzero(agents['is_follower'])
    if len(followers):
        some = rng.choice(followers, size=len(followers) // 4, replace=False)
        agents['index_leader'][some] = rng.choice(leaders, size=len(some))
    door = 1.5
    y_lo, y_hi = side / 2 - door / 2, side / 2 + door / 2
    segs = [((0.0, 0.0), (side, 0.0)), ((0.0, side), (side, side)),
            ((0.0, 0.0), (0.0, y_lo)), ((0.0, y_hi), (0.0, side)), ((side, 0.0), (side, y_lo)), ((side, y_hi), (side, side))]
    for _ in range(inner_walls):
        p0 = rng.uniform(0.1 * side, 0.9 * side, 2)
        segs.append((tuple(p0), tuple(p0 + rng.uniform(-0.2 * side, 0.2 * side, 2))))
    obstacles = np.zeros(len(segs), dtype=obstacle_type_linear)
    for k, (p0, p1) in enumerate(segs):
        obstacles[k]['p0'], obstacles[k]['p1'] = p0, p1
    center_door = np.array([(0.0, side / 2), (side, side / 2)] + [tuple(rng.uniform(0, side, 2)) for _ in range(n_doors - 2)])
    return agents, obstacles, center_door[:n_doors], side
