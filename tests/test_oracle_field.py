"""Navigation-field construction, CPU side: the oracle (oracle/field_oracle.{c,py}) against the
reference's own functions where they can run (golden vectors of direction_map / obstacle_handling, written by
tests/golden/generate_field.py), against closed-form distances, and against scipy for the nearest-neighbour fill.
skfmm / shapely / skimage are absent: the eikonal solver, the obstacle buffer and the rasterisation are "parity unpinned"."""
import os

import numpy as np
import pytest

from oracle import field_oracle as F

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'field_reference.npz')


def test_direction_map_and_obstacle_handling_reproduce_the_reference():
    g = np.load(GOLD)
    u, v = F.direction_map(np.ma.MaskedArray(g['dmap'].copy(), g['mask']))
    m = np.ma.getmaskarray(u) | np.ma.getmaskarray(v)
    assert (m == g['dm_mask']).all()
    assert np.array_equal(np.ma.getdata(u)[~m], g['dm_u_masked'][~m], equal_nan=True)
    assert np.array_equal(np.ma.getdata(v)[~m], g['dm_v_masked'][~m], equal_nan=True)
    up, vp = F.direction_map(g['dmap'].copy())
    assert np.array_equal(up, g['dm_u_plain'], equal_nan=True) and np.array_equal(vp, g['dm_v_plain'], equal_nan=True)
    for tag in 'ab':
        radius, strength = g['oh_par_' + tag]
        uo, vo = F.obstacle_handling(g['dmap_obs'], (g['dir_obs_u'], g['dir_obs_v']), (up, vp), radius, strength)
        np.testing.assert_allclose(uo, g['oh_u_' + tag], rtol=0, atol=2e-15, equal_nan=True)
        np.testing.assert_allclose(vo, g['oh_v_' + tag], rtol=0, atol=2e-15, equal_nan=True)


@pytest.mark.parametrize('step', [0.2, 0.1, 0.05])
def test_distance_map_converges_to_closed_form_distances(step):
    bounds = (0.0, 0.0, 12.0, 9.0)
    ny, nx = F.grid_shape(step, *bounds)
    xs, ys = step * np.arange(nx), step * np.arange(ny)
    X, Y = np.meshgrid(xs, ys)
    # a point target and a line target in free space: |distance_map| -> Euclidean distance with first-order error O(h)
    point = F.raster_segments([(6.0, 4.0, 6.0, 4.0)], ny, nx, 0.0, 0.0, step)
    d = F.distance_map(point, None, step)
    px, py = np.argwhere(point)[0][1] * step, np.argwhere(point)[0][0] * step
    exact = np.hypot(X - px, Y - py)
    far = exact > 3 * step
    assert (d[far] < 0).all() and d[point.astype(bool)].min() > 0
    assert np.abs(-d - exact)[far].max() <= 1.5 * step
    line = F.raster_segments([(12.0, 3.0, 12.0, 6.0)], ny, nx, 0.0, 0.0, step)
    lx = np.argwhere(line)[:, 1].max() * step
    ly0, ly1 = np.argwhere(line)[:, 0].min() * step, np.argwhere(line)[:, 0].max() * step
    d = F.distance_map(line, None, step)
    exact = np.hypot(X - lx, Y - np.clip(Y, ly0, ly1))
    assert np.abs(-d - exact)[exact > 3 * step].max() <= 1.5 * step
    # the discrete solution is the fixed point of the upwind update
    L = F._lib()
    T = np.abs(d)
    y, x = ny // 3, nx // 4
    a, b = min(T[y, x - 1], T[y, x + 1]), min(T[y - 1, x], T[y + 1, x])
    assert T[y, x] == L.oracle_eikonal_update(a, b, step)


def test_error_decreases_with_the_step():
    errs = []
    for step in (0.2, 0.1, 0.05):
        ny, nx = F.grid_shape(step, 0.0, 0.0, 10.0, 10.0)
        X, Y = np.meshgrid(step * np.arange(nx), step * np.arange(ny))
        r = F.raster_segments([(5.0, 5.0, 5.0, 5.0)], ny, nx, 0.0, 0.0, step)
        py, px = np.argwhere(r)[0] * step
        d = F.distance_map(r, None, step)
        errs.append(np.abs(-d - np.hypot(X - px, Y - py)).max())
    assert errs[0] > errs[1] > errs[2] and errs[2] < 0.6 * errs[0]


def test_walls_are_impassable_and_the_field_points_around_them():
    bounds, step, radius = (0.0, 0.0, 10.0, 8.0), 0.1, 0.5
    walls = [(5.0, 0.0, 5.0, 5.0)]                     # a wall from the bottom edge: the way to the right leads over it
    target = [(10.0, 1.0, 10.0, 2.0)]
    dmap, (U, V) = F.navigation_to_target(target, walls, bounds, step, radius, 0.3)
    ny, nx = dmap.shape
    mask = F.buffer_mask(walls, radius, ny, nx, 0.0, 0.0, step).astype(bool)
    assert np.isnan(dmap[mask]).all() and np.isfinite(dmap[~mask]).all()
    # left of the wall, low: geodesic distance is longer than the straight line through the wall
    iy, ix = int(1.0 / step), int(2.0 / step)
    assert -dmap[iy, ix] > np.hypot(10.0 - 2.0, 0.5) + 2.0
    assert V[iy, ix] > 0.3                              # heading up, around the wall's end
    ok = np.isfinite(U)
    assert np.abs(np.hypot(U[ok], V[ok]) - 1.0).max() < 1e-12
    assert ok.all()                                     # the buffer zone (wall lines included) was filled
    # close to the wall the field leans away from it (obstacle_handling)
    assert U[int(2.0 / step), int(4.7 / step)] < U[int(2.0 / step), int(3.0 / step)]


def test_nearest_fill_matches_scipy_where_the_nearest_cell_is_unique():
    from scipy.interpolate import NearestNDInterpolator
    rng = np.random.default_rng(3)
    ny, nx = 30, 40
    dirmask = np.zeros((ny, nx), dtype=bool)
    dirmask[10:20, 12:30] = True
    U, V = rng.normal(size=(ny, nx)), rng.normal(size=(ny, nx))
    U[dirmask], V[dirmask] = np.nan, np.nan
    fill = dirmask.copy()
    fu, fv = F.fill_missing(fill, dirmask, U, V)
    b = F.find_boundaries_outer(dirmask)
    Y, X = np.mgrid[0:ny, 0:nx]
    pts = np.stack((Y[b], X[b])).T
    ip = NearestNDInterpolator(pts, U[b], rescale=False)
    by, bx = Y[b], X[b]
    checked = 0
    for y, x in zip(*np.nonzero(fill)):
        d = np.sort((by - y) ** 2 + (bx - x) ** 2)
        if d[0] < d[1]:
            assert fu[y, x] == ip((y, x))
            checked += 1
    assert checked > 50 and np.isfinite(fu).all() and np.isfinite(fv).all()
