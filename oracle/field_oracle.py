"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's navigation-field construction
(simulation/field.py:155-164 -> core/steering/quickest_path.py:54-197, core/steering/obstacle_handling.py:15-74,106+).

What follows the reference's own code (numpy / numba, pinned by tests/golden/field_reference.npz, which
tests/golden/generate_field.py writes by executing the reference functions): ``meshgrid``, ``direction_map``,
``obstacle_handling``.  What replaces third-party packages that are not installed here -- **parity unpinned** against them:

* ``skfmm.distance`` (scikit-fmm 0.0.9)  -> first-order fast marching, oracle/field_oracle.c;
* ``shapely`` ``obstacles.buffer(radius)`` + ``skimage.draw.polygon`` -> grid points within ``radius`` of a segment;
* ``skimage.draw.line``                   -> Bresenham between the truncated indices of the end points;
* ``skimage.segmentation.find_boundaries(mode='outer')`` -> unmasked cells with a masked 4-neighbour;
* ``scipy.interpolate.NearestNDInterpolator`` (installed) -> brute-force nearest boundary cell, first in row-major order on
  ties (checked against scipy where the nearest cell is unique).
"""
import ctypes as C

import numpy as np

from . import crowd_oracle as _O


def _lib():
    L = _O.lib()
    if not getattr(L, '_field_ready', False):
        L.oracle_distance_map.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_double, C.c_void_p]
        L.oracle_eikonal_update.restype = C.c_double
        L.oracle_eikonal_update.argtypes = [C.c_double, C.c_double, C.c_double]
        L._field_ready = True
    return L


def grid_shape(step, minx, miny, maxx, maxy):
    """quickest_path.py:36-39: np.arange(min, max + step, step) per axis -> (ny, nx)"""
    return len(np.arange(miny, maxy + step, step=step)), len(np.arange(minx, maxx + step, step=step))


def indicer(points, minx, miny, step):
    """quickest_path.py:41-44"""
    return ((np.asarray(points, dtype=np.float64) - np.array((minx, miny))) / step).astype(np.int64)


def draw_line(r0, c0, r1, c1):
    """skimage.draw.line (Bresenham), as draw_geom uses it for LineStrings (core/geometry.py:112-116)."""
    r, c = int(r0), int(c0)
    dr, dc = abs(int(r1) - r), abs(int(c1) - c)
    sr = 1 if (int(r1) - r) > 0 else -1
    sc = 1 if (int(c1) - c) > 0 else -1
    steep = dr > dc
    if steep:
        c, r, dc, dr, sc, sr = r, c, dr, dc, sr, sc
    d = 2 * dr - dc
    rr, cc = [], []
    for _ in range(dc):
        if steep:
            rr.append(c); cc.append(r)
        else:
            rr.append(r); cc.append(c)
        while d >= 0:
            r += sr
            d -= 2 * dc
        c += sc
        d += 2 * dr
    rr.append(int(r1)); cc.append(int(c1))
    return np.array(rr), np.array(cc)


def raster_segments(segments, ny, nx, minx, miny, step):
    grid = np.zeros((ny, nx), dtype=np.uint8)
    for p0x, p0y, p1x, p1y in np.asarray(segments, dtype=np.float64).reshape(-1, 4):
        (r0, c0), (r1, c1) = indicer([(p0x, p0y), (p1x, p1y)], minx, miny, step)
        x, y = draw_line(r0, c0, r1, c1)
        ok = (x >= 0) & (x < nx) & (y >= 0) & (y < ny)
        grid[y[ok], x[ok]] = 1
    return grid


def buffer_mask(segments, radius, ny, nx, minx, miny, step):
    xs = minx + step * np.arange(nx, dtype=np.float64)
    ys = miny + step * np.arange(ny, dtype=np.float64)
    X, Y = np.meshgrid(xs, ys, indexing='xy')
    mask = np.zeros((ny, nx), dtype=bool)
    for ax, ay, bx, by in np.asarray(segments, dtype=np.float64).reshape(-1, 4):
        ex, ey = bx - ax, by - ay
        l2 = ex * ex + ey * ey
        t = ((X - ax) * ex + (Y - ay) * ey) / l2 if l2 > 0 else np.zeros_like(X)
        t = np.clip(t, 0.0, 1.0)
        qx, qy = X - (ax + t * ex), Y - (ay + t * ey)
        mask |= qx * qx + qy * qy <= radius * radius
    return mask.astype(np.uint8)


def distance_map(raster, mask, step):
    """quickest_path.py:54-117 with skfmm.distance replaced by oracle/field_oracle.c; returns an ndarray, NaN where masked."""
    raster = np.ascontiguousarray(raster, dtype=np.uint8)
    ny, nx = raster.shape
    out = np.empty((ny, nx), dtype=np.float64)
    m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
    rc = _lib().oracle_distance_map(raster.ctypes.data, None if m is None else m.ctypes.data, ny, nx, float(step), out.ctypes.data)
    assert rc == 0
    return out


def direction_map(dmap):
    """quickest_path.py:144-163, verbatim semantics: dmap may be a masked array."""
    u, v = np.gradient(dmap)
    l = np.hypot(u, v)
    l[l == 0] = np.nan
    return v / l, u / l


def find_boundaries_outer(mask):
    """skimage.segmentation.find_boundaries(mask, mode='outer') for a boolean image: background (False) cells with a
    foreground 4-neighbour."""
    m = np.asarray(mask, dtype=bool)
    nb = np.zeros_like(m)
    nb[1:, :] |= m[:-1, :]; nb[:-1, :] |= m[1:, :]; nb[:, 1:] |= m[:, :-1]; nb[:, :-1] |= m[:, 1:]
    return nb & ~m


def fill_missing(fill, dirmask, U, V):
    """quickest_path.py:168-181 on plain arrays: cells in ``fill`` take the value of the nearest boundary cell."""
    by, bx = np.nonzero(find_boundaries_outer(dirmask))
    U, V = U.copy(), V.copy()
    if len(by) == 0:
        return U, V
    order = np.lexsort((bx, by))                 # row-major: first on ties
    by, bx = by[order], bx[order]
    for y, x in zip(*np.nonzero(fill)):
        d = (by - y) ** 2 + (bx - x) ** 2
        k = int(np.argmin(d))                    # first minimum
        U[y, x], V[y, x] = U[by[k], bx[k]], V[by[k], bx[k]]
    return U, V


def obstacle_handling(dmap_obs, dir_map_obs, dir_map_targets, radius, strength):
    """obstacle_handling.py:15-74 (vectorised; same operations per cell)."""
    u1, v1 = dir_map_obs
    u2, v2 = dir_map_targets
    u_out, v_out = np.array(u2, dtype=np.float64), np.array(v2, dtype=np.float64)
    x = -np.asarray(dmap_obs)
    near = (0 < x) & (x < radius)
    with np.errstate(invalid='ignore', divide='ignore'):
        p = np.power(strength, x[near] / radius)
        u_out[near] = -p * u1[near] + (1 - p) * u2[near]
        v_out[near] = -p * v1[near] + (1 - p) * v2[near]
        l = np.hypot(u_out, v_out)
        return u_out / l, v_out / l


def navigation_to_target(target_segments, obstacle_segments, bounds, step, radius, strength):
    """field.py:155-164 for line-segment geometry -> (dmap_targets, (U, V)); dmap NaN inside the buffered obstacles."""
    minx, miny, maxx, maxy = bounds
    ny, nx = grid_shape(step, *bounds)
    target = raster_segments(target_segments, ny, nx, minx, miny, step)
    walls = obstacle_segments is not None and len(np.asarray(obstacle_segments).reshape(-1, 4))
    mask = buffer_mask(obstacle_segments, radius, ny, nx, minx, miny, step) if walls else None
    dmap = distance_map(target, mask, step)
    md = np.ma.MaskedArray(dmap, mask.astype(bool)) if walls else dmap
    U, V = direction_map(md)
    if not walls:
        return dmap, (np.asarray(U), np.asarray(V))
    dirmask = np.ma.getmaskarray(U)
    Ud, Vd = np.where(dirmask, np.nan, U.data), np.where(dirmask, np.nan, V.data)
    obst = raster_segments(obstacle_segments, ny, nx, minx, miny, step).astype(bool)
    # the reference fills logical_xor(obst, dirmask) and leaves the cells ON the obstacle lines masked (undefined data);
    # they are filled here as well -- the one deliberate difference, see csrc/field_kernels.cuh k_fill_missing
    Ud, Vd = fill_missing(dirmask, dirmask, Ud, Vd)
    dmap_obs = distance_map(obst.astype(np.uint8), None, step)
    dir_obs = direction_map(dmap_obs)
    return dmap, obstacle_handling(dmap_obs, dir_obs, (Ud, Vd), radius, strength)
