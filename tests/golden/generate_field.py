"""Writes tests/golden/field_reference.npz with outputs of the REFERENCE's own navigation-field functions that can run here:
``direction_map`` (core/steering/quickest_path.py:144-163, numpy) and ``obstacle_handling``
(core/steering/obstacle_handling.py:15-74, numba) -- executed unmodified from /root/reference (their modules cannot be
imported as a whole: skfmm / shapely / skimage / loggingtools are missing; the two functions need none of them).
Run where /root/reference exists:  python tests/golden/generate_field.py"""
import ast
import os
import sys

import numba
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.environ.get('CROWD_REFERENCE', '/root/reference/crowddynamics')


def extract(path, name):
    """source of a top-level function WITHOUT its decorators (log_with / numba.jit signatures); the body is untouched"""
    src = open(path).read()
    for node in ast.parse(src).body:
        if getattr(node, 'name', None) == name:
            lines = src.splitlines()[node.lineno - 1:node.end_lineno]
            return '\n'.join(lines)
    raise KeyError(name)


def main():
    ns = {'np': np, 'DistanceMap': np.ndarray}       # the annotation alias of quickest_path.py:19
    exec(compile(extract(os.path.join(REF, 'core', 'steering', 'quickest_path.py'), 'direction_map'), 'reference:quickest_path.py', 'exec'), ns)
    exec(compile(extract(os.path.join(REF, 'core', 'steering', 'obstacle_handling.py'), 'obstacle_handling'), 'reference:obstacle_handling.py', 'exec'), ns)
    direction_map = ns['direction_map']
    obstacle_handling = numba.njit(ns['obstacle_handling'])      # the reference jit-compiles it in nopython mode too
    rng = np.random.default_rng(7)
    ny, nx = 23, 31
    Y, X = np.mgrid[0:ny, 0:nx].astype(np.float64)
    dmap = -np.hypot(X - 25.3, Y - 4.2) * 0.1 + 0.01 * rng.normal(size=(ny, nx))
    dmap[5, 5:9] = dmap[5, 4]                       # a flat spot: zero gradient along x
    mask = np.zeros((ny, nx), dtype=bool)
    mask[8:15, 10:12] = True; mask[0, 0] = True; mask[ny - 1, 7] = True; mask[3, nx - 1] = True
    masked = np.ma.MaskedArray(dmap.copy(), mask)
    u_m, v_m = direction_map(masked)
    u_p, v_p = direction_map(dmap.copy())
    dmap_obs = -np.abs(X - 10.5) * 0.1 + 0.003 * rng.normal(size=(ny, nx))
    dmap_obs[:, 10:12] = np.abs(rng.normal(size=(ny, 2))) * 0.05
    dir_obs = direction_map(dmap_obs.copy())
    out = {}
    for radius, strength, tag in ((0.5, 0.3, 'a'), (1.2, 0.7, 'b')):
        uo, vo = obstacle_handling(dmap_obs, (np.ascontiguousarray(dir_obs[0]), np.ascontiguousarray(dir_obs[1])),
                                   (np.ascontiguousarray(u_p), np.ascontiguousarray(v_p)), radius, strength)
        out['oh_u_' + tag], out['oh_v_' + tag] = uo, vo
        out['oh_par_' + tag] = np.array([radius, strength])
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'field_reference.npz'), dmap=dmap, mask=mask,
                        dm_u_masked=np.ma.getdata(u_m), dm_v_masked=np.ma.getdata(v_m), dm_mask=np.ma.getmaskarray(u_m) | np.ma.getmaskarray(v_m),
                        dm_u_plain=u_p, dm_v_plain=v_p, dmap_obs=dmap_obs, dir_obs_u=dir_obs[0], dir_obs_v=dir_obs[1], **out)
    print('written', {k: v.shape for k, v in out.items()})


if __name__ == '__main__':
    sys.exit(main())
