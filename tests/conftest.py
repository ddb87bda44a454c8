import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')
    # a fresh checkout has no built artefacts (*.so is git-ignored): build them once (nvcc cross-compiles without a GPU)
    from crowddynamics_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()


def _have_gpu():
    try:
        import ctypes as C
        from crowddynamics_b200 import _lib
        n = C.c_int(0)
        return _lib.load().cdb_device_count(C.byref(n)) == 0 and n.value > 0
    except Exception:
        return False


HAVE_GPU = None


def pytest_collection_modifyitems(config, items):
    global HAVE_GPU
    if not any('gpu' in item.keywords for item in items):
        return
    if HAVE_GPU is None:
        HAVE_GPU = _have_gpu()
    if HAVE_GPU:
        return
    skip = pytest.mark.skip(reason='no CUDA device in this container')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name)))


def from_raw(raw, dtype):
    """(n, itemsize) uint8 rows -> structured array (copy)."""
    return np.ascontiguousarray(raw).view(dtype).reshape(-1).copy()


def rel_err_fields(a, b, fields=None):
    """max over fields of max |a-b| / max(|b|, tiny); returns (worst, field)."""
    worst, wf = 0.0, None
    for f in fields or a.dtype.names:
        x, y = a[f].astype(np.float64), b[f].astype(np.float64)
        if x.size == 0:
            continue
        same = (x == y) | (np.isnan(x) & np.isnan(y))
        if same.all():
            continue
        err = np.max(np.where(same, 0.0, np.abs(x - y) / np.maximum(np.abs(y), 1e-300)))
        if err > worst:
            worst, wf = float(err), f
    return worst, wf


def vec_rel_err(x, y, floor=1e-12):
    """Per-agent vector error |x - y| / (|y| + floor) for (n, 2) or (n,) arrays -> max."""
    x, y = np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64)
    if x.size == 0:
        return 0.0
    if x.ndim == 1:
        num, den = np.abs(x - y), np.abs(y)
    else:
        num, den = np.hypot(*(x - y).T), np.hypot(*y.T)
    return float(np.max(num / (den + floor)))
