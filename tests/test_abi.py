"""The C-ABI library loads and exports every symbol include/crowd_b200.h declares (no compute calls: CPU only)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from crowddynamics_b200 import _lib
from crowddynamics_b200.exceptions import CrowdDynamicsException, InvalidType, ExtensionMissing


def _declared_symbols():
    with open(_lib.HEADER_PATH) as f:
        src = f.read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(cdb_[a-z0-9_]+)\s*\(', src)))


def test_header_symbols_are_exported():
    L = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 35
    for name in names:
        assert hasattr(L, name), 'libcrowd_b200.so does not export %s' % name


def test_ctypes_signatures_cover_header():
    L = _lib.load()
    assert sorted(L._signatures) == _declared_symbols()


def test_version_and_error_channel():
    L = _lib.load()
    assert L.cdb_version() >= 100
    assert isinstance(L.cdb_last_error(), bytes)
    with pytest.raises(InvalidType):
        _lib.check(_lib.CDB_ERR_INVALID_TYPE)
    with pytest.raises(CrowdDynamicsException):
        _lib.check(_lib.CDB_ERR_CAPACITY)
    assert issubclass(InvalidType, TypeError) and issubclass(InvalidType, CrowdDynamicsException)


def test_null_handles_are_rejected_without_a_device():
    L = _lib.load()
    assert L.cdb_reset(None) == _lib.CDB_ERR_INVALID_VALUE
    assert b'NULL' in L.cdb_last_error()
    assert L.cdb_num_agents(None) == -1


def test_missing_extension_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, '_lib', None)
    monkeypatch.setattr(_lib, 'LIB_PATH', str(tmp_path / 'nope.so'))
    with pytest.raises(ExtensionMissing):
        _lib.load()


def test_product_does_not_import_oracle():
    """The product package must never route through the CPU oracle."""
    root = os.path.dirname(_lib.__file__)
    for dirpath, _, files in os.walk(root):
        for fn in files:
            if fn.endswith(('.py', '.cu', '.cuh', '.h')):
                with open(os.path.join(dirpath, fn)) as f:
                    txt = f.read()
                assert 'import oracle' not in txt and 'from oracle' not in txt and 'liboracle' not in txt, fn
