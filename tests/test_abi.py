"""CPU: libgennbv_b200.so loads and exports exactly the entry points include/gennbv_b200.h declares
(no compute call is made: there is no GPU here)."""
import ctypes
import os
import re

import pytest

from gennbv_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "gennbv_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gnbv_[a-z0-9_]+)\s*\(", text)))


def test_library_is_built_and_loads():
    assert os.path.exists(_lib.LIB_PATH), "run `python -m gennbv_b200.build` (or __graft_entry__.build())"
    assert _lib.lib().gnbv_abi_version() == _lib.ABI_VERSION


def test_every_declared_symbol_is_exported_and_bound():
    names = header_functions()
    assert len(names) >= 6
    h = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(h, n), f"{n} declared in include/gennbv_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == names, "gennbv_b200/_lib.py bindings out of sync with the header"


def test_workspace_size_and_error_reporting():
    l = _lib.lib()
    assert l.gnbv_voxelize_workspace_bytes(256, 64) >= 2 * 256 * 64 ** 3 // 8
    assert l.gnbv_voxelize_workspace_bytes(0, 64) == 0
    # argument validation happens before any CUDA call: a null pointer is rejected with a message
    rc = l.gnbv_gae(None, None, None, None, None, 0.99, 0.95, 4, 4, None, None, None)
    assert rc == -1 and b"null" in l.gnbv_last_error()


def test_ops_refuse_cpu_tensors():
    import torch
    from gennbv_b200 import ops
    z = torch.zeros(2, 3)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        ops.gae(z, z, z.byte(), z[0], z[0].byte(), 0.99, 0.95, z.clone(), z.clone())
