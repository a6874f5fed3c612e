"""CPU: libu2pl_b200.so loads and exports every symbol include/u2pl_b200.h declares, and the ctypes
table of u2pl_b200/_lib.py covers exactly that set (no compute calls: there is no GPU here)."""
import ctypes
import os
import re

from u2pl_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "u2pl_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return set(re.findall(r"\b(u2pl_[a-z0-9_]+)\s*\(", text))


def test_header_symbols_exported_and_bound():
    path = build.build()
    lib = ctypes.CDLL(path)
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/u2pl_b200.h but not exported"
    assert names == set(_lib.SIGNATURES), (names ^ set(_lib.SIGNATURES))


def test_library_is_sm100a_only_and_has_no_cpu_path():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", build.build()], capture_output=True, text=True).stdout
    assert "sm_100a" in out and "sm_90" not in out and "sm_80" not in out
    lib = _lib.load(build_if_missing=False)
    assert lib.u2pl_abi_version() == 1 and lib.u2pl_launch_count() == 0


def test_ops_refuse_cpu_tensors():
    import pytest
    import torch
    from u2pl_b200 import ops
    with pytest.raises(_lib.U2PLNativeError):
        ops.entropy_thresholds(torch.zeros(1, 3, 4, 4), torch.zeros(1, 4, 4, dtype=torch.long), [50.0])
