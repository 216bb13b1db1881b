"""The drop-in boundary is the C ABI of include/resql_b200.h: the shared library must load without a
GPU and export every function the header declares (no compute call is made here)."""
import os
import re

from common import ROOT
from resql_b200 import native as N


def _declared():
    src = open(os.path.join(ROOT, "include", "resql_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rq_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_entry_point():
    names = _declared()
    assert len(names) >= 14, names
    lib = N.load()
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in include/resql_b200.h but not exported: {missing}"
    # the binding's own list covers the header
    assert set(N.ABI_SYMBOLS) == set(names), (sorted(set(names) ^ set(N.ABI_SYMBOLS)))


def test_calls_fail_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        return
    lib = N.load()
    assert lib.rq_init(0) != 0                      # no CPU fallback: the engine refuses to start
    assert b"" != lib.rq_last_error()
