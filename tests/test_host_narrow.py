"""CPU-only: the host side of transfer narrowing (resql_b200/csrc/host_narrow.h). A chunk of 8-byte values is
converted to 1 or 4 bytes only if EVERY value fits; the exact value range comes back with it."""
import ctypes as C

import numpy as np
import pytest

from resql_b200 import native as N


def convert(values, width):
    lib = N.load()
    a = np.ascontiguousarray(values, dtype=np.int64)
    out = np.zeros(max(len(a), 1) * width, dtype=np.uint8)
    lo, hi = C.c_int64(0), C.c_int64(0)
    rc = lib.rq_debug_convert_chunk(a.ctypes.data_as(C.POINTER(C.c_int64)), C.c_int64(len(a)), C.c_int32(width),
                                    out.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(lo), C.byref(hi))
    return rc, out[: len(a) * width].view(np.uint8 if width == 1 else np.int32), lo.value, hi.value


@pytest.mark.parametrize("n", [1, 7, 31, 32, 33, 1000, 65537])
def test_values_that_fit_are_converted_exactly(n):
    rng = np.random.default_rng(n)
    v = rng.integers(0, 256, n)
    rc, out, lo, hi = convert(v, 1)
    assert rc == 1 and np.array_equal(out, v.astype(np.uint8)) and (lo, hi) == (int(v.min()), int(v.max()))
    w = rng.integers(-2**31, 2**31, n)
    rc, out, lo, hi = convert(w, 4)
    assert rc == 1 and np.array_equal(out, w.astype(np.int32)) and (lo, hi) == (int(w.min()), int(w.max()))


@pytest.mark.parametrize("pos", [0, 1, 500, 998, 999])
@pytest.mark.parametrize("bad,width", [(256, 1), (-1, 1), (2**31, 4), (-2**31 - 1, 4), (2**40, 1), (-2**62, 4), (2**63 - 1, 4)])
def test_one_value_that_does_not_fit_is_noticed_wherever_it_sits(pos, bad, width):
    v = np.full(1000, 7, dtype=np.int64)
    v[pos] = bad
    rc, _, _, _ = convert(v, width)
    assert rc == 0


def test_boundary_values_fit():
    assert convert([0, 255], 1)[0] == 1
    assert convert([-2**31, 2**31 - 1], 4)[0] == 1


def convert32(values):
    lib = N.load()
    a = np.ascontiguousarray(values, dtype=np.int32)
    out = np.zeros(max(len(a), 1), dtype=np.uint8)
    lo, hi = C.c_int64(0), C.c_int64(0)
    rc = lib.rq_debug_convert_chunk(a.ctypes.data_as(C.POINTER(C.c_int64)), C.c_int64(len(a)), C.c_int32(-1),
                                    out.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(lo), C.byref(hi))
    return rc, out[: len(a)], lo.value, hi.value


@pytest.mark.parametrize("n", [1, 31, 33, 1000, 65537])
def test_int_columns_that_fit_a_byte(n):
    v = np.random.default_rng(n).integers(0, 256, n)
    rc, out, lo, hi = convert32(v)
    assert rc == 1 and np.array_equal(out, v.astype(np.uint8)) and (lo, hi) == (int(v.min()), int(v.max()))


@pytest.mark.parametrize("pos", [0, 501, 999])
@pytest.mark.parametrize("bad", [256, -1, 2**31 - 1, -2**31])
def test_int_value_that_does_not_fit_a_byte_is_noticed(pos, bad):
    v = np.full(1000, 3, dtype=np.int32)
    v[pos] = bad
    assert convert32(v)[0] == 0
