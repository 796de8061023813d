"""walt_stdsort.cuh (the device builder's replay of libstdc++ std::sort, which fixes the order of
tied suffixes inside a .dbindex bucket, reference.cpp:290-300) against the real std::sort of
this toolchain, on the CPU: same source, compiled for the host by tests/emu."""
import ctypes as C

import numpy as np
import pytest

from emu import emu


def _both(cls):
    L = emu.lib()
    cls = np.ascontiguousarray(cls, dtype=np.uint32)
    ours = np.zeros(cls.size, np.uint32)
    std = np.zeros(cls.size, np.uint32)
    fb = L.emu_stdsort_both(cls.ctypes.data_as(C.c_void_p), C.c_uint32(cls.size),
                            ours.ctypes.data_as(C.c_void_p), std.ctypes.data_as(C.c_void_p))
    return ours, std, fb


def _levels(cls):
    L = emu.lib()
    cls = np.ascontiguousarray(cls, dtype=np.uint32)
    out = np.zeros(cls.size, np.uint32)
    lv = C.c_uint32(0)
    L.emu_stdsort_levels(cls.ctypes.data_as(C.c_void_p), C.c_uint32(cls.size), out.ctypes.data_as(C.c_void_p), C.byref(lv))
    return out, int(lv.value)


def _killer(n):
    out = np.zeros(n, np.uint32)
    emu.lib().emu_antiqsort(C.c_uint32(n), out.ctypes.data_as(C.c_void_p))
    return out


@pytest.mark.parametrize("n", [0, 1, 2, 3, 15, 16, 17, 18, 31, 32, 33, 100, 257, 1000, 4097, 50000])
@pytest.mark.parametrize("n_classes", [1, 2, 3, 7, 50, 10 ** 9])
def test_random_with_ties(n, n_classes):
    rng = np.random.default_rng(n * 31 + n_classes % 1000)
    for _ in range(3):
        cls = rng.integers(0, n_classes, size=n, dtype=np.uint64).astype(np.uint32)
        ours, std, _ = _both(cls)
        assert np.array_equal(ours, std)
        assert np.all(np.diff(cls[ours].astype(np.int64)) >= 0)
        assert np.array_equal(_levels(cls)[0], std)     # the task-tree form the device runs


@pytest.mark.parametrize("n", [17, 64, 1000, 30000])
def test_structured_inputs(n):
    i = np.arange(n, dtype=np.uint32)
    cases = [i, i[::-1].copy(), i // 5, (i[::-1] // 3).copy(), np.minimum(i, n - 1 - i),   # organ pipe
             (i % 2), (i * 7919) % 13, np.where(i < n // 2, 1, 0)]
    for cls in cases:
        ours, std, _ = _both(cls)
        assert np.array_equal(ours, std)
        assert np.array_equal(_levels(cls)[0], std)


@pytest.mark.parametrize("n", [200, 3000, 40000])
@pytest.mark.parametrize("div", [1, 2, 4, 9])
def test_depth_limit_fallback(n, div):
    """The adversarial permutation drives the introsort into its heap-sort fallback; dividing
    the values creates ties inside it."""
    cls = _killer(n) // div
    ours, std, fb = _both(cls)
    assert np.array_equal(ours, std)
    by_levels, n_levels = _levels(cls)
    assert np.array_equal(by_levels, std)
    if div == 1:
        assert n_levels > 2 * int(np.log2(n)), "the depth limit was not reached in the task-tree form"
    if div == 1:
        assert fb > 0, "the adversary did not reach the depth limit: the fallback path is untested"


def test_heapsort_by_position_restores_bucket_order():
    rng = np.random.default_rng(5)
    for n in (0, 1, 2, 17, 1000, 33333):
        pos = rng.permutation(10 * n + 1)[:n].astype(np.uint32)
        cls = rng.integers(0, 9, size=n).astype(np.uint32)
        want = np.argsort(pos, kind="stable")
        p, c = pos.copy(), cls.copy()
        emu.lib().emu_heapsort_by_pos(p.ctypes.data_as(C.c_void_p), c.ctypes.data_as(C.c_void_p), C.c_uint32(n))
        assert np.array_equal(p, pos[want]) and np.array_equal(c, cls[want])
