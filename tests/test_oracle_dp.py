"""CPU: the oracle's DP restatements against the real reference functions (ref_shim.c).
Pins kswx_extend_align_core / _shift_core (kswx.h:101-335), ksw_global2 (ksw.c:503), hz_align_hzmo
(hzm_aln.h:278) and sort_array (sort.h:104) bit-exactly on seeded random inputs + edge cases."""
import ctypes as C

import numpy as np
import pytest

from conftest import call_ext, call_global, mutate


def _cases(rng, n):
    for _ in range(n):
        ln = int(rng.integers(1, 400))
        a = rng.integers(0, 4, ln).astype(np.uint8)
        kind = rng.integers(0, 4)
        if kind == 0:
            b = mutate(rng, a)
        elif kind == 1:
            b = rng.integers(0, 4, int(rng.integers(1, 400))).astype(np.uint8)
        elif kind == 2:
            b = np.concatenate([mutate(rng, a[: ln // 2]), rng.integers(0, 4, int(rng.integers(0, 300))).astype(np.uint8)])
        else:
            b = np.concatenate([rng.integers(0, 4, int(rng.integers(0, 60))).astype(np.uint8), mutate(rng, a)])
        if len(b) == 0:
            b = np.array([1], np.uint8)
        yield a, b


@pytest.mark.parametrize("mode,fn", [(0, "ref_extend_core"), (1, "ref_extend_shift_core")])
def test_extend_matches_reference(ref_lib, oracle_lib, mode, fn):
    rng = np.random.default_rng(11 + mode)
    n = 0
    for a, b in _cases(rng, 300):
        init = int(rng.choice([0, 0, 17, 200, 2500, -5]))
        W = int(rng.choice([50, 50, 5, 1, 120, -800, -30, -7, -1]))
        exp = call_ext(ref_lib, fn, None, a, b, init, W)
        got = call_ext(oracle_lib, "orc_extend", mode, a, b, init, W)
        assert exp == got, (mode, len(a), len(b), init, W)
        n += 1
    assert n == 300


def test_extend_edge_cases(ref_lib, oracle_lib):
    one = np.array([2], np.uint8)
    same = np.array([0, 1, 2, 3] * 10, np.uint8)
    for mode, fn in ((0, "ref_extend_core"), (1, "ref_extend_shift_core")):
        for q, t, init, W in [(one, one, 0, 50), (one, same, 0, -800), (same, one, 100, -800), (same, same, 0, 50),
                              (same, (3 - same).astype(np.uint8), 0, 50), (same, same[::-1].copy(), 40, -3)]:
            assert call_ext(ref_lib, fn, None, q, t, init, W) == call_ext(oracle_lib, "orc_extend", mode, q, t, init, W)


def test_global_matches_reference(ref_lib, oracle_lib):
    rng = np.random.default_rng(5)
    for a, b in _cases(rng, 300):
        w = int(rng.choice([50, 100, 7, 400]))
        while w < abs(len(a) - len(b)):
            w <<= 1
        assert call_global(ref_lib, "ref_global2", a, b, w) == call_global(oracle_lib, "orc_global2", a, b, w)
    z = np.zeros(0, np.uint8)
    a = np.array([0, 1, 2, 3, 3], np.uint8)
    assert call_global(ref_lib, "ref_global2", a, a[:1], 50) == call_global(oracle_lib, "orc_global2", a, a[:1], 50)
    assert call_global(ref_lib, "ref_global2", a[:1], a, 50) == call_global(oracle_lib, "orc_global2", a[:1], a, 50)


def test_runlen_align_matches_reference(ref_lib, oracle_lib):
    rng = np.random.default_rng(9)

    def run(lib, name, x, y):
        out = (C.c_int * 10)()
        cig = (C.c_uint32 * 64)()
        n = getattr(lib, name)(x.ctypes.data_as(C.c_void_p), len(x), y.ctypes.data_as(C.c_void_p), len(y), 2, -3, -3, -1, out, cig, 64)
        return list(out), list(cig[:n])

    for _ in range(300):
        runs = rng.integers(0, 4, int(rng.integers(1, 12)))
        runs = runs[np.insert(np.diff(runs) != 0, 0, True)]
        x = np.repeat(runs, rng.integers(1, 4, len(runs))).astype(np.uint8)
        y = np.repeat(runs, rng.integers(1, 4, len(runs))).astype(np.uint8)
        if rng.random() < 0.1:
            y = y.copy()
            y[-1] = (y[-1] + 1) & 3
        assert run(ref_lib, "ref_hz_align", x, y) == run(oracle_lib, "orc_hz_align", x, y)


@pytest.mark.parametrize("name", ["sort_u64_asc", "sort_u64_lo32_desc", "sort_u64_hi32_asc", "sort_u64_hi32_desc"])
def test_sort_array_permutation(ref_lib, oracle_lib, name):
    rng = np.random.default_rng(3)
    for n in list(range(0, 40)) + [63, 64, 65, 200, 1000, 5000]:
        for nkeys in (2, 5, 50, 1 << 20):
            keys = rng.integers(0, nkeys, n).astype(np.uint64)
            if "hi32" in name:
                arr = (keys << np.uint64(32)) | np.arange(n, dtype=np.uint64)
            elif "lo32" in name:
                arr = (np.arange(n, dtype=np.uint64) << np.uint64(32)) | keys
            else:
                arr = keys.copy()
            a = arr.copy()
            b = arr.copy()
            getattr(ref_lib, "ref_" + name)(a.ctypes.data_as(C.c_void_p), C.c_size_t(n))
            getattr(oracle_lib, "orc_" + name)(b.ctypes.data_as(C.c_void_p), C.c_size_t(n))
            assert (a == b).all(), (name, n, nkeys)
