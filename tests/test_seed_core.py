"""CPU: pair-seeding stage.  (1) the oracle restatement against the real reference functions
(ref_shim.c: index_single_read_seeds .. chaining_wtseedv), (2) the product's per-pair device logic
(smartdenovo_b200/csrc/zmo_seed_core.cuh, compiled for the host by tests/hostsim -- test-only)
against the oracle.  Bit-exact on match counts, chain weights, windows and anchors."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import REPO, mutate


@pytest.fixture(scope="module")
def sim_lib():
    out = os.path.join(REPO, "tests", "_build", "libseed_host.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", out, os.path.join(REPO, "tests", "hostsim", "seed_host.cpp")], check=True)
    return C.CDLL(out)


def run_pw(lib, name, a, b, zsize=10, hz=1, zcut=64, kvar=2, kwin=800, kstep=400, zovl=200, ztot=300, W=3200):
    a = np.ascontiguousarray(a, np.uint8)
    b = np.ascontiguousarray(b, np.uint8)
    nh = C.c_int(0)
    na = C.c_int(0)
    ovl = (C.c_int * 2)()
    wcap, acap = 4096, 1 << 18
    wins = (C.c_int * (7 * wcap))()
    anc = (C.c_int * (6 * acap))()
    nw = getattr(lib, name)(a.ctypes.data_as(C.c_void_p), len(a), b.ctypes.data_as(C.c_void_p), len(b), zsize, hz, zcut, kvar, kwin, kstep,
                            zovl, ztot, W, C.byref(nh), ovl, wins, wcap, anc, acap, C.byref(na))
    assert 0 <= nw <= wcap and na.value <= acap
    return nh.value, list(ovl), list(wins[: 7 * nw]), list(anc[: 6 * na.value])


def pairs(seed, n):
    rng = np.random.default_rng(seed)
    for i in range(n):
        g = rng.integers(0, 4, 9000).astype(np.uint8)
        if i % 5 == 4:   # tandem/low-complexity stretch to exercise the -Z cap and repeated z-mers
            unit = rng.integers(0, 4, 37).astype(np.uint8)
            g[3000:4500] = np.tile(unit, 41)[:1500]
        s1, s2 = int(rng.integers(0, 3000)), int(rng.integers(0, 3000))
        a = mutate(rng, g[s1:s1 + int(rng.integers(2000, 6000))])
        b = mutate(rng, g[s2:s2 + int(rng.integers(2000, 6000))])
        if rng.random() < 0.5:
            b = (3 - b[::-1]).astype(np.uint8)
        if i % 7 == 6:
            b = rng.integers(0, 4, 3000).astype(np.uint8)   # unrelated pair
        yield a, b


def test_oracle_pair_windows_matches_reference(ref_lib, oracle_lib):
    nwin = 0
    for a, b in pairs(100, 40):
        exp = run_pw(ref_lib, "ref_pair_windows", a, b)
        got = run_pw(oracle_lib, "orc_pair_windows", a, b)
        assert exp == got
        nwin += len(exp[2]) // 7
    assert nwin > 50
    for a, b in pairs(101, 10):   # non-default parameters
        kw = dict(zsize=12, zcut=16, kvar=1, kwin=500, kstep=250, zovl=100, ztot=200, W=1000)
        assert run_pw(ref_lib, "ref_pair_windows", a, b, **kw) == run_pw(oracle_lib, "orc_pair_windows", a, b, **kw)
        kw = dict(zsize=8, hz=0, zcut=255, kvar=0)
        assert run_pw(ref_lib, "ref_pair_windows", a, b, **kw) == run_pw(oracle_lib, "orc_pair_windows", a, b, **kw)


def test_device_seed_core_matches_oracle(oracle_lib, sim_lib):
    nwin = 0
    for a, b in pairs(200, 60):
        exp = run_pw(oracle_lib, "orc_pair_windows", a, b)
        got = run_pw(sim_lib, "sim_pair_windows", a, b)
        assert exp == got
        nwin += len(exp[2]) // 7
    assert nwin > 80
    for a, b in pairs(201, 10):
        for kw in (dict(zsize=12, zcut=16, kvar=1, kwin=500, kstep=250, zovl=100, ztot=200, W=1000), dict(zsize=8, hz=0, zcut=255, kvar=0), dict(zsize=16)):
            assert run_pw(oracle_lib, "orc_pair_windows", a, b, **kw) == run_pw(sim_lib, "sim_pair_windows", a, b, **kw)
    # degenerate inputs: reads shorter than z, identical reads
    tiny = np.array([0, 1, 2], np.uint8)
    a = next(pairs(5, 1))[0]
    for x, y in ((tiny, a), (a, tiny), (a, a)):
        assert run_pw(oracle_lib, "orc_pair_windows", x, y) == run_pw(sim_lib, "sim_pair_windows", x, y)


def run_dot(lib, name, a, b, zcut=16, xvar=128, yvar=64, mbl=160, dev=1.0, gap=0.05):
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    out = (C.c_int * 6)()
    n = getattr(lib, name)(a.ctypes.data_as(C.c_void_p), len(a), b.ctypes.data_as(C.c_void_p), len(b), 10, 1, zcut, 2, xvar, yvar, mbl, 2 * xvar,
                           C.c_float(dev), C.c_float(gap), out)
    return n, list(out)


def test_device_dot_core_matches_oracle(oracle_lib, sim_lib):
    hits = 0
    for a, b in pairs(400, 60):
        exp = run_dot(oracle_lib, "orc_pair_dotmatrix", a, b)
        assert exp == run_dot(sim_lib, "sim_pair_dotmatrix", a, b)
        hits += exp[1][0] > 0
    assert hits > 15
    for a, b in pairs(401, 10):
        kw = dict(zcut=64, xvar=256, yvar=32, mbl=300, dev=0.1, gap=0.01)
        assert run_dot(oracle_lib, "orc_pair_dotmatrix", a, b, **kw) == run_dot(sim_lib, "sim_pair_dotmatrix", a, b, **kw)


def test_dotmatrix_oracle_matches_reference(ref_lib, oracle_lib):
    run = run_dot
    hits = 0
    for a, b in pairs(300, 40):
        a = np.ascontiguousarray(a)
        b = np.ascontiguousarray(b)
        exp = run(ref_lib, "ref_pair_dotmatrix", a, b)
        assert exp == run(oracle_lib, "orc_pair_dotmatrix", a, b)
        hits += exp[1][0] > 0
    assert hits > 10


def test_chunk_scan_equals_full_scan(sim_lib):
    """zmo_scan_kmers_chunk over disjoint chunks == zmo_scan_kmers over the read (used by the parallel scans)"""
    rng = np.random.default_rng(77)
    for trial in range(60):
        n = int(rng.integers(1, 3000))
        s = rng.integers(0, 4, n).astype(np.uint8)
        if trial % 3 == 0:     # long homopolymers and dinucleotide repeats
            s = np.repeat(rng.integers(0, 4, max(1, n // 7)), rng.integers(1, 15, max(1, n // 7))).astype(np.uint8)[:n]
        s = np.ascontiguousarray(s)
        for k, hp in ((16, 1), (10, 1), (5, 1), (32, 1), (16, 0), (10, 0)):
            for chunk in (1, 7, 64, 128, 1000):
                assert sim_lib.sim_scan_chunks_equal(s.ctypes.data_as(C.c_void_p), len(s), k, hp, chunk) == 1, (trial, len(s), k, hp, chunk)
