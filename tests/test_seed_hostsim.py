"""CPU: the pair-seeding KERNEL k_p_seed (smartdenovo_b200/csrc/zmo_seed_kernels.cuh: one warp per pair, register-chunk scan of
the match list, warp-cooperative span searches with bitonic key sorts and ballot compaction from zmo_seed_warp.cuh, serial
fallbacks from zmo_seed_core.cuh, chain, window/anchor export) compiled for the host by tests/hostsim (test-only, warps run as
cooperative fibers) against the oracle's restatement of merge_paired_kmers_window .. chaining_wtseedv (hzm_aln.h:316-713):
match counts, chain weights, kept windows and anchors bit-exact.  tests/test_gpu_stages.py runs the same stage on the device."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import REPO
from test_seed_core import pairs, run_pw


@pytest.fixture(scope="module")
def seedk_sim():
    out = os.path.join(REPO, "tests", "_build", "libseedk_host.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    subprocess.run(["g++", "-O1", "-std=c++17", "-I" + os.path.join(REPO, "tests", "hostsim", "emu"), "-fPIC", "-shared", "-o", out,
                    os.path.join(REPO, "tests", "hostsim", "seedk_host.cpp")], check=True)
    return C.CDLL(out)


def run_kernel(lib, a, b, copies=1, force_tie=0, F=2, zsize=10, hz=1, zcut=64, kvar=2, kwin=800, kstep=400, zovl=200, ztot=300, W=3200):
    a = np.ascontiguousarray(a, np.uint8)
    b = np.ascontiguousarray(b, np.uint8)
    nh, na = C.c_int(0), C.c_int(0)
    ovl = (C.c_int * 2)()
    wcap, acap = 4096, 1 << 18
    wins = (C.c_int * (7 * wcap))()
    anc = (C.c_int * (6 * acap))()
    nw = lib.simk_pair_windows(a.ctypes.data_as(C.c_void_p), len(a), b.ctypes.data_as(C.c_void_p), len(b), zsize, hz, zcut, kvar, kwin, kstep, zovl, ztot, W,
                               copies, force_tie, F, C.byref(nh), ovl, wins, wcap, anc, acap, C.byref(na))
    assert nw != -2, "copies of the same pair disagree"
    if nw == -1:
        return None       # scratch overflow at this capacity factor: the product retries with 4 x F
    assert 0 <= nw <= wcap and na.value <= acap
    return nh.value, list(ovl), list(wins[: 7 * nw]), list(anc[: 6 * na.value])


def expected(orc, a, b, **kw):
    """oracle windows, restricted to what the kernel exports: strands whose chain weight reaches ztot (wtzmo.c:896-914); the
    kernel reports no chain weights at all when the match list is too short to reach ztot (hzm_aln.h:1184)"""
    ztot = kw.get("ztot", 300)
    nh, ovl, wins, anc = run_pw(orc, "orc_pair_windows", a, b, **kw)
    w2, a2, k = [], [], 0
    for i in range(len(wins) // 7):
        na = wins[7 * i + 6]
        if ovl[wins[7 * i]] >= ztot:
            w2 += wins[7 * i: 7 * i + 7]
            a2 += anc[6 * k: 6 * (k + na)]
        k += na
    return nh, ovl, w2, a2


def check(lib, orc, a, b, **kw):
    simkw = {k: v for k, v in kw.items() if k not in ("copies", "force_tie")}
    exp = expected(orc, a, b, **simkw)
    got = None
    for F in (2, 8, 32, 128):
        got = run_kernel(lib, a, b, F=F, **kw)
        if got is not None:
            break
    assert got is not None, "overflow at every capacity factor"
    assert got == exp
    return len(exp[2]) // 7


@pytest.mark.parametrize("lanes", [0, 4, 32])
def test_seed_kernel_matches_oracle(seedk_sim, oracle_lib, lanes):
    seedk_sim.simk_set_seed_lanes(lanes)      # 0: k_p_seed (a pair per warp); G: k_p_seed_lanes with G pairs per warp (zmo_seed_lanes.cuh)
    nwin = 0
    for a, b in pairs(500, 24):
        nwin += check(seedk_sim, oracle_lib, a, b)
    assert nwin > 30


@pytest.mark.parametrize("lanes", [0, 4, 32])
def test_seed_kernel_parameters_and_degenerate_inputs(seedk_sim, oracle_lib, lanes):
    seedk_sim.simk_set_seed_lanes(lanes)      # 0: k_p_seed (a pair per warp); G: k_p_seed_lanes with G pairs per warp (zmo_seed_lanes.cuh)
    for a, b in pairs(501, 5):
        for kw in (dict(zsize=12, zcut=16, kvar=1, kwin=500, kstep=250, zovl=100, ztot=200, W=1000), dict(zsize=8, hz=0, zcut=255, kvar=0), dict(zsize=16)):
            check(seedk_sim, oracle_lib, a, b, **kw)
    tiny = np.array([0, 1, 2], np.uint8)
    a = next(pairs(5, 1))[0]
    for x, y in ((tiny, a), (a, tiny), (a, a)):
        check(seedk_sim, oracle_lib, x, y)


@pytest.mark.parametrize("lanes", [0, 4, 32])
def test_seed_kernel_tie_path_and_work_loop(seedk_sim, oracle_lib, lanes):
    """force_tie: the kernel first rebuilds the reference's emission order, then runs the exact sort_array emulation -- the same
    bytes as the radix-sorted fast path; 50 copies of a pair on 2 CTAs x 22 warps run the work-counter loop"""
    seedk_sim.simk_set_seed_lanes(lanes)      # 0: k_p_seed (a pair per warp); G: k_p_seed_lanes with G pairs per warp (zmo_seed_lanes.cuh)
    for i, (a, b) in enumerate(pairs(502, 6)):
        check(seedk_sim, oracle_lib, a, b, force_tie=1)
        if i < 2:
            check(seedk_sim, oracle_lib, a, b, copies=50)
    # palindromic stretches: a q position matching both strands at the same c coordinate gives genuinely tied keys
    rng = np.random.default_rng(3)
    g = rng.integers(0, 4, 3000).astype(np.uint8)
    half = rng.integers(0, 4, 40).astype(np.uint8)
    pal = np.concatenate([half, (3 - half[::-1]).astype(np.uint8)])
    for k in range(200, 2800, 300):
        g[k: k + 80] = pal
    check(seedk_sim, oracle_lib, g, g.copy())
    check(seedk_sim, oracle_lib, g, (3 - g[::-1]).astype(np.uint8))


def run_dot_kernel(lib, a, b, copies=1, force_tie=0, ztot=0, zcut=16, xvar=128, yvar=64, mbl=160, dev=1.0, gap=0.05):
    a = np.ascontiguousarray(a, np.uint8)
    b = np.ascontiguousarray(b, np.uint8)
    out = (C.c_int * 6)()
    n = lib.simk_pair_dotmatrix(a.ctypes.data_as(C.c_void_p), len(a), b.ctypes.data_as(C.c_void_p), len(b), 10, 1, zcut, 2, xvar, yvar, mbl, 2 * xvar,
                                C.c_float(dev), C.c_float(gap), ztot, copies, force_tie, out)
    assert n != -2, "copies of the same pair disagree"
    return n, list(out)


def test_dot_kernel_matches_oracle(seedk_sim, oracle_lib):
    """k_p_dot (zmo_dot_kernels.cuh): lane 0 runs the serial dot-matrix logic of zmo_dot_core.cuh, its sort_array emulations are staged
    through shared memory by the helper lanes (mailbox + paired __syncwarp()s); input sorted by diagonal like the device front end"""
    from test_seed_core import run_dot
    hits = 0
    for i, (a, b) in enumerate(pairs(600, 30)):
        exp = run_dot(oracle_lib, "orc_pair_dotmatrix", a, b)
        assert exp == run_dot_kernel(seedk_sim, a, b)
        if i % 6 == 0:
            assert exp == run_dot_kernel(seedk_sim, a, b, force_tie=1)
        if i % 10 == 0:
            assert exp == run_dot_kernel(seedk_sim, a, b, copies=20)
        hits += exp[1][0] > 0
    assert hits > 8
    for a, b in pairs(601, 6):
        kw = dict(zcut=64, xvar=256, yvar=32, mbl=300, dev=0.1, gap=0.01)
        assert run_dot(oracle_lib, "orc_pair_dotmatrix", a, b, **kw) == run_dot_kernel(seedk_sim, a, b, **kw)
    # below the -r threshold the kernel reports an empty hit without looking at the list (hzm_aln.h:1184)
    a, b = next(pairs(602, 1))
    n, out = run_dot_kernel(seedk_sim, a, b[:300], ztot=10 ** 6)
    assert out[0] == 0 and out[5] == 0


@pytest.mark.parametrize("lanes", [0, 4, 32])
def test_seeding_stage_batch(seedk_sim, oracle_lib, lanes):
    """the whole device side of zmo_pair_windows for a batch of pairs sharing query reads: z-index of the batch (k_z_scan, k_z_heads,
    k_z_slots with the hashed slot filter, k_z_ranges), chunk-parallel hits (k_hit), rank cap + expansion (k_expand), unpacking with tie
    flags (k_unpack, k_pair_offsets) -- zmo_seedfront_kernels.cuh, with std:: sorts / scans where seed_prepare calls CUB -- then k_p_seed;
    every pair must give the oracle's match count, chain weights, windows and anchors"""
    seedk_sim.simk_set_seed_lanes(lanes)      # 0: k_p_seed (a pair per warp); G: k_p_seed_lanes with G pairs per warp (zmo_seed_lanes.cuh)
    reads = []
    for a, b in pairs(800, 3):
        reads += [a, b]
    rng = np.random.default_rng(9)
    reads.append(rng.integers(0, 4, 7).astype(np.uint8))          # shorter than z: no z-mers at all
    plist = [(0, 1), (2, 3), (4, 5), (0, 3), (1, 0), (2, 1), (0, 6), (6, 0), (5, 4), (0, 0)]
    seqs = np.ascontiguousarray(np.concatenate(reads), np.uint8)
    lens = (C.c_int * len(reads))(*[len(r) for r in reads])
    flat = (C.c_int * (2 * len(plist)))(*[v for p in plist for v in p])
    nwin = 0
    for zcut in (64, 3):          # -Z 3: the per-slot rank cap cuts most repeated z-mers
        for which, (q, c) in enumerate(plist):
            exp = expected(oracle_lib, reads[q], reads[c], zcut=zcut)
            got = None
            for F in (2, 8, 32):
                nh, na = C.c_int(0), C.c_int(0)
                ovl = (C.c_int * 2)()
                wcap, acap = 4096, 1 << 18
                wins = (C.c_int * (7 * wcap))()
                anc = (C.c_int * (6 * acap))()
                tie = (C.c_int * len(plist))()
                nz = (C.c_int * len(plist))()
                nw = seedk_sim.simk_batch_windows(seqs.ctypes.data_as(C.c_void_p), lens, len(reads), flat, len(plist), 10, 1, zcut, 2, 800, 400, 200, 300, 3200,
                                                  F, which, tie, nz, C.byref(nh), ovl, wins, wcap, anc, acap, C.byref(na))
                if nw >= 0:
                    got = (nh.value, list(ovl), list(wins[: 7 * nw]), list(anc[: 6 * na.value]))
                    assert nz[which] == nh.value
                    break
            assert got is not None
            assert got == exp, (zcut, which, q, c)
            nwin += len(exp[2]) // 7
    assert nwin > 10
