"""CPU: the product's DP executor KERNELS (smartdenovo_b200/csrc/zmo_dp_kernels.cuh: k_ext_warp, k_ext_cta<64|128,7>,
k_ext_cta<128,13>, k_glb_warp, k_glb_cta; zmo_winalign.cuh: k_window_align; with the register-resident sweeps of zmo_dpr.cuh and the shared/global-row
fallback of zmo_dp.cuh underneath) compiled for the host by tests/hostsim (test-only; thread blocks run as cooperative
fibers, warp collectives and barriers emulated) against the oracle: score, end point, counts and CIGAR bit-exact.
The same comparisons run on the real device in tests/test_gpu_dp.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import REPO, call_ext, call_global, mutate

SC = (2, -5, -3, -1, -50)   # -M -X -O -E -T defaults (wtzmo.c:1574-1578)
# compile-time variants of the kernels (ZMO_SIM_DEFINES="-D...") are checked with the same tests
SIM_DEFINES = os.environ.get("ZMO_SIM_DEFINES", "").split()


@pytest.fixture(scope="module")
def dp_sim():
    out = os.path.join(REPO, "tests", "_build", "libdp_host.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    subprocess.run(["g++", "-O1", "-std=c++17"] + SIM_DEFINES + ["-I" + os.path.join(REPO, "tests", "hostsim", "emu"), "-fPIC", "-shared", "-o", out,
                    os.path.join(REPO, "tests", "hostsim", "dp_host.cpp")], check=True)
    return C.CDLL(out)


def sim_ext(sim, mode, cls, copies, q, t, init, W, sc=SC):
    q = np.ascontiguousarray(q, np.uint8)
    t = np.ascontiguousarray(t, np.uint8)
    out = (C.c_int * 10)()
    cap = len(q) + len(t) + 8
    cig = (C.c_uint32 * cap)()
    M, X, O, E, T = sc
    n = sim.sim_dp_extend(mode, cls, copies, q.ctypes.data_as(C.c_void_p), len(q), t.ctypes.data_as(C.c_void_p), len(t), init, W, M, X, O, E, T, out, cig, cap)
    assert n >= 0, "simulation refused the problem (%d)" % n
    out = list(out)
    cells, out[1] = out[1], 0
    return out, list(cig[:n]), cells


def check_ext(sim, orc, mode, cls, q, t, init, W, sc=SC, copies=1):
    got, gcig, cells = sim_ext(sim, mode, cls, copies, q, t, init, W, sc)
    M, X, O, E, T = sc
    exp, ecig = call_ext(orc, "orc_extend", mode, q, t, init, W, sc=(M, X, O, O, E, T))
    assert got == exp, (mode, cls, len(q), len(t), init, W, got, exp)
    assert gcig == ecig, (mode, cls, len(q), len(t), init, W)
    return cells


def related(rng, n, err=1.0, shift=0):
    """a noisy copy pair: q = genome slice, t = mutated copy (PacBio-like profile scaled by err)"""
    g = rng.integers(0, 4, n + abs(shift) + 50).astype(np.uint8)
    q = g[:n]
    t = mutate(rng, g[max(shift, 0): max(shift, 0) + n], ins=0.0825 * err, dele=0.045 * err, sub=0.0225 * err)
    return q, t


@pytest.mark.parametrize("mode", [0, 1])
def test_kernels_by_executor_class(dp_sim, oracle_lib, mode):
    """every executor class on bands it serves: warp x 7 (<= 211 columns), 64 x 7 (<= 435), 128 x 7 (<= 883), 128 x 13 (<= 1639)"""
    rng = np.random.default_rng(10 + mode)
    total = 0
    for cls, Ws, nmax in ((0, [-3, -20, -50, -105, 50], 500), (1, [-120, -217], 420), (2, [-250, -441], 380), (3, [-500, -819], 330)):
        for W in Ws:
            for rep in range(2):
                n = int(rng.integers(nmax // 2, nmax))
                q, t = related(rng, n, err=float(rng.choice([0.3, 1.0])))
                total += check_ext(dp_sim, oracle_lib, mode, -1, q, t, int(rng.choice([0, 0, 150])), W)
    assert total > 10 ** 6


def test_band_wider_than_the_register_executors(dp_sim, oracle_lib):
    """> 1639 columns: the class-3 kernel falls back to the chunked sweep with H/E rows in shared (2W+3 <= 4096... per kernel) or global memory"""
    rng = np.random.default_rng(5)
    q, t = related(rng, 2100, err=0.5)
    for mode in (0, 1):
        check_ext(dp_sim, oracle_lib, mode, -1, q[:260], t, 0, -900)
        check_ext(dp_sim, oracle_lib, mode, -1, q[:200], t, 0, -1500)


def test_larger_class_gives_the_same_bytes(dp_sim, oracle_lib):
    """a band may run on any executor class that holds it (the pipeline's class choice is a performance decision only)"""
    rng = np.random.default_rng(6)
    q, t = related(rng, 300)
    for mode in (0, 1):
        for cls in (0, 1, 2, 3):
            check_ext(dp_sim, oracle_lib, mode, cls, q, t, 0, -40)


def test_persistent_executor_loop(dp_sim, oracle_lib):
    """several jobs on a 2-CTA grid: the first job of an executor is its own index, the rest come from the work counter"""
    rng = np.random.default_rng(7)
    q, t = related(rng, 200)
    for cls, copies in ((0, 11), (1, 5), (3, 4)):
        check_ext(dp_sim, oracle_lib, 1, cls, q, t, 0, -30, copies=copies)


def test_end_rules_and_degenerate_shapes(dp_sim, oracle_lib):
    """early termination (row maximum <= 0), the T rule (kswx.h:200-204), W > 0 clamped by max_gap (kswx.h:115-121), length
    truncation, negative init, 1 x n and n x 1, unrelated sequences, other scoring parameters"""
    rng = np.random.default_rng(8)
    q, t = related(rng, 400)
    u = rng.integers(0, 4, 400).astype(np.uint8)
    for mode in (0, 1):
        check_ext(dp_sim, oracle_lib, mode, -1, q, u, 0, -50)                 # unrelated: stops after a few rows
        check_ext(dp_sim, oracle_lib, mode, -1, q, u, 300, -50)
        check_ext(dp_sim, oracle_lib, mode, -1, q, t, -9, 50)                 # negative init clamps to 0; W > 0
        check_ext(dp_sim, oracle_lib, mode, -1, q, t, 3000, 800)
        check_ext(dp_sim, oracle_lib, mode, -1, q[:1], t, 0, -800)
        check_ext(dp_sim, oracle_lib, mode, -1, q, t[:1], 100, -800)
        check_ext(dp_sim, oracle_lib, mode, -1, q[:1], t[:1], 0, -1)
        check_ext(dp_sim, oracle_lib, mode, -1, q[:120], t, 0, -30)           # tl truncated to ql + W
        check_ext(dp_sim, oracle_lib, mode, -1, q, t[:120], 0, -30)           # ql truncated to tl + W
        check_ext(dp_sim, oracle_lib, mode, -1, q, np.concatenate([t[:200], u[:150]]), 0, -60)   # good prefix, then noise
        for sc in ((1, -2, -2, -1, -20), (3, -7, -6, -2, -100), (2, -5, -3, -1, 0)):
            check_ext(dp_sim, oracle_lib, mode, -1, q, t, 0, -45, sc=sc)
    lowc = np.repeat(rng.integers(0, 4, 150), 3).astype(np.uint8)             # homopolymer triples: many score ties
    for mode in (0, 1):
        check_ext(dp_sim, oracle_lib, mode, -1, lowc, lowc[3:], 0, 50)
        check_ext(dp_sim, oracle_lib, mode, -1, lowc, mutate(rng, lowc), 0, -25)


def sim_glb(sim, wide, q, t, w, wmax=0, sc=(2, -5, -3, -1)):
    q = np.ascontiguousarray(q, np.uint8)
    t = np.ascontiguousarray(t, np.uint8)
    score, wused = C.c_int(0), C.c_int(0)
    cnt = (C.c_int * 4)()
    cap = len(q) + len(t) + 8
    cig = (C.c_uint32 * cap)()
    M, X, O, E = sc
    n = sim.sim_dp_global(wide, q.ctypes.data_as(C.c_void_p), len(q), t.ctypes.data_as(C.c_void_p), len(t), M, X, O, E, w, wmax, C.byref(score), C.byref(wused), cnt, cig, cap)
    return score.value, wused.value, list(cnt), list(cig[:n])


def check_glb(sim, orc, wide, q, t, w, wmax=0):
    score, wused, cnt, cig = sim_glb(sim, wide, q, t, w, wmax)
    ww = w
    while True:       # hzm_aln.h:1400-1418: widen until the band holds the length difference, then while the score is negative
        while ww < abs(len(q) - len(t)):
            ww <<= 1
        escore, ecig = call_global(orc, "orc_global2", q, t, ww, sc=(2, -5, 3, 1, 3, 1))
        if wmax > 0 and escore < 0 and ww < wmax and ww < max(len(q), len(t)):
            ww <<= 1
        else:
            break
    assert (score, wused) == (escore, ww), (wide, len(q), len(t), w, wmax)
    assert cig == ecig, (wide, len(q), len(t), w)
    x1 = x2 = mat = mis = ins = dele = 0
    for op in ecig:
        ln, o = op >> 4, op & 15
        if o == 0:
            same = int((q[x1:x1 + ln] == t[x2:x2 + ln]).sum())
            mat += same; mis += ln - same; x1 += ln; x2 += ln
        elif o == 1:
            ins += ln; x1 += ln
        else:
            dele += ln; x2 += ln
    assert cnt == [mat, mis, ins, dele]


def test_gap_kernels(dp_sim, oracle_lib):
    """ksw_global2 through k_glb_warp (1/2/4/7 columns per lane by band width) and k_glb_cta, with the reference's band doubling"""
    rng = np.random.default_rng(9)
    for n, w in ((30, 50), (60, 7), (150, 50), (250, 100), (400, 50), (420, 400)):
        q, t = related(rng, n)
        check_glb(dp_sim, oracle_lib, 0, q, t, w)
    q, t = related(rng, 300)
    check_glb(dp_sim, oracle_lib, 0, q[:200], t, 7)                 # band doubles until it holds the length difference
    check_glb(dp_sim, oracle_lib, 0, q, rng.integers(0, 4, 280).astype(np.uint8), 8, wmax=3200)   # negative score: retry with 2w
    for n, w in ((500, 300), (700, 400)):
        q, t = related(rng, n, err=0.5)
        check_glb(dp_sim, oracle_lib, 1, q, t, w)
    check_glb(dp_sim, oracle_lib, 1, q[:40], t[:60], 50)            # small problem on the wide kernel


def _windows(orc, a, b):
    """kept windows + anchors of the pair (a = q, b = c) from the oracle's seeding stage"""
    from test_seed_core import run_pw
    nh, ovl, wins, anc = run_pw(orc, "orc_pair_windows", a, b)
    out, k = [], 0
    for i in range(len(wins) // 7):
        d, b0, e0, b1, e1, _, na = wins[7 * i: 7 * i + 7]
        out.append((d, e0 - b0, e1 - b1, anc[6 * k: 6 * (k + na)]))
        k += na
    return out


@pytest.mark.parametrize("w", [50, 20, 120])
def test_window_align_kernel(dp_sim, oracle_lib, w):
    """k_window_align (zmo_winalign.cuh: anchor walk, fixed-band bridges with 1/2/4/7 columns per lane, traceback in shared memory
    for short bridges, D/I padding, run-length anchor alignment, CIGAR splicing, region filter) against the oracle's restatement of
    fast_seeds_align_hzmo (hzm_aln.h:1247-1302) on the windows and anchors of real read pairs, both strands"""
    from test_seed_core import pairs
    n_win = 0
    for a, b in pairs(40 + w, 6):
        wl = _windows(oracle_lib, a, b)
        for d in (0, 1):
            ws = [x for x in wl if x[0] == d]
            if not ws:
                continue
            c_strand = b if d == 0 else (3 - b[::-1]).astype(np.uint8)
            win = (C.c_int * (3 * len(ws)))(*[v for x in ws for v in (x[1], x[2], len(x[3]) // 6)])
            flat = [v for x in ws for v in x[3]]
            anc = (C.c_int * max(len(flat), 1))(*flat)
            out = (C.c_int * (11 * len(ws)))()
            cap = sum(x[1] + x[2] + 16 + len(x[3]) // 3 for x in ws)
            cig = (C.c_uint32 * cap)()
            cn = (C.c_int * len(ws))()
            qa = np.ascontiguousarray(a, np.uint8)
            cb = np.ascontiguousarray(b, np.uint8)
            tot = dp_sim.sim_window_align(qa.ctypes.data_as(C.c_void_p), len(qa), cb.ctypes.data_as(C.c_void_p), len(cb), d, win, len(ws), anc,
                                          w, 2, -5, -3, -1, -50, 200, C.c_float(0.6), out, cig, cap, cn)
            assert tot >= 0
            pos = 0
            for i, x in enumerate(ws):
                eo = (C.c_int * 10)()
                ecap = x[1] + x[2] + 16 + len(x[3]) // 3
                ec = (C.c_uint32 * ecap)()
                ea = (C.c_int * len(x[3]))(*x[3])
                cs = np.ascontiguousarray(c_strand, np.uint8)
                en = oracle_lib.orc_window_align(qa.ctypes.data_as(C.c_void_p), cs.ctypes.data_as(C.c_void_p), ea, len(x[3]) // 6, w, 2, -5, -3, -1, -50, eo, ec, ecap)
                got = list(out[11 * i: 11 * i + 10])
                assert got == list(eo), (d, i, got, list(eo))
                assert list(cig[pos: pos + cn[i]]) == list(ec[:en]), (d, i)
                kept = not (eo[5] * 2 < 200 or np.float32(eo[6]) < np.float32(eo[5]) * np.float32(0.6))
                assert out[11 * i + 10] == int(kept)
                pos += cn[i]
                n_win += 1
    assert n_win >= 8


@pytest.mark.parametrize("w,sc,copies,acap", [(50, (2, -5, -3, -1, -50), 1, 22), (20, (2, -5, -3, -1, -50), 9, 22), (91, (2, -5, -3, -1, -50), 1, 22), (50, (1, -3, -2, -2, -10), 1, 22), (50, (2, -5, -3, -1, -20), 1, 22), (50, (2, -5, -3, -1, -50), 1, 6), (50, (2, -5, -3, -1, -50), 0, 22),
                                                    (50, (2, -5, -3, -1, -50), -12, 22), (20, (2, -5, -3, -1, -50), -7, 22), (91, (2, -5, -3, -1, -50), -30, 22),
                                                    (30, (3, -4, -5, -2, -30), -5, 22), (10, (1, -1, -1, -1, -5), -9, 22), (64, (5, -9, -7, -3, -200), -16, 22), (50, (2, -30, -3, -1, -50), -12, 22)])
def test_window_align_bridge_pipeline(dp_sim, oracle_lib, w, sc, copies, acap):
    """bridge-level window alignment (zmo_winbridge.cuh: k_wb_prep / k_wb_sweep / k_wb_ends / k_wb_walk / k_wb_stitch + k_window_align for the
    windows left out) against the oracle's fast_seeds_align_hzmo on the windows and anchors of real read pairs, both strands.  The sweep runs
    every bridge with init = 0; the test is the claim that this shift changes nothing.  -T -20: max_gap(0) < w for short bridges, so those
    windows must take the sequential kernel (likewise anchors with more CIGAR ops than acap); copies < 0: anchors thinned out so that bridges are longer than the band is wide; copies = 9: more than 32 bridges per warp and several rounds; copies = 0: one window per pass of the pipeline"""
    from test_seed_core import pairs
    M, X, O, E, T = sc
    n_win = n_fb = 0
    thin, copies = (-copies, 1) if copies < 0 else (1, copies)
    for a, b in pairs(40 + w, 6):
        wl = _windows(oracle_lib, a, b)
        if thin > 1:
            # keep every thin-th anchor: bridges several hundred bases long, wider than the band (the sweep's column ring wraps)
            wl = [(x[0], x[1], x[2], [v for k in range(0, len(x[3]) // 6, thin) for v in x[3][6 * k: 6 * k + 6]]) for x in wl]
        for d in (0, 1):
            ws = [x for x in wl if x[0] == d]
            if not ws:
                continue
            c_strand = b if d == 0 else (3 - b[::-1]).astype(np.uint8)
            win = (C.c_int * (3 * len(ws)))(*[v for x in ws for v in (x[1], x[2], len(x[3]) // 6)])
            flat = [v for x in ws for v in x[3]]
            anc = (C.c_int * max(len(flat), 1))(*flat)
            out = (C.c_int * (11 * len(ws)))()
            cap = sum(x[1] + x[2] + 16 + len(x[3]) // 3 for x in ws)
            cig = (C.c_uint32 * cap)()
            cn = (C.c_int * len(ws))()
            nfb = C.c_int(0)
            qa = np.ascontiguousarray(a, np.uint8)
            cb = np.ascontiguousarray(b, np.uint8)
            tot = dp_sim.sim_window_align_bridge(qa.ctypes.data_as(C.c_void_p), len(qa), cb.ctypes.data_as(C.c_void_p), len(cb), d, win, len(ws), anc,
                                                 w, M, X, O, E, T, 200, C.c_float(0.6), copies, acap, out, cig, cap, cn, C.byref(nfb))
            assert tot >= 0, tot
            n_fb += nfb.value
            pos = 0
            for i, x in enumerate(ws):
                eo = (C.c_int * 10)()
                ecap = x[1] + x[2] + 16 + len(x[3]) // 3
                ec = (C.c_uint32 * ecap)()
                ea = (C.c_int * len(x[3]))(*x[3])
                cs = np.ascontiguousarray(c_strand, np.uint8)
                en = oracle_lib.orc_window_align(qa.ctypes.data_as(C.c_void_p), cs.ctypes.data_as(C.c_void_p), ea, len(x[3]) // 6, w, M, X, O, E, T, eo, ec, ecap)
                got = list(out[11 * i: 11 * i + 10])
                assert got == list(eo), (d, i, got, list(eo))
                assert list(cig[pos: pos + cn[i]]) == list(ec[:en]), (d, i)
                pos += cn[i]
                n_win += 1
    assert n_win >= 8
    if T == -20 or acap == 6:
        assert n_fb > 0          # the guards send them to the sequential kernel
    elif sc == SC and w <= 50:
        assert n_fb == 0         # default scores: every window takes the bridge pipeline


def _bridge_vs_oracle(dp_sim, oracle_lib, a, b, ws, w=50, sc=SC):
    """forward-strand windows ws of (a = q, b = c) through the bridge pipeline and through the oracle; returns the number of windows left to k_window_align"""
    M, X, O, E, T = sc
    win = (C.c_int * (3 * len(ws)))(*[v for x in ws for v in (x[1], x[2], len(x[3]) // 6)])
    flat = [v for x in ws for v in x[3]]
    anc = (C.c_int * max(len(flat), 1))(*flat)
    out = (C.c_int * (11 * len(ws)))()
    cap = sum(x[1] + x[2] + 16 + len(x[3]) // 3 for x in ws)
    cig = (C.c_uint32 * cap)()
    cn = (C.c_int * len(ws))()
    nfb = C.c_int(0)
    qa = np.ascontiguousarray(a, np.uint8)
    cb = np.ascontiguousarray(b, np.uint8)
    tot = dp_sim.sim_window_align_bridge(qa.ctypes.data_as(C.c_void_p), len(qa), cb.ctypes.data_as(C.c_void_p), len(cb), 0, win, len(ws), anc,
                                         w, M, X, O, E, T, 200, C.c_float(0.6), 1, 22, out, cig, cap, cn, C.byref(nfb))
    assert tot >= 0, tot
    pos = 0
    res = []
    for i, x in enumerate(ws):
        eo = (C.c_int * 10)()
        ecap = x[1] + x[2] + 16 + len(x[3]) // 3
        ec = (C.c_uint32 * ecap)()
        ea = (C.c_int * len(x[3]))(*x[3])
        en = oracle_lib.orc_window_align(qa.ctypes.data_as(C.c_void_p), cb.ctypes.data_as(C.c_void_p), ea, len(x[3]) // 6, w, M, X, O, E, T, eo, ec, ecap)
        assert list(out[11 * i: 11 * i + 10]) == list(eo), (i, list(out[11 * i: 11 * i + 10]), list(eo))
        assert list(cig[pos: pos + cn[i]]) == list(ec[:en]), i
        pos += cn[i]
        res.append(list(eo))
    return nfb.value, res


def test_window_align_bridge_stopped_sweeps_and_broken_anchors(dp_sim, oracle_lib):
    """the two rare paths of the bridge pipeline: (1) a bridge whose sweep the reference stops early (row maximum <= 0 without improvement:
    the sequence between two anchors replaced by noise while the running score is still small), where k_wb_ends must replay the rows with the
    real init instead of using the sweep's init-free summary; (2) an anchor whose run bases differ (the window is truncated after the bridge
    in front of it, hzm_aln.h:1288-1291)"""
    from test_seed_core import pairs
    rng = np.random.default_rng(77)
    n1 = n2 = 0
    for a, b in pairs(90, 6):
        wl = [x for x in _windows(oracle_lib, a, b) if x[0] == 0 and len(x[3]) // 6 >= 30]
        if not wl:
            continue
        # (1) every 12th anchor; noise between the first two kept anchors
        ws = [(x[0], x[1], x[2], [v for k in range(0, len(x[3]) // 6, 12) for v in x[3][6 * k: 6 * k + 6]]) for x in wl]
        b1 = b.copy()
        for x in ws:
            a0, a1 = x[3][0:6], x[3][6:12]
            lo, hi = a0[1] + a0[3] + 2, a1[1] - 2            # on c, strictly between the two anchors
            if hi - lo > 60:
                b1[lo:hi] = rng.integers(0, 4, hi - lo).astype(np.uint8)
        _, res = _bridge_vs_oracle(dp_sim, oracle_lib, a, b1, ws)
        _, res0 = _bridge_vs_oracle(dp_sim, oracle_lib, a, b, ws)
        n1 += sum(1 for r, r0 in zip(res, res0) if r != r0)
        # (2) one base inside the 5th anchor changed
        b2 = b.copy()
        for x in wl:
            p = x[3][24:30]
            b2[p[1] + p[3] // 2] = (b2[p[1] + p[3] // 2] + 1) & 3
        _, res2 = _bridge_vs_oracle(dp_sim, oracle_lib, a, b2, wl)
        _, res3 = _bridge_vs_oracle(dp_sim, oracle_lib, a, b, wl)
        n2 += sum(1 for r, r0 in zip(res2, res3) if r[5] < r0[5] // 2)      # truncated: far fewer aligned columns
    assert n1 >= 2 and n2 >= 2, (n1, n2)


def _refine_both(sim, orc, q, c, d, tb, qb, cig, W=50):
    """refine the alignment (start tb on q, qb on c's strand d, CIGAR cig) with the simulated kernels and with the oracle"""
    q = np.ascontiguousarray(q, np.uint8)
    c = np.ascontiguousarray(c, np.uint8)
    cs = np.ascontiguousarray(c if d == 0 else (3 - c[::-1]), np.uint8)
    ql = sum(op >> 4 for op in cig if (op & 15) in (0, 1))
    tl = sum(op >> 4 for op in cig if (op & 15) in (0, 2))
    cin = (C.c_uint32 * max(len(cig), 1))(*cig)
    cap = ql + tl + 8
    eo, go = (C.c_int * 10)(), (C.c_int * 10)()
    ec, gc = (C.c_uint32 * cap)(), (C.c_uint32 * cap)()
    cls = C.c_int(-1)
    gn = sim.sim_refine(q.ctypes.data_as(C.c_void_p), len(q), c.ctypes.data_as(C.c_void_p), len(c), d, tb, tb + tl, qb, qb + ql, cin, len(cig),
                        W, 2, -5, -3, -1, go, gc, cap, C.byref(cls))
    assert gn >= 0, gn
    en = orc.orc_refine(cs.ctypes.data_as(C.c_void_p), qb, q.ctypes.data_as(C.c_void_p), tb, W, 2, -5, -3, -1, cin, len(cig), eo, ec, cap)
    assert list(go) == list(eo), (list(go), list(eo))
    assert list(gc[:gn]) == list(ec[:en])
    return cls.value


def test_refine_kernels(dp_sim, oracle_lib):
    """-n: k_refine_size / k_refine_band (per-row band from the CIGAR, widened around indel runs, made monotone), the warp / CTA
    executors of the variable-band sweep (reg_refine) and the wide-band fallback k_refine_wide against the oracle's restatement of kswx_refine_alignment (kswx.h:483-659)"""
    rng = np.random.default_rng(21)
    seen = set()
    for trial in range(10):
        n = int(rng.integers(150, 700))
        g = rng.integers(0, 4, n + 400).astype(np.uint8)
        tb, qb = int(rng.integers(0, 100)), int(rng.integers(0, 100))
        q = g.copy()                                                  # pb1
        body = mutate(rng, g[tb: tb + n])
        c_strand = np.concatenate([rng.integers(0, 4, qb).astype(np.uint8), body, rng.integers(0, 4, 50).astype(np.uint8)])
        d = trial & 1
        c = c_strand if d == 0 else (3 - c_strand[::-1]).astype(np.uint8)
        # a banded global alignment of the two segments supplies a realistic CIGAR (I = base of c only, D = base of q only)
        score, cig = call_global(oracle_lib, "orc_global2", body, g[tb: tb + n], 60)
        seen.add(_refine_both(dp_sim, oracle_lib, q, c, d, tb, qb, cig, W=int(rng.choice([50, 20, 5]))))
        # a crude CIGAR (what a stitch with long pads looks like): the band widens around the long runs
        ql, tl = len(body), n
        k = min(ql, tl) - 40
        run = int(rng.integers(60, 260))
        crude = [(k // 2) << 4, (run << 4) | 1, (run << 4) | 2, ((k - k // 2 - run) << 4)] if k - k // 2 - run > 0 else [(k << 4)]
        used_q = sum(op >> 4 for op in crude if (op & 15) in (0, 1))
        used_t = sum(op >> 4 for op in crude if (op & 15) in (0, 2))
        crude += [((ql - used_q) << 4) | 1, ((tl - used_t) << 4) | 2]
        seen.add(_refine_both(dp_sim, oracle_lib, q, c, d, tb, qb, crude, W=50))
    assert {0, 1} <= seen          # both register executor classes ran
    # indel runs of ~800+ bases need a band beyond the 1,639 columns of the CTA executor: wide-band fallback (band_refine, zmo_dp.cuh),
    # including bands of more than one 1,792-column chunk
    g = rng.integers(0, 4, 6000).astype(np.uint8)
    h = mutate(rng, g)
    for d in (0, 1):
        c = h if d == 0 else (3 - h[::-1]).astype(np.uint8)
        assert _refine_both(dp_sim, oracle_lib, g, c, d, 0, 0, [(100 << 4), (900 << 4) | 1, (900 << 4) | 2, (1100 << 4)]) == 2
        assert _refine_both(dp_sim, oracle_lib, g, c, d, 30, 10, [(300 << 4), (1500 << 4) | 2, (200 << 4), (1400 << 4) | 1, (500 << 4)], W=20) == 2


def _rand_ops(rng, n, first_op=None):
    """n run-length ops with alternating types (a valid CIGAR never repeats an op), optionally starting with first_op"""
    ops, prev = [], None
    for i in range(n):
        op = first_op if (i == 0 and first_op is not None) else int(rng.choice([o for o in (0, 1, 2) if o != prev]))
        ops.append((int(rng.integers(1, 300)) << 4) | op)
        prev = op
    return ops


def _cat(dst, block):
    """kswx_push_cigars (kswx.h:46-52): only the first op of an appended block may merge with the previous last op"""
    if not block:
        return
    if dst and (dst[-1] & 15) == (block[0] & 15):
        dst[-1] += block[0] & 0xFFFFFFF0
        dst.extend(block[1:])
    else:
        dst.extend(block)


@pytest.mark.parametrize("warp", [1])
def test_finish_kernels(dp_sim, warp):
    """k_finish_warp (zmo_stitch_kernels.cuh, one warp per task) against an independent restatement of the
    stitch of global_align_regs_hzmo (hzm_aln.h:1345-1486): [left extension] + region 0 + sum([gap] + region i) + [right extension],
    gap and right-extension CIGARs stored in walk order (reversed), block-wise merging at every seam, counts from the right job"""
    rng = np.random.default_rng(33 + warp)
    nt = 70
    arena, regs, jobs, tasks, tsv, exp_recs, exp_cigs, out_off = [], [], [], [], [], [], [], []

    def put(ops):
        off = len(arena)
        arena.extend(ops)
        arena.extend([0xABCDEF] * int(rng.integers(0, 3)))      # slack between segments
        return off

    def job(ops_walk_order, vals):
        jobs.append((put(ops_walk_order), vals + [len(ops_walk_order)]))
        return len(jobs) - 1

    total = 0
    for t in range(nt):
        item_off, n_item = len(regs), int(rng.integers(1, 7))
        base = [int(x) for x in rng.integers(0, 5000, 10)]          # score tb te qb qe aln mat mis ins del
        if t % 9 == 8:                                              # a task whose windows were all filtered out
            for _ in range(n_item):
                regs.append((0, 0, 0))
            tasks.append((item_off, n_item)); tsv.append([0, -1, -1, -1] + base); exp_recs.append([0] * 12); exp_cigs.append([]); out_off.append(total)
            continue
        cig, first, prev_last = [], -1, None
        left = -1
        if rng.random() < 0.6:
            ops = _rand_ops(rng, int(rng.integers(0, 40)))
            left = job(ops, [int(x) for x in rng.integers(0, 900, 7)])      # stored in alignment order (already flipped by the walk of a backward extension)
            _cat(cig, ops)
        for k in range(n_item):
            if k and rng.random() < 0.3:
                regs.append((0, 0, 0))                              # a region dropped by the per-window filter
                continue
            force = (cig[-1] & 15) if (cig and rng.random() < 0.5) else None     # make seams that merge
            rops = _rand_ops(rng, int(rng.integers(1, 120)), force)
            if first < 0:
                first = len(regs)
                regs.append((1, put(rops), len(rops)))
                _cat(cig, rops)
            else:
                gops = _rand_ops(rng, int(rng.integers(0, 30)), (cig[-1] & 15) if rng.random() < 0.5 else None)
                gid = job(gops[::-1], [int(x) for x in rng.integers(0, 900, 7)])
                regs.append((2 + gid, put(rops), len(rops)))
                _cat(cig, gops)
                _cat(cig, rops)
        rec = [1] + base + [0]
        right = -1
        if rng.random() < 0.6:
            ops = _rand_ops(rng, int(rng.integers(0, 60)), (cig[-1] & 15) if rng.random() < 0.5 else None)
            vals = [int(x) for x in rng.integers(0, 900, 7)]        # score qe te mat mis ins del
            right = job(ops[::-1], vals)
            _cat(cig, ops)
            rec[1] = vals[0]; rec[5] += vals[1]; rec[3] += vals[2]  # score replaced; qe, te advanced
            rec[6] += sum(vals[3:7]); rec[7] += vals[3]; rec[8] += vals[4]; rec[9] += vals[5]; rec[10] += vals[6]
        rec[11] = len(cig)
        tasks.append((item_off, n_item)); tsv.append([1, first, left, right] + base); exp_recs.append(rec); exp_cigs.append(cig); out_off.append(total)
        total += len(cig) + int(rng.integers(0, 4))
    ia = lambda xs: (C.c_int * max(len(xs), 1))(*xs)
    out_cig = (C.c_uint32 * (total + 8))()
    recs = (C.c_int * (12 * nt))()
    rc = dp_sim.sim_finish(warp, nt, ia([x[0] for x in tasks]), ia([x[1] for x in tasks]), ia([v for x in tsv for v in x]), len(regs), ia([x[0] for x in regs]),
                           ia([x[1] for x in regs]), ia([x[2] for x in regs]), len(jobs), ia([x[0] for x in jobs]), ia([v for x in jobs for v in x[1]]),
                           (C.c_uint32 * max(len(arena), 1))(*arena), ia(out_off), out_cig, recs)
    assert rc == 0
    for t in range(nt):
        assert list(recs[12 * t: 12 * t + 12]) == exp_recs[t], t
        assert list(out_cig[out_off[t]: out_off[t] + len(exp_cigs[t])]) == exp_cigs[t], t


@pytest.fixture(scope="module")
def align_sim():
    out = os.path.join(REPO, "tests", "_build", "libalign_host.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    subprocess.run(["g++", "-O1", "-std=c++17"] + SIM_DEFINES + ["-I" + os.path.join(REPO, "tests", "hostsim", "emu"), "-fPIC", "-shared", "-o", out,
                    os.path.join(REPO, "tests", "hostsim", "align_host.cpp")], check=True)
    return C.CDLL(out)


def _pair_align_both(sim, orc, a, b, refine=0, finish_warp=0, w=50, ew=800):
    """every strand with kept windows of the pair (a = q, b = c): the simulated alignment stage against the oracle's pair_align"""
    results = []
    wl = _windows(orc, a, b)
    qa = np.ascontiguousarray(a, np.uint8)
    cb = np.ascontiguousarray(b, np.uint8)
    for d in (0, 1):
        ws = [x for x in wl if x[0] == d]
        if not ws:
            continue
        cs = np.ascontiguousarray(b if d == 0 else (3 - b[::-1]), np.uint8)
        flat = [v for x in ws for v in x[3]]
        anc = (C.c_int * max(len(flat), 1))(*flat)
        win3 = (C.c_int * (3 * len(ws)))(*[v for x in ws for v in (x[1], x[2], len(x[3]) // 6)])
        win1 = (C.c_int * len(ws))(*[len(x[3]) // 6 for x in ws])
        cap = len(a) + len(b) + 64
        go, eo = (C.c_int * 10)(), (C.c_int * 10)()
        gc, ec = (C.c_uint32 * cap)(), (C.c_uint32 * cap)()
        stats = (C.c_int * 7)()
        gn = sim.sim_pair_align(qa.ctypes.data_as(C.c_void_p), len(qa), cb.ctypes.data_as(C.c_void_p), len(cb), d, win3, len(ws), anc,
                                w, ew, 3200, 200, C.c_float(0.6), 2, -5, -3, -1, -50, refine, finish_warp, go, gc, cap, stats)
        en = orc.orc_pair_align(qa.ctypes.data_as(C.c_void_p), len(qa), cs.ctypes.data_as(C.c_void_p), len(cs), win1, len(ws), anc,
                                w, ew, 3200, 200, C.c_float(0.6), 2, -5, -3, -1, -50, refine, eo, ec, cap)
        assert gn > -2, gn
        assert gn == en, (d, gn, en)
        if en >= 0:
            assert list(go) == list(eo), (d, list(go), list(eo))
            assert list(gc[:gn]) == list(ec[:en]), d
        results.append((en, list(stats)))
    return results


def test_alignment_stage_end_to_end(align_sim, oracle_lib):
    """zmo_pair_align's device side for real read pairs: window alignment, job planning, the six DP job lists run from sorted orders with
    executor slabs, right extensions, stitch -- and the same with -n refinement and with the opt-in warp stitch"""
    from test_seed_core import pairs
    done, jobs = 0, np.zeros(6, int)
    for i, (a, b) in enumerate(pairs(700, 8)):
        for en, stats in _pair_align_both(align_sim, oracle_lib, a, b, refine=int(i % 3 == 2), finish_warp=int(i % 2)):
            done += en >= 0
            jobs += np.array(stats[:6])
    assert done >= 5
    assert jobs[:4].sum() >= 8 and jobs[4:].sum() >= 3, jobs          # end extensions and gap fills really ran
    # a narrow end-extension band (-e 60) sends the extensions to the warp class, a wide one (-e 800) to the CTA classes
    a, b = next(pairs(701, 1))
    narrow = _pair_align_both(align_sim, oracle_lib, a, b, ew=60)
    assert any(st[0] > 0 for en, st in narrow)
