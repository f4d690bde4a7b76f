"""GPU: the stand-alone DP operators of the C ABI (zmo_dp_extend / zmo_dp_global, the same kernels
the pipeline runs) against the oracle / reference functions, bit-exact on score, end points, counts
and CIGAR.  Covers both extension modes, forward/backward/complemented views, narrow (warp) and
wide (CTA) bands, multi-chunk bands, early termination, the T end rule and degenerate shapes."""
import numpy as np
import pytest

from conftest import call_ext, call_global, mutate

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx_reads():
    from smartdenovo_b200 import Zmo
    rng = np.random.default_rng(2024)
    reads = []
    base = rng.integers(0, 4, 6000).astype(np.uint8)
    reads.append(base)
    for _ in range(5):
        reads.append(mutate(rng, base))
    reads.append(rng.integers(0, 4, 3000).astype(np.uint8))
    reads.append(np.array([1], np.uint8))
    reads.append(np.repeat(rng.integers(0, 4, 300), 3).astype(np.uint8))
    z = Zmo()
    z.upload_seqs(reads)
    yield z, reads
    z.close()


def _view(reads, rid, start, step, comp, n):
    idx = start + step * np.arange(n)
    s = reads[rid][idx]
    return (s ^ 3) if comp else s


def _problems(rng, reads, n, modes):
    from smartdenovo_b200.api import DP_PROBLEM
    probs = np.zeros(n, DP_PROBLEM)
    for i in range(n):
        qr, tr = int(rng.integers(0, 7)), int(rng.integers(0, 7))
        kind = int(rng.integers(0, 10))
        if kind == 0:
            qlen, tlen = int(rng.integers(1, 40)), int(rng.integers(1, 40))
        elif kind < 7:
            qlen, tlen = int(rng.integers(20, 300)), int(rng.integers(20, 300))
        else:
            qlen, tlen = int(rng.integers(300, 2500)), int(rng.integers(300, 2500))
        qlen = min(qlen, len(reads[qr]))
        tlen = min(tlen, len(reads[tr]))
        corr = rng.random() < 0.7       # correlated slices (same region of homologous reads)
        qstep = 1 if rng.random() < 0.6 else -1
        tstep = qstep if corr else (1 if rng.random() < 0.5 else -1)
        pos = int(rng.integers(0, 2500))
        def start(step, ln, rlen):
            lo, hi = (0, rlen - ln) if step == 1 else (ln - 1, rlen - 1)
            p = min(max(pos + int(rng.integers(-30, 30)) if corr else int(rng.integers(lo, hi + 1)), lo), hi)
            return p
        comp = int(rng.random() < 0.3)
        probs[i] = (qr, tr, start(qstep, qlen, len(reads[qr])), qstep, comp, qlen,
                    start(tstep, tlen, len(reads[tr])), tstep, comp if corr else int(rng.random() < 0.5), tlen,
                    int(rng.choice([0, 0, 37, 200, 3000, -9])), int(rng.choice(modes)))
    return probs


def _check_ext(z, reads, oracle_lib, mode, probs):
    res, cig = z.dp_extend(mode, probs)
    for i, p in enumerate(probs):
        q = _view(reads, p["q_rid"], p["q_start"], p["q_step"], p["q_comp"], p["qlen"])
        t = _view(reads, p["t_rid"], p["t_start"], p["t_step"], p["t_comp"], p["tlen"])
        exp, ecig = call_ext(oracle_lib, "orc_extend", mode, q, t, int(p["init_score"]), int(p["W"]))
        r = res[i]
        got = [int(r["score"]), 0, int(r["te"]), 0, int(r["qe"]), int(r["aln"]), int(r["mat"]), int(r["mis"]), int(r["ins"]), int(r["del"])]
        gcig = [int(x) for x in cig[int(r["cigar_off"]): int(r["cigar_off"]) + int(r["n_cigar"])]]
        assert exp == got, (i, mode, p)
        assert ecig == gcig, (i, mode, p)


def test_extend_fixed_band(ctx_reads, oracle_lib):
    z, reads = ctx_reads
    rng = np.random.default_rng(1)
    _check_ext(z, reads, oracle_lib, 0, _problems(rng, reads, 400, [50, 50, 50, 10, 3, 120, 300, -40]))


def test_extend_shifting_band(ctx_reads, oracle_lib):
    z, reads = ctx_reads
    rng = np.random.default_rng(2)
    _check_ext(z, reads, oracle_lib, 1, _problems(rng, reads, 400, [-800, -800, -100, -30, -7, -1, 50, -1000]))


def test_extend_very_wide_band_uses_global_rows(ctx_reads, oracle_lib):
    """bands wider than the shared-memory row capacity (2W+3 > 2048) fall back to global-memory rows"""
    z, reads = ctx_reads
    rng = np.random.default_rng(3)
    probs = _problems(rng, reads, 12, [-1500, -2500])
    probs["qlen"] = np.minimum(np.maximum(probs["qlen"], 1200), 2500)
    probs["tlen"] = 2900
    probs["q_start"] = np.where(probs["q_step"] == 1, 100, 2900)
    probs["t_start"] = np.where(probs["t_step"] == 1, 50, 2950)
    probs["q_rid"] = 1
    probs["t_rid"] = 2
    _check_ext(z, reads, oracle_lib, 1, probs)


def test_extend_degenerate(ctx_reads, oracle_lib):
    from smartdenovo_b200.api import DP_PROBLEM
    z, reads = ctx_reads
    probs = np.zeros(6, DP_PROBLEM)
    probs[0] = (0, 1, 0, 1, 0, 0, 0, 1, 0, 10, 5, 50)          # qlen 0
    probs[1] = (0, 1, 0, 1, 0, 10, 0, 1, 0, 0, -3, 50)         # tlen 0, negative init clamps to 0
    probs[2] = (7, 7, 0, 1, 0, 1, 0, 1, 0, 1, 0, -800)         # 1x1
    probs[3] = (7, 0, 0, 1, 0, 1, 10, 1, 0, 500, 0, -800)      # 1 x 500
    probs[4] = (0, 7, 10, 1, 0, 500, 0, 1, 0, 1, 100, -800)    # 500 x 1
    probs[5] = (8, 8, 0, 1, 0, 900, 3, 1, 0, 890, 0, 50)       # low-complexity repeats
    for mode in (0, 1):
        _check_ext(z, reads, oracle_lib, mode, probs)


def test_global(ctx_reads, oracle_lib):
    from smartdenovo_b200.api import DP_PROBLEM
    z, reads = ctx_reads
    rng = np.random.default_rng(4)
    probs = _problems(rng, reads, 300, [0])
    w = rng.choice([50, 50, 100, 7, 400], size=len(probs)).astype(np.int32)
    # degenerate shapes: empty query / empty target / both
    extra = np.zeros(3, DP_PROBLEM)
    extra[0] = (0, 1, 5, 1, 0, 0, 7, 1, 0, 30, 0, 0)
    extra[1] = (0, 1, 5, 1, 0, 25, 7, 1, 0, 0, 0, 0)
    extra[2] = (0, 1, 5, 1, 0, 0, 7, 1, 0, 0, 0, 0)
    probs = np.concatenate([probs, extra])
    w = np.concatenate([w, np.array([50, 50, 50], np.int32)])
    res, cig = z.dp_global(probs, w)
    for i, p in enumerate(probs):
        q = _view(reads, p["q_rid"], p["q_start"], p["q_step"], p["q_comp"], p["qlen"])
        t = _view(reads, p["t_rid"], p["t_start"], p["t_step"], p["t_comp"], p["tlen"])
        ww = int(w[i])
        while ww < abs(len(q) - len(t)):
            ww <<= 1
        escore, ecig = call_global(oracle_lib, "orc_global2", q, t, ww)
        r = res[i]
        gcig = [int(x) for x in cig[int(r["cigar_off"]): int(r["cigar_off"]) + int(r["n_cigar"])]]
        assert escore == int(r["score"]), (i, p, ww)
        assert ecig == gcig, (i, p, ww)
        # counts re-derived from the CIGAR must match what the kernel reports (hzm_aln.h:1423-1432)
        x1 = x2 = mat = mis = ins = dele = 0
        for op in ecig:
            ln, o = op >> 4, op & 15
            if o == 0:
                mat += int((q[x1:x1 + ln] == t[x2:x2 + ln]).sum()); mis += ln - int((q[x1:x1 + ln] == t[x2:x2 + ln]).sum()); x1 += ln; x2 += ln
            elif o == 1:
                ins += ln; x1 += ln
            else:
                dele += ln; x2 += ln
        assert (mat, mis, ins, dele) == (int(r["mat"]), int(r["mis"]), int(r["ins"]), int(r["del"])), (i, p)
