"""GPU: stage-level entry points of the C ABI on a small read set -- device-formatted CIGAR text
(zmo_pair_align_text) against the binary ops of zmo_pair_align formatted the way kswx_cigar2string does
(kswx.h:1093-1120), and context clones (zmo_ctx_clone) against the root context."""
import numpy as np
import pytest

from conftest import mutate

pytestmark = pytest.mark.gpu


def _fmt(ops):
    return "".join("%d%s" % (int(o) >> 4, "MID"[int(o) & 3]) for o in ops if int(o) >> 4)


@pytest.fixture(scope="module")
def staged():
    from smartdenovo_b200 import Zmo
    from smartdenovo_b200.api import PAIR, TASK
    rng = np.random.default_rng(11)
    genome = rng.integers(0, 4, 50000).astype(np.uint8)
    reads = []
    for _ in range(100):
        s = int(rng.integers(0, 50000 - 9000))
        r = mutate(rng, genome[s:s + int(rng.integers(6000, 9000))])
        reads.append((3 - r[::-1]).astype(np.uint8) if rng.random() < 0.5 else r)
    reads.sort(key=len, reverse=True)
    z = Zmo()
    z.upload_seqs(reads)
    z.index_build()
    off, ev = z.candidates(np.arange(len(reads), dtype=np.uint32))
    pairs = []
    for q in range(len(reads)):
        for e in ev[int(off[q]):int(off[q + 1])]:
            c = int(e["tkey"]) >> 1
            if c != q and (q, c) not in pairs:
                pairs.append((q, c))
    pairs = np.array(pairs, PAIR)
    seeds, wins = z.pair_windows(pairs)
    tasks = []
    for i, s in enumerate(seeds):
        for d in (0, 1):
            if s["n_win"][d]:
                tasks.append((i, d))
    tasks = np.array(tasks, TASK)
    assert len(tasks) > 5, (len(pairs), len(tasks))
    yield z, pairs, seeds, wins, tasks
    z.close()


def test_cigar_text_matches_binary(staged):
    z, pairs, seeds, wins, tasks = staged
    rb, ops = z.pair_align(tasks)
    rt, txt = z.pair_align_text(tasks, text_cap=64)      # forces the capacity retry path too
    n_ok = 0
    for a, b in zip(rb, rt):
        for f in ("ok", "score", "tb", "te", "qb", "qe", "aln", "mat", "mis", "ins", "del"):
            assert a[f] == b[f]
        if not a["ok"]:
            assert b["n_cigar"] == 0
            continue
        want = _fmt(ops[int(a["cigar_off"]):int(a["cigar_off"]) + int(a["n_cigar"])])
        got = bytes(txt[int(b["cigar_off"]):int(b["cigar_off"]) + int(b["n_cigar"])]).decode()
        assert got == want
        n_ok += 1
    assert n_ok > 3


def test_clone_shares_reads_and_index(staged):
    z, pairs, seeds, wins, tasks = staged
    c = z.clone()
    try:
        off0, ev0 = z.candidates(np.arange(8, dtype=np.uint32))
        off1, ev1 = c.candidates(np.arange(8, dtype=np.uint32))
        assert np.array_equal(off0, off1) and np.array_equal(ev0, ev1)
        s1, w1 = c.pair_windows(pairs)
        # kept windows are bump-allocated on the device: win_off (arena order) is not reproducible, the windows of a pair are
        for f in ("n_zpair", "ovl", "n_win"):
            assert np.array_equal(s1[f], seeds[f])
        assert len(w1) == len(wins)
        for a, b in zip(seeds, s1):
            for d in (0, 1):
                n = int(a["n_win"][d])
                assert np.array_equal(wins[int(a["win_off"][d]):int(a["win_off"][d]) + n], w1[int(b["win_off"][d]):int(b["win_off"][d]) + n])
        r0, t0 = z.pair_align_text(tasks)
        r1, t1 = c.pair_align_text(tasks)
        for f in r0.dtype.names:
            if f != "cigar_off" and not f.startswith("_"):
                assert np.array_equal(r0[f], r1[f]), f
        for a, b in zip(r0, r1):
            assert bytes(t0[int(a["cigar_off"]):int(a["cigar_off"]) + int(a["n_cigar"])]) == bytes(t1[int(b["cigar_off"]):int(b["cigar_off"]) + int(b["n_cigar"])])
        from smartdenovo_b200.api import ZmoError
        with pytest.raises(ZmoError):
            c.index_build()
    finally:
        c.close()
