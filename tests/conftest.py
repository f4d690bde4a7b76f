import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
ORACLE_DIR = os.path.join(REPO, "oracle")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 GPU (run with -m gpu)")


def _make(target):
    subprocess.run(["make", "-C", ORACLE_DIR, target], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


@pytest.fixture(scope="session")
def oracle_lib():
    """oracle/_ref/libzmo_oracle.so: our CPU restatement (test infrastructure)."""
    _make("_ref/libzmo_oracle.so")
    return C.CDLL(os.path.join(REF_DIR, "libzmo_oracle.so"))


@pytest.fixture(scope="session")
def ref_lib():
    """oracle/_ref/libzmo_ref.so: the real reference functions behind a shim; built where
    /root/reference exists, prebuilt copy used elsewhere (GPU box)."""
    p = os.path.join(REF_DIR, "libzmo_ref.so")
    if os.path.isdir("/root/reference"):
        _make("ref")
    if not os.path.exists(p):
        pytest.skip("reference shim not available (no /root/reference and no prebuilt oracle/_ref)")
    return C.CDLL(p)


@pytest.fixture(scope="session")
def oracle_bin():
    _make("_ref/zmo_oracle")
    return os.path.join(REF_DIR, "zmo_oracle")


@pytest.fixture(scope="session")
def ref_bin():
    p = os.path.join(REF_DIR, "wtzmo")
    if os.path.isdir("/root/reference"):
        _make("ref")
    if not os.path.exists(p):
        pytest.skip("reference binary not available")
    return p


@pytest.fixture(scope="session")
def gen_reads():
    out = os.path.join(REPO, "tools", "_build", "gen_reads")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    subprocess.run(["gcc", "-O2", "-o", out, os.path.join(REPO, "tools", "gen_reads.c"), "-lm"], check=True)
    return out


def mutate(rng, seq, ins=0.0825, dele=0.045, sub=0.0225):
    """PacBio-like noisy copy of a 0..3 numpy sequence."""
    out = []
    for b in seq:
        while rng.random() < ins:
            out.append(rng.integers(0, 4))
        u = rng.random()
        if u < dele:
            continue
        if u < dele + sub:
            out.append((int(b) + int(rng.integers(1, 4))) & 3)
        else:
            out.append(int(b))
    return np.array(out, dtype=np.uint8)


def call_ext(lib, fn_name, mode, q, t, init, W, sc=(2, -5, -3, -3, -1, -50)):
    """Run an extension DP through a shim-shaped entry point; returns (10 ints, cigar list)."""
    q = np.ascontiguousarray(q, np.uint8)
    t = np.ascontiguousarray(t, np.uint8)
    out = (C.c_int * 10)()
    cap = len(q) + len(t) + 8
    cig = (C.c_uint32 * cap)()
    M, X, I, D, E, T = sc
    fn = getattr(lib, fn_name)
    args = [len(q), q.ctypes.data_as(C.c_void_p), len(t), t.ctypes.data_as(C.c_void_p), 1, init, W, M, X, I, D, E, T, out, cig, cap]
    if mode is not None:
        args = [mode] + args
    n = fn(*args)
    return list(out), list(cig[:n])


def call_global(lib, fn_name, q, t, w, sc=(2, -5, 3, 1, 3, 1)):
    q = np.ascontiguousarray(q, np.uint8)
    t = np.ascontiguousarray(t, np.uint8)
    score = C.c_int(0)
    cap = len(q) + len(t) + 8
    cig = (C.c_uint32 * cap)()
    M, X, od, ed, oi, ei = sc
    n = getattr(lib, fn_name)(len(q), q.ctypes.data_as(C.c_void_p), len(t), t.ctypes.data_as(C.c_void_p), M, X, od, ed, oi, ei, w, C.byref(score), cig, cap)
    return score.value, list(cig[:n])


def write_long_indel_reads(path, seed=5):
    """five reads over one 12 kb genome, 4% noise; read w1 carries a 1,000-base insertion and w2 a 900-base deletion, so that alignments
    against them hold indel runs of ~800-1,000 bases (with -n: refinement bands beyond the 1,639 columns of the register executors)"""
    rng = np.random.default_rng(seed)

    def mut(s, r=0.04):
        o = []
        for ch in s:
            u = rng.random()
            if u < r / 3:
                continue
            o.append("ACGT"[rng.integers(0, 4)] if u < 2 * r / 3 else ch)
            if rng.random() < r / 3:
                o.append("ACGT"[rng.integers(0, 4)])
        return "".join(o)
    g = "".join("ACGT"[i] for i in rng.integers(0, 4, 12000))
    ins = "".join("ACGT"[i] for i in rng.integers(0, 4, 1000))
    reads = [mut(g[0:9000]), mut(g[500:4500] + ins + g[4500:9500]), mut(g[1000:5000] + g[5900:10900]), mut(g[2000:11000]), mut(g[0:8000])]
    with open(path, "w") as f:
        for i, s in enumerate(reads):
            f.write(">w%d\n%s\n" % (i, s))
