"""CPU: the k-mer index and candidate-query KERNELS (smartdenovo_b200/csrc/zmo_index_kernels.cuh: hp-k-mer scans with sub-sampling,
run-length encode of the sorted k-mers, saturated counts and the automatic K, filter flags, posting gather; chunk-parallel query scan,
index look-up, posting expansion with the self / length filters, union length per (target, strand) with the uint32 wrap rule, event
emission) compiled for the host by tests/hostsim (test-only; std:: sorts / scans where zmo_index.cu calls CUB) against the oracle's
restatement of index_wtzmo / query_wtzmo (wtzmo.c:227-562): K, index size and every query's (target<<1|strand, ol) event stream."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import REPO, mutate


@pytest.fixture(scope="module")
def index_sim():
    out = os.path.join(REPO, "tests", "_build", "libindex_host.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    subprocess.run(["g++", "-O1", "-std=c++17", "-I" + os.path.join(REPO, "tests", "hostsim", "emu"), "-fPIC", "-shared", "-o", out,
                    os.path.join(REPO, "tests", "hostsim", "index_host.cpp")], check=True)
    return C.CDLL(out)


def read_set(seed, n=40, L=2500, G=14000):
    rng = np.random.default_rng(seed)
    g = rng.integers(0, 4, G).astype(np.uint8)
    g[3000:3600] = g[9000:9600]                      # a repeat: k-mers with doubled counts
    g[5000:5200] = np.tile(rng.integers(0, 4, 5).astype(np.uint8), 40)      # low complexity: homopolymer compression, high counts
    reads = []
    for i in range(n):
        s = int(rng.integers(0, G - L))
        r = mutate(rng, g[s:s + int(rng.integers(L // 2, L))], ins=0.03, dele=0.02, sub=0.01)
        if rng.random() < 0.5:
            r = (3 - r[::-1]).astype(np.uint8)
        reads.append(r)
    reads.append(rng.integers(0, 4, 9).astype(np.uint8))          # shorter than k
    reads.append(np.zeros(400, np.uint8))                          # one homopolymer
    return reads


def both(sim, orc, reads, beg, end, qids, ksize=16, hk=1, ksave=4, kovl=300, kcut=0):
    seqs = np.ascontiguousarray(np.concatenate(reads), np.uint8)
    lens = (C.c_int * len(reads))(*[len(r) for r in reads])
    cap = 1 << 16
    k1 = C.c_uint32(kcut)
    st1 = (C.c_ulonglong * 2)()
    evoff = (C.c_ulonglong * (len(qids) + 1))()
    ev1 = (C.c_uint32 * (2 * cap))()
    q = (C.c_int * len(qids))(*qids)
    n1 = sim.sim_index_candidates(seqs.ctypes.data_as(C.c_void_p), lens, len(reads), beg, end, q, len(qids), ksize, hk, ksave, kovl, C.byref(k1), st1, evoff, ev1, cap)
    assert 0 <= n1 <= cap and evoff[len(qids)] == n1
    total = 0
    for i, qid in enumerate(qids):
        k2 = C.c_uint32(kcut)
        st2 = (C.c_ulonglong * 2)()
        ev2 = (C.c_uint32 * (2 * cap))()
        n2 = orc.orc_candidates(seqs.ctypes.data_as(C.c_void_p), lens, len(reads), beg, end, qid, ksize, hk, ksave, kovl, C.byref(k2), st2, ev2, cap)
        assert (k1.value, list(st1)) == (k2.value, list(st2)), (beg, end, ksize, ksave)
        got = list(ev1[2 * evoff[i]: 2 * evoff[i + 1]])
        assert got == list(ev2[: 2 * n2]), (qid, beg, end, ksize, ksave, kovl)
        total += n2
    return total


def test_index_and_candidates_match_oracle(index_sim, oracle_lib):
    reads = read_set(1)
    n = len(reads)
    qids = list(range(0, n, 3)) + [n - 2, n - 1]
    total = both(index_sim, oracle_lib, reads, 0, n, qids)
    assert total > 40
    total += both(index_sim, oracle_lib, reads, 0, n, qids, ksize=13, ksave=1, kovl=100)
    total += both(index_sim, oracle_lib, reads, 0, n, qids, ksize=16, hk=0, ksave=2, kovl=150)
    total += both(index_sim, oracle_lib, reads, 0, n, qids, ksize=20, ksave=1, kovl=100, kcut=3)      # explicit -K 3: most shared k-mers are filtered as too frequent
    # a range-partitioned index (-G): reads [10, 30) only, queried by reads inside and outside the range
    total += both(index_sim, oracle_lib, reads, 10, 30, [0, 5, 12, 20, 29, 35], ksave=1, kovl=100)
    # partition without any k-mer: the two degenerate reads at the end
    assert both(index_sim, oracle_lib, reads, n - 2, n - 1, [0, n - 1], ksave=1) == 0
    assert total > 300
