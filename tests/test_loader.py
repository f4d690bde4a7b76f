"""CPU: the parallel FASTA loader of the C host (wtzmo_main.c: rs_load_parallel) against its serial reader -- same reads, same
2-bit bank, same names, and the same lrand48() & 3 values for non-ACGT bases in file order (dna.h:405) -- on awkward inputs: multi-line
records, header comments, lower case, N runs, CRLF line ends, empty records, reads below -J, no final newline, more threads than records."""
import ctypes as C
import os
import random

import pytest

from conftest import REPO


@pytest.fixture(scope="module")
def host():
    import __graft_entry__ as ge
    ge.build()
    lib = C.CDLL(os.path.join(REPO, "smartdenovo_b200", "lib", "libwtzmo_host.so"))
    lib.wz_load_digest.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.c_int, C.c_int, C.POINTER(C.c_uint64)]
    return lib


def digest(lib, files, min_len, threads):
    arr = (C.c_char_p * len(files))(*[f.encode() for f in files])
    out = (C.c_uint64 * 3)()
    assert lib.wz_load_digest(len(files), arr, min_len, threads, out) == 0
    return tuple(out)


def write_fasta(path, n, seed, width=70, crlf=False, final_newline=True):
    rng = random.Random(seed)
    nl = "\r\n" if crlf else "\n"
    with open(path, "w", newline="") as f:
        for i in range(n):
            ln = rng.choice([0, 1, 5, 31, 32, 33, 64, 500, 3000, 9000]) if i % 7 else rng.randrange(0, 12000)
            s = [rng.choice("ACGT") for _ in range(ln)]
            if i % 3 == 0:
                s = [c.lower() for c in s]
            for k in range(rng.randrange(0, 4)):
                if ln:
                    p = rng.randrange(ln)
                    s[p:p + rng.randrange(1, 40)] = "N" * min(ln - p, rng.randrange(1, 40))
            s = "".join(s)[:ln]
            f.write(">r%d some comment\tmore%s" % (i, nl))
            w = rng.choice([width, 1, 60, 100000])
            for k in range(0, len(s), w):
                f.write(s[k:k + w] + nl)
            if i % 11 == 0:
                f.write(nl)           # an empty line inside the record
        if not final_newline:
            f.write(">last\nACGTNNACGT")


@pytest.mark.parametrize("case", ["plain", "crlf", "nofinal", "tiny"])
def test_parallel_loader_equals_serial(host, tmp_path, case):
    a, b = str(tmp_path / "a.fa"), str(tmp_path / "b.fa")
    write_fasta(a, 5 if case == "tiny" else 700, 1, crlf=case == "crlf", final_newline=case != "nofinal")
    write_fasta(b, 3 if case == "tiny" else 300, 2)
    for files in ([a], [a, b]):
        for min_len in (0, 400):
            ref = digest(host, files, min_len, 1)
            assert ref[0] > 0
            for th in (2, 3, 8, 16):
                assert digest(host, files, min_len, th) == ref, (case, files, min_len, th)


def test_gzip_and_fastq_take_the_serial_reader(host, tmp_path):
    import subprocess
    a = str(tmp_path / "a.fa")
    write_fasta(a, 200, 5)
    ref = digest(host, [a], 0, 1)
    subprocess.run(["gzip", "-kf", a], check=True)
    assert digest(host, [a + ".gz"], 0, 8) == ref
    q = str(tmp_path / "q.fq")
    with open(q, "w") as f:
        for i in range(50):
            f.write("@q%d\nACGTNACGT%s\n+\n%s\n" % (i, "A" * i, "I" * (9 + i)))
    assert digest(host, [q], 0, 8) == digest(host, [q], 0, 1)
