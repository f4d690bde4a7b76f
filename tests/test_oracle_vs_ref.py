"""CPU: whole-program parity of the oracle restatement against the UNMODIFIED reference binary
(oracle/_ref/wtzmo, built from /root/reference where that exists) -- .ovl byte-for-byte, .contained
byte-for-byte, -9 pairs as a set -- and against committed golden digests (tests/golden/) for boxes
without the reference tree."""
import hashlib
import json
import os
import subprocess

import pytest

from conftest import REPO

GOLDEN = os.path.join(REPO, "tests", "golden", "ovl_digests.json")
CASES = {
    "sw": (["-n", "150", "-L", "5000", "-G", "50000", "-s", "3"], ["-k", "16", "-s", "200", "-m", "0.6"]),
    "ont": (["-n", "200", "-L", "4000", "-G", "60000", "-s", "7", "-m", "ont"], ["-k", "16"]),
    "dot": (["-n", "200", "-L", "4000", "-G", "60000", "-s", "7", "-m", "ont"], ["-k", "16", "-z", "10", "-Z", "16", "-U", "-1", "-m", "0.1", "-A", "1000"]),
    "shard": (["-n", "200", "-L", "4000", "-G", "60000", "-s", "9"], ["-k", "16", "-P", "2", "-p", "1"]),
    "gpart": (["-n", "200", "-L", "4000", "-G", "60000", "-s", "9"], ["-k", "16", "-G", "2"]),
    "seeds": (["-n", "150", "-L", "4000", "-G", "50000", "-s", "5"], ["-N", "-k", "16"]),
    "refine": (["-n", "150", "-L", "5000", "-G", "50000", "-s", "3"], ["-k", "16", "-s", "200", "-m", "0.6", "-n"]),
    "refine_ont": (["-n", "150", "-L", "4000", "-G", "60000", "-s", "17", "-m", "ont"], ["-k", "16", "-n", "-w", "20"]),
    "dot_seeds_only": (["-n", "200", "-L", "4000", "-G", "60000", "-s", "7", "-m", "ont"], ["-k", "16", "-z", "10", "-Z", "16", "-U", "-1", "-m", "0.1", "-N"]),
    "params": (["-n", "150", "-L", "4000", "-G", "50000", "-s", "11"], ["-k", "15", "-S", "2", "-z", "12", "-Z", "32", "-y", "600", "-R", "150", "-r", "250", "-w", "30", "-e", "300", "-W", "800", "-m", "0.55", "-s", "150", "-A", "50", "-B", "20"]),
}


def _run(exe, fa, out, extra):
    r = subprocess.run([exe, "-t", "1", "-i", fa, "-f", "-o", out, "-9", out + ".pairs"] + extra, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr[-1000:]


def _digest(out):
    d = {"ovl": hashlib.sha256(open(out, "rb").read()).hexdigest(), "lines": sum(1 for _ in open(out))}
    if os.path.exists(out + ".contained"):
        d["contained"] = hashlib.sha256(open(out + ".contained", "rb").read()).hexdigest()
    d["pairs"] = hashlib.sha256("\n".join(sorted(open(out + ".pairs").read().split("\n"))).encode()).hexdigest()
    return d


@pytest.mark.parametrize("case", sorted(CASES))
def test_oracle_matches_reference_binary(case, tmp_path, gen_reads, oracle_bin, ref_bin):
    gen_args, extra = CASES[case]
    fa = str(tmp_path / "r.fa")
    subprocess.run([gen_reads] + gen_args + ["-o", fa], check=True)
    _run(ref_bin, fa, str(tmp_path / "ref.ovl"), extra)
    _run(oracle_bin, fa, str(tmp_path / "orc.ovl"), extra)
    assert open(tmp_path / "ref.ovl", "rb").read() == open(tmp_path / "orc.ovl", "rb").read()
    assert _digest(str(tmp_path / "ref.ovl")) == _digest(str(tmp_path / "orc.ovl"))


@pytest.mark.parametrize("case", sorted(CASES))
def test_oracle_matches_golden_digest(case, tmp_path, gen_reads, oracle_bin):
    """golden digests were produced by the reference binary (tests/golden/make_golden.py)"""
    gold = json.load(open(GOLDEN))
    gen_args, extra = CASES[case]
    fa = str(tmp_path / "r.fa")
    subprocess.run([gen_reads] + gen_args + ["-o", fa], check=True)
    _run(oracle_bin, fa, str(tmp_path / "orc.ovl"), extra)
    assert _digest(str(tmp_path / "orc.ovl")) == gold[case]
