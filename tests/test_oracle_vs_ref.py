"""CPU: whole-program parity of the oracle restatement against the UNMODIFIED reference binary
(oracle/_ref/wtzmo, built from /root/reference where that exists) -- .ovl byte-for-byte, .contained
byte-for-byte, -9 pairs as a set -- and against committed golden digests (tests/golden/) for boxes
without the reference tree."""
import hashlib
import json
import os
import subprocess

import pytest

from conftest import REPO, write_long_indel_reads

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


# side inputs of wtzmo.c:1732-1769: -L tried pairs, -F excluded reads, -b clip table (applied after the length sort), -I query-only reads,
# -J minimum length, -C (only suppresses the .contained file, wtzmo.c:1781)
SIDE_GEN = ["-n", "250", "-L", "5000", "-G", "70000", "-s", "21"]
SIDE_CASES = ["side_L", "side_F", "side_b", "side_I", "side_J", "side_C"]


def _side_inputs(tmp_path, gen_reads, case):
    """writes the read set and the side file of one case; returns (fasta, extra wtzmo arguments)"""
    fa = str(tmp_path / "base.fa")
    subprocess.run([gen_reads] + SIDE_GEN + ["-o", fa], check=True)
    names = [l[1:].strip() for l in open(fa) if l.startswith(">")]
    seqs = [l.strip() for l in open(fa) if not l.startswith(">")]
    base = ["-k", "16", "-s", "200", "-m", "0.6"]
    if case == "side_L":
        with open(tmp_path / "L.pairs", "w") as f:
            for i in range(0, 60, 2):
                f.write("%s\t%s\n" % (names[i], names[i + 1]))
        return fa, base + ["-L", str(tmp_path / "L.pairs")]
    if case == "side_F":
        with open(tmp_path / "F.names", "w") as f:
            f.write("# comment\n" + "\n".join(names[5:25]) + "\n")
        return fa, base + ["-F", str(tmp_path / "F.names")]
    if case == "side_b":
        with open(tmp_path / "B.clip", "w") as f:
            for i in range(0, 250, 9):
                f.write("%s\t%d\t%d\t%d\n" % (names[i], 100, len(seqs[i]) - 300, len(seqs[i])))
        return fa, base + ["-b", str(tmp_path / "B.clip")]
    if case == "side_I":
        q = str(tmp_path / "q.fa")
        subprocess.run([gen_reads, "-n", "30", "-L", "4000", "-G", "70000", "-s", "21", "-o", q], check=True)
        return fa, base + ["-I", q]
    if case == "side_J":
        return fa, base + ["-J", "4500"]
    return fa, base + ["-C"]


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


@pytest.mark.parametrize("case", SIDE_CASES)
def test_oracle_side_inputs_match_reference_binary(case, tmp_path, gen_reads, oracle_bin, ref_bin):
    fa, extra = _side_inputs(tmp_path, gen_reads, case)
    _run(ref_bin, fa, str(tmp_path / "ref.ovl"), extra)
    _run(oracle_bin, fa, str(tmp_path / "orc.ovl"), extra)
    assert open(tmp_path / "ref.ovl", "rb").read() == open(tmp_path / "orc.ovl", "rb").read()
    assert os.path.exists(tmp_path / "ref.ovl.contained") == os.path.exists(tmp_path / "orc.ovl.contained") == (case != "side_C")
    assert _digest(str(tmp_path / "ref.ovl")) == _digest(str(tmp_path / "orc.ovl"))


@pytest.mark.parametrize("case", SIDE_CASES)
def test_oracle_side_inputs_match_golden_digest(case, tmp_path, gen_reads, oracle_bin):
    """golden digests were produced by the reference binary (tests/golden/make_golden.py)"""
    gold = json.load(open(GOLDEN))
    fa, extra = _side_inputs(tmp_path, gen_reads, case)
    _run(oracle_bin, fa, str(tmp_path / "orc.ovl"), extra)
    assert _digest(str(tmp_path / "orc.ovl")) == gold[case]


LONG_INDEL_ARGS = ["-k", "16", "-s", "200", "-m", "0.5", "-n"]


def test_oracle_refine_with_long_indel_runs(tmp_path, oracle_bin, ref_bin):
    """-n on alignments holding indel runs of 800-1,000 bases (kswx_refine_alignment widens its band by the run length, kswx.h:524-560)"""
    import re
    fa = str(tmp_path / "w.fa")
    write_long_indel_reads(fa)
    _run(ref_bin, fa, str(tmp_path / "ref.ovl"), LONG_INDEL_ARGS)
    _run(oracle_bin, fa, str(tmp_path / "orc.ovl"), LONG_INDEL_ARGS)
    ref = open(tmp_path / "ref.ovl", "rb").read()
    assert ref == open(tmp_path / "orc.ovl", "rb").read()
    assert sum(1 for l in ref.split(b"\n") if l and max(int(n) for n in re.findall(rb"(\d+)[ID]", l.split(b"\t")[16])) >= 800) >= 4
    assert _digest(str(tmp_path / "ref.ovl")) == json.load(open(GOLDEN))["refine_long_indel"]


def test_oracle_refine_with_long_indel_runs_golden(tmp_path, oracle_bin):
    fa = str(tmp_path / "w.fa")
    write_long_indel_reads(fa)
    _run(oracle_bin, fa, str(tmp_path / "orc.ovl"), LONG_INDEL_ARGS)
    assert _digest(str(tmp_path / "orc.ovl")) == json.load(open(GOLDEN))["refine_long_indel"]


@pytest.mark.parametrize("name", ["g4_2000", "cfg1_dot"])
def test_oracle_matches_scale_golden(name, tmp_path, gen_reads, oracle_bin):
    """the larger reference digests of tests/golden/scale_digests.json that the oracle reproduces in well under a minute: -G 4 with 1,370
    records (candidate carry-over between four index partitions) and BASELINE configs[0]'s reads in dot-matrix mode (11,921 records)"""
    gold = json.load(open(os.path.join(REPO, "tests", "golden", "scale_digests.json")))[name]
    fa = str(tmp_path / "r.fa")
    subprocess.run([gen_reads] + gold["gen"] + ["-o", fa], check=True)
    out = str(tmp_path / "orc.ovl")
    r = subprocess.run([oracle_bin, "-t", "1", "-i", fa, "-f", "-o", out] + gold["args"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr[-1000:]
    data = open(out, "rb").read()
    assert data.count(b"\n") == gold["lines"] and hashlib.md5(data).hexdigest() == gold["md5"]
    assert hashlib.md5(open(out + ".contained", "rb").read()).hexdigest() == gold["contained_md5"]
