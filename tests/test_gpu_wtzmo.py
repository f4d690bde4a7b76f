"""GPU: whole-program parity.  The product binary smartdenovo_b200/bin/wtzmo (C host + libzmo_b200.so)
against the unmodified reference binary (oracle/_ref/wtzmo, when prebuilt) or the CPU oracle, on the
same synthetic reads: .ovl byte-for-byte (all 17 columns incl. CIGAR), .contained byte-for-byte, -9 pair
file as a set."""
import os
import subprocess

import pytest

from conftest import REF_DIR, REPO

pytestmark = pytest.mark.gpu
EXE = os.path.join(REPO, "smartdenovo_b200", "bin", "wtzmo")


def _checker(oracle_bin):
    ref = os.path.join(REF_DIR, "wtzmo")
    return ref if os.path.exists(ref) else oracle_bin


def _run(exe, fa, out, extra, env=None):
    cmd = [exe, "-t", "1", "-i", fa, "-f", "-o", out, "-9", out + ".pairs"] + extra
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stderr


def _compare(tmp_path, gen_reads, oracle_bin, gen_args, extra, env=None):
    fa = str(tmp_path / "reads.fa")
    subprocess.run([gen_reads] + gen_args + ["-o", fa], check=True)
    _run(_checker(oracle_bin), fa, str(tmp_path / "ref.ovl"), extra)
    log = _run(EXE, fa, str(tmp_path / "gpu.ovl"), extra, env=env)
    ref = open(tmp_path / "ref.ovl", "rb").read()
    got = open(tmp_path / "gpu.ovl", "rb").read()
    if ref != got:
        rl, gl = ref.split(b"\n"), got.split(b"\n")
        first = next((i for i, (x, y) in enumerate(zip(rl, gl)) if x != y), min(len(rl), len(gl)))
        raise AssertionError("ovl differs: %d vs %d lines, first difference at line %d\nref: %s\ngpu: %s\n%s" % (
            len(rl), len(gl), first, rl[first][:300] if first < len(rl) else b"<eof>", gl[first][:300] if first < len(gl) else b"<eof>", log[-1500:]))
    assert len(ref) > 0
    if os.path.exists(tmp_path / "ref.ovl.contained"):
        assert open(tmp_path / "ref.ovl.contained", "rb").read() == open(tmp_path / "gpu.ovl.contained", "rb").read()
    assert sorted(open(tmp_path / "ref.ovl.pairs").read().split("\n")) == sorted(open(tmp_path / "gpu.ovl.pairs").read().split("\n"))
    return ref.count(b"\n")


def test_sw_small(tmp_path, gen_reads, oracle_bin):
    n = _compare(tmp_path, gen_reads, oracle_bin, ["-n", "200", "-L", "6000", "-G", "60000", "-s", "1"], ["-k", "16", "-s", "200", "-m", "0.6"])
    assert n > 100


def test_sw_small_batches(tmp_path, gen_reads, oracle_bin):
    """batch size 1 (== the reference's read-by-read order) and odd batch sizes give the same bytes"""
    for br in ("1", "7"):
        env = dict(os.environ, ZMO_BATCH_READS=br)
        _compare(tmp_path, gen_reads, oracle_bin, ["-n", "120", "-L", "5000", "-G", "50000", "-s", "3"], ["-k", "16", "-s", "200", "-m", "0.6"], env=env)


def test_sw_ont_repeats(tmp_path, gen_reads, oracle_bin):
    n = _compare(tmp_path, gen_reads, oracle_bin, ["-n", "400", "-L", "5000", "-G", "80000", "-s", "7", "-m", "ont"], ["-k", "16", "-s", "200", "-m", "0.6"])
    assert n > 300


def test_sw_job_shard_and_partitioned_index(tmp_path, gen_reads, oracle_bin):
    _compare(tmp_path, gen_reads, oracle_bin, ["-n", "300", "-L", "5000", "-G", "70000", "-s", "11"], ["-k", "16", "-P", "2", "-p", "1"])
    _compare(tmp_path, gen_reads, oracle_bin, ["-n", "300", "-L", "5000", "-G", "70000", "-s", "11"], ["-k", "16", "-G", "2"])


def test_seed_only_mode(tmp_path, gen_reads, oracle_bin):
    _compare(tmp_path, gen_reads, oracle_bin, ["-n", "200", "-L", "5000", "-G", "60000", "-s", "5"], ["-N", "-k", "16"])


def test_nondefault_parameters(tmp_path, gen_reads, oracle_bin):
    _compare(tmp_path, gen_reads, oracle_bin, ["-n", "200", "-L", "5000", "-G", "60000", "-s", "9"],
             ["-k", "15", "-S", "2", "-z", "12", "-Z", "32", "-y", "600", "-R", "150", "-r", "250", "-w", "30", "-e", "300", "-W", "800", "-m", "0.55", "-s", "150", "-A", "50", "-B", "20"])
