"""GPU: whole-program parity.  The product binary smartdenovo_b200/bin/wtzmo (C host + libzmo_b200.so)
against the unmodified reference binary (oracle/_ref/wtzmo, when prebuilt) or the CPU oracle, on the
same synthetic reads: .ovl byte-for-byte (all 17 columns incl. CIGAR), .contained byte-for-byte, -9 pair
file as a set."""
import os
import subprocess

import pytest

from conftest import REF_DIR, REPO, write_long_indel_reads

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


def test_ovl_16_columns_mode(tmp_path, gen_reads, oracle_bin):
    """ZMO_OVL_COLS=16: the file the pipeline makes with `wtzmo ... -fo - | cut -f1-16` (smartdenovo.pl:58), written directly -- records only
    through zmo_pair_align_records, no CIGAR text formatted or copied; with and without -n, and in dot-matrix mode (where column 17 is `0M`)"""
    fa = str(tmp_path / "reads.fa")
    subprocess.run([gen_reads, "-n", "200", "-L", "6000", "-G", "60000", "-s", "1", "-o", fa], check=True)
    for extra in (["-k", "16", "-s", "200", "-m", "0.6"], ["-k", "16", "-s", "200", "-m", "0.6", "-n"], ["-k", "16", "-z", "10", "-Z", "16", "-U", "-1", "-m", "0.1", "-A", "1000"]):
        _run(_checker(oracle_bin), fa, str(tmp_path / "ref.ovl"), extra)
        _run(EXE, fa, str(tmp_path / "g16.ovl"), extra, env=dict(os.environ, ZMO_OVL_COLS="16"))
        want = b"".join(b"\t".join(l.split(b"\t")[:16]) + b"\n" for l in open(tmp_path / "ref.ovl", "rb").read().splitlines())
        got = open(tmp_path / "g16.ovl", "rb").read()
        assert len(want) > 0 and got == want, extra
        assert open(tmp_path / "ref.ovl.contained", "rb").read() == open(tmp_path / "g16.ovl.contained", "rb").read()


def test_sw_small_batches(tmp_path, gen_reads, oracle_bin):
    """batch size 1 (== the reference's read-by-read order) and odd batch sizes give the same bytes"""
    for br in ("1", "7"):
        env = dict(os.environ, ZMO_BATCH_READS=br)
        _compare(tmp_path, gen_reads, oracle_bin, ["-n", "120", "-L", "5000", "-G", "50000", "-s", "3"], ["-k", "16", "-s", "200", "-m", "0.6"], env=env)


def test_refine_n_long_indel_runs(tmp_path, oracle_bin):
    """-n with refinement bands beyond the register executors (indel runs of 800-1,000 bases): k_refine_wide / band_refine, bit-exact in
    the host simulation (tests/test_dp_hostsim.py::test_refine_kernels); before this fallback existed such a run was rejected with an error"""
    fa = str(tmp_path / "w.fa")
    write_long_indel_reads(fa)
    extra = ["-k", "16", "-s", "200", "-m", "0.5", "-n"]
    _run(_checker(oracle_bin), fa, str(tmp_path / "ref.ovl"), extra)
    _run(EXE, fa, str(tmp_path / "gpu.ovl"), extra)
    ref = open(tmp_path / "ref.ovl", "rb").read()
    assert ref == open(tmp_path / "gpu.ovl", "rb").read() and ref.count(b"\n") >= 8


def test_overfull_batches_are_split(tmp_path, gen_reads, oracle_bin):
    """a batch whose pair list exceeds the per-call limit hands its tail back to the read cursor (SW and dot-matrix mode)"""
    env = dict(os.environ, ZMO_CALL_PAIRS="300")
    _compare(tmp_path, gen_reads, oracle_bin, ["-n", "200", "-L", "5000", "-G", "50000", "-s", "29"], ["-k", "16", "-s", "200", "-m", "0.6"], env=env)
    _compare(tmp_path, gen_reads, oracle_bin, ["-n", "200", "-L", "4000", "-G", "60000", "-s", "7", "-m", "ont"], ["-k", "16", "-z", "10", "-Z", "16", "-U", "-1", "-m", "0.1", "-A", "1000"], env=env)


def test_refine_n(tmp_path, gen_reads, oracle_bin):
    """-n: kswx_refine_alignment after the stitch (wtzmo.c:1031-1034), PacBio-like and ONT-like error models, narrow -w too"""
    n = _compare(tmp_path, gen_reads, oracle_bin, ["-n", "200", "-L", "6000", "-G", "60000", "-s", "1"], ["-k", "16", "-s", "200", "-m", "0.6", "-n"])
    assert n > 100
    _compare(tmp_path, gen_reads, oracle_bin, ["-n", "150", "-L", "4000", "-G", "60000", "-s", "17", "-m", "ont"], ["-k", "16", "-n", "-w", "20"])
    env = dict(os.environ, ZMO_BATCH_READS="5")
    _compare(tmp_path, gen_reads, oracle_bin, ["-n", "100", "-L", "5000", "-G", "50000", "-s", "23"], ["-k", "16", "-s", "100", "-m", "0.5", "-n", "-w", "120"], env=env)


def test_sw_ont_repeats(tmp_path, gen_reads, oracle_bin):
    n = _compare(tmp_path, gen_reads, oracle_bin, ["-n", "400", "-L", "5000", "-G", "80000", "-s", "7", "-m", "ont"], ["-k", "16", "-s", "200", "-m", "0.6"])
    assert n > 300


def test_sw_job_shard_and_partitioned_index(tmp_path, gen_reads, oracle_bin):
    _compare(tmp_path, gen_reads, oracle_bin, ["-n", "300", "-L", "5000", "-G", "70000", "-s", "11"], ["-k", "16", "-P", "2", "-p", "1"])
    _compare(tmp_path, gen_reads, oracle_bin, ["-n", "300", "-L", "5000", "-G", "70000", "-s", "11"], ["-k", "16", "-G", "2"])


def test_dot_matrix_with_seed_only_flag_prints_nothing(tmp_path, gen_reads, oracle_bin):
    """-U with -N: the reference prints hits only in step with the (empty) seed list (wtzmo.c:1176-1186)"""
    fa = str(tmp_path / "reads.fa")
    subprocess.run([gen_reads, "-n", "200", "-L", "4000", "-G", "60000", "-s", "7", "-m", "ont", "-o", fa], check=True)
    extra = ["-k", "16", "-z", "10", "-Z", "16", "-U", "-1", "-m", "0.1", "-N"]
    _run(_checker(oracle_bin), fa, str(tmp_path / "ref.ovl"), extra)
    _run(EXE, fa, str(tmp_path / "gpu.ovl"), extra)
    assert open(tmp_path / "ref.ovl", "rb").read() == open(tmp_path / "gpu.ovl", "rb").read() == b""


def test_seed_only_mode(tmp_path, gen_reads, oracle_bin):
    _compare(tmp_path, gen_reads, oracle_bin, ["-n", "200", "-L", "5000", "-G", "60000", "-s", "5"], ["-N", "-k", "16"])


def test_nondefault_parameters(tmp_path, gen_reads, oracle_bin):
    _compare(tmp_path, gen_reads, oracle_bin, ["-n", "200", "-L", "5000", "-G", "60000", "-s", "9"],
             ["-k", "15", "-S", "2", "-z", "12", "-Z", "32", "-y", "600", "-R", "150", "-r", "250", "-w", "30", "-e", "300", "-W", "800", "-m", "0.55", "-s", "150", "-A", "50", "-B", "20"])


def test_dot_matrix_mode(tmp_path, gen_reads, oracle_bin):
    """smartdenovo.pl:48 / run_dmo.sh flags: -z 10 -Z 16 -U -1 -m 0.1 -A 1000 (records end with 0M)"""
    n = _compare(tmp_path, gen_reads, oracle_bin, ["-n", "400", "-L", "5000", "-G", "80000", "-s", "7", "-m", "ont"], ["-k", "16", "-z", "10", "-Z", "16", "-U", "-1", "-m", "0.1", "-A", "1000"])
    assert n > 300
    _compare(tmp_path, gen_reads, oracle_bin, ["-n", "200", "-L", "6000", "-G", "60000", "-s", "1"], ["-k", "16", "-U", "256", "-U", "32", "-U", "300", "-U", "0.5", "-U", "0.02", "-m", "0.1"])


def test_side_inputs(tmp_path, gen_reads, oracle_bin):
    """-L tried pairs, -F excluded reads, -b clipping, -I query-only reads, -J minimum length"""
    fa = str(tmp_path / "base.fa")
    subprocess.run([gen_reads, "-n", "250", "-L", "5000", "-G", "70000", "-s", "21", "-o", fa], check=True)
    names = [l[1:].strip() for l in open(fa) if l.startswith(">")]
    seqs = [l.strip() for l in open(fa) if not l.startswith(">")]
    with open(tmp_path / "L.pairs", "w") as f:
        for i in range(0, 60, 2):
            f.write("%s\t%s\n" % (names[i], names[i + 1]))
    with open(tmp_path / "F.names", "w") as f:
        f.write("# comment\n" + "\n".join(names[5:25]) + "\n")
    with open(tmp_path / "B.clip", "w") as f:
        for i in range(0, 250, 9):
            f.write("%s\t%d\t%d\t%d\n" % (names[i], 100, len(seqs[i]) - 300, len(seqs[i])))
    q = str(tmp_path / "q.fa")
    subprocess.run([gen_reads, "-n", "30", "-L", "4000", "-G", "70000", "-s", "21", "-o", q], check=True)
    base = ["-k", "16", "-s", "200", "-m", "0.6"]
    for extra in (["-L", str(tmp_path / "L.pairs")], ["-F", str(tmp_path / "F.names")], ["-b", str(tmp_path / "B.clip")], ["-I", q], ["-J", "4500"], ["-C"]):
        for f in ("ref.ovl", "gpu.ovl", "ref.ovl.contained", "gpu.ovl.contained"):
            if os.path.exists(tmp_path / f):
                os.remove(tmp_path / f)
        _run(_checker(oracle_bin), fa, str(tmp_path / "ref.ovl"), base + extra)
        _run(EXE, fa, str(tmp_path / "gpu.ovl"), base + extra)
        assert open(tmp_path / "ref.ovl", "rb").read() == open(tmp_path / "gpu.ovl", "rb").read(), extra
        assert os.path.exists(tmp_path / "ref.ovl.contained") == os.path.exists(tmp_path / "gpu.ovl.contained"), extra
        if os.path.exists(tmp_path / "ref.ovl.contained"):
            assert open(tmp_path / "ref.ovl.contained", "rb").read() == open(tmp_path / "gpu.ovl.contained", "rb").read(), extra
        assert sorted(open(tmp_path / "ref.ovl.pairs").read().split("\n")) == sorted(open(tmp_path / "gpu.ovl.pairs").read().split("\n")), extra


def test_edge_reads(tmp_path, oracle_bin):
    """reads shorter than k / z, equal-length reads (sort ties), multi-line FASTA with header comments, FASTQ input"""
    import random
    random.seed(5)
    g = "".join(random.choice("ACGT") for _ in range(30000))

    def mut(s):
        o = []
        for ch in s:
            r = random.random()
            if r < 0.05:
                continue
            o.append(random.choice("ACGT") if r < 0.08 else ch)
            if random.random() < 0.06:
                o.append(random.choice("ACGT"))
        return "".join(o)
    recs = []
    for i in range(150):
        st = random.randrange(0, 27000)
        s = mut(g[st:st + 3000])[:random.choice([2000, 2500, 2500, 2500, 3000])]
        if random.random() < 0.5:
            s = s[::-1].translate(str.maketrans("ACGT", "TGCA"))
        recs.append(s)
    for L in (3, 9, 10, 15, 16, 17, 40):
        recs.append(g[100:100 + L])
    fa = tmp_path / "ties.fa"
    with open(fa, "w") as f:
        for n, s in enumerate(recs):
            f.write(">t%d some comment\n" % n)
            for k in range(0, len(s), 70):
                f.write(s[k:k + 70] + "\n")
    fq = tmp_path / "ties.fq"
    with open(fq, "w") as f:
        for n, s in enumerate(recs):
            f.write("@t%d x\n%s\n+\n%s\n" % (n, s, "I" * len(s)))
    # lower case, non-ACGT bases (lrand48() & 3 in file order, dna.h:405) and gzip input (file_reader.c:66-72)
    fn = tmp_path / "mixed.fa"
    with open(fn, "w") as f:
        for n, s in enumerate(recs):
            t = list(s.lower() if n % 3 == 0 else s)
            for k in range(7, len(t), 401):
                t[k] = "N" if n % 2 else "n"
            f.write(">t%d\n%s\n" % (n, "".join(t)))
    subprocess.run(["gzip", "-kf", str(fn)], check=True)
    extra = ["-k", "16", "-s", "100", "-m", "0.5", "-r", "200", "-R", "100", "-d", "100"]
    for src in (fa, fq, fn, str(fn) + ".gz"):
        _run(_checker(oracle_bin), str(src), str(tmp_path / "ref.ovl"), extra)
        _run(EXE, str(src), str(tmp_path / "gpu.ovl"), extra)
        ref = open(tmp_path / "ref.ovl", "rb").read()
        assert ref == open(tmp_path / "gpu.ovl", "rb").read() and ref.count(b"\n") > 200
        assert open(tmp_path / "ref.ovl.contained", "rb").read() == open(tmp_path / "gpu.ovl.contained", "rb").read()


def _golden_run(tmp_path, gen_reads, name, env=None, keep=None):
    """product binary on the seeded read set of tests/golden/scale_digests.json[name] (digests of the UNMODIFIED reference `-t 1`, made by
    tests/golden/make_scale_golden.py where /root/reference exists): .ovl and .contained must be md5-identical"""
    import hashlib
    import json
    gold = json.load(open(os.path.join(REPO, "tests", "golden", "scale_digests.json")))[name]
    base = "/dev/shm" if os.path.isdir("/dev/shm") else str(tmp_path)
    fa = os.path.join(base, "zmo_gold_%s.fa" % hashlib.md5(" ".join(gold["gen"]).encode()).hexdigest()[:10])
    out = os.path.join(base, "zmo_gold_%s_%d.ovl" % (name, os.getpid()))
    try:
        if not os.path.exists(fa):
            subprocess.run([gen_reads] + gold["gen"] + ["-o", fa + ".tmp"], check=True)
            os.replace(fa + ".tmp", fa)
        r = subprocess.run([EXE, "-t", "1", "-i", fa, "-f", "-o", out] + gold["args"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env)
        assert r.returncode == 0, r.stderr[-2000:]
        data = open(out, "rb").read()
        assert data.count(b"\n") == gold["lines"], (name, data.count(b"\n"), gold["lines"])
        assert hashlib.md5(data).hexdigest() == gold["md5"], name
        assert hashlib.md5(open(out + ".contained", "rb").read()).hexdigest() == gold["contained_md5"], name
        return r.stderr
    finally:
        for f in (out, out + ".contained") + (() if keep else (fa,)):
            if os.path.exists(f):
                os.remove(f)


def test_scale_200k_reads_shard_matches_reference_golden(tmp_path, gen_reads):
    """2.09 Gbp read set (200,000 x 10 kb): one query shard against the full index (~8 min of CPU for the reference)"""
    _golden_run(tmp_path, gen_reads, "big200k_P400_p0")


def test_cfg1_full_run_matches_reference_golden(tmp_path, gen_reads):
    """BASELINE.json configs[0], the whole job: 2,000 PacBio-like reads x 8 kb, 7,443 records incl. every CIGAR; and the same reads in
    dot-matrix mode (smartdenovo.pl:48 flags), 11,921 records"""
    _golden_run(tmp_path, gen_reads, "cfg1_full", keep=True)
    # window alignment: the bridge-level pipeline in passes of 1 MB of scratch (a large wave is swept in several passes over item ranges), and
    # the sequential warp-per-window kernel alone -- the same bytes
    _golden_run(tmp_path, gen_reads, "cfg1_full", env=dict(os.environ, ZMO_WB_CHUNK_MB="1"), keep=True)
    _golden_run(tmp_path, gen_reads, "cfg1_full", env=dict(os.environ, ZMO_WA_BRIDGE="0"), keep=True)
    _golden_run(tmp_path, gen_reads, "cfg1_dot")


def test_cfg2_bench_workload_shards_match_reference_golden(tmp_path, gen_reads):
    """BASELINE.json configs[1] = the workload bench.py times (50,000 reads x 10 kb): query shards `-P 160 -p 0` and `-p 7` against the
    full index, byte-identical to the reference `-t 1` (SURVEY 8d's sub-shard recipe); one of them again with the batch pipeline off"""
    _golden_run(tmp_path, gen_reads, "cfg2_P160_p0", keep=True)
    _golden_run(tmp_path, gen_reads, "cfg2_P160_p7", keep=True)
    _golden_run(tmp_path, gen_reads, "cfg2_P160_p0", env=dict(os.environ, ZMO_PIPELINE="0"))


def test_cfg2_full_bench_step_matches_reference_golden(tmp_path, gen_reads):
    """one complete bench step (`-P 10 -p 0`: 5,000 query reads of cfg2) against the reference `-t 1` digest"""
    import json
    if "cfg2_P10_p0" not in json.load(open(os.path.join(REPO, "tests", "golden", "scale_digests.json"))):
        pytest.skip("golden not generated (tests/golden/make_scale_golden.py cfg2_P10_p0: ~1 h of CPU)")
    _golden_run(tmp_path, gen_reads, "cfg2_P10_p0")


def test_cfg3_shaped_dot_matrix_shards_match_reference_golden(tmp_path, gen_reads):
    """BASELINE.json configs[2] shape at 1/10 of the reads (20,000 ONT-like reads x 15 kb, 30x, `-U -1 -m 0.1 -A 1000`): two `-P 4` query
    shards, ~147,000 records each"""
    _golden_run(tmp_path, gen_reads, "cfg3s_dot_P4_p0", keep=True)
    _golden_run(tmp_path, gen_reads, "cfg3s_dot_P4_p1")


def test_partitioned_index_G4_matches_reference_golden(tmp_path, gen_reads):
    """-G 4 (index built in four read-range partitions, candidates carried between them, wtzmo.c:1276-1303): 1,370 records"""
    _golden_run(tmp_path, gen_reads, "g4_2000")


def test_two_jobs_concurrently_on_one_gpu_match_their_shards(tmp_path, gen_reads, oracle_bin):
    """GPU g of n == reference job `-P n -p g`: both jobs of a -P 2 split run AT THE SAME TIME as two processes sharing the
    device (separate contexts, streams and arenas), each byte-identical to the checker's `-P 2 -p g`"""
    fa = str(tmp_path / "reads.fa")
    subprocess.run([gen_reads, "-n", "400", "-L", "6000", "-G", "80000", "-s", "31", "-o", fa], check=True)
    base = ["-k", "16", "-s", "200", "-m", "0.6", "-P", "2"]
    procs = [subprocess.Popen([EXE, "-t", "1", "-i", fa, "-f", "-o", str(tmp_path / ("gpu%d.ovl" % g))] + base + ["-p", str(g)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for g in range(2)]
    for g, p in enumerate(procs):
        _, err = p.communicate(timeout=600)
        assert p.returncode == 0, err[-2000:]
    total = 0
    for g in range(2):
        _run(_checker(oracle_bin), fa, str(tmp_path / ("ref%d.ovl" % g)), base + ["-p", str(g)])
        ref = open(tmp_path / ("ref%d.ovl" % g), "rb").read()
        assert ref == open(tmp_path / ("gpu%d.ovl" % g), "rb").read(), g
        assert open(tmp_path / ("ref%d.ovl.contained" % g), "rb").read() == open(tmp_path / ("gpu%d.ovl.contained" % g), "rb").read()
        total += ref.count(b"\n")
    assert total > 300


def test_multi_gpu_job_needs_distinct_gpus(tmp_path, gen_reads):
    """ZMO_GPUS=2 on a box with one GPU (or the same ordinal twice) stops before any work is done"""
    fa = str(tmp_path / "reads.fa")
    subprocess.run([gen_reads, "-n", "50", "-L", "3000", "-G", "30000", "-s", "2", "-o", fa], check=True)
    r = subprocess.run([EXE, "-t", "1", "-i", fa, "-f", "-o", str(tmp_path / "x.ovl"), "-k", "16"], env=dict(os.environ, ZMO_GPUS="2", ZMO_DEVICES="0,0"), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 2 and "twice" in r.stderr


def test_multi_gpu_job_equals_cat_of_reference_jobs(tmp_path, gen_reads, oracle_bin):
    """product-level multi-GPU (ZMO_GPUS=n, one process, one host thread per GPU, records gathered on GPU 0 over NCCL): the output is
    `cat` of the checker's jobs `-P n -p 0..n-1`, .contained their union in read-id order, -9 the union of their pairs"""
    import torch
    n = min(torch.cuda.device_count(), 4)
    if n < 2:
        pytest.skip("needs at least two GPUs (run under `gpurun --gpus 2`)")
    fa = str(tmp_path / "reads.fa")
    subprocess.run([gen_reads, "-n", "600", "-L", "6000", "-G", "100000", "-s", "37", "-o", fa], check=True)
    base = ["-k", "16", "-s", "200", "-m", "0.6"]
    _run(EXE, fa, str(tmp_path / "gpu.ovl"), base, env=dict(os.environ, ZMO_GPUS=str(n)))
    cat, contained, pairs = b"", set(), set()
    for g in range(n):
        _run(_checker(oracle_bin), fa, str(tmp_path / ("ref%d.ovl" % g)), base + ["-P", str(n), "-p", str(g)])
        cat += open(tmp_path / ("ref%d.ovl" % g), "rb").read()
        contained |= set(open(tmp_path / ("ref%d.ovl.contained" % g)).read().split())
        pairs |= set(open(tmp_path / ("ref%d.ovl.pairs" % g)).read().split("\n"))
    assert open(tmp_path / "gpu.ovl", "rb").read() == cat and cat.count(b"\n") > 500
    got = open(tmp_path / "gpu.ovl.contained").read().split()
    assert set(got) == contained and len(got) == len(contained)
    lens = {l[1:].strip(): len(s.strip()) for l, s in zip(*[iter(open(fa))] * 2)}
    assert all(lens[a] >= lens[b] for a, b in zip(got, got[1:]))          # read-id order = length descending
    assert set(open(tmp_path / "gpu.ovl.pairs").read().split("\n")) == pairs


def test_cfg2_full_size_shard_properties(tmp_path, gen_reads):
    """BASELINE.json configs[1] at full size (50,000 PacBio-like reads x 10 kb, the bench workload): one `-P 10 -p 3` query shard, too
    large for a CPU run inside the suite, checked through the size-independent properties of tests/ovl_props.py -- every record inside its
    reads, counts consistent with the spans, thresholds honoured, pairs unique, and for every 32nd record the CIGAR walked over the actual
    bases must reproduce mat / mis / ins / del and the reported coordinates on the strand shown"""
    from ovl_props import check_ovl, load_fasta
    base = "/dev/shm" if os.path.isdir("/dev/shm") else str(tmp_path)
    fa = os.path.join(base, "zmo_cfg2_props.fa")
    out = os.path.join(base, "zmo_cfg2_props.ovl")
    try:
        subprocess.run([gen_reads, "-n", "50000", "-L", "10000", "-G", "4600000", "-m", "pacbio", "-s", "20240603", "-o", fa], check=True)
        r = subprocess.run([EXE, "-t", "1", "-i", fa, "-f", "-o", out, "-k", "16", "-s", "200", "-m", "0.6", "-P", "10", "-p", "3"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        assert r.returncode == 0, r.stderr[-2000:]
        reads = load_fasta(fa)
        assert len(reads) == 50000
        n, walked, cols = check_ovl(reads, out, min_score=200, min_id=0.6, walk_every=32)
        assert n > 30000 and walked >= n // 32 and cols > 3 * 10 ** 8
        contained = open(out + ".contained", "rb").read().split()
        assert len(contained) == len(set(contained)) > 100 and all(x in reads for x in contained)
    finally:
        for f in (fa, out, out + ".contained"):
            if os.path.exists(f):
                os.remove(f)
