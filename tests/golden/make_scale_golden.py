#!/usr/bin/env python
"""Golden digests of the UNMODIFIED reference binary (`oracle/_ref/wtzmo -t 1`, the only deterministic mode) on the
BASELINE.json-shaped inputs that are too large for the CPU suite to re-run: tests/golden/scale_digests.json.

  python tests/golden/make_scale_golden.py [case ...]        (where /root/reference exists; minutes to tens of minutes per case)

Every case is (generator arguments, wtzmo arguments); the read sets are regenerated from the seed (tools/gen_reads.c), never
committed.  Recorded per case: md5 + line count of the .ovl, md5 of .contained, aligned columns (cols 13-16, or max span in
dot-matrix mode), the reference's wall time here and its overlap-phase seconds.  Consumers: tests/test_gpu_wtzmo.py (-m gpu,
byte parity of the product binary at BASELINE sizes) and bench.py (parity of the timed workload, per-rank `-P n -p g` parity).
"""
import hashlib
import json
import os
import subprocess
import sys
import time
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(HERE, "scale_digests.json")

SW = ["-k", "16", "-s", "200", "-m", "0.6"]
DOT = ["-k", "16", "-z", "10", "-Z", "16", "-U", "-1", "-m", "0.1", "-A", "1000"]
CFG1 = ["-n", "2000", "-L", "8000", "-G", "500000", "-m", "pacbio", "-s", "20240602"]
CFG2 = ["-n", "50000", "-L", "10000", "-G", "4600000", "-m", "pacbio", "-s", "20240603"]
CFG3S = ["-n", "20000", "-L", "15000", "-G", "10000000", "-m", "ont", "-s", "20240604"]     # cfg3 shape (ONT 15 kb, 30x) at 1/10 of the reads
CASES = {
    "big200k_P400_p0": (["-n", "200000", "-L", "10000", "-G", "20000000", "-m", "pacbio", "-s", "20240605"], SW + ["-P", "400", "-p", "0"]),
    "cfg1_full": (CFG1, SW),
    "cfg1_dot": (CFG1, DOT),
    "cfg3s_dot_P4_p0": (CFG3S, DOT + ["-P", "4", "-p", "0"]),
    "cfg3s_dot_P4_p1": (CFG3S, DOT + ["-P", "4", "-p", "1"]),
    "g4_2000": (["-n", "2000", "-L", "6000", "-G", "400000", "-m", "pacbio", "-s", "20240611"], SW + ["-G", "4"]),
    "cfg2_P10_p0": (CFG2, SW + ["-P", "10", "-p", "0"]),
}
for _g in range(8):
    CASES["cfg2_P160_p%d" % _g] = (CFG2, SW + ["-P", "160", "-p", str(_g)])


def aligned_cols(path, dot):
    cols = 0
    with open(path) as f:
        for line in f:
            c = line.split("\t")
            cols += max(int(c[4]) - int(c[3]), int(c[9]) - int(c[8])) if dot else int(c[12]) + int(c[13]) + int(c[14]) + int(c[15])
    return cols


def reads_of(gen_args, tmp):
    gen = os.path.join(REPO, "tools", "_build", "gen_reads")
    fa = os.path.join(tmp, "gold_%s.fa" % hashlib.md5(" ".join(gen_args).encode()).hexdigest()[:10])
    if not os.path.exists(fa):
        subprocess.run([gen] + gen_args + ["-o", fa + ".tmp"], check=True)
        os.replace(fa + ".tmp", fa)
    return fa


def run_case(name, tmp):
    gen_args, args = CASES[name]
    ref = os.path.join(REPO, "oracle", "_ref", "wtzmo")
    fa = reads_of(gen_args, tmp)
    out = os.path.join(tmp, "gold_%s.ovl" % name)
    t0 = time.time()
    r = subprocess.run(["nice", "-n", "10", ref, "-t", "1", "-i", fa, "-f", "-o", out] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    wall = time.time() - t0
    if r.returncode != 0:
        raise RuntimeError(name + ": " + r.stderr[-500:])
    data = open(out, "rb").read()
    rec = {"gen": gen_args, "args": args, "md5": hashlib.md5(data).hexdigest(), "lines": data.count(b"\n"),
           "contained_md5": hashlib.md5(open(out + ".contained", "rb").read()).hexdigest(),
           "aligned_cols": aligned_cols(out, "-U" in args), "ref_wall_s": round(wall, 1)}
    os.remove(out)
    os.remove(out + ".contained")
    return name, rec


def main():
    names = sys.argv[1:] or [n for n in CASES if n != "big200k_P400_p0"]
    tmp = os.environ.get("ZMO_GOLD_TMP", "/dev/shm")
    subprocess.run(["make", "-C", os.path.join(REPO, "oracle"), "ref"], check=True, stdout=subprocess.DEVNULL)
    gold = json.load(open(OUT)) if os.path.exists(OUT) else {}
    # generate the read sets first (one writer per file), then the reference runs in parallel
    for n in names:
        reads_of(CASES[n][0], tmp)
    with ThreadPoolExecutor(max_workers=int(os.environ.get("ZMO_GOLD_JOBS", "4"))) as ex:
        for name, rec in ex.map(lambda n: run_case(n, tmp), names):
            gold[name] = rec
            print(name, json.dumps(rec), flush=True)
            json.dump(gold, open(OUT, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
