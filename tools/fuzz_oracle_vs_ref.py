#!/usr/bin/env python
"""CPU fuzz: random small read sets x random wtzmo parameters, CPU oracle (oracle/_ref/zmo_oracle) against the unmodified
reference binary (oracle/_ref/wtzmo), byte-for-byte.  usage: tools/fuzz_oracle_vs_ref.py [n_cases] [seed] [gpu]
With a third argument the PRODUCT binary (smartdenovo_b200/bin/wtzmo, needs a B200) is checked instead of the oracle."""
import os, random, subprocess, sys, tempfile, hashlib
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
gen = os.path.join(REPO, "tools", "_build", "gen_reads"); ref = os.path.join(REPO, "oracle", "_ref", "wtzmo"); orc = os.path.join(REPO, "oracle", "_ref", "zmo_oracle")
if len(sys.argv) > 3:
    orc = os.path.join(REPO, "smartdenovo_b200", "bin", "wtzmo")
n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 20
rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
bad = 0
with tempfile.TemporaryDirectory() as d:
    for c in range(n_cases):
        # n is divisible by every -G used below: the reference reads out of bounds otherwise (wtzmo.c:1283)
        n = rng.choice([60, 100, 160]); L = rng.choice([2500, 4000, 6000]); G = rng.choice([15000, 30000, 50000]); model = rng.choice(["pacbio", "ont"])
        fa = os.path.join(d, "r.fa")
        subprocess.run([gen, "-n", str(n), "-L", str(L), "-G", str(G), "-m", model, "-s", str(rng.randrange(1, 10 ** 6)), "-o", fa], check=True)
        args = ["-k", str(rng.choice([13, 15, 16, 17, 20])), "-S", str(rng.choice([1, 2, 4])), "-z", str(rng.choice([8, 10, 12])), "-Z", str(rng.choice([8, 16, 64, 300])),
                "-y", str(rng.choice([400, 800, 1200])), "-R", str(rng.choice([100, 200])), "-r", str(rng.choice([150, 300])), "-l", str(rng.choice([1, 2, 4])),
                "-d", str(rng.choice([100, 300])), "-A", str(rng.choice([5, 50, 500])), "-B", str(rng.choice([3, 20, 100])), "-w", str(rng.choice([10, 50, 120])),
                "-e", str(rng.choice([100, 800])), "-W", str(rng.choice([400, 3200])), "-s", str(rng.choice([50, 200])), "-m", str(rng.choice([0.4, 0.5, 0.6])),
                "-M", str(rng.choice([1, 2, 3])), "-X", str(rng.choice([-2, -5, -7])), "-O", str(rng.choice([-2, -3, -6])), "-E", str(rng.choice([-1, -2])), "-T", str(rng.choice([-20, -50, -100])),
                "-H", str(rng.choice([0, 1, 2, 3])), "-q", str(rng.choice([20, 100]))]
        if rng.random() < 0.3: args += ["-n"]
        if rng.random() < 0.25: args += ["-U", "-1"]
        if rng.random() < 0.2: args += ["-P", "2", "-p", str(rng.randrange(2))]
        if rng.random() < 0.2: args += ["-G", str(rng.choice([2, 4]))]
        if rng.random() < 0.15: args += ["-N"]
        outs = []
        for exe, tag in ((ref, "ref"), (orc, "orc")):
            o = os.path.join(d, tag + ".ovl")
            r = subprocess.run([exe, "-t", "1", "-i", fa, "-f", "-o", o] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
            outs.append((r.returncode, open(o, "rb").read() if os.path.exists(o) else b""))
        ok = outs[0] == outs[1]
        print("case %d %s n=%d L=%d %s lines=%d %s" % (c, "ok " if ok else "DIFF", n, L, model, outs[0][1].count(b"\n"), " ".join(args)), flush=True)
        bad += not ok
        if not ok:
            import shutil
            keep = os.path.join(tempfile.gettempdir(), 'zmo_fuzz_case%d' % c); os.makedirs(keep, exist_ok=True)
            for f in ('r.fa', 'ref.ovl', 'orc.ovl'):
                if os.path.exists(os.path.join(d, f)): shutil.copy(os.path.join(d, f), keep)
            open(os.path.join(keep, 'args.txt'), 'w').write(' '.join(args) + '\nrc ref=%d orc=%d\n' % (outs[0][0], outs[1][0]))
            print('  kept in', keep)
print("mismatches:", bad)
sys.exit(1 if bad else 0)
