#!/usr/bin/env python
"""CPU fuzz: random small read sets x random wtzmo parameters, CPU oracle (oracle/_ref/zmo_oracle) against the unmodified
reference binary (oracle/_ref/wtzmo), byte-for-byte.  usage: tools/fuzz_oracle_vs_ref.py [n_cases] [seed] [gpu]
With a third argument the PRODUCT binary (smartdenovo_b200/bin/wtzmo, needs a B200) is checked instead of the oracle."""
import os, random, subprocess, sys, tempfile, hashlib
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
gen = os.path.join(REPO, "tools", "_build", "gen_reads"); ref = os.path.join(REPO, "oracle", "_ref", "wtzmo"); orc = os.path.join(REPO, "oracle", "_ref", "zmo_oracle")
if len(sys.argv) > 3:
    orc = os.path.join(REPO, "smartdenovo_b200", "bin", "wtzmo")
wide = os.environ.get("ZMO_FUZZ_WIDE", "0") != "0"   # ZMO_FUZZ_WIDE=1: also side inputs / -K / -J / -C / explicit -U, and .contained + -9 in the comparison
n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 20
rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
bad = 0
with tempfile.TemporaryDirectory() as d:
    for c in range(n_cases):
        # n is divisible by every -G used below: the reference reads out of bounds otherwise (wtzmo.c:1283)
        n = rng.choice([60, 100, 160]); L = rng.choice([2500, 4000, 6000]); G = rng.choice([15000, 30000, 50000]); model = rng.choice(["pacbio", "ont"])
        if wide: G = n * L // rng.choice([8, 15, 30])   # enough coverage for every case to produce records
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
        if wide:
            # second generation of cases: side inputs, -K, -J, -C, explicit -U values; .contained and -9 are compared as well
            names = [l[1:].split()[0] for l in open(fa) if l.startswith(">")]; lens = [len(l.strip()) for l in open(fa) if not l.startswith(">")]
            if rng.random() < 0.2: args += ["-K", str(rng.choice([3, 10, 40, 1000]))]
            if rng.random() < 0.2: args += ["-J", str(rng.choice([L // 2, L, L + L // 5]))]
            if rng.random() < 0.15: args += ["-C"]
            if "-U" not in args and rng.random() < 0.15:
                for v in (rng.choice([64, 128, 256]), rng.choice([16, 64, 100]), rng.choice([80, 160, 300]), rng.choice([0.5, 1.0, 2.0]), rng.choice([0.01, 0.05, 0.2])): args += ["-U", str(v)]
            if rng.random() < 0.2:
                with open(os.path.join(d, "L.pairs"), "w") as f:
                    for _ in range(rng.randrange(1, 80)): f.write("%s\t%s\n" % (rng.choice(names), rng.choice(names)))
                args += ["-L", os.path.join(d, "L.pairs")]
            if rng.random() < 0.2:
                with open(os.path.join(d, "F.names"), "w") as f:
                    f.write("".join(x + "\n" for x in rng.sample(names, rng.randrange(1, n // 3))))
                args += ["-F", os.path.join(d, "F.names")]
            if rng.random() < 0.2:
                with open(os.path.join(d, "B.clip"), "w") as f:
                    for i in rng.sample(range(n), rng.randrange(1, n // 2)):
                        b = rng.randrange(0, lens[i] // 3); e = rng.randrange(lens[i] // 2, lens[i] + 1)
                        f.write("%s\t%d\t%d\t%d\n" % (names[i], b, e, lens[i]))
                args += ["-b", os.path.join(d, "B.clip")]
            if rng.random() < 0.15:
                q = os.path.join(d, "q.fa")
                subprocess.run([gen, "-n", str(rng.choice([10, 30])), "-L", str(L), "-G", str(G), "-m", model, "-s", str(rng.randrange(1, 10 ** 6)), "-o", q], check=True)
                args += ["-I", q]
        outs = []
        for exe, tag in ((ref, "ref"), (orc, "orc")):
            o = os.path.join(d, tag + ".ovl")
            for f in (o, o + ".contained", o + ".pairs"):
                if os.path.exists(f): os.remove(f)
            r = subprocess.run([exe, "-t", "1", "-i", fa, "-f", "-o", o] + (["-9", o + ".pairs"] if wide else []) + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
            res = (r.returncode, open(o, "rb").read() if os.path.exists(o) else b"")
            if wide:
                res += (open(o + ".contained", "rb").read() if os.path.exists(o + ".contained") else None, sorted(open(o + ".pairs").read().split("\n")) if os.path.exists(o + ".pairs") else None)
            outs.append(res)
        ok = outs[0] == outs[1]
        if outs[0][0] < 0:   # the reference died on a signal: its behaviour is undefined there (e.g. -G n with a read count not divisible by n, wtzmo.c:1283; -J changes the count)
            print("case %d REFCRASH(%d) %s" % (c, outs[0][0], " ".join(args)), flush=True); continue
        print("case %d %s n=%d L=%d %s lines=%d %s" % (c, "ok " if ok else "DIFF", n, L, model, outs[0][1].count(b"\n"), " ".join(args)), flush=True)
        bad += not ok
        if not ok:
            import shutil
            keep = os.path.join(tempfile.gettempdir(), 'zmo_fuzz_case%d' % c); os.makedirs(keep, exist_ok=True)
            for f in ('r.fa', 'ref.ovl', 'orc.ovl', 'ref.ovl.contained', 'orc.ovl.contained', 'ref.ovl.pairs', 'orc.ovl.pairs', 'L.pairs', 'F.names', 'B.clip', 'q.fa'):
                if os.path.exists(os.path.join(d, f)): shutil.copy(os.path.join(d, f), keep)
            open(os.path.join(keep, 'args.txt'), 'w').write(' '.join(args) + '\nrc ref=%d orc=%d\n' % (outs[0][0], outs[1][0]))
            print('  kept in', keep)
print("mismatches:", bad)
sys.exit(1 if bad else 0)
