#!/usr/bin/env python
"""Regenerate tests/golden/ovl_digests.json by running the UNMODIFIED reference binary
(oracle/_ref/wtzmo, built from /root/reference by oracle/Makefile) on the seeded synthetic inputs of
tests/test_oracle_vs_ref.py.  Run where /root/reference exists: python tests/golden/make_golden.py"""
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(REPO, "tests"))
import test_oracle_vs_ref as t  # noqa: E402

subprocess.run(["make", "-C", os.path.join(REPO, "oracle"), "ref"], check=True)
gen = os.path.join(REPO, "tools", "_build", "gen_reads")
os.makedirs(os.path.dirname(gen), exist_ok=True)
subprocess.run(["gcc", "-O2", "-o", gen, os.path.join(REPO, "tools", "gen_reads.c"), "-lm"], check=True)
ref = os.path.join(REPO, "oracle", "_ref", "wtzmo")
out = {}
with tempfile.TemporaryDirectory() as d:
    for case, (gen_args, extra) in sorted(t.CASES.items()):
        fa = os.path.join(d, case + ".fa")
        subprocess.run([gen] + gen_args + ["-o", fa], check=True)
        t._run(ref, fa, os.path.join(d, case + ".ovl"), extra)
        out[case] = t._digest(os.path.join(d, case + ".ovl"))
    import pathlib
    for case in t.SIDE_CASES:
        sub = pathlib.Path(d) / case
        sub.mkdir()
        fa, extra = t._side_inputs(sub, gen, case)
        t._run(ref, fa, str(sub / "ref.ovl"), extra)
        out[case] = t._digest(str(sub / "ref.ovl"))
    sub = pathlib.Path(d) / "long_indel"
    sub.mkdir()
    from conftest import write_long_indel_reads
    write_long_indel_reads(str(sub / "w.fa"))
    t._run(ref, str(sub / "w.fa"), str(sub / "ref.ovl"), t.LONG_INDEL_ARGS)
    out["refine_long_indel"] = t._digest(str(sub / "ref.ovl"))
json.dump(out, open(os.path.join(HERE, "ovl_digests.json"), "w"), indent=1, sort_keys=True)
print(json.dumps(out, indent=1))
