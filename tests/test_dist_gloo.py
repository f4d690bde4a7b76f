"""CPU, world_size 2, gloo: the multi-rank plumbing used by `bench.py --gpus N` -- job sharding identical to the
reference's `-P n -p i`, record gather equivalent to `cat part*.ovl`, max/sum reductions of the timing rule.
The shard semantics themselves are checked against the oracle: the union of the two ranks' oracle runs with
-P 2 -p rank equals what the gather returns."""
import os
import subprocess
import sys

import pytest

from conftest import REPO

WORKER = r'''
import os, sys, subprocess
sys.path.insert(0, sys.argv[1])
import torch.distributed as dist
from smartdenovo_b200 import dist as zd
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % os.environ["MASTER_PORT"], rank=rank, world_size=world)
oracle, fa, outdir = sys.argv[2], sys.argv[3], sys.argv[4]
assert [zd.shard_of(s, rank, world, 2 * world) for s in range(4)] == [(s * world + rank) % (2 * world) for s in range(4)]
out = os.path.join(outdir, "part%d.ovl" % rank)
subprocess.run([oracle, "-t", "1", "-i", fa, "-f", "-o", out, "-k", "16", "-P", str(world), "-p", str(zd.shard_of(0, rank, world, world))], check=True, stderr=subprocess.DEVNULL)
parts = zd.gather_records(open(out, "rb").read())
mx = zd.max_over_ranks([float(rank + 1), 5.0])
sm = zd.sum_over_ranks([float(rank + 1)])
assert mx == [float(world), 5.0] and sm == [world * (world + 1) / 2.0]
tot, blob = zd.gather_record_file(out)
assert tot == sum(len(x) for x in parts) and (b"".join(v.tobytes() for v in blob) == b"".join(parts) if rank == 0 else blob is None)
if rank == 0:
    open(os.path.join(outdir, "gathered.ovl"), "wb").write(b"".join(parts))
# an empty payload on one rank must not break the gather
parts2 = zd.gather_records(b"" if rank == 1 else b"x\n")
assert parts2 == [b"x\n", b""]
dist.barrier()
dist.destroy_process_group()
'''


def test_two_rank_shards_and_gather(tmp_path, gen_reads, oracle_bin):
    fa = str(tmp_path / "r.fa")
    subprocess.run([gen_reads, "-n", "160", "-L", "5000", "-G", "40000", "-s", "13", "-o", fa], check=True)
    worker = tmp_path / "worker.py"
    worker.write_text(WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=port)
        procs.append(subprocess.Popen([sys.executable, str(worker), REPO, oracle_bin, fa, str(tmp_path)], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    for p in procs:
        out, err = p.communicate(timeout=300)
        assert p.returncode == 0, err[-2000:]
    gathered = open(tmp_path / "gathered.ovl", "rb").read()
    assert gathered == open(tmp_path / "part0.ovl", "rb").read() + open(tmp_path / "part1.ovl", "rb").read()
    assert gathered.count(b"\n") > 4
