#!/bin/bash
# round 2, GPU call F (2 GPUs): cfg4 at full size -- 1,000,000 PacBio-like reads x 12 kb over a 140 Mb genome (12 Gbp, 86x) -- query-sharded:
# two jobs of 1,000 query reads each (-P 1000 -p 0/1) one after the other on one GPU, then both at once on two GPUs (ZMO_GPUS=2)
set -u
out=gpurun_out/r2f; mkdir -p "$out"
G=tools/_build/gen_reads
python -c 'import __graft_entry__ as g; g.build()' > "$out/build.log" 2>&1
df -h /dev/shm > "$out/env.txt"; free -g >> "$out/env.txt"; nproc >> "$out/env.txt"
( time $G -n 1000000 -L 12000 -G 140000000 -m pacbio -s 20240605 -o /dev/shm/cfg4.fa ) 2> "$out/gen4.log"
ls -la /dev/shm/cfg4.fa >> "$out/env.txt"
W=smartdenovo_b200/bin/wtzmo
A="-t 1 -i /dev/shm/cfg4.fa -f -k 16 -s 200 -m 0.6"
for p in 0 1; do
  ( time ZMO_STATS=$out/stats_1gpu_p$p.json $W $A -o /dev/shm/c4_p$p.ovl -P 1000 -p $p ) 2> "$out/run_1gpu_p$p.err"
  nvidia-smi --query-gpu=index,memory.used --format=csv,noheader >> "$out/run_1gpu_p$p.err"
done
( time ZMO_STATS=$out/stats_1gpu_G4_p0.json $W $A -o /dev/shm/c4_G4_p0.ovl -P 1000 -p 0 -G 4 ) 2> "$out/run_1gpu_G4_p0.err"
( time ZMO_GPUS=2 ZMO_STATS=$out/stats_2gpu.json $W $A -o /dev/shm/c4_m2.ovl -P 500 -p 0 ) 2> "$out/run_2gpu.err"
cat /dev/shm/c4_p0.ovl /dev/shm/c4_p1.ovl | md5sum > "$out/md5.txt"; md5sum /dev/shm/c4_m2.ovl >> "$out/md5.txt"; wc -l /dev/shm/c4_p0.ovl /dev/shm/c4_p1.ovl /dev/shm/c4_m2.ovl >> "$out/md5.txt"
cat "$out/md5.txt"; tail -4 "$out/gen4.log"
python - <<'PY'
import json
for f in ("stats_1gpu_p0","stats_1gpu_p1","stats_1gpu_G4_p0","stats_2gpu"):
    try:
        d=json.load(open("gpurun_out/r2f/%s.json"%f)); print(f, "records",d["records"],"cols",d["aligned_cols"],"overlap_s",d["overlap_s"],"total_s",d["total_s"],"load_s",d["load_s"],"gather",d["gather_wall_s"],{k:round(v) for k,v in d["stage_ms"].items()}, d["alloc"])
    except Exception as e: print(f, "failed", e)
PY
tail -n 3 "$out/run_1gpu_p0.err"; tail -n 12 "$out/run_1gpu_G4_p0.err"; tail -n 5 "$out/run_2gpu.err"; wc -l /dev/shm/c4_G4_p0.ovl
