#!/bin/bash
# round 2, final state at cfg4 size on one GPU: 1,000,000 reads x 12 kb; the two 1,000-query jobs of profiles/r02_cfg4_full_md5.txt again
set -u
out=gpurun_out/r2c4; mkdir -p "$out"
G=tools/_build/gen_reads
python -c 'import __graft_entry__ as g; g.build()' > "$out/build.log" 2>&1 || { echo BUILD FAILED; exit 9; }
( time timeout 400 $G -n 1000000 -L 12000 -G 140000000 -m pacbio -s 20240605 -o /dev/shm/cfg4.fa ) 2> "$out/gen4.log"; tail -3 "$out/gen4.log"
W=smartdenovo_b200/bin/wtzmo
A="-t 1 -i /dev/shm/cfg4.fa -f -k 16 -s 200 -m 0.6"
for p in 0 1; do
  ( time ZMO_STATS=$out/stats_1gpu_p$p.json timeout 300 $W $A -o /dev/shm/c4_p$p.ovl -P 1000 -p $p ) 2> "$out/run_1gpu_p$p.err"; tail -4 "$out/run_1gpu_p$p.err"
  nvidia-smi --query-gpu=index,memory.used --format=csv,noheader
done
cat /dev/shm/c4_p0.ovl /dev/shm/c4_p1.ovl | md5sum | tee "$out/md5.txt"; wc -l /dev/shm/c4_p0.ovl /dev/shm/c4_p1.ovl | tee -a "$out/md5.txt"
echo "recorded before the bridge pipeline: $(head -1 profiles/r02_cfg4_full_md5.txt)"
python - <<'PY'
import json
for f in ("stats_1gpu_p0","stats_1gpu_p1"):
    try:
        d=json.load(open("gpurun_out/r2c4/%s.json"%f)); print(f, "records",d["records"],"overlap_s",d["overlap_s"],"total_s",d["total_s"],"load_s",d["load_s"],{k:round(v) for k,v in d["stage_ms"].items()}, d["alloc"])
    except Exception as e: print(f, "failed", e)
PY
rm -f /dev/shm/cfg4.fa
