#!/bin/bash
# round 2, GPU call E (1 GPU): full -m gpu suite after the clean-up, cold-start anatomy (allocation timers), warp stitch kernel time,
# cfg3 at full size (200,000 ONT-like reads x 15 kb, dot-matrix mode): one query shard
set -u
out=gpurun_out/r2e; mkdir -p "$out"
G=tools/_build/gen_reads
python -c 'import __graft_entry__ as g; g.build()' > "$out/build.log" 2>&1
( $G -n 200000 -L 15000 -G 100000000 -m ont -s 20240604 -o /dev/shm/cfg3.fa; echo gen3 done ) > "$out/gen3.log" 2>&1 &
timeout 1200 python -m pytest tests -q -m gpu -x > "$out/pytest_gpu.log" 2>&1; echo "pytest rc=$?" >> "$out/pytest_gpu.log"
tail -3 "$out/pytest_gpu.log"
FA=/dev/shm/cfg2.fa
[ -f $FA ] || $G -n 50000 -L 10000 -G 4600000 -m pacbio -s 20240603 -o $FA
W=smartdenovo_b200/bin/wtzmo
for i in 1 2; do ( time ZMO_STATS=$out/stats_cold_P10_$i.json $W -t 1 -i $FA -f -o /dev/shm/o.ovl -k 16 -s 200 -m 0.6 -P 10 -p 0 ) 2> "$out/run_cold_P10_$i.err"; done
python - <<'PY'
import json
for i in (1,2):
    d=json.load(open("gpurun_out/r2e/stats_cold_P10_%d.json"%i)); print("cold P10 run",i,"overlap_s",d["overlap_s"],"total_s",d["total_s"],"load_s",d["load_s"],d["alloc"],{k:round(v) for k,v in d["stage_ms"].items()})
PY
ARGS="-t 1 -i $FA -f -o /dev/shm/o.ovl -k 16 -s 200 -m 0.6 -P 40 -p 0"
ZMO_PIPELINE=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/launches_cfg2_P40.csv $W $ARGS >/dev/null 2>&1
python tools/launch_summary.py $out/launches_cfg2_P40.csv 14
timeout 900 python bench.py --steps 3 --warmup 3 > "$out/bench_default.json" 2> "$out/bench_default.err"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2e/bench_default.json").read().strip().splitlines()[-1])
print("bench value", d["value"], "ms", d["ms_per_step"], "parity", d["parity_checked"], {k:round(v) for k,v in d["stage_ms_per_step"].items()}, "cli", d.get("cli_whole_job"))
PY
wait
cat "$out/gen3.log"
( time ZMO_STATS=$out/stats_cfg3_P160_p0.json $W -t 1 -i /dev/shm/cfg3.fa -f -o /dev/shm/cfg3.ovl -k 16 -z 10 -Z 16 -U -1 -m 0.1 -A 1000 -P 160 -p 0 ) 2> "$out/run_cfg3.err"
tail -6 "$out/run_cfg3.err"; cat $out/stats_cfg3_P160_p0.json; wc -l /dev/shm/cfg3.ovl; md5sum /dev/shm/cfg3.ovl > $out/cfg3_P160_p0.md5
nvidia-smi --query-gpu=memory.used,memory.total --format=csv >> "$out/run_cfg3.err"
