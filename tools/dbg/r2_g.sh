#!/bin/bash
# round 2, GPU call G (1 GPU): full -m gpu suite, bench, launch list after the seeding changes (warp-parallel sub-window scan, chunk-parallel z-scan)
set -u
out=gpurun_out/r2h; mkdir -p "$out"
G=tools/_build/gen_reads
python -c "import __graft_entry__ as g; g.build()" > "$out/build.log" 2>&1 || { echo BUILD FAILED; tail -5 "$out/build.log"; exit 9; }
timeout 1200 python -m pytest tests -q -m gpu -x > "$out/pytest_gpu.log" 2>&1; echo "pytest rc=$?" >> "$out/pytest_gpu.log"
tail -3 "$out/pytest_gpu.log"
FA=/dev/shm/cfg2.fa
[ -f $FA ] || $G -n 50000 -L 10000 -G 4600000 -m pacbio -s 20240603 -o $FA
W=smartdenovo_b200/bin/wtzmo
ARGS="-t 1 -i $FA -f -o /dev/shm/o.ovl -k 16 -s 200 -m 0.6 -P 40 -p 0"
ZMO_PIPELINE=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/launches_cfg2_P40.csv $W $ARGS >/dev/null 2>&1
python tools/launch_summary.py $out/launches_cfg2_P40.csv 16
for i in 1 2; do ( time ZMO_STATS=$out/stats_cold_P10_$i.json $W -t 1 -i $FA -f -o /dev/shm/o.ovl -k 16 -s 200 -m 0.6 -P 10 -p 0 ) 2> "$out/run_cold_P10_$i.err"; done
python - <<'PY'
import json
for i in (1,2):
    d=json.load(open("gpurun_out/r2h/stats_cold_P10_%d.json"%i)); print("cold P10 run",i,"overlap_s",d["overlap_s"],"total_s",d["total_s"],"load_s",d["load_s"],d["alloc"])
PY
timeout 900 python bench.py --steps 3 --warmup 3 > "$out/bench_default.json" 2> "$out/bench_default.err"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2h/bench_default.json").read().strip().splitlines()[-1])
print("bench value", d["value"], "e2e", d["e2e"]["value"], "ms", d["ms_per_step"], "parity", d["parity_checked"], {k:round(v) for k,v in d["stage_ms_per_step"].items()})
print("cli", d.get("cli_whole_job")); print({k:(v.get("value"), v.get("ms_per_step"), v.get("parity_checked")) for k,v in d["sub"].items()})
PY
