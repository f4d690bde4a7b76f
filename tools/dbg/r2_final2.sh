#!/bin/bash
# round 2: final state on one GPU -- full -m gpu suite, smoke, bench (both arms)
set -u
out=gpurun_out/r2final2; mkdir -p "$out"
python -c "import __graft_entry__ as g; g.build()" > "$out/build.log" 2>&1 || { echo BUILD FAILED; tail -5 "$out/build.log"; exit 9; }
timeout 1500 python -m pytest tests -q -m gpu -x --durations=8 > "$out/pytest_gpu.log" 2>&1; echo "pytest rc=$?" >> "$out/pytest_gpu.log"
tail -12 "$out/pytest_gpu.log"
python -c "import __graft_entry__ as g; g.smoke()" > "$out/smoke.log" 2>&1; tail -1 "$out/smoke.log"
timeout 900 python bench.py --steps 3 --warmup 3 > "$out/bench_default.json" 2> "$out/bench_default.err"
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 3 > "$out/bench_reference.json" 2> "$out/bench_reference.err" ) 2> "$out/bench_reference.time"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2final2/bench_default.json").read().strip().splitlines()[-1])
r=json.loads(open("gpurun_out/r2final2/bench_reference.json").read().strip().splitlines()[-1])
print("ours value", d["value"], "e2e", d["e2e"]["value"], "ms", d["ms_per_step"], "parity", d["parity_checked"], "launches", d["gpu_launches"])
print("stage", {k:round(v) for k,v in d["stage_ms_per_step"].items()})
print("roofline", d["roofline"]["stage"], d["roofline"]["frac"], "clocks", d["clocks"])
print("sub", {k:(v.get("value"), v.get("ms_per_step"), v.get("parity_checked")) for k,v in d["sub"].items()})
print("cli", d["cli_whole_job"]); print("cpu_baseline", d["cpu_baseline"]["value"])
print("reference value", r["value"], r["reference_phases_s_per_step"], "over process wall", r["value_over_process_wall"])
print("ratio e2e", d["e2e"]["value"]/r["value"])
PY
cat "$out/bench_reference.time"
G=tools/_build/gen_reads; FA=/dev/shm/cfg2.fa
[ -f $FA ] || $G -n 50000 -L 10000 -G 4600000 -m pacbio -s 20240603 -o $FA
W=smartdenovo_b200/bin/wtzmo
ARGS="-t 1 -i $FA -f -o /dev/shm/o.ovl -k 16 -s 200 -m 0.6 -P 40 -p 0"
export ZMO_PIPELINE=0
$W $ARGS 2>/dev/null; md5sum /dev/shm/o.ovl
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/launches_cfg2_P40.csv $W $ARGS >/dev/null 2>&1
python tools/launch_summary.py $out/launches_cfg2_P40.csv > $out/launches_cfg2_P40.md 2>/dev/null; head -30 $out/launches_cfg2_P40.md
for K in k_wb_sweep k_wb_stitch k_wb_walk k_wb_prep; do
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$K" -s 1 -c 1 -f -o "$out/prof_$K" $W $ARGS >/dev/null 2>$out/ncu_err_$K.txt
done
