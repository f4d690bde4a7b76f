#!/bin/bash
set -u
out=gpurun_out/r2bench; mkdir -p "$out"
python -c "import __graft_entry__ as g; g.build()" > "$out/build.log" 2>&1 || { echo BUILD FAILED; tail -5 "$out/build.log"; exit 9; }
( time timeout 900 python bench.py > "$out/bench_default.json" 2> "$out/bench_default.err" ) 2>&1 | tail -3
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2bench/bench_default.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "ms", d["ms_per_step"], "parity", d["parity_checked"], "launches", d["gpu_launches"])
print("cli", {k: d["cli_whole_job"].get(k) for k in ("process_wall_s","overlap_phase_s","records")})
print("cli16", {k: d["cli_whole_job_16_columns"].get(k) for k in ("command","process_wall_s","overlap_phase_s","records")})
print("roofline", d["roofline"]["stage"], d["roofline"]["frac"], d["roofline"]["traffic"])
PY
