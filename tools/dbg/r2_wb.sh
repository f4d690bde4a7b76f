#!/bin/bash
# round 2: bridge-level window alignment (zmo_winbridge.cuh) -- whole-program golden parity, A/B bench, launch list
set -u
out=gpurun_out/r2wb; mkdir -p "$out"
python -c "import __graft_entry__ as g; g.build()" > "$out/build.log" 2>&1 || { echo BUILD FAILED; tail -5 "$out/build.log"; exit 9; }
timeout 1200 python -m pytest tests/test_gpu_wtzmo.py -q -m gpu -x --durations=5 > "$out/pytest_gpu.log" 2>&1; echo "pytest rc=$?" >> "$out/pytest_gpu.log"
tail -9 "$out/pytest_gpu.log"
for m in 1 0; do
  ZMO_WA_BRIDGE=$m timeout 600 python bench.py --steps 3 --warmup 3 --no-sub > "$out/bench_wb$m.json" 2> "$out/bench_wb$m.err" || tail -5 "$out/bench_wb$m.err"
done
python - <<'PY'
import json
for m in (1,0):
    try:
        d=json.loads(open("gpurun_out/r2wb/bench_wb%d.json"%m).read().strip().splitlines()[-1])
        print("WB",m,"value", d["value"], "e2e", d["e2e"]["value"], "ms", d["ms_per_step"], "parity", d["parity_checked"], d["parity"].get("ok"), "launches", d["gpu_launches"])
        print("   stage", {k:round(v) for k,v in d["stage_ms_per_step"].items()})
    except Exception as e: print("WB",m,"failed",e)
PY
