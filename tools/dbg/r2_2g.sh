#!/bin/bash
set -u
out=gpurun_out/r2_2g; mkdir -p "$out"
python -c 'import __graft_entry__ as g; g.build()' > "$out/build.log" 2>&1 || { echo BUILD FAILED; exit 9; }
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 3 --warmup 3 > "$out/bench_2gpu.json" 2> "$out/bench_2gpu.err"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_2g/bench_2gpu.json").read().strip().splitlines()[-1])
print("bench 2gpu value", d["value"], "e2e", d["e2e"]["value"], "ms", d["ms_per_step"], "parity", d["parity_checked"], d["parity"])
PY
