#!/bin/bash
set -u
out=gpurun_out/r2sw; mkdir -p "$out"
python -c "import __graft_entry__ as g; g.build()" > "$out/build.log" 2>&1 || { echo BUILD FAILED; tail -5 "$out/build.log"; exit 9; }
bash tools/dbg/sweep.sh "A=1" "ZMO_BATCH_PAIRS=30000" "ZMO_BATCH_PAIRS=60000" "ZMO_BATCH_READS=448 ZMO_BATCH_PAIRS=50000" "ZMO_DRAIN_MIN=32" "ZMO_DRAIN_DIV=2" "ZMO_WAVE0=24" "ZMO_WAVE0=48" "A=1" 2>&1 | tee "$out/sweep.txt"
