#!/bin/bash
set -u
out=gpurun_out/r2pred; mkdir -p "$out"
python -c "import __graft_entry__ as g; g.build()" > "$out/build.log" 2>&1 || { echo BUILD FAILED; tail -5 "$out/build.log"; exit 9; }
bash tools/dbg/sweep.sh "ZMO_WAVE_PREDICT=2 ZMO_WAVE0=32" "ZMO_WAVE_PREDICT=1 ZMO_WAVE0=32" "ZMO_WAVE_PREDICT=3 ZMO_WAVE0=32" "ZMO_WAVE_PREDICT=2 ZMO_WAVE0=64" "ZMO_WAVE_PREDICT=2 ZMO_WAVE0=128" "ZMO_WAVE_PREDICT=2 ZMO_WAVE0=32 ZMO_WAVE_GROWTH=8" "ZMO_WAVE_PREDICT=1 ZMO_WAVE0=64" "ZMO_WAVE_PREDICT=2 ZMO_WAVE0=32 ZMO_DEPTH=3" "ZMO_WAVE_PREDICT=2 ZMO_WAVE0=32" 2>&1 | tee "$out/sweep.txt"
