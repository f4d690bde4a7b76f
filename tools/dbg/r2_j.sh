#!/bin/bash
# round 2, GPU call J (1 GPU): pipeline ramp / drain sweep
set -u
out=gpurun_out/r2j; mkdir -p "$out"
python -c "import __graft_entry__ as g; g.build()" > "$out/build.log" 2>&1 || { echo BUILD FAILED; tail -5 "$out/build.log"; exit 9; }
timeout 1800 bash tools/dbg/sweep.sh "ZMO_RAMP=96" "ZMO_RAMP=96 ZMO_DRAIN_DIV=4" "ZMO_RAMP=96 ZMO_DRAIN_DIV=6" "ZMO_RAMP=96 ZMO_DRAIN_DIV=8 ZMO_DRAIN_MIN=32" "ZMO_RAMP=48 ZMO_DRAIN_DIV=6" "ZMO_RAMP=96 ZMO_DRAIN_DIV=6 ZMO_BATCH_READS=512" "ZMO_RAMP=96 ZMO_DRAIN_DIV=6 ZMO_BATCH_READS=256" "ZMO_RAMP=96 ZMO_DRAIN_DIV=6 ZMO_WAVE0=6" "ZMO_RAMP=96 ZMO_DRAIN_DIV=6 ZMO_DEPTH=3" > "$out/sweep.log" 2>&1
cat "$out/sweep.log"
