#!/bin/bash
set -u
out=gpurun_out/r2sl3; mkdir -p "$out"
python -c "import __graft_entry__ as g; g.build()" > "$out/build.log" 2>&1 || { echo BUILD FAILED; tail -5 "$out/build.log"; exit 9; }
bash tools/dbg/sweep.sh "ZMO_SEED_WARPS=22" "ZMO_SEED_WARPS=11" "ZMO_SEED_WARPS=11 ZMO_SEED_CTAS=2" "ZMO_SEED_WARPS=8 ZMO_SEED_CTAS=2" "ZMO_SEED_WARPS=16" "ZMO_SEED_LANES=4 ZMO_SEED_WARPS=11" "ZMO_SEED_LANES=4 ZMO_SEED_WARPS=11 ZMO_SEED_CTAS=2" "ZMO_SEED_WARPS=22 ZMO_DEPTH=3" "ZMO_SEED_WARPS=11 ZMO_DEPTH=3" 2>&1 | tee "$out/sweep.txt"
