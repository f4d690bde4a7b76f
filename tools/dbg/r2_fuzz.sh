#!/bin/bash
set -u
out=gpurun_out/r2fuzz; mkdir -p "$out"
python -c "import __graft_entry__ as g; g.build()" > "$out/build.log" 2>&1 || { echo BUILD FAILED; tail -5 "$out/build.log"; exit 9; }
( time timeout 900 python tools/fuzz_oracle_vs_ref.py 110 2026 gpu ) > "$out/fuzz_a.log" 2>&1; tail -4 "$out/fuzz_a.log"
( time ZMO_FUZZ_WIDE=1 timeout 600 python tools/fuzz_oracle_vs_ref.py 50 777 gpu ) > "$out/fuzz_b.log" 2>&1; tail -4 "$out/fuzz_b.log"
