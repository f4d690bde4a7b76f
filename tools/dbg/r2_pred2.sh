#!/bin/bash
set -u
out=gpurun_out/r2pred2; mkdir -p "$out"
python -c "import __graft_entry__ as g; g.build()" > "$out/build.log" 2>&1 || { echo BUILD FAILED; tail -5 "$out/build.log"; exit 9; }
timeout 900 python -m pytest tests/test_gpu_wtzmo.py -q -m gpu -x -k "sw_small or cfg1 or cfg2_bench or cfg2_full_bench or nondefault or edge or refine or side_inputs" > "$out/pytest_gpu.log" 2>&1; echo "pytest rc=$?" >> "$out/pytest_gpu.log"
tail -3 "$out/pytest_gpu.log"
bash tools/dbg/sweep.sh "A=1" "ZMO_BATCH_READS=512" "ZMO_BATCH_READS=320" "ZMO_RAMP=128" "ZMO_RAMP=64" "ZMO_BATCH_READS=448 ZMO_RAMP=112" "A=1" 2>&1 | tee "$out/sweep.txt"
G=tools/_build/gen_reads; FA=/dev/shm/cfg2.fa
[ -f $FA ] || $G -n 50000 -L 10000 -G 4600000 -m pacbio -s 20240603 -o $FA
W=smartdenovo_b200/bin/wtzmo
ARGS="-t 1 -i $FA -f -o /dev/shm/o.ovl -k 16 -s 200 -m 0.6"
for cfg in "A=1" "ZMO_WAVE_PREDICT=-1 ZMO_WAVE0=8"; do
 ( for kv in $cfg; do export "$kv"; done; ZMO_STATS=$out/stats.json $W $ARGS > /dev/null 2> $out/err.txt; echo "[$cfg] rc=$? $(md5sum < /dev/shm/o.ovl)"; python -c "
import json; d=json.load(open('$out/stats.json')); print('   overlap_s', d['overlap_s'], 'total_s', d['total_s'], 'launches', d['launches'], 'tasks', d['tasks'], 'used', d['tasks_used'])" )
done
