#!/bin/bash
# round 2: k_p_seed_lanes with L1 prefetch of the list streams; warp-per-window stitch + sorted walk of the bridge pipeline
set -u
out=gpurun_out/r2sl2; mkdir -p "$out"
python -c "import __graft_entry__ as g; g.build()" > "$out/build.log" 2>&1 || { echo BUILD FAILED; tail -5 "$out/build.log"; exit 9; }
timeout 900 python -m pytest tests/test_gpu_wtzmo.py -q -m gpu -x -k "sw_small or cfg1 or cfg2_bench or nondefault or edge" > "$out/pytest_gpu.log" 2>&1; echo "pytest rc=$?" >> "$out/pytest_gpu.log"
tail -3 "$out/pytest_gpu.log"
bash tools/dbg/sweep.sh "ZMO_SEED_LANES=0" "ZMO_SEED_LANES=4" "ZMO_SEED_LANES=8" "ZMO_SEED_LANES=0" 2>&1 | tee "$out/sweep.txt"
G=tools/_build/gen_reads; FA=/dev/shm/cfg2.fa
[ -f $FA ] || $G -n 50000 -L 10000 -G 4600000 -m pacbio -s 20240603 -o $FA
W=smartdenovo_b200/bin/wtzmo
ARGS="-t 1 -i $FA -f -o /dev/shm/o.ovl -k 16 -s 200 -m 0.6 -P 40 -p 0"
export ZMO_PIPELINE=0
$W $ARGS 2>/dev/null; md5sum /dev/shm/o.ovl
for g in 0 4 8; do
ZMO_SEED_LANES=$g ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/launches_g$g.csv $W $ARGS >/dev/null 2>&1
python tools/launch_summary.py $out/launches_g$g.csv 2>/dev/null | grep -E "k_p_seed|k_wb_|total" | head -8
done
