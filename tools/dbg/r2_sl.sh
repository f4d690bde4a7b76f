#!/bin/bash
# round 2: pair seeding with several pairs per warp (k_p_seed_lanes) -- golden tests, A/B bench over ZMO_SEED_LANES, launch list
set -u
out=gpurun_out/r2sl; mkdir -p "$out"
python -c "import __graft_entry__ as g; g.build()" > "$out/build.log" 2>&1 || { echo BUILD FAILED; tail -5 "$out/build.log"; exit 9; }
timeout 900 python -m pytest tests/test_gpu_wtzmo.py -q -m gpu -x -k "sw_small or cfg1 or cfg2_bench or nondefault or edge or side_inputs" > "$out/pytest_gpu.log" 2>&1; echo "pytest rc=$?" >> "$out/pytest_gpu.log"
tail -3 "$out/pytest_gpu.log"
bash tools/dbg/sweep.sh "ZMO_SEED_LANES=0" "ZMO_SEED_LANES=2" "ZMO_SEED_LANES=4" "ZMO_SEED_LANES=8" "ZMO_SEED_LANES=16" "ZMO_SEED_LANES=32" "ZMO_SEED_LANES=-1" 2>&1 | tee "$out/sweep.txt"
G=tools/_build/gen_reads; FA=/dev/shm/cfg2.fa
[ -f $FA ] || $G -n 50000 -L 10000 -G 4600000 -m pacbio -s 20240603 -o $FA
W=smartdenovo_b200/bin/wtzmo
ARGS="-t 1 -i $FA -f -o /dev/shm/o.ovl -k 16 -s 200 -m 0.6 -P 40 -p 0"
export ZMO_PIPELINE=0
$W $ARGS 2>/dev/null; md5sum /dev/shm/o.ovl
for g in 4 8 32; do
ZMO_SEED_LANES=$g ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/launches_g$g.csv $W $ARGS >/dev/null 2>&1
grep "k_p_seed" $out/launches_g$g.csv | python -c "
import sys,csv
t=[float(r[-1].replace(',',''))/1e6 for r in csv.reader(sys.stdin)]
print('G=$g k_p_seed launches',len(t),'total ms',sum(t),'max',max(t))
"
done
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:k_p_seed_lanes" -s 1 -c 1 -f -o "$out/prof_k_p_seed_lanes" $W $ARGS >/dev/null 2>$out/ncu_err.txt
ls $out
