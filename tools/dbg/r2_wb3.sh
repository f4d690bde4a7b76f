#!/bin/bash
# round 2: pipeline parameters with the bridge-level window alignment in place; ncu --set full of the sweep
set -u
out=gpurun_out/r2wb3; mkdir -p "$out"
python -c "import __graft_entry__ as g; g.build()" > "$out/build.log" 2>&1 || { echo BUILD FAILED; tail -5 "$out/build.log"; exit 9; }
bash tools/dbg/sweep.sh "ZMO_DEPTH=2" "ZMO_DEPTH=3" "ZMO_DEPTH=4" "ZMO_BATCH_READS=512" "ZMO_BATCH_READS=256 ZMO_DEPTH=3" "ZMO_WAVE0=5" "ZMO_WAVE0=12" "ZMO_RAMP=64" "ZMO_RAMP=160" "ZMO_DEPTH=2" 2>&1 | tee "$out/sweep.txt"
G=tools/_build/gen_reads; FA=/dev/shm/cfg2.fa
[ -f $FA ] || $G -n 50000 -L 10000 -G 4600000 -m pacbio -s 20240603 -o $FA
W=smartdenovo_b200/bin/wtzmo
ARGS="-t 1 -i $FA -f -o /dev/shm/o.ovl -k 16 -s 200 -m 0.6 -P 40 -p 0"
export ZMO_PIPELINE=0
$W $ARGS 2>/dev/null
for K in k_wb_sweep k_wb_stitch k_wb_walk k_wb_prep; do
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$K" -s 1 -c 1 -f -o "$out/prof_$K" $W $ARGS >/dev/null 2>$out/ncu_err_$K.txt
done
ls -la $out
