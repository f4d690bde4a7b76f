#!/bin/bash
set -u
out=gpurun_out/r2pred3; mkdir -p "$out"
python -c "import __graft_entry__ as g; g.build()" > "$out/build.log" 2>&1 || { echo BUILD FAILED; tail -5 "$out/build.log"; exit 9; }
G=tools/_build/gen_reads; FA=/dev/shm/cfg2.fa
[ -f $FA ] || $G -n 50000 -L 10000 -G 4600000 -m pacbio -s 20240603 -o $FA
W=smartdenovo_b200/bin/wtzmo
ARGS="-t 1 -i $FA -f -o /dev/shm/o.ovl -k 16 -s 200 -m 0.6 -P 10 -p 0"
for cfg in "A=1" "ZMO_WAVE_PREDICT_B=4" "ZMO_WAVE_PREDICT=1 ZMO_WAVE_PREDICT_B=4" "ZMO_WAVE_PREDICT=3 ZMO_WAVE_PREDICT_B=1"; do
 ( for kv in $cfg; do export "$kv"; done; ZMO_WAVE_DEBUG=1 $W $ARGS > /dev/null 2> $out/err.txt; echo "[$cfg] rc=$? $(md5sum < /dev/shm/o.ovl)"; grep "DP wave\|Done," $out/err.txt | cut -c1-200 )
done
bash tools/dbg/sweep.sh "A=1" "ZMO_WAVE_PREDICT_B=4" "ZMO_WAVE_PREDICT=1 ZMO_WAVE_PREDICT_B=4" "ZMO_WAVE_PREDICT=1 ZMO_WAVE_PREDICT_B=3" "ZMO_WAVE_PREDICT=3 ZMO_WAVE_PREDICT_B=1" "ZMO_WAVE_PREDICT_B=6" "A=1" 2>&1 | tee "$out/sweep.txt"
