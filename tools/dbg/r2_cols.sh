#!/bin/bash
set -u
out=gpurun_out/r2cols; mkdir -p "$out"
python -c "import __graft_entry__ as g; g.build()" > "$out/build.log" 2>&1 || { echo BUILD FAILED; tail -5 "$out/build.log"; exit 9; }
timeout 600 python -m pytest tests/test_gpu_wtzmo.py -q -m gpu -x -k "ovl_16 or sw_small or cfg1" > "$out/pytest_gpu.log" 2>&1; echo "pytest rc=$?" >> "$out/pytest_gpu.log"
tail -3 "$out/pytest_gpu.log"
G=tools/_build/gen_reads; FA=/dev/shm/cfg2.fa
[ -f $FA ] || $G -n 50000 -L 10000 -G 4600000 -m pacbio -s 20240603 -o $FA
W=smartdenovo_b200/bin/wtzmo
ARGS="-t 1 -i $FA -f -k 16 -s 200 -m 0.6 -P 10 -p 0"
$W $ARGS -o /dev/shm/w.ovl > /dev/null 2>&1
for cfg in "A=1" "ZMO_OVL_COLS=16" "A=1" "ZMO_OVL_COLS=16"; do
 ( for kv in $cfg; do export "$kv"; done; ZMO_STATS=$out/stats.json $W $ARGS -o /dev/shm/o_$cfg.ovl > /dev/null 2> $out/err.txt; echo "[$cfg] rc=$? $(wc -c < /dev/shm/o_$cfg.ovl) bytes"; python -c "
import json; d=json.load(open('$out/stats.json')); print('   overlap_s', d['overlap_s'], 'total_s', d['total_s'], 'replay_s', d['replay_s'], 'd2h', d['counters']['d2h_bytes'], 'copy ms', d['stage_ms']['copy'])" )
done
cut -f1-16 "/dev/shm/o_A=1.ovl" | md5sum; md5sum "/dev/shm/o_ZMO_OVL_COLS=16.ovl"
