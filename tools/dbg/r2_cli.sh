#!/bin/bash
set -u
out=gpurun_out/r2cli; mkdir -p "$out"
python -c "import __graft_entry__ as g; g.build()" > "$out/build.log" 2>&1 || { echo BUILD FAILED; tail -5 "$out/build.log"; exit 9; }
G=tools/_build/gen_reads; FA=/dev/shm/cfg2.fa
[ -f $FA ] || $G -n 50000 -L 10000 -G 4600000 -m pacbio -s 20240603 -o $FA
W=smartdenovo_b200/bin/wtzmo
ARGS="-t 1 -i $FA -f -o /dev/shm/o.ovl -k 16 -s 200 -m 0.6"
i=0
for cfg in "A=1" "ZMO_WB_CHUNK_MB=64" "ZMO_DEPTH=2" "A=1"; do
  i=$((i+1))
  ( for kv in $cfg; do export "$kv"; done; s=$(date +%s.%N); ZMO_STATS=$out/stats_$i.json $W $ARGS > /dev/null 2> $out/err_$i.txt; rc=$?; e=$(date +%s.%N); echo "run $i [$cfg] rc=$rc $(md5sum < /dev/shm/o.ovl) $(wc -l < /dev/shm/o.ovl) lines wall $(echo "$e - $s" | bc)"; grep -v "^\[wtzmo" $out/err_$i.txt | head -5; nvidia-smi --query-gpu=memory.used --format=csv,noheader
    python -c "
import json; d=json.load(open('$out/stats_$i.json')); print('   overlap_s', d['overlap_s'], 'total_s', d['total_s'], 'alloc', d.get('alloc'))" )
done
timeout 600 python -m pytest tests/test_gpu_wtzmo.py -q -m gpu -x -k "cfg2_full_bench or sw_small or cfg1" 2>&1 | tail -2
