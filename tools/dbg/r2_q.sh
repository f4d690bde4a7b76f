#!/bin/bash
set -u
out=gpurun_out/r2q; mkdir -p "$out"
python -c "import __graft_entry__ as g; g.build()" > "$out/build.log" 2>&1 || { echo BUILD FAILED; tail -5 "$out/build.log"; exit 9; }
bash tools/dbg/sweep.sh "ZMO_WB_CHUNK_MB=6144" "ZMO_WB_CHUNK_MB=2048" "ZMO_WB_CHUNK_MB=16384" "ZMO_WB_CHUNK_MB=6144" 2>&1 | tee "$out/sweep.txt"
G=tools/_build/gen_reads; FA=/dev/shm/cfg2.fa
[ -f $FA ] || $G -n 50000 -L 10000 -G 4600000 -m pacbio -s 20240603 -o $FA
W=smartdenovo_b200/bin/wtzmo
ARGS="-t 1 -i $FA -f -o /dev/shm/o.ovl -k 16 -s 200 -m 0.6"
for cfg in "ZMO_WB_CHUNK_MB=6144" "ZMO_WB_CHUNK_MB=2048" "ZMO_WB_CHUNK_MB=16384"; do
 ( for kv in $cfg; do export "$kv"; done; ZMO_STATS=$out/stats.json $W $ARGS > /dev/null 2> $out/err.txt; echo "[$cfg] rc=$? $(md5sum < /dev/shm/o.ovl)"; python -c "
import json; d=json.load(open('$out/stats.json')); print('   overlap_s', d['overlap_s'], 'total_s', d['total_s'], 'launches', d['launches'], 'alloc', d.get('alloc'))" )
done
