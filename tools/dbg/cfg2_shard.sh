#!/bin/bash
# one cfg2 bench step through the product binary with stats (usage: tools/dbg/cfg2_shard.sh <tag> [env...])
TAG=$1; shift
FA=/dev/shm/c2.fa
[ -f $FA ] || tools/_build/gen_reads -n 50000 -L 10000 -G 4600000 -m pacbio -s 20240603 -o $FA
for kv in "$@"; do export "$kv"; done
W=smartdenovo_b200/bin/wtzmo
$W -t 1 -i $FA -f -o /dev/shm/o.ovl -k 16 -s 200 -m 0.6 -P 10 -p 0 2>/dev/null
ZMO_STATS=gpurun_out/stats_$TAG.json $W -t 1 -i $FA -f -o /dev/shm/o.ovl -k 16 -s 200 -m 0.6 -P 10 -p 1 2>/dev/null
cat gpurun_out/stats_$TAG.json
