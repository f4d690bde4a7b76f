#!/bin/bash
# smoke at > 2^32 bases: 600,000 reads x 10 kb (6.3 Gbp), one small query shard
mkdir -p gpurun_out
FA=/dev/shm/huge.fa
( time tools/_build/gen_reads -n 600000 -L 10000 -G 60000000 -m pacbio -s 20240606 -o $FA ) 2>&1 | grep real
ls -la $FA
W=smartdenovo_b200/bin/wtzmo
( time ZMO_STATS=gpurun_out/stats_huge.json $W -t 1 -i $FA -fo /dev/shm/huge.ovl -k 16 -s 200 -m 0.6 -P 2000 -p 0 ) 2>&1 | grep -E "Done|real|index|wtzmo\(b200\)"
wc -l /dev/shm/huge.ovl; cut -f1-16 /dev/shm/huge.ovl | head -3
cat gpurun_out/stats_huge.json
nvidia-smi --query-gpu=memory.used --format=csv,noheader
