#!/bin/bash
# parity + timing at 200k-read scale on one shard: product binary vs the unmodified reference (-t all cores)
set -u
mkdir -p gpurun_out
FA=/dev/shm/big.fa
N=${N:-200000}; G=${G:-20000000}; P=${P:-400}
time tools/_build/gen_reads -n $N -L 10000 -G $G -m pacbio -s 20240605 -o $FA
ls -la $FA
W=smartdenovo_b200/bin/wtzmo
( time ZMO_STATS=gpurun_out/stats_big_sw.json $W -t 1 -i $FA -fo /dev/shm/gpu_sw.ovl -k 16 -s 200 -m 0.6 -P $P -p 0 ) 2>&1 | tail -8
( time oracle/_ref/wtzmo -t $(nproc) -i $FA -fo /dev/shm/ref_sw.ovl -k 16 -s 200 -m 0.6 -P $P -p 0 ) 2>&1 | tail -4
sort /dev/shm/gpu_sw.ovl | md5sum; sort /dev/shm/ref_sw.ovl | md5sum; wc -l /dev/shm/gpu_sw.ovl /dev/shm/ref_sw.ovl
( time ZMO_STATS=gpurun_out/stats_big_dot.json $W -t 1 -i $FA -fo /dev/shm/gpu_dot.ovl -k 16 -z 10 -Z 16 -U -1 -m 0.1 -A 1000 -P $P -p 0 ) 2>&1 | tail -5
( time oracle/_ref/wtzmo -t $(nproc) -i $FA -fo /dev/shm/ref_dot.ovl -k 16 -z 10 -Z 16 -U -1 -m 0.1 -A 1000 -P $P -p 0 ) 2>&1 | tail -4
sort /dev/shm/gpu_dot.ovl | md5sum; sort /dev/shm/ref_dot.ovl | md5sum; wc -l /dev/shm/gpu_dot.ovl /dev/shm/ref_dot.ovl
cat gpurun_out/stats_big_sw.json; cat gpurun_out/stats_big_dot.json
nvidia-smi --query-gpu=memory.used --format=csv
