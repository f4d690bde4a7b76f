#!/bin/bash
# product binary on one shard of the 200k-read set; leaves digests + the 16-column records in gpurun_out/
mkdir -p gpurun_out
FA=/dev/shm/big.fa
tools/_build/gen_reads -n 200000 -L 10000 -G 20000000 -m pacbio -s 20240605 -o $FA
W=smartdenovo_b200/bin/wtzmo
( time $W -t 1 -i $FA -fo /dev/shm/gpu_sw.ovl -k 16 -s 200 -m 0.6 -P 400 -p 0 ) 2>&1 | grep -E "Done|real"
md5sum /dev/shm/gpu_sw.ovl | tee gpurun_out/big_gpu.md5; cut -f1-16 /dev/shm/gpu_sw.ovl | md5sum | tee -a gpurun_out/big_gpu.md5
cut -f1-16 /dev/shm/gpu_sw.ovl | gzip > gpurun_out/big_gpu_sw.ovl16.gz
