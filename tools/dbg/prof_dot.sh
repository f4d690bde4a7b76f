#!/bin/bash
# launch list of one dot-matrix (-U) shard: 5,000 ONT-like reads x 10 kb
mkdir -p gpurun_out
FA=/dev/shm/c3s.fa
tools/_build/gen_reads -n 5000 -L 10000 -G 460000 -m ont -s 20240604 -o $FA
W=smartdenovo_b200/bin/wtzmo
ARGS="-t 1 -i $FA -f -o /dev/shm/o.ovl -k 16 -z 10 -Z 16 -U -1 -m 0.1 -A 1000 -P 5 -p 0"
ZMO_PIPELINE=0 $W $ARGS 2>&1 | tail -1
ZMO_PIPELINE=0 ZMO_STATS=gpurun_out/stats_dot.json $W $ARGS 2>/dev/null; cat gpurun_out/stats_dot.json
ZMO_PIPELINE=0 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_dot.csv $W $ARGS >/dev/null 2>&1
ZMO_PIPELINE=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_p_dot -c 1 -f -o gpurun_out/prof_dot_k_p_dot $W $ARGS >/dev/null 2>&1
