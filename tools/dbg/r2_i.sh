#!/bin/bash
# round 2, GPU call I (1 GPU): anchor cache parity + kernel time, pipeline fill/drain: shard size and batch ramp
set -u
out=gpurun_out/r2i; mkdir -p "$out"
G=tools/_build/gen_reads
python -c "import __graft_entry__ as g; g.build()" > "$out/build.log" 2>&1 || { echo BUILD FAILED; tail -5 "$out/build.log"; exit 9; }
timeout 900 python -m pytest tests/test_gpu_wtzmo.py tests/test_gpu_dp.py -q -m gpu -x -k "not scale_200k and not properties and not cfg3 and not full_bench_step" > "$out/pytest_gpu.log" 2>&1; echo "pytest rc=$?" >> "$out/pytest_gpu.log"
tail -3 "$out/pytest_gpu.log"
FA=/dev/shm/cfg2.fa
[ -f $FA ] || $G -n 50000 -L 10000 -G 4600000 -m pacbio -s 20240603 -o $FA
W=smartdenovo_b200/bin/wtzmo
ARGS="-t 1 -i $FA -f -o /dev/shm/o.ovl -k 16 -s 200 -m 0.6 -P 40 -p 0"
ZMO_PIPELINE=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/launches_cfg2_P40.csv $W $ARGS >/dev/null 2>&1
python tools/launch_summary.py $out/launches_cfg2_P40.csv 6
timeout 1500 bash tools/dbg/sweep.sh "ZMO_RAMP=0" "ZMO_RAMP=64" "ZMO_RAMP=96" "ZMO_RAMP=128" "ZMO_RAMP=96 ZMO_DEPTH=3" "ZMO_BENCH_SHARDS=5" "ZMO_BENCH_SHARDS=5 ZMO_RAMP=96" "ZMO_BENCH_SHARDS=2" > "$out/sweep.log" 2>&1
cat "$out/sweep.log"
