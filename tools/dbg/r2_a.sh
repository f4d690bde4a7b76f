#!/bin/bash
# round 2, GPU call A: state check (full -m gpu suite incl. the new reference goldens), the new bench line, host-policy sweep, launch list
set -u
out=gpurun_out/r2a; mkdir -p "$out"
python -c 'import __graft_entry__ as g; g.build()' > "$out/build.log" 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > "$out/gpu.txt"
nproc >> "$out/gpu.txt"
timeout 1500 python -m pytest tests -q -m gpu -x --durations=12 > "$out/pytest_gpu.log" 2>&1; echo "pytest rc=$?" >> "$out/pytest_gpu.log"
timeout 900 python bench.py --steps 3 --warmup 3 > "$out/bench_default.json" 2> "$out/bench_default.err"
sed -i 's/--no-cpu-baseline 2/--no-cpu-baseline --no-sub 2/' tools/dbg/sweep.sh
timeout 900 bash tools/dbg/sweep.sh "ZMO_FINISH_WARP=1" "ZMO_WAVE_MASKCHECK=1" "ZMO_DEPTH=1" "ZMO_BATCH_READS=256" "ZMO_BATCH_READS=192 ZMO_DEPTH=1" "ZMO_BATCH_READS=256 ZMO_WAVE_MASKCHECK=1 ZMO_FINISH_WARP=1" > "$out/sweep.log" 2>&1
FA=$(ls /dev/shm/zmo_bench/reads_50000_*.fa | head -1)
W=smartdenovo_b200/bin/wtzmo
ARGS="-t 1 -i $FA -f -o /dev/shm/o.ovl -k 16 -s 200 -m 0.6 -P 40 -p 0"
ZMO_STATS=$out/stats_P40.json $W $ARGS 2> "$out/run_P40.err"
ZMO_PIPELINE=0 ZMO_STATS=$out/stats_P40_nopipe.json $W $ARGS 2>> "$out/run_P40.err"
ZMO_PIPELINE=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/launches_cfg2_P40.csv $W $ARGS >/dev/null 2>&1
tail -4 "$out/pytest_gpu.log"; cat "$out/sweep.log"; head -c 1500 "$out/bench_default.json"
