#!/bin/bash
# round 2, GPU call B: the lane-per-window kernel on the device: parity, A/B against the warp kernel, tuning knobs, ncu capture
set -u
out=gpurun_out/r2b; mkdir -p "$out"
python -c 'import __graft_entry__ as g; g.build()' > "$out/build.log" 2>&1
timeout 1200 python -m pytest tests -q -m gpu -x > "$out/pytest_gpu.log" 2>&1; echo "pytest rc=$?" >> "$out/pytest_gpu.log"
tail -3 "$out/pytest_gpu.log"
timeout 1500 bash tools/dbg/sweep.sh "ZMO_WA_WARP=0" "ZMO_WA_WARP=1" "ZMO_WL_EPI=4" "ZMO_WL_EPI=16" "ZMO_WL_EPI=1" "ZMO_WL_CTAS=2" > "$out/sweep.log" 2>&1
cat "$out/sweep.log"
FA=$(ls /dev/shm/zmo_bench/reads_50000_*.fa | head -1)
W=smartdenovo_b200/bin/wtzmo
ARGS="-t 1 -i $FA -f -o /dev/shm/o.ovl -k 16 -s 200 -m 0.6 -P 40 -p 0"
ZMO_PIPELINE=0 ZMO_STATS=$out/stats_P40_nopipe.json $W $ARGS 2> "$out/run_P40.err"
ZMO_PIPELINE=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/launches_cfg2_P40.csv $W $ARGS >/dev/null 2>&1
ZMO_PIPELINE=0 timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:k_wa_lane" -s 2 -c 1 -f -o "$out/prof_wa_lane" $W $ARGS >/dev/null 2>"$out/ncu_err.txt"
python tools/launch_summary.py $out/launches_cfg2_P40.csv 12
ls -la $out
