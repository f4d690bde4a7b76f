#!/bin/bash
# round 2, GPU call L (1 GPU): lane-per-pair dot-matrix kernel A/B on cfg3s (bench line incl. golden parity), k_ext_cta capture
set -u
out=gpurun_out/r2l; mkdir -p "$out"
python -c "import __graft_entry__ as g; g.build()" > "$out/build.log" 2>&1 || { echo BUILD FAILED; tail -5 "$out/build.log"; exit 9; }
for v in 0 1; do
  ZMO_DOT_LANE=$v timeout 600 python bench.py --workload cfg3s --steps 2 --warmup 1 --no-sub --no-cpu-baseline > "$out/bench_cfg3s_lane$v.json" 2> "$out/bench_cfg3s_lane$v.err"
  python - "$v" <<'PY'
import json,sys
d=json.loads(open("gpurun_out/r2l/bench_cfg3s_lane%s.json"%sys.argv[1]).read().strip().splitlines()[-1])
print("DOT_LANE="+sys.argv[1], "value", round(d["value"],4), "ms", round(d["ms_per_step"]), "parity", d["parity_checked"], {k:round(v) for k,v in d["stage_ms_per_step"].items() if v>1})
PY
done
ZMO_DOT_LANE=1 timeout 600 python -m pytest tests/test_gpu_wtzmo.py -q -m gpu -x -k "dot or cfg1_full or cfg3" > "$out/pytest_dot_lane.log" 2>&1; echo "pytest rc=$?" >> "$out/pytest_dot_lane.log"; tail -3 "$out/pytest_dot_lane.log"
W=smartdenovo_b200/bin/wtzmo
FA3=$(ls /dev/shm/zmo_bench/reads_20000_15000_*.fa | head -1)
ARGS3="-t 1 -i $FA3 -f -o /dev/shm/o3.ovl -k 16 -z 10 -Z 16 -U -1 -m 0.1 -A 1000 -P 16 -p 0"
ZMO_DOT_LANE=1 ZMO_PIPELINE=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/launches_cfg3s_dot_P16_lane.csv $W $ARGS3 >/dev/null 2>&1
python tools/launch_summary.py $out/launches_cfg3s_dot_P16_lane.csv 4
ZMO_DOT_LANE=1 ZMO_PIPELINE=0 timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:k_p_dot_lane" -s 1 -c 1 -f -o "$out/prof_k_p_dot_lane" $W $ARGS3 >/dev/null 2>$out/ncu_err_dotlane.txt
FA=/dev/shm/cfg2.fa; [ -f $FA ] || tools/_build/gen_reads -n 50000 -L 10000 -G 4600000 -m pacbio -s 20240603 -o $FA
ARGS="-t 1 -i $FA -f -o /dev/shm/o.ovl -k 16 -s 200 -m 0.6 -P 40 -p 0"
ZMO_PIPELINE=0 ZMO_RAMP=0 timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:k_ext_cta" -s 8 -c 4 -f -o "$out/prof_k_ext_cta" $W $ARGS >/dev/null 2>$out/ncu_err_ext.txt
ls -la $out | tail -8
