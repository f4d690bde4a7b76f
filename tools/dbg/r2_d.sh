#!/bin/bash
# round 2, GPU call D (1 GPU): compile-time experiments of k_window_align timed on one cfg2 shard (serialised stages, CUDA events), then
# source-level ncu captures of the two heaviest kernels
set -u
out=gpurun_out/r2d; mkdir -p "$out"
G=tools/_build/gen_reads
python -c 'import __graft_entry__ as g; g.build()' > "$out/build.log" 2>&1
FA=/dev/shm/cfg2.fa
[ -f $FA ] || $G -n 50000 -L 10000 -G 4600000 -m pacbio -s 20240603 -o $FA
W=smartdenovo_b200/bin/wtzmo
ARGS="-t 1 -i $FA -f -o /dev/shm/o.ovl -k 16 -s 200 -m 0.6 -P 40 -p 0"
run2(){ # tag
  ZMO_PIPELINE=0 $W $ARGS 2>/dev/null; ZMO_PIPELINE=0 ZMO_STATS=$out/stats_$1.json $W $ARGS 2>/dev/null
  md5sum /dev/shm/o.ovl | cut -c1-32 > $out/md5_$1.txt
  python - "$1" <<'PY'
import json,sys
d=json.load(open("gpurun_out/r2d/stats_%s.json"%sys.argv[1])); print(sys.argv[1], open("gpurun_out/r2d/md5_%s.txt"%sys.argv[1]).read().strip()[:8], "overlap_s", d["overlap_s"], {k:round(v,1) for k,v in d["stage_ms"].items()})
PY
}
run2 default
for defs in "-DZMO_EXP_WA_C4" "-DZMO_EXP_WALK_RUNS" "-DZMO_EXP_ANCHOR_REGS" "-DZMO_EXP_WALK_RUNS -DZMO_EXP_ANCHOR_REGS" "-DZMO_EXP_WALK_RUNS -DZMO_EXP_ANCHOR_REGS -DZMO_EXP_WA_C4"; do
  tag=$(echo "$defs" | tr -d ' ' | tr -c 'A-Za-z0-9_\n' '_')
  ZMO_NVCC_DEFINES="$defs" python -c 'import __graft_entry__ as g; g.build()' > "$out/build$tag.log" 2>&1
  run2 "x$tag"
done
python -c 'import __graft_entry__ as g; g.build()' >> "$out/build.log" 2>&1
ZMO_FINISH_WARP=1 run2 finish_warp
for K in k_p_seed k_window_align; do
  ZMO_PIPELINE=0 timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$K" -s 2 -c 1 -f -o "$out/prof_$K" $W $ARGS >/dev/null 2>"$out/ncu_err_$K.txt"
done
ls -la $out | tail -20
