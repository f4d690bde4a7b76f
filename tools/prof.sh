#!/bin/bash
# GPU-box profiling recipe (B200_PROFILING.md): launch list + `ncu --set full` captures of the top kernels on one query shard of cfg2
# (50,000 reads x 10 kb: `wtzmo -P 40 -p 0`, 1,250 query reads) and one of cfg3s (dot-matrix mode).
# usage: tools/prof.sh <tag> [kernel-regex:skip:count ...]      (regex matched against the demangled kernel name)
set -u
TAG=${1:-r02}; shift || true
G=tools/_build/gen_reads
FA=/dev/shm/cfg2.fa
[ -f $FA ] || $G -n 50000 -L 10000 -G 4600000 -m pacbio -s 20240603 -o $FA
W=smartdenovo_b200/bin/wtzmo
ARGS="-t 1 -i $FA -f -o /dev/shm/o.ovl -k 16 -s 200 -m 0.6 -P 40 -p 0"
out=gpurun_out/prof_$TAG; mkdir -p $out
export ZMO_PIPELINE=0      # one batch at a time under the profiler: kernels are serialised anyway
$W $ARGS 2>/dev/null
ZMO_STATS=$out/stats.json $W $ARGS 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/launches_cfg2_P40.csv $W $ARGS >/dev/null 2>&1
for SPEC in "$@"; do
  K="${SPEC%%:*}"; REST="${SPEC#*:}"; SKIP="${REST%%:*}"; CNT="${REST#*:}"
  N=$(echo "$K" | tr -c 'A-Za-z0-9_' '_')
  timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$K" -s "$SKIP" -c "$CNT" -f -o "$out/prof_$N" $W $ARGS >/dev/null 2>$out/ncu_err_$N.txt
done
# dot-matrix mode: cfg3s shard
FA3=/dev/shm/cfg3s.fa
[ -f $FA3 ] || $G -n 20000 -L 15000 -G 10000000 -m ont -s 20240604 -o $FA3
ARGS3="-t 1 -i $FA3 -f -o /dev/shm/o3.ovl -k 16 -z 10 -Z 16 -U -1 -m 0.1 -A 1000 -P 16 -p 0"
$W $ARGS3 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/launches_cfg3s_dot_P16.csv $W $ARGS3 >/dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:k_p_dot" -s 1 -c 1 -f -o "$out/prof_k_p_dot" $W $ARGS3 >/dev/null 2>$out/ncu_err_k_p_dot.txt
ls -la $out
