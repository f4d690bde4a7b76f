#!/bin/bash
# GPU-box profiling recipe (B200_PROFILING.md): launch list + full captures of the top kernels on one cfg2-shaped shard.
# usage: tools/prof.sh <tag> [kernel-regex:skip:count ...]      (regex matched against the demangled kernel name)
set -u
TAG=${1:-r01}; shift || true
G=tools/_build/gen_reads
FA=/dev/shm/c2s.fa
[ -f $FA ] || $G -n 5000 -L 10000 -G 460000 -m pacbio -s 20240603 -o $FA
W=smartdenovo_b200/bin/wtzmo
ARGS="-t 1 -i $FA -f -o /dev/shm/o.ovl -k 16 -s 200 -m 0.6 -P 5 -p 0"
mkdir -p gpurun_out
export ZMO_DEPTH=${ZMO_DEPTH:-0}     # one batch at a time under the profiler: kernels are serialised anyway
ZMO_PIPELINE=0 $W $ARGS 2>/dev/null
ZMO_PIPELINE=0 ZMO_STATS=gpurun_out/stats_$TAG.json $W $ARGS 2>/dev/null
ZMO_PIPELINE=0 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv $W $ARGS >/dev/null 2>&1
for SPEC in "$@"; do
  K="${SPEC%%:*}"; REST="${SPEC#*:}"; SKIP="${REST%%:*}"; CNT="${REST#*:}"
  N=$(echo "$K" | tr -c 'A-Za-z0-9_' '_')
  ZMO_PIPELINE=0 timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$K" -s "$SKIP" -c "$CNT" -f -o "gpurun_out/prof_${TAG}_$N" $W $ARGS >/dev/null 2>gpurun_out/ncu_err_$N.txt
done
ls -la gpurun_out | tail -12
