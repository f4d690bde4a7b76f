#!/bin/bash
set -u
out=gpurun_out/r2cli2; mkdir -p "$out"
python -c "import __graft_entry__ as g; g.build()" > "$out/build.log" 2>&1 || { echo BUILD FAILED; tail -5 "$out/build.log"; exit 9; }
G=tools/_build/gen_reads; FA=/dev/shm/cfg2.fa
[ -f $FA ] || $G -n 50000 -L 10000 -G 4600000 -m pacbio -s 20240603 -o $FA
W=smartdenovo_b200/bin/wtzmo
ARGS="-t 1 -i $FA -f -o /dev/shm/o.ovl -k 16 -s 200 -m 0.6"
$W $ARGS -P 40 -p 0 >/dev/null 2>&1
for cfg in "ZMO_WAVE_PREDICT=-1 ZMO_WAVE0=8" "A=1" "ZMO_WAVE_PREDICT=2 ZMO_WAVE0=8" "ZMO_WAVE_PREDICT=2 ZMO_WAVE0=16" "ZMO_WAVE_PREDICT=-1 ZMO_WAVE0=8" "A=1"; do
 ( for kv in $cfg; do export "$kv"; done; ZMO_STATS=$out/stats.json $W $ARGS > /dev/null 2> $out/err.txt; echo "[$cfg] rc=$? $(md5sum < /dev/shm/o.ovl)"; python -c "
import json; d=json.load(open('$out/stats.json')); print('   overlap_s', d['overlap_s'], 'total_s', d['total_s'], 'device_call_s', d['device_call_s'], 'replay_s', d['replay_s'], 'launches', d['launches'], 'tasks', d['tasks'], 'batches', d['batches']); print('   ', {k:round(v) for k,v in d['stage_ms'].items()}, d.get('alloc'))" )
done
