#!/bin/bash
# round 2: bridge pipeline iteration -- golden test subset, bench, launch list of the cfg2 -P 40 shard
set -u
out=gpurun_out/r2wb2; mkdir -p "$out"
python -c "import __graft_entry__ as g; g.build()" > "$out/build.log" 2>&1 || { echo BUILD FAILED; tail -5 "$out/build.log"; exit 9; }
timeout 900 python -m pytest tests/test_gpu_wtzmo.py -q -m gpu -x -k "sw_small or cfg1 or cfg2_bench or nondefault or edge" > "$out/pytest_gpu.log" 2>&1; echo "pytest rc=$?" >> "$out/pytest_gpu.log"
tail -3 "$out/pytest_gpu.log"
timeout 600 python bench.py --steps 3 --warmup 3 --no-sub > "$out/bench.json" 2> "$out/bench.err" || tail -5 "$out/bench.err"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2wb2/bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "ms", d["ms_per_step"], "parity", d["parity_checked"], d["parity"].get("ok"), "launches", d["gpu_launches"])
print("   stage", {k:round(v) for k,v in d["stage_ms_per_step"].items()})
PY
G=tools/_build/gen_reads; FA=/dev/shm/cfg2.fa
[ -f $FA ] || $G -n 50000 -L 10000 -G 4600000 -m pacbio -s 20240603 -o $FA
W=smartdenovo_b200/bin/wtzmo
ARGS="-t 1 -i $FA -f -o /dev/shm/o.ovl -k 16 -s 200 -m 0.6 -P 40 -p 0"
export ZMO_PIPELINE=0
$W $ARGS 2>/dev/null; md5sum /dev/shm/o.ovl
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/launches_cfg2_P40.csv $W $ARGS >/dev/null 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(l for l in open("gpurun_out/r2wb2/launches_cfg2_P40.csv") if l.startswith('"'))]
h=rows[0]; ki=h.index("Kernel Name"); vi=h.index("Metric Value")
t=collections.defaultdict(lambda:[0,0.0,0.0])
for r in rows[1:]:
    v=float(r[vi].replace(",",""))/1e6; k=r[ki][:60]; t[k][0]+=1; t[k][1]+=v; t[k][2]=max(t[k][2],v)
tot=sum(v[1] for v in t.values())
for k,v in sorted(t.items(), key=lambda kv:-kv[1][1])[:16]: print("%-60s %5d %9.2f %8.2f %5.1f%%"%(k,v[0],v[1],v[2],100*v[1]/tot))
print("total",tot)
PY
