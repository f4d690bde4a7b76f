#!/bin/bash
# round 2, GPU call C (2 GPUs): product-level multi-GPU mode + the torchrun bench with per-rank golden parity
set -u
out=gpurun_out/r2c2; mkdir -p "$out"
python -c 'import __graft_entry__ as g; g.build()' > "$out/build.log" 2>&1
nvidia-smi -L > "$out/gpus.txt"
timeout 600 python -m pytest tests/test_gpu_wtzmo.py -q -m gpu -x -k "multi_gpu or two_jobs" > "$out/pytest_multi.log" 2>&1; echo "pytest rc=$?" >> "$out/pytest_multi.log"
tail -3 "$out/pytest_multi.log"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 2 > "$out/bench_2gpu.json" 2> "$out/bench_2gpu.err"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2c2/bench_2gpu.json").read().strip().splitlines()[-1])
print("bench 2gpu value", d["value"], "ms", d["ms_per_step"], "parity", d["parity_checked"], d["parity"], "gathered", d.get("gathered_bytes"))
PY
FA=$(ls /dev/shm/zmo_bench/reads_50000_*.fa | head -1)
W=smartdenovo_b200/bin/wtzmo
# product binary: two bench-sized shards (-P 10 -p 0,1) on one GPU each vs the same two jobs one after another on one GPU
( time ZMO_GPUS=2 ZMO_STATS=$out/stats_multi2.json $W -t 1 -i $FA -f -o /dev/shm/m2.ovl -k 16 -s 200 -m 0.6 -P 5 -p 0 ) 2> "$out/run_multi2.err"
( time ZMO_STATS=$out/stats_single_p0.json $W -t 1 -i $FA -f -o /dev/shm/s0.ovl -k 16 -s 200 -m 0.6 -P 10 -p 0 ) 2> "$out/run_single_p0.err"
( time ZMO_STATS=$out/stats_single_p1.json $W -t 1 -i $FA -f -o /dev/shm/s1.ovl -k 16 -s 200 -m 0.6 -P 10 -p 1 ) 2> "$out/run_single_p1.err"
cat /dev/shm/s0.ovl /dev/shm/s1.ovl | md5sum > "$out/md5.txt"; md5sum /dev/shm/m2.ovl >> "$out/md5.txt"; cat "$out/md5.txt"
python - <<'PY'
import json, hashlib
g=json.load(open("tests/golden/scale_digests.json"))["cfg2_P10_p0"]
print("single -P 10 -p 0 vs reference golden:", hashlib.md5(open("/dev/shm/s0.ovl","rb").read()).hexdigest()==g["md5"])
PY
tail -4 "$out/run_multi2.err"; cat "$out/stats_multi2.json"
