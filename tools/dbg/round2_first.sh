#!/bin/bash
# First GPU call of the next round (written after round 1's GPU budget was spent): everything prepared on the CPU that still needs a
# device run, in the order of what gates what.  usage: gpurun --timeout 1500 -- 'bash tools/dbg/round2_first.sh'
# Results: gpurun_out/r2_first/*.log
set -u
out=gpurun_out/r2_first; mkdir -p "$out"
python -c 'import __graft_entry__ as g; g.build()' > "$out/build.log" 2>&1
# 1. parity: the GPU suite (incl. k_finish_warp, the wide-band -n fallback and the cfg2 full-size property test, all green once at the end of round 1)
timeout 1200 python -m pytest tests -q -m gpu -x > "$out/pytest_gpu.log" 2>&1; echo "pytest rc=$?" >> "$out/pytest_gpu.log"
# 2. baseline bench line, then the opt-in stitch kernel (4% of the GPU time in r01_ncu_final.md was k_finish)
timeout 600 python bench.py --steps 3 --warmup 3 > "$out/bench_default.json" 2> "$out/bench_default.err"
ZMO_FINISH_WARP=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > "$out/bench_finish_warp.json" 2> "$out/bench_finish_warp.err"
# 3. speculation lag (+36% seeding at batch 384 x depth 2): smaller batches / different depth, with and without the warp stitch
timeout 1500 bash tools/dbg/sweep.sh "ZMO_BATCH_READS=384" "ZMO_BATCH_READS=256" "ZMO_BATCH_READS=192" "ZMO_BATCH_READS=128" \
  "ZMO_BATCH_READS=192 ZMO_DEPTH=3" "ZMO_BATCH_READS=256 ZMO_DEPTH=1" "ZMO_BATCH_READS=256 ZMO_FINISH_WARP=1" "ZMO_WAVE_MASKCHECK=1" "ZMO_WAVE_MASKCHECK=1 ZMO_BATCH_READS=256" > "$out/sweep.log" 2>&1
# 4. compile-time experiment: diagonal runs of the traceback walk taken in one go (reg_walk, zmo_dpr.cuh; bit-exact in the host simulation
#    with ZMO_SIM_DEFINES=-DZMO_EXP_WALK_RUNS): parity on the device first, then the bench; the plain build() at the end restores the default
ZMO_NVCC_DEFINES="-DZMO_EXP_WALK_RUNS" python -c 'import __graft_entry__ as g; g.build()' > "$out/build_walk_runs.log" 2>&1
timeout 900 python -m pytest tests/test_gpu_dp.py tests/test_gpu_wtzmo.py -q -m gpu -x -k "not scale and not cfg2" > "$out/pytest_walk_runs.log" 2>&1; echo "pytest rc=$?" >> "$out/pytest_walk_runs.log"
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > "$out/bench_walk_runs.json" 2> "$out/bench_walk_runs.err"
# 5. second compile-time experiment: window anchors aligned from registers (zmo_winalign.cuh), alone and together with the first
for defs in "-DZMO_EXP_ANCHOR_REGS" "-DZMO_EXP_ANCHOR_REGS -DZMO_EXP_WALK_RUNS"; do
  tag=$(echo "$defs" | tr -d ' ' | tr -c 'A-Za-z0-9_\n' '_')
  ZMO_NVCC_DEFINES="$defs" python -c 'import __graft_entry__ as g; g.build()' > "$out/build$tag.log" 2>&1
  timeout 900 python -m pytest tests/test_gpu_wtzmo.py -q -m gpu -x -k "not scale and not cfg2" > "$out/pytest$tag.log" 2>&1; echo "pytest rc=$?" >> "$out/pytest$tag.log"
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > "$out/bench$tag.json" 2> "$out/bench$tag.err"
done
python -c 'import __graft_entry__ as g; g.build()' >> "$out/build_walk_runs.log" 2>&1
tail -3 "$out/pytest_gpu.log" "$out/pytest_walk_runs.log"; cat "$out/sweep.log"
