#!/usr/bin/env python
"""bench.py -- aligned Gbp/s of the all-vs-all overlap hot path (BASELINE.json metric).

A *step* is one complete overlap job `wtzmo -P n -p i` (the reference's own query sharding, wtzmo.c:1291,1314)
over the resident read set: k-mer index build + candidate query + z-mer seeding + banded DP + state replay +
.ovl text, for the 1/n of the query reads with rd_id % n == i.  Steps use consecutive shard indices, so every
step is different work of the same shape.  The read set is synthetic (tools/gen_reads.c, seed 20240601+cfg).

  value : whole-job aligned bp / time, reads already resident in HBM when the timed region starts
  e2e   : same, but the timed region of every step also re-uploads the packed reads from host memory and, like
          `value`, brings every record + CIGAR back to the host and formats the 17-column .ovl text
  roofline : dominant kernel (banded end-extension DP): algorithmic bytes = 0.5 B per DP cell (4 traceback bits)
          + packed sequence bytes, over the CUDA-event time of that kernel inside the timed steps
  cpu_baseline : the unmodified reference binary (oracle/_ref/wtzmo -t <cores>) on a bounded sub-shard

`--impl reference` times only the reference CPU binary on the same configuration.
Under torchrun (N>1) every rank owns a GPU and its own shard sequence -- the shard size is the same for every N (weak
scaling, no data-path collective); the last step's records are gathered to rank 0 with one NCCL all-gather at the end of
the timed region.
"""
import argparse
import ctypes as C
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

# configs[1] of BASELINE.json: 50k PacBio-like reads x 10 kb over a 4.6 Mb genome (108x), -k 16 -s 200 -m 0.6
WORKLOADS = {
    "cfg1": dict(n=2000, L=8000, G=500000, model="pacbio", seed=20240602, flags=["-k", "16", "-s", "200", "-m", "0.6"], shards=1),
    "cfg2": dict(n=50000, L=10000, G=4600000, model="pacbio", seed=20240603, flags=["-k", "16", "-s", "200", "-m", "0.6"], shards=int(os.environ.get("ZMO_BENCH_SHARDS", "10"))),
    "cfg2s": dict(n=5000, L=10000, G=460000, model="pacbio", seed=20240603, flags=["-k", "16", "-s", "200", "-m", "0.6"], shards=5),
}


def sh(cmd, **kw):
    return subprocess.run(cmd, check=True, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, **kw)


def ensure_reads(wl, tmpdir):
    gen = os.path.join(REPO, "tools", "_build", "gen_reads")
    fa = os.path.join(tmpdir, "reads_%d_%d_%d_%s_%d.fa" % (wl["n"], wl["L"], wl["G"], wl["model"], wl["seed"]))
    if not os.path.exists(fa):
        sh([gen, "-n", str(wl["n"]), "-L", str(wl["L"]), "-G", str(wl["G"]), "-m", wl["model"], "-s", str(wl["seed"]), "-o", fa + ".tmp"])
        os.replace(fa + ".tmp", fa)
    return fa


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.stop_flag = gpu, [], False

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,utilization.gpu"
        while not self.stop_flag:
            try:
                r = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits"], stdout=subprocess.PIPE, text=True, timeout=5)
                self.rows.append([x.strip() for x in r.stdout.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        ut = sorted(int(r[6]) for r in self.rows if len(r) > 6 and r[6].isdigit())
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(self.rows),
                "gpu_util_pct_median": ut[len(ut) // 2] if ut else None}


def run_reference(fa, flags, n_job, i_job, threads, outdir, tag):
    """unmodified reference binary on one shard; returns (aligned bp, overlap seconds)"""
    ref = os.path.join(REPO, "oracle", "_ref", "wtzmo")
    out = os.path.join(outdir, "ref_%s.ovl" % tag)
    t0 = time.time()
    r = subprocess.run([ref, "-t", str(threads), "-i", fa, "-f", "-o", out, "-P", str(n_job), "-p", str(i_job)] + flags, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    wall = time.time() - t0
    if r.returncode != 0:
        raise RuntimeError("reference wtzmo failed: " + r.stderr[-500:])
    cols = 0
    with open(out) as f:
        for line in f:
            c = line.split("\t")
            cols += int(c[12]) + int(c[13]) + int(c[14]) + int(c[15])
    # overlap phase = "calculating overlaps" -> "Done" (wtzmo.c:1777-1780); the stamps have 1 s resolution, so
    # subtract the measured load phase instead: time from start to the "calculating overlaps" line is not
    # recoverable exactly, so report whole-process wall for small samples (load is < 2% of it)
    return cols, wall


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("ZMO_BENCH_WORKLOAD", "cfg2"))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cores = os.cpu_count() or 1
    tmpdir = os.environ.get("ZMO_BENCH_TMP", "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir())
    tmpdir = os.path.join(tmpdir, "zmo_bench")
    os.makedirs(tmpdir, exist_ok=True)
    metric = "aligned Gbp/sec (all-vs-all overlap)"
    config = {"workload": "%s: %d synthetic %s reads x %d bp, genome %d bp, wtzmo %s; step = one `-P %d -p i` query shard incl. index build" % (
        args.workload, wl["n"], wl["model"], wl["L"], wl["G"], " ".join(wl["flags"]), max(wl["shards"], world)),
        "n_reads": wl["n"], "read_len": wl["L"], "genome": wl["G"], "shards": max(wl["shards"], world),
        "l2": "inputs larger than L2 per step (every step streams a different shard: new candidates, match lists and traceback)"}

    import __graft_entry__ as ge
    if rank == 0:
        ge.build()
    if args.impl == "reference":
        if rank != 0:
            return 0
        fa = ensure_reads(wl, tmpdir)
        # bounded sample: a sub-shard of one step's shard so that K+W steps finish within minutes
        sub = int(os.environ.get("ZMO_REF_SUBSHARD", "2"))       # 1 = exactly the shard a step of our arm processes
        n_job = wl["shards"] * sub
        times, bp = [], 0
        for s in range(args.warmup + args.steps):
            cols, wall = run_reference(fa, wl["flags"], n_job, s % n_job, cores, tmpdir, "r%d" % s)
            if s >= args.warmup:
                times.append(wall)
                bp += cols
        total = sum(times)
        val = bp / total / 1e9 if total > 0 else 0.0
        line = {"impl": "reference", "metric": metric, "value": val, "unit": "Gbp/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * total / max(1, len(times)), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
                "config": config,
                "cpu_baseline": {"value": val, "unit": "Gbp/s", "cores": cores, "kind": "reference", "sample": "each step = `oracle/_ref/wtzmo -t %d -P %d -p i`: 1/%d of the query shard of one step of our arm, full index rebuilt per step like ours" % (cores, n_job, sub)},
                "e2e": {"value": val, "unit": "Gbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    from smartdenovo_b200 import dist as zdist
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist.barrier()        # rank 0 finished build()
    torch.cuda.set_device(local_rank)
    os.environ["ZMO_DEVICE"] = str(local_rank)
    fa = ensure_reads(wl, tmpdir) if rank == 0 else None
    if world > 1:
        dist.barrier()
    fa = ensure_reads(wl, tmpdir)

    host = C.CDLL(os.path.join(REPO, "smartdenovo_b200", "lib", "libwtzmo_host.so"))
    host.wz_open.restype = C.c_void_p
    host.wz_open.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int)]
    host.wz_upload.argtypes = [C.c_void_p]
    host.wz_run.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_char_p]
    host.wz_stats.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    host.wz_close.argtypes = [C.c_void_p]
    out_path = os.path.join(tmpdir, "bench_rank%d.ovl" % rank)
    argv = [b"wtzmo", b"-t", b"1", b"-i", fa.encode(), b"-f", b"-o", out_path.encode()] + [x.encode() for x in wl["flags"]]
    arr = (C.c_char_p * len(argv))(*argv)
    rc = C.c_int(0)
    S = host.wz_open(len(argv), arr, C.byref(rc))
    if not S:
        raise RuntimeError("wz_open failed rc=%d (no GPU / build missing: there is no CPU fallback)" % rc.value)
    n_job = max(wl["shards"], world)      # the shard size (work per GPU per step) does not depend on the number of GPUs: weak scaling

    def stats():
        a = (C.c_double * host.wz_stats_n())()
        host.wz_stats(S, a)
        return list(a)

    def step(idx, reupload):
        if reupload and host.wz_upload(S):
            raise RuntimeError("upload failed")
        shard = zdist.shard_of(idx, rank, world, n_job)
        if host.wz_run(S, n_job, shard, out_path.encode()):
            raise RuntimeError("wz_run failed")
        return stats()

    if host.wz_upload(S):
        raise RuntimeError("upload failed")

    def timed(nsteps, first_idx, reupload):
        """K steps bracketed by barrier + synchronize; device-side wall via CUDA events on the default stream + host wall"""
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        st0 = stats()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t0 = time.perf_counter()
        bp = rec = 0
        d2h0, h2d0 = st0[24], st0[23]
        for k in range(nsteps):
            st = step(first_idx + k, reupload)
            bp += st[1]
            rec += st[0]
        gathered = None
        if world > 1:
            gathered, _blob = zdist.gather_record_file(out_path, device="cuda")     # one NCCL all-gather of sizes + one of records, to rank 0
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        wall = time.perf_counter() - t0
        st1 = stats()
        if world > 1:
            wall = zdist.max_over_ranks([wall], device="cuda")[0]
            bp, rec = zdist.sum_over_ranks([float(bp), float(rec)], device="cuda")
        return dict(wall=wall, bp=bp, rec=rec, st0=st0, st1=st1, h2d=(st1[23] - h2d0) / nsteps, d2h=(st1[24] - d2h0) / nsteps, gathered=gathered)

    for w in range(args.warmup):
        step(w, False)
    if world > 1:
        zdist.gather_record_file(out_path, device="cuda")      # warm-up of the gather path too (NCCL channels, pinned staging)
    sampler = ClockSampler(local_rank)
    sampler.start()
    r_val = timed(args.steps, args.warmup, False)
    r_e2e = timed(args.steps, args.warmup + args.steps, True)
    sampler.stop_flag = True
    sampler.join(timeout=2)
    value = r_val["bp"] / r_val["wall"] / 1e9
    e2e = r_e2e["bp"] / r_e2e["wall"] / 1e9
    # roofline of the dominant kernel over the `value` region (stage timers are CUDA events on the library stream)
    names = ["index", "candidates", "pair_windows", "window_align", "gap_global", "end_extend", "dotmatrix", "copy"]
    st0, st1 = r_val["st0"], r_val["st1"]
    stage_ms = {n: st1[10 + i] - st0[10 + i] for i, n in enumerate(names)}
    stage_ms["dp_phase_wall"] = st1[31] - st0[31]
    cells = {"end_extend": st1[18] - st0[18], "window_align": st1[19] - st0[19], "gap_global": st1[20] - st0[20]}
    # dominant DP kernel group: the end-extension + gap-fill executors run CONCURRENTLY (one stream per executor class), so
    # their cost is the wall time of that phase (CUDA events on the library stream), not the sum of overlapping kernels
    groups = {"dp_phase": (stage_ms["dp_phase_wall"], cells["end_extend"] + cells["gap_global"], "k_ext_cta<64|128|256,1> + k_ext_warp<1> + k_glb_warp/k_glb_cta (concurrent)"),
              "window_align": (stage_ms["window_align"], cells["window_align"], "k_window_align")}
    dom = max(groups, key=lambda k: groups[k][0])
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    launches = st1[9] - st0[9]
    dom_s = groups[dom][0] / 1e3
    alg_bytes = 0.5 * groups[dom][1]
    achieved = alg_bytes / dom_s / 1e9 if dom_s > 0 else 0.0
    roof = {"bound": "hbm", "kernel": groups[dom][2],
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if peak else None,
            # DRAM bytes (read + write) of ONE captured launch of the group's kernel: ncu --set full, profiles/r01_ncu_final.md
            # (k_window_align launch id 0: 144.5 + 233.3 MB; k_ext_cta<128,13,1>: 26.2 + 207.4 MB); see traffic_profile for its algorithmic bytes
            "traffic": {"window_align": 377.8e6, "dp_phase": 233.6e6}[dom],
            "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 (of fallback)",
            "alg_bytes_per_cell": 0.5, "cells_per_step": groups[dom][1] / args.steps, "kernel_ms_per_step": groups[dom][0] / args.steps,
            "gcells_per_s": groups[dom][1] / dom_s / 1e9 if dom_s > 0 else 0.0,
            "traffic_profile": {"kernel": "k_window_align", "dram_bytes_per_launch": 377.8e6, "launch_ms": 17.7, "alg_bytes_per_launch_est": 0.85e9,
                                "source": "profiles/r01_ncu_final.md (ncu --set full, one launch on the cfg2s shard; short bridges keep their traceback in shared memory)"},
            "note": "integer DP is ALU/latency bound, not HBM bound (see DESIGN.md): gcells_per_s is the number to optimise; kernel time = CUDA-event stage time summed over the contexts in flight"}
    line = {"metric": metric, "value": value, "unit": "Gbp/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * r_val["wall"] / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": config, "clocks": sampler.summary(),
            "e2e": {"value": e2e, "unit": "Gbp/s", "h2d_bytes_per_step": r_e2e["h2d"], "d2h_bytes_per_step": r_e2e["d2h"], "ms_per_step": 1e3 * r_e2e["wall"] / args.steps},
            "gpu_launches": int(launches), "roofline": roof,
            "stage_ms_per_step": {k: v / args.steps for k, v in stage_ms.items()},
            "last_step_host_ms": {"device_calls": 1e3 * st1[3], "replay_format": 1e3 * st1[4]},
            "records_per_step": r_val["rec"] / args.steps, "aligned_bp_per_step": r_val["bp"] / args.steps,
            "last_step_work": {"batches": st1[5], "pairs_seeded": st1[6], "pairs_aligned": st1[7], "alignments_consumed": st1[8], "demand_waves": st1[29], "demand_wave_tasks": st1[30],
                               # speculation of the batch pipeline: reads put into batches, reads masked by the time of their turn, candidates of those reads (seeded for nothing)
                               "reads_batched": st1[32], "reads_late_masked": st1[33], "cands_late_masked": st1[34]}}
    if world > 1:
        line["gathered_bytes"] = r_e2e["gathered"]
    if rank == 0 and world == 1 and not args.no_cpu_baseline and os.path.exists(os.path.join(REPO, "oracle", "_ref", "wtzmo")):
        sub = int(os.environ.get("ZMO_REF_SUBSHARD", "16"))
        try:
            cols, wall = run_reference(fa, wl["flags"], n_job * sub, 0, cores, tmpdir, "cpu")
            line["cpu_baseline"] = {"value": cols / wall / 1e9, "unit": "Gbp/s", "cores": cores, "kind": "reference",
                                    "sample": "`oracle/_ref/wtzmo -t %d -P %d -p 0` = 1/%d of one bench step's shard, full index build included, %.1f s" % (cores, n_job * sub, sub, wall)}
        except Exception as e:   # noqa: BLE001
            line["cpu_baseline"] = {"value": None, "unit": "Gbp/s", "cores": cores, "kind": "reference", "sample": "failed: %s" % e}
    host.wz_close(S)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
