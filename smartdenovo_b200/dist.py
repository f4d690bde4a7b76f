"""Multi-GPU plumbing for the overlap path: query reads shard across ranks exactly like the reference's
`-P n -p i` jobs (wtzmo.c:1291,1314, usage :1431-1433) -- rank r of N is job r -- with the k-mer index replicated,
so there is NO collective during compute.  The only exchange is the final gather of the variable-length record
text to rank 0 (one all-gather of sizes + one gather of padded byte tensors to the root), equivalent to `cat part*.ovl`.
Works with any torch.distributed backend (NCCL on GPUs, gloo in the CPU tests)."""
import os

import torch
import torch.distributed as dist


def shard_of(step, rank, world, n_jobs):
    """job index (-p) processed by `rank` at bench step `step` when the job space has n_jobs = shards * world entries"""
    return (step * world + rank) % n_jobs


def gather_records(payload: bytes, device="cpu"):
    """All-gather variable-length byte strings; returns the list ordered by rank on every rank."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [payload]
    world = dist.get_world_size()
    n = torch.tensor([len(payload)], dtype=torch.int64, device=device)
    sizes = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(sizes, n)
    mx = max(1, int(max(int(s.item()) for s in sizes)))
    buf = torch.zeros(mx, dtype=torch.uint8, device=device)
    if payload:
        buf[: len(payload)] = torch.frombuffer(bytearray(payload), dtype=torch.uint8).to(device)
    outs = [torch.empty(mx, dtype=torch.uint8, device=device) for _ in range(world)]
    dist.all_gather(outs, buf)
    return [bytes(o[: int(s.item())].cpu().numpy().tobytes()) for o, s in zip(outs, sizes)]


def gather_record_file(path, device="cpu", dst_rank=0):
    """Gather the record files of all ranks on `dst_rank` (`cat part*.ovl` in rank order) with ONE all-gather of sizes and
    ONE gather of the padded byte buffers to dst_rank.  The file is read straight into a page-locked staging tensor, the
    gathered buffer is copied back to the host on dst_rank only.  Returns (total bytes over all ranks, list of per-rank
    uint8 numpy views of the host copy on dst_rank or None elsewhere); the staging tensors come from torch's caching
    pinned allocator, so repeated gathers do not pay cudaHostAlloc again."""
    n = os.path.getsize(path)
    on_gpu = str(device).startswith("cuda")
    if not dist.is_initialized() or dist.get_world_size() == 1:
        import numpy as np
        return n, [np.fromfile(path, dtype=np.uint8)]
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = torch.zeros(world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(sizes, torch.tensor([n], dtype=torch.int64, device=device))
    sizes = [int(x) for x in sizes.cpu()]
    mx = max(1, max(sizes))
    stage = torch.empty(mx, dtype=torch.uint8, pin_memory=on_gpu)
    if n:
        with open(path, "rb", buffering=0) as f:
            mv = memoryview(stage.numpy())[:n]
            got = 0
            while got < n:
                r = f.readinto(mv[got:])
                if not r:
                    raise IOError("short read of %s" % path)
                got += r
    buf = stage.to(device, non_blocking=True)
    # gather to dst_rank only (the other ranks send and receive nothing): 1/world of the traffic of an all-gather
    outs = [torch.empty(mx, dtype=torch.uint8, device=device) for _ in range(world)] if rank == dst_rank else None
    dist.gather(buf, gather_list=outs, dst=dst_rank)
    if rank != dst_rank:
        return sum(sizes), None
    host = torch.empty(world * mx, dtype=torch.uint8, pin_memory=on_gpu)
    for r in range(world):
        host[r * mx: r * mx + sizes[r]].copy_(outs[r][: sizes[r]], non_blocking=True)
    if on_gpu:
        torch.cuda.current_stream().synchronize()
    a = host.numpy()
    return sum(sizes), [a[r * mx: r * mx + sizes[r]] for r in range(world)]


def max_over_ranks(values, device="cpu"):
    """element-wise MAX of a list of floats over all ranks (timing rule: max over ranks)"""
    t = torch.tensor(values, dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t]


def sum_over_ranks(values, device="cpu"):
    t = torch.tensor(values, dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(x) for x in t]
