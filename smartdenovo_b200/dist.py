"""Multi-GPU plumbing for the overlap path: query reads shard across ranks exactly like the reference's
`-P n -p i` jobs (wtzmo.c:1291,1314, usage :1431-1433) -- rank r of N is job r -- with the k-mer index replicated,
so there is NO collective during compute.  The only exchange is the final gather of the variable-length record
text to rank 0 (one all-gather of sizes + one all-gather of padded byte tensors), equivalent to `cat part*.ovl`.
Works with any torch.distributed backend (NCCL on GPUs, gloo in the CPU tests)."""
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
