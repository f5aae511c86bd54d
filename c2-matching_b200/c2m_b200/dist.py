"""Multi-GPU plumbing: the path shards by independent (LR, Ref) pairs — one process per GPU,
weights replicated, pair list split `rank::world`, no collective on the data path; the only
exchange is the final gather of per-image metric rows."""
import torch
import torch.distributed as dist


def shard_indices(n, rank, world):
    """Indices of the pairs rank `rank` of `world` processes owns."""
    return list(range(rank, n, world))


def gather_rows(rows):
    """rows: [k, m] float64 tensor of this rank (k may differ per rank) -> all ranks' rows
    concatenated in rank order.  Works with NCCL (CUDA tensors) and gloo (CPU tensors)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return rows
    world = dist.get_world_size()
    n = torch.tensor([rows.shape[0]], device=rows.device, dtype=torch.int64)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n)
    cap = max(int(c.item()) for c in counts)
    padded = torch.zeros(cap, rows.shape[1], dtype=rows.dtype, device=rows.device)
    padded[:rows.shape[0]] = rows
    parts = [torch.zeros_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded)
    return torch.cat([p[:int(c.item())] for p, c in zip(parts, counts)])


def max_over_ranks(value, device):
    """Scalar max-reduce (bench timing: the job is as slow as its slowest rank)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], device=device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
