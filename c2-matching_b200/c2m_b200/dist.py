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


# ------------------------------------------------------------------------------------------------
# Optional latency mode: ONE image's correlation split over the ranks of a group.
def ref_row_slab(rh, rank, world):
    """Rows [r0, r1) of the Ref PATCH grid (rh rows) owned by `rank`: contiguous, balanced."""
    base, rem = divmod(rh, world)
    r0 = rank * base + min(rank, rem)
    return r0, r0 + base + (1 if rank < rem else 0)


def combine_argmax(val, idx, group=None):
    """All ranks hold (val, idx) of the same queries over DIFFERENT Ref slabs (idx already global).
    Returns the global maximum with the reference's tie rule (lowest index wins): an all-gather of
    12 bytes per query — latency-bound, so no fused kernel — followed by a lexicographic max."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return val, idx
    world = dist.get_world_size(group)
    vals = [torch.empty_like(val) for _ in range(world)]
    idxs = [torch.empty_like(idx) for _ in range(world)]
    dist.all_gather(vals, val.contiguous(), group=group)
    dist.all_gather(idxs, idx.contiguous(), group=group)
    v, i = torch.stack(vals), torch.stack(idxs)
    vmax = v.max(dim=0).values
    cand = torch.where(v == vmax, i, torch.full_like(i, torch.iinfo(i.dtype).max))
    return vmax, cand.min(dim=0).values


def corr_argmax_ref_sharded(feat_in, feat_ref, patch_size=3, is_norm=True, norm_input=False, l2norm=False, group=None):
    """`corr_argmax` (stride 1) with the Ref patch grid's rows split across the ranks of `group`: every
    rank searches its slab (exact rescoring included) and the partial maxima are combined.  The result
    is identical on all ranks and bit-identical to the single-GPU search."""
    from . import ops
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    hr, wr = feat_ref.shape[2:]
    rh, rw = hr - patch_size + 1, wr - patch_size + 1
    r0, r1 = ref_row_slab(rh, rank, world)
    if l2norm:          # normalise per pixel BEFORE slicing so every rank sees the same values
        feat_ref = torch.nn.functional.normalize(feat_ref, dim=1)
        feat_in = torch.nn.functional.normalize(feat_in, dim=1)
    if r1 > r0:
        slab = feat_ref[:, :, r0:r1 + patch_size - 1].contiguous()
        idx, val = ops.corr_argmax(feat_in, slab, patch_size, 1, 1, is_norm, norm_input)
        idx = idx + r0 * rw
    else:               # more ranks than Ref rows: contribute nothing
        gh, gw = feat_in.shape[2] - patch_size + 1, feat_in.shape[3] - patch_size + 1
        idx = torch.zeros(feat_in.shape[0], gh, gw, dtype=torch.int64, device=feat_in.device)
        val = torch.full((feat_in.shape[0], gh, gw), float('-inf'), device=feat_in.device)
    val, idx = combine_argmax(val, idx, group)
    return idx, val
