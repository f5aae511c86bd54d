"""Samplers of the evaluation path (reference: mmsr/data/data_sampler.py:8-69 has only the training-side
`DistIterSampler`; its validation loop runs every image on every rank, sr_model.py:160-162).

`ShardedEvalSampler` partitions the pair list `rank::world` with NO padding or repetition, so a rank decodes
only the pairs it will run; `ShapeBucketBatchSampler` groups a rank's pairs into batches of equal tensor shape
(CUFED5 pairs differ in size), so the batched kernels see B > 1 without resizing anything."""
from collections import OrderedDict

import torch.distributed as dist
from torch.utils.data.sampler import Sampler


class ShardedEvalSampler(Sampler):

    def __init__(self, dataset, num_replicas=None, rank=None):
        if num_replicas is None:
            num_replicas = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        if rank is None:
            rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
        if not 0 <= rank < num_replicas:
            raise ValueError(f'rank {rank} outside world of {num_replicas}')
        self.indices = list(range(rank, len(dataset), num_replicas))

    def __iter__(self):
        return iter(self.indices)

    def __len__(self):
        return len(self.indices)


class ShapeBucketBatchSampler(Sampler):
    """Batches of <= batch_size indices that share `shape_fn(index)`, in first-seen order; a bucket's ragged tail is
    emitted as a smaller batch (nothing is dropped or duplicated)."""

    def __init__(self, indices, shape_fn, batch_size):
        if batch_size < 1:
            raise ValueError('batch_size must be >= 1')
        buckets = OrderedDict()
        for i in indices:
            buckets.setdefault(shape_fn(i), []).append(i)
        self.batches = [b[k:k + batch_size] for b in buckets.values() for k in range(0, len(b), batch_size)]

    def __iter__(self):
        return iter(self.batches)

    def __len__(self):
        return len(self.batches)
