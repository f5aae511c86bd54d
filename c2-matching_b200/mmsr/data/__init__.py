"""Dataset / dataloader factories (reference: mmsr/data/__init__.py:25-93), test phase only."""
import importlib
import os

import torch.utils.data

_here = os.path.dirname(os.path.abspath(__file__))
_dataset_modules = [importlib.import_module(f'mmsr.data.{n[:-3]}')
                    for n in sorted(os.listdir(_here)) if n.endswith('_dataset.py')]


def create_dataset(dataset_opt):
    for m in _dataset_modules:
        cls = getattr(m, dataset_opt['type'], None)
        if cls is not None:
            return cls(dataset_opt)
    raise ValueError(f"Dataset {dataset_opt['type']} is not found.")


def collate_pairs(samples):
    """Stack the image tensors of equally shaped samples; keep paths / flags / sizes as per-sample lists."""
    out = {}
    for k in samples[0]:
        v = [s[k] for s in samples]
        out[k] = torch.stack(v) if torch.is_tensor(v[0]) else v
    return out


def create_dataloader(dataset, dataset_opt, num_gpu=1, dist=False, sampler=None):
    """Test-phase loader (reference: data/__init__.py:86-93, batch 1 / 1 worker / every rank sees everything).
    Here: the pair list is sharded `rank::world` at the INDEX level (a rank decodes only its pairs), pairs of equal
    shape are batched (`batch_size`, default 1 = the reference's behaviour), several workers decode ahead."""
    if dataset_opt.get('phase', 'test') == 'train':
        raise NotImplementedError('training loaders are outside the B200 hot-path scope')
    from .data_sampler import ShapeBucketBatchSampler, ShardedEvalSampler
    if sampler is None:
        sampler = ShardedEvalSampler(dataset) if dist else ShardedEvalSampler(dataset, 1, 0)
    batch = int(dataset_opt.get('batch_size') or 1)
    shape_fn = getattr(dataset, 'pair_shape', None) if batch > 1 else None
    batches = ShapeBucketBatchSampler(list(sampler), shape_fn or (lambda i: i), batch if shape_fn else 1)
    workers = int(dataset_opt.get('num_workers', 2) or 0)
    return torch.utils.data.DataLoader(dataset, batch_sampler=batches, num_workers=workers, pin_memory=True,
                                       collate_fn=collate_pairs, persistent_workers=False,
                                       prefetch_factor=(int(dataset_opt.get('prefetch_factor', 2)) if workers else None))
