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


def create_dataloader(dataset, dataset_opt, num_gpu=1, dist=False, sampler=None):
    if dataset_opt.get('phase', 'test') == 'train':
        raise NotImplementedError('training loaders are outside the B200 hot-path scope')
    return torch.utils.data.DataLoader(dataset, batch_size=1, shuffle=False,
                                       num_workers=int(dataset_opt.get('num_workers', 2) or 0), pin_memory=True)
