"""Arch registry: every `*_arch.py` next to this file is imported and searched by class name
(reference behaviour: mmsr/models/archs/__init__.py:8-18, there via mmcv.scandir)."""
import importlib
import os

_here = os.path.dirname(os.path.abspath(__file__))
_arch_modules = [
    importlib.import_module(f'mmsr.models.archs.{name[:-3]}')
    for name in sorted(os.listdir(_here)) if name.endswith('_arch.py')
]
