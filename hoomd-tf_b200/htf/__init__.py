"""htf -- B200-native implementation of hoomd-tf's nlist -> forces+virial path.

Mirrors the public surface of ``hoomd.htf`` (/root/reference htf/__init__.py:8-13) for
that path.  The compute path is libhtf_b200.so (hand-written sm_100a CUDA behind a C
ABI); torch provides device memory, streams and torch.distributed only.
"""
from . import _lib
from .context import HtfContext
from . import synthetic

__version__ = "0.1.0"
