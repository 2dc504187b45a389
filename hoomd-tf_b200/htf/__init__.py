"""htf -- B200-native implementation of hoomd-tf's nlist -> forces+virial path.

Mirrors the public surface of ``hoomd.htf`` (/root/reference htf/__init__.py:8-13) for
that path.  The compute path is libhtf_b200.so (hand-written sm_100a CUDA behind a C
ABI); torch provides device memory, streams and torch.distributed only.
"""
from . import _lib
from . import synthetic
from .context import HtfContext
from . import ops
from .simmodel import (SimModel, compute_nlist_forces, compute_positions_forces, nlist_rinv, safe_norm, box_size,
                       wrap_vector, compute_rdf, masked_nlist, rdf_from_hist, Mean, MeanTensor)
from .layers import RBFExpansion, WCARepulsion, EDSLayer
from .tensorflowcompute import tfcompute
from .utils import compute_nlist, compute_pairwise, iter_from_trajectory
from . import models
from . import sim
from . import parallel
from .ops import lj_forces, rdf_hist

__version__ = "0.1.0"
