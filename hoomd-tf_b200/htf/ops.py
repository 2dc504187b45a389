"""Functional front end of the fused sm_100a kernels for tensors that are already on the device.

These are what the built-in models call instead of a TensorFlow graph
(/root/reference htf/simmodel.py:526-555, :618-669): ``lj_forces`` is the closed form of
LJModel -> nlist_rinv -> compute_nlist_forces -> _compute_virial, ``rdf_hist`` is
compute_rdf's histogram.  One ``HtfContext`` per device is created lazily and shared.
"""
import torch

from .context import HtfContext

_CTX = {}


def default_context(device=None):
    """The per-device context used by the functional ops (cutoff-independent entry points only)."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    if idx not in _CTX:
        _CTX[idx] = HtfContext(1, 1, 1.0, device=torch.device("cuda", idx))
    return _CTX[idx]


def _as_nlist(nlist):
    if not (torch.is_tensor(nlist) and nlist.is_cuda):
        raise ValueError("nlist must be a CUDA tensor: the hot path has no CPU fallback")
    if nlist.dim() != 3 or nlist.shape[2] != 4:
        raise ValueError("nlist must be [N, NN, 4]")
    nl = nlist.detach()
    if nl.dtype != torch.float32 or not nl.is_contiguous():
        nl = nl.to(torch.float32).contiguous()
    return nl


def lj_forces(nlist, virial=False, counts=None):
    """LJ (epsilon = sigma = 1) forces+energy [N,4] (and the virial [N,3,3]) of the built-in LJ model.

    Same numbers as ``compute_nlist_forces(nlist, sum_j 2 (rinv^12 - rinv^6), virial)`` with
    ``rinv = nlist_rinv(nlist)`` (htf/test-py/build_examples.py:67-77, :104-115), computed by one
    streaming pass of libhtf_b200 over the neighbor tensor.
    """
    nl = _as_nlist(nlist)
    ctx = default_context(nl.device)
    if not virial:
        return ctx.lj_forces(nl, counts=counts)
    fe, v9 = ctx.lj_forces(nl, virial=True, virial_components=9, counts=counts)
    return fe, v9.view(-1, 3, 3)


def rdf_hist(nlist, r_range, nbins=100, type_tensor=None, type_i=None, type_j=None, bins=None):
    """int64[nbins+2] histogram of |d| over the neighbor tensor (htf/simmodel.py:657-662)."""
    nl = _as_nlist(nlist)
    ctx = default_context(nl.device)
    tt = None
    if type_tensor is not None and type_i is not None:
        tt = type_tensor.detach()
        if tt.dim() == 2 and tt.shape[1] == 1:
            tt = tt[:, 0]
        if tt.dtype != torch.float32:
            tt = tt.to(torch.float32)
    return ctx.rdf_hist(nl, r_range, nbins=nbins, type_tensor=tt, type_i=type_i if tt is not None else None,
                        type_j=type_j if type_tensor is not None else None, bins=bins)


def lj_cv_forces(nlist, r0, virial=False, rdf_range=None, nbins=100, bins=None, cv_sum=None, counts=None):
    """One pass: LJ forces (+virial [N,6]) + the smooth coordination CV of BASELINE config 5 (+ RDF histogram).

    Returns ``(forces[N,4], virial6 or None, cv_row[N,4], cv_sum float64[1], bins or None)`` where
    ``cv_row = (sum_j ds/dd_ij (x,y,z), cn_i)`` and ``cv_sum = sum_i cn_i`` (see include/htf_b200.h).
    ``counts`` int32[N] (the builder's neighbors per row) lets the pass skip the zero padding.
    """
    nl = _as_nlist(nlist)
    ctx = default_context(nl.device)
    rows = nl.shape[0]
    cv_row = torch.empty((rows, 4), dtype=torch.float32, device=nl.device)
    if cv_sum is None:
        cv_sum = torch.zeros(1, dtype=torch.float64, device=nl.device)
    vir = torch.empty((rows, 6), dtype=torch.float32, device=nl.device) if virial else None
    if rdf_range is not None and bins is None:
        bins = torch.zeros(nbins + 2, dtype=torch.int64, device=nl.device)
    fe = ctx.lj_cv_forces(nl, r0, cv_row, cv_sum, virial_out=vir, bins=bins,
                          r_range=rdf_range if rdf_range is not None else (0.0, 1.0), nbins=nbins, counts=counts)
    return fe, vir, cv_row, cv_sum, bins


def mlp_forces(nlist, raw_params, rbf_high, packed=None):
    """Pairwise-MLP forces+energy [N,4] from the raw fp32 parameter blob (packed to bf16 operands on the fly)."""
    nl = _as_nlist(nlist)
    ctx = default_context(nl.device)
    packed = ctx.mlp_pack(raw_params.detach().to(torch.float32).contiguous(), packed)
    return ctx.mlp_forces(nl, packed, rbf_high)
