"""Trajectory-side helpers of hoomd-tf on the cell-list kernel
(/root/reference htf/utils.py: compute_nlist :75-161, compute_pairwise :164-201,
iter_from_trajectory :627-749).  The O(N^2) distance matrix + top_k of the reference is
replaced by libhtf_b200's neighbor build; only the per-row ordering / top-k bookkeeping runs
in torch."""
import numpy as np
import torch

from .context import HtfContext


_NLIST_CTX = {}


def _nlist_context(dev, n, K, r_cut):
    """One persistent context per device for the trajectory helpers (the reference rebuilds nothing per frame
    either: its compute_nlist is a traced TF function, htf/utils.py:75-161).  Cutoff, K and box are re-set per call."""
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    ctx = _NLIST_CTX.get(idx)
    if ctx is None:
        ctx = HtfContext(max(n, 1), K, float(r_cut), device=torch.device("cuda", idx))
        _NLIST_CTX[idx] = ctx
    else:
        ctx.set_cutoff(float(r_cut), K)
    return ctx


def _wrap_into_box(xyz, lo, L):
    return xyz - torch.floor((xyz - lo) / L) * L


def compute_nlist(positions, r_cut, NN, box_size, sorted=False, return_types=False, exclusion_matrix=None):
    """Neighbor lists of arbitrary positions: [N, NN, 4] = (dx, dy, dz, index-or-type), zero padded.

    Semantics of htf/utils.py:75-161: pairs with 5e-4 <= r <= r_cut; ``sorted=True`` keeps the NN
    nearest in ascending distance, ``sorted=False`` keeps the NN farthest (the reference's ``top_k``
    of the masked distances) in no particular order; the last column is the neighbor index
    (``return_types=False``) or its type; ``exclusion_matrix[i,j] | exclusion_matrix[j,i]`` removes a pair.
    """
    if not torch.is_tensor(positions):
        positions = torch.as_tensor(np.asarray(positions), dtype=torch.float32)
    if not positions.is_cuda:
        if not torch.cuda.is_available():
            raise RuntimeError("compute_nlist needs a CUDA device: the hot path has no CPU fallback")
        positions = positions.cuda()
    if return_types and positions.shape[1] == 3:
        raise ValueError("Cannot return type if positions does not have type. Make sure positions is N x 4")
    dev = positions.device
    n = positions.shape[0]
    L = torch.as_tensor(np.asarray(box_size, dtype=np.float32), device=dev).reshape(3)
    lo = torch.zeros(3, device=dev)
    pos4 = torch.zeros((n, 4), dtype=torch.float32, device=dev)
    pos4[:, :3] = _wrap_into_box(positions[:, :3].to(torch.float32), lo, L)
    pos4[:, :3] = torch.where(pos4[:, :3] >= L, pos4[:, :3] - L, pos4[:, :3])
    if positions.shape[1] > 3:
        pos4[:, 3] = positions[:, 3]
    hi = L.cpu().numpy()
    K = int(NN)
    while True:
        ctx = _nlist_context(dev, n, K, r_cut)
        ctx.set_box([0.0, 0.0, 0.0], hi)
        nl, idx, cnt = ctx.build_nlist(pos4, want_idx=True, want_count=True)
        cmax = int(cnt.max().item()) if n > 0 else 0
        ctx.overflow()                # reset the sticky "a row is full" flag of the shared context
        if cmax <= K:
            break
        K = cmax                      # more candidates than NN: rebuild wide, then pick NN of them below
    d = nl[:, :, :3]
    r = torch.linalg.norm(d, dim=2)
    valid = (idx >= 0) & (r >= 5e-4)
    if exclusion_matrix is not None:
        em = torch.as_tensor(np.asarray(exclusion_matrix), device=dev).bool()
        em = em | em.t()
        rows = torch.arange(n, device=dev)[:, None].expand_as(idx)
        valid &= ~em[rows, idx.clamp(min=0).long()]
    if sorted:
        key = torch.where(valid, r, torch.full_like(r, float("inf")))
        order = torch.argsort(key, dim=1, stable=True)[:, :NN]
    else:
        key = torch.where(valid, r, torch.full_like(r, -1.0))
        order = torch.topk(key, k=min(NN, key.shape[1]), dim=1, sorted=False).indices
    take = lambda t: torch.gather(t, 1, order)
    v = take(valid)
    dsel = torch.gather(d, 1, order[:, :, None].expand(-1, -1, 3))
    last = take(nl[:, :, 3]) if return_types else take(idx).to(torch.float32)
    out = torch.cat([dsel, last[:, :, None]], dim=-1) * v[:, :, None].to(torch.float32)
    if out.shape[1] < NN:
        out = torch.cat([out, torch.zeros((n, NN - out.shape[1], 4), device=dev)], dim=1)
    return out


def compute_pairwise(model, r, type_i=0, type_j=0):
    """Model output for a 2-particle system at separations ``r`` (htf/utils.py:164-201)."""
    NN = model.nneighbor_cutoff
    dev = torch.device("cuda")
    output = None
    positions = torch.zeros((2, 4), device=dev)
    positions[0, -1] = type_i
    positions[1, -1] = type_j
    box = torch.tensor([[0.0, 0, 0], [1e10, 1e10, 1e10], [0, 0, 0]], device=dev)
    for ri in np.asarray(r, dtype=np.float64):
        nlist = torch.zeros((2, NN, 4), device=dev)
        nlist[0, :, -1] = type_j
        nlist[1, :, -1] = type_i
        nlist[0, 0, 1] = float(ri)
        nlist[1, 0, 1] = -float(ri)
        result = model([nlist, positions, box], False)
        vals = [o.detach().cpu().numpy()[np.newaxis, ...] for o in result]
        output = vals if output is None else [np.append(o, v, axis=0) for o, v in zip(output, vals)]
    return output


def iter_from_trajectory(nneighbor_cutoff, universe, selection="all", r_cut=10.0, period=1, start=0.0, end=None,
                         static_nlist=False):
    """Yield ``([nlist, positions, box], timestep)`` for the frames of an MDAnalysis-style universe
    (htf/utils.py:627-749).  Any object with ``.select_atoms(sel)`` -> group(``.positions``, ``.atoms.types``),
    ``.dimensions`` and an iterable ``.trajectory`` of frames with ``.frame`` works; MDAnalysis is optional.

    The reference computes the neighbor list once, before the frame loop (:717-721); pass
    ``static_nlist=True`` to reproduce that, the default recomputes it for every frame as documented.
    """
    box = np.asarray(universe.dimensions, dtype=np.float64)
    gamma, beta, alpha = np.deg2rad(box[5]), np.deg2rad(box[4]), np.deg2rad(box[3])
    xy = 1.0 / np.tan(gamma)
    xz = np.cos(beta)
    yz = np.cos(alpha) - xy * xz
    dev = torch.device("cuda")
    hoomd_box = torch.tensor(np.array([[0, 0, 0], [box[0], box[1], box[2]], [xy, xz, yz]]), dtype=torch.float32,
                             device=dev)
    atom_group = universe.select_atoms(selection)
    try:
        types = list(np.unique(atom_group.atoms.types))
        type_array = np.array([types.index(i) for i in atom_group.atoms.types]).reshape(-1, 1)
    except Exception:
        type_array = np.zeros(len(atom_group)).reshape(-1, 1)

    def nl_of(group):
        return compute_nlist(torch.as_tensor(np.asarray(group.positions), dtype=torch.float32, device=dev),
                             r_cut=r_cut, NN=nneighbor_cutoff, box_size=box[:3])

    nlist = nl_of(atom_group) if static_nlist else None
    if end is None:
        end = getattr(universe.trajectory, "totaltime", float("inf"))
    for i, ts in enumerate(universe.trajectory):
        if ts.frame >= start and ts.frame <= end and i % period == 0:
            pos = np.concatenate((np.asarray(atom_group.positions), type_array), axis=1)
            yield [nlist if static_nlist else nl_of(atom_group),
                   torch.as_tensor(pos, dtype=torch.float32, device=dev), hoomd_box], ts
