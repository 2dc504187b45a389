"""Particle-row sharding over the GPUs of one node (one process per GPU, torch.distributed).

The reference has no multi-GPU code of its own (it inherits HOOMD's MPI domain decomposition,
/root/reference htf/test-py/test_mpi_tensorflow.py:59-80); its row batching
(htf/TensorflowCompute.cc:143-150,188-194) is what shards here: rank g builds and evaluates
rows [g*N/G, (g+1)*N/G).  Per step the path has ONE exchange: the all-gather of the position
shards; the RDF histogram (int64, order independent, so still bit-exact) and scalar collective
variables are all-reduced.  NCCL on GPUs, gloo in the CPU tests.
"""
import torch
import torch.distributed as dist


def row_shard(n, world, rank):
    """Contiguous row range of ``rank``: equal slabs, the last rank takes the remainder."""
    per = n // world
    lo = rank * per
    hi = n if rank == world - 1 else lo + per
    return lo, hi


def allgather_positions(shard, out=None, group=None):
    """All ranks' position shards [n_r,4] -> the full [N,4] array in rank order.

    Equal shards use one all_gather_into_tensor (NCCL ring/NVLS over NVSwitch); a ragged last
    shard falls back to all_gather on a list."""
    world = dist.get_world_size(group)
    sizes = [None] * world
    dist.all_gather_object(sizes, int(shard.shape[0]), group=group) if out is None else None
    if out is not None and out.shape[0] == shard.shape[0] * world:
        dist.all_gather_into_tensor(out, shard.contiguous(), group=group)
        return out
    if out is None:
        total = sum(sizes)
        out = torch.empty((total, shard.shape[1]), dtype=shard.dtype, device=shard.device)
    else:
        dist.all_gather_object(sizes, int(shard.shape[0]), group=group)
    if len(set(sizes)) == 1:
        dist.all_gather_into_tensor(out, shard.contiguous(), group=group)
        return out
    parts = [torch.empty((s, shard.shape[1]), dtype=shard.dtype, device=shard.device) for s in sizes]
    dist.all_gather(parts, shard.contiguous(), group=group)
    torch.cat(parts, dim=0, out=out)
    return out


def allreduce_bins(bins, group=None):
    """Sum the per-rank RDF histograms (int64: exact and order independent)."""
    dist.all_reduce(bins, op=dist.ReduceOp.SUM, group=group)
    return bins


def allreduce_scalar(x, group=None):
    """Sum a per-rank partial of a collective variable (fp64 to keep the sum rank-count independent)."""
    t = x.detach().to(torch.float64).reshape(1).clone()
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t[0]


def roi_for_rows(pos_rows, box_lo, box_hi, r_cut, skin=0.4):
    """Region of interest for ``HtfContext.set_roi``: per axis, the bounding interval of this rank's rows
    widened by ``r_cut + skin``.  Axes whose widened interval covers the whole box get half-width -1 (no
    restriction).  ``skin`` plays the role of HOOMD's ``r_buff``: the region stays valid until a local
    particle has moved more than ``skin`` out of the interval it was computed for."""
    import numpy as np
    p = pos_rows.detach().cpu().numpy() if torch.is_tensor(pos_rows) else np.asarray(pos_rows)
    lo = np.asarray(box_lo, dtype=np.float64)
    L = np.asarray(box_hi, dtype=np.float64) - lo
    a = p[:, :3].min(axis=0).astype(np.float64)
    b = p[:, :3].max(axis=0).astype(np.float64)
    center = 0.5 * (a + b)
    half = 0.5 * (b - a) + float(r_cut) + float(skin)
    half = np.where(2.0 * half >= L, -1.0, half)
    return center.astype(np.float32), half.astype(np.float32)
