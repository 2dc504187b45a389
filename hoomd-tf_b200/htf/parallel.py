"""Particle-row sharding over the GPUs of one node (one process per GPU, torch.distributed).

The reference has no multi-GPU code of its own (it inherits HOOMD's MPI domain decomposition,
/root/reference htf/test-py/test_mpi_tensorflow.py:59-80); its row batching
(htf/TensorflowCompute.cc:143-150,188-194) is what shards here: rank g builds and evaluates
rows [g*N/G, (g+1)*N/G).  Per step the path has ONE exchange: the two faces of the rank's slab go to
its two neighbours -- through the library's peer-memory windows (``htf_comm_*``: no NCCL, no host) or,
where peers cannot be mapped, NCCL send/recv; ``allgather_positions`` keeps the literal all-gather.  The RDF
histogram (int64, order independent, so still bit-exact) and scalar collective variables are all-reduced.
NCCL on GPUs, gloo in the CPU tests.
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


class SlabExchange:
    """Halo exchange for row shards that are slabs along one axis (particles spatially sorted along it).

    Instead of all-gathering every position (16 B x N_total per rank and step), each rank sends the particles
    within ``width = r_cut + skin`` of its two slab faces to the two neighbouring ranks (periodic) and bins
    ``[own rows | halo from below | halo from above]``.  The buffers have a fixed capacity; unused entries
    carry a sentinel that the region-of-interest test of the binning rejects, so no count ever has to reach the
    host.  ``libhtf_b200`` packs the halos (stable, index order); NCCL send/recv moves them.
    """

    def __init__(self, ctx, n_local, axis, lo_face, hi_face, width, capacity, group=None, transport="auto"):
        """``transport``: "p2p" = the library's peer-memory exchange (fused pack + send into the neighbours' windows
        over NVLink, flags instead of a collective; everything is one stream-ordered, graph-capturable call),
        "nccl" = pack kernels + NCCL send/recv, "auto" = p2p when every rank can map its peers, else nccl."""
        self.ctx, self.n_local, self.axis, self.cap = ctx, int(n_local), int(axis), int(capacity)
        self.lo_thr, self.hi_thr = float(lo_face) + float(width), float(hi_face) - float(width)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        dev = ctx.device
        extra = 2 * self.cap if self.world > 1 else 0
        self.local = torch.empty((self.n_local + extra, 4), dtype=torch.float32, device=dev)
        self.transport = "nccl" if self.world > 1 else "none"
        self.transport_note = ""
        if self.world > 1 and transport in ("auto", "p2p"):
            self._connect_p2p(require=(transport == "p2p"))
        if self.transport != "p2p":
            self.send_lo = torch.empty((self.cap, 4), dtype=torch.float32, device=dev)
            self.send_hi = torch.empty((self.cap, 4), dtype=torch.float32, device=dev)

    def _connect_p2p(self, require):
        """Create this rank's window, swap the IPC handles over the process group, map the peers.  Every rank takes
        part in both collectives whatever happened locally, and the ranks agree on the outcome."""
        from ._lib import HtfError
        handle, err = None, ""
        try:
            handle = self.ctx.comm_create(self.rank, self.world, self.cap)
        except HtfError as ex:
            err = str(ex)
        handles = [None] * self.world
        dist.all_gather_object(handles, handle, group=self.group)
        ok = all(h is not None for h in handles)
        if ok:
            try:
                self.ctx.comm_connect(handles)
            except HtfError as ex:
                ok, err = False, str(ex)
        flags = [None] * self.world
        dist.all_gather_object(flags, bool(ok), group=self.group)
        if all(flags):
            self.transport = "p2p"
            return
        try:
            self.ctx.comm_destroy()
        except HtfError:
            pass
        self.transport_note = err or "a peer could not map the windows"
        if require:
            raise RuntimeError("peer-memory halo exchange unavailable: " + self.transport_note)

    @property
    def own(self):
        """View of this rank's rows inside the local array (write the shard's positions here)."""
        return self.local[:self.n_local]

    def pack(self):
        """Device-side part of the exchange (graph-capturable).  p2p: the WHOLE exchange (pack straight into the
        neighbours' windows, signal, wait, gather); nccl: both faces into the two send buffers."""
        if self.world <= 1:
            return
        if self.transport == "p2p":
            self.ctx.comm_exchange_halo(self.local, self.n_local, self.axis, self.lo_thr, self.hi_thr)
        else:
            self.ctx.pack_halo_pair(self.own, self.axis, self.lo_thr, self.hi_thr, self.send_lo, self.send_hi)

    def exchange(self):
        """Pack both faces, swap with the neighbours, return the local array ``[own | halo | halo]``."""
        self.pack()
        return self.swap()

    def swap(self):
        """NCCL half of the exchange: send the packed faces to the two neighbours, receive theirs (nothing left to
        do for the peer-memory transport)."""
        if self.world == 1 or self.transport == "p2p":
            return self.local
        prev, nxt = (self.rank - 1) % self.world, (self.rank + 1) % self.world
        a = self.local[self.n_local:self.n_local + self.cap]
        b = self.local[self.n_local + self.cap:]
        ops = [dist.P2POp(dist.isend, self.send_lo, prev, self.group), dist.P2POp(dist.isend, self.send_hi, nxt, self.group),
               dist.P2POp(dist.irecv, a, nxt, self.group), dist.P2POp(dist.irecv, b, prev, self.group)]
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        return self.local


def slab_plan(pos_rows, axis, r_cut, skin=0.4, slack=1.5):
    """(lo_face, hi_face, width, capacity) of a slab shard from its current positions (host side, at set-up or
    whenever the region of interest is refreshed)."""
    import numpy as np
    p = pos_rows.detach().cpu().numpy() if torch.is_tensor(pos_rows) else np.asarray(pos_rows)
    x = p[:, axis]
    lo_face, hi_face = float(x.min()), float(x.max())
    width = float(r_cut) + float(skin)
    n_lo = int((x < lo_face + width).sum())
    n_hi = int((x > hi_face - width).sum())
    cap = int(max(n_lo, n_hi) * slack) + 256
    return lo_face, hi_face, width, (cap + 255) // 256 * 256
