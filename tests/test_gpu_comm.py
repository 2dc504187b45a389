"""Peer-memory exchange behind the C ABI (htf_comm_*), two ranks.

Runs on ONE GPU: two processes share the device, map each other's windows with CUDA IPC exactly as two GPUs of a
node would, and go through the same kernels (fused pack + send, flags, gather; mailbox all-reduce).  The handles
travel over a gloo process group.  What a sharded run must reproduce (the intent of the reference's MPI test,
/root/reference htf/test-py/test_mpi_tensorflow.py:59-80): the neighbor tensor of every rank's rows equals the
single-domain one bit for bit, and reductions give every rank the same exact sums.
"""
import os
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _sorted_rows(a):
    u = np.ascontiguousarray(a).view(np.uint32).astype(np.uint64)
    k1 = (u[..., 0] << np.uint64(32)) | u[..., 1]
    k2 = (u[..., 2] << np.uint64(32)) | u[..., 3]
    order = np.lexsort((k2, k1), axis=1)
    return np.take_along_axis(a, order[:, :, None], axis=1)


def _worker(rank, world, port, out_dir):
    for p in (ROOT, os.path.join(ROOT, "hoomd-tf_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    import htf
    import oracle
    from htf import synthetic
    from htf.parallel import SlabExchange, roi_for_rows, slab_plan
    dist.init_process_group("gloo", rank=rank, world_size=world, init_method="tcp://127.0.0.1:%d" % port)
    torch.cuda.set_device(0)
    result = {"ok": False}
    try:
        pos, lo, hi = synthetic.lattice_fluid((12, 12, 24), 0.7, seed=31)
        n, K, r_cut = pos.shape[0], 64, 2.5
        per = n // world
        a, b = rank * per, (rank + 1) * per
        ctx = htf.HtfContext(n, K, r_cut)
        ctx.set_box(lo, hi)
        ctx.set_roi(*roi_for_rows(pos[a:b], lo, hi, r_cut))
        lo_face, hi_face, width, cap = slab_plan(pos[a:b], 2, r_cut)
        caps = [None] * world
        dist.all_gather_object(caps, cap)
        xch = SlabExchange(ctx, per, 2, lo_face, hi_face, width, max(caps), transport="p2p")
        assert xch.transport == "p2p"
        rng = np.random.default_rng(7)                      # same stream on both ranks: same global moves
        cur = pos.copy()
        for it in range(3):
            if it:
                cur[:, :3] += rng.normal(0.0, 0.02, (n, 3)).astype(np.float32)
                L = (hi - lo)
                cur[:, :3] = (cur[:, :3] - lo) % L + lo
                cur[:, :3] = np.where(cur[:, :3] >= hi, cur[:, :3] - L, cur[:, :3]).astype(np.float32)
            xch.own.copy_(torch.from_numpy(cur[a:b]).cuda())
            local = xch.exchange()
            torch.cuda.synchronize()
            assert ctx.comm_status() == 0 and ctx.overflow() == 0
            # expected faces, from the global array: stable order, sentinel padded
            nxt, prv = (rank + 1) % world, (rank - 1) % world
            pn, pp = cur[nxt * per:(nxt + 1) * per], cur[prv * per:(prv + 1) * per]
            planes = [None] * world
            dist.all_gather_object(planes, (xch.lo_thr, xch.hi_thr))
            from_next = pn[pn[:, 2] < np.float32(planes[nxt][0])]
            from_prev = pp[pp[:, 2] > np.float32(planes[prv][1])]
            got = local.cpu().numpy()
            c = xch.cap
            assert np.array_equal(got[per:per + len(from_next)], from_next)
            assert np.all(got[per + len(from_next):per + c, 0] == np.float32(1e30))
            assert np.array_equal(got[per + c:per + c + len(from_prev)], from_prev)
            assert np.all(got[per + c + len(from_prev):, 0] == np.float32(1e30))
            # neighbor tensor of the own rows from the exchanged array == single-domain oracle, bit for bit
            nl = ctx.build_nlist(local, 0, per)
            nl_o, _, _ = oracle.nlist(cur, lo, hi, r_cut, K, a, b, cells=True, want_idx=False)
            assert np.array_equal(_sorted_rows(nl.cpu().numpy()).view(np.uint32), _sorted_rows(nl_o).view(np.uint32))
            # all-reduces: exact int64 sums, rank-ordered fp64 sums (same bits on every rank)
            bins = ctx.rdf_hist(nl, (0.0, r_cut), 100)
            want_bins = oracle.rdf_hist(oracle.nlist(cur, lo, hi, r_cut, K, cells=True, want_idx=False)[0], (0.0, r_cut), 100)
            ctx.comm_allreduce(bins)
            v = torch.tensor([0.1 * (rank + 1), 1e-17 * (rank + 1), float(it)], dtype=torch.float64, device="cuda")
            ctx.comm_allreduce(v)
            gvec = torch.arange(10497, dtype=torch.float32, device="cuda") * (rank + 1) * 1e-3   # weight-gradient sized
            ctx.comm_allreduce(gvec)
            torch.cuda.synchronize()
            want_g = np.zeros(10497, dtype=np.float32)
            for r in range(world):
                want_g = want_g + (np.arange(10497, dtype=np.float32) * np.float32(r + 1) * np.float32(1e-3))
            assert np.array_equal(gvec.cpu().numpy(), want_g)
            assert np.array_equal(bins.cpu().numpy(), want_bins)
            want_v = np.zeros(3)
            for r in range(world):
                want_v = want_v + np.array([0.1 * (r + 1), 1e-17 * (r + 1), float(it)])
            assert np.array_equal(v.cpu().numpy(), want_v)
        # the same exchange captured into a CUDA graph advances its own epochs
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            xch.exchange()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            xch.pack()
        for _ in range(4):
            g.replay()
        torch.cuda.synchronize()
        assert ctx.comm_status() == 0
        assert np.array_equal(xch.local.cpu().numpy(), got)
        ctx.comm_destroy()
        result["ok"] = True
    except BaseException as ex:                              # noqa: BLE001 -- reported to the parent
        import traceback
        result["error"] = "".join(traceback.format_exception(type(ex), ex, ex.__traceback__))[-3000:]
    np.save(os.path.join(out_dir, "rank%d.npy" % rank), np.array([repr(result)]))
    try:
        dist.destroy_process_group()
    except Exception:
        pass


def test_peer_memory_halo_exchange_and_allreduce_two_ranks(tmp_path):
    import torch.multiprocessing as mp
    world, port = 2, 29000 + os.getpid() % 2000
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_worker, args=(r, world, port, str(tmp_path))) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(240)
    hung = [p for p in procs if p.is_alive()]
    for p in hung:
        p.kill()
    assert not hung, "a rank did not finish (exchange dead-locked?)"
    for r in range(world):
        res = eval(str(np.load(os.path.join(str(tmp_path), "rank%d.npy" % r))[0]))     # noqa: S307 -- our own repr
        assert res.get("ok"), "rank %d: %s" % (r, res.get("error"))
