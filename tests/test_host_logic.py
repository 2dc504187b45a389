"""CPU tests of the host side: C-ABI surface, SimModel plumbing, EDS layer vs the oracle, and the
world_size-2 row sharding over gloo (the N>1 path without GPUs)."""
import ctypes
import os
import re
import sys

import numpy as np
import pytest
import torch

import htf
from htf import synthetic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_abi_library_exports_every_declared_symbol():
    """the library loads and exports every function include/htf_b200.h declares (no compute calls here)."""
    hdr = open(os.path.join(ROOT, "include", "htf_b200.h")).read()
    declared = set(re.findall(r"\b(htf_[a-z0-9_]+)\s*\(", hdr))
    assert {"htf_create", "htf_build_nlist", "htf_lj_forces", "htf_rdf_hist", "htf_lj_step"} <= declared
    lib = htf._lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(htf._lib.SYMBOLS), "ctypes table out of sync with the header"
    assert lib.htf_abi_version() == htf._lib.ABI_VERSION == int(re.search(r"HTF_ABI_VERSION (\d+)", hdr).group(1))
    # error convention: negative code + message, never a throw; no GPU here -> create must fail cleanly
    if not torch.cuda.is_available():
        h = ctypes.c_void_p()
        rc = lib.htf_create(ctypes.byref(h), 0, 16, 8, 1.0, 0)
        assert rc < 0 and not h.value and len(lib.htf_last_error(None)) > 0
    h = ctypes.c_void_p()
    assert lib.htf_create(ctypes.byref(h), 0, 16, 0, 1.0, 0) == htf._lib.EINVAL       # K < 1
    assert b"nneighbor_cutoff" in lib.htf_last_error(None)


def test_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        htf.HtfContext(16, 8, 1.0)
    with pytest.raises(ValueError):
        htf.lj_forces(torch.zeros((4, 8, 4)))
    with pytest.raises(ValueError):
        htf.compute_rdf(torch.zeros((4, 8, 4)), [0, 1])


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "hoomd-tf_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "htf_oracle" not in src, f


def test_simmodel_argument_introspection():
    """htf/simmodel.py:51-68: 1-3 positional args, optional trailing `training`."""
    class A(htf.SimModel):
        def compute(self, nlist):
            return nlist.sum(dim=(1, 2))

    class B(htf.SimModel):
        def compute(self, nlist, positions, training):
            return positions[:, :3] * (2.0 if training else 1.0)

    class C(htf.SimModel):
        def setup(self, scale):
            self.scale = scale

        def compute(self, nlist, positions, box):
            return htf.box_size(box) * self.scale

    a, b, c = A(4), B(4, output_forces=False), C(4, scale=3.0)
    assert (a._arg_count, a._pass_training) == (1, False)
    assert (b._arg_count, b._pass_training) == (2, True)
    assert (c._arg_count, c._pass_training) == (3, False)
    nl, pos = torch.ones((5, 4, 4)), torch.ones((5, 4))
    box = torch.tensor([[-1.0, -2, -3], [1, 2, 3], [0, 0, 0]])
    assert a([nl, pos, box], False)[0].shape == (5,)
    assert torch.equal(b([nl, pos, box], True)[0], 2 * pos[:, :3])
    assert torch.equal(c([nl, pos, box], False)[0], torch.tensor([6.0, 12.0, 18.0]))
    with pytest.raises(AttributeError):
        htf.SimModel(4)
    with pytest.raises(ValueError):
        a.mapped_nlist(nl)
    assert a.get_config()["nneighbor_cutoff"] == 4


def test_autograd_force_math_matches_oracle(oracle_mod):
    """compute_nlist_forces / _compute_virial / nlist_rinv on torch (CPU tensors) vs the C oracle."""
    pos, lo, hi = synthetic.lattice_fluid((6, 6, 6), 0.7, seed=2)
    nl, _, _ = oracle_mod.nlist(pos, lo, hi, 2.5, 64)
    fe_o, v9_o, _ = oracle_mod.lj(nl)
    m = htf.models.LJVirialModelAutograd(64, virial=True)
    f, v = m([torch.from_numpy(nl), torch.from_numpy(pos), torch.zeros(3, 3)], False)
    np.testing.assert_allclose(f.detach().numpy(), fe_o, rtol=2e-4, atol=2e-4)
    np.testing.assert_allclose(v.detach().numpy().reshape(-1, 9), v9_o, rtol=2e-4, atol=2e-4)
    s = htf.models.SimplePotential(64)
    fs = s([torch.from_numpy(nl), torch.from_numpy(pos)], False)[0]
    d = nl[:, :, :3].astype(np.float64)
    r = np.sqrt((d ** 2).sum(-1, keepdims=True))
    want = np.where(r > 0, -d / np.where(r > 0, r, 1), 0).sum(1)
    np.testing.assert_allclose(fs.detach().numpy(), want, atol=1e-4)


def test_wrap_vector_and_rdf_normalisation():
    box = torch.tensor([[-5.0, -5, -5], [5, 5, 5], [0, 0, 0]])
    r = torch.tensor([9.0, -7.0, 2.0])
    assert torch.allclose(htf.wrap_vector(r, box), torch.tensor([-1.0, 3.0, 2.0]))
    hist = torch.arange(102, dtype=torch.int64)
    rdf, rs = htf.rdf_from_hist(hist, (0.0, 2.5), 100)
    assert rdf.shape == (100,) and rs.shape == (100,)
    shell = np.linspace(0, 2.5, 101, dtype=np.float32)
    np.testing.assert_allclose(rdf.numpy(), np.arange(1, 101) / (shell[1:] ** 3 - shell[:-1] ** 3), rtol=1e-5)


def test_eds_layer_matches_oracle(oracle_mod):
    """torch EDSLayer (masked, device resident) == scalar oracle restatement of htf/layers.py:142-195."""
    rng = np.random.default_rng(3)
    lay = htf.EDSLayer(4.0, 10, learning_rate=0.5, cv_scale=2.0)
    ora = oracle_mod.EDSLayer(4.0, 10, learning_rate=0.5, cv_scale=2.0)
    for i in range(95):
        cv = np.float32(5.0 + 0.3 * rng.standard_normal())
        a = float(lay(torch.tensor(cv)))
        b = float(ora(cv))
        assert abs(a - b) <= 1e-5 * max(1.0, abs(b)), (i, a, b)
    assert int(lay.n) == ora.n and float(lay.adam_t) == ora.t
    with pytest.raises(ValueError):
        htf.EDSLayer(4, 10)


def test_row_shard_partition():
    for n, w in [(10, 2), (10, 3), (1048576, 8), (7, 8)]:
        rows = [htf.parallel.row_shard(n, w, r) for r in range(w)]
        assert rows[0][0] == 0 and rows[-1][1] == n
        assert all(rows[i][1] == rows[i + 1][0] for i in range(w - 1))


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "hoomd-tf_b200"))
    import oracle
    from htf import parallel, synthetic as syn
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pos, lo, hi = syn.lattice_fluid((8, 8, 9), 0.7, seed=4)            # 576 particles: ragged over 2 ranks? no: even
    n = pos.shape[0]
    a, b = parallel.row_shard(n, world, rank)
    shard = torch.from_numpy(pos[a:b].copy())
    full = parallel.allgather_positions(shard, out=torch.empty((n, 4)))
    assert np.array_equal(full.numpy(), pos)
    nl, idx, cnt = oracle.nlist(full.numpy(), lo, hi, 2.5, 64, a, b)     # this rank's rows only
    fe, _, v6 = oracle.lj(nl)
    bins = torch.from_numpy(oracle.rdf_hist(nl, (0.0, 2.5), 100))
    parallel.allreduce_bins(bins)
    cv = parallel.allreduce_scalar(torch.tensor(float((nl[:, :, :3] ** 2).sum())))
    q.put((rank, a, b, fe, bins.numpy(), float(cv)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_row_sharding_gloo(oracle_mod):
    """world_size 2 over gloo: all-gather of the position shards, per-rank rows, all-reduced RDF bins."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    pos, lo, hi = synthetic.lattice_fluid((8, 8, 9), 0.7, seed=4)
    nl, _, _ = oracle_mod.nlist(pos, lo, hi, 2.5, 64)
    fe, _, _ = oracle_mod.lj(nl)
    got = np.concatenate([r[3] for r in res])
    assert np.array_equal(got, fe)                                        # row sharding changes nothing
    bins = oracle_mod.rdf_hist(nl, (0.0, 2.5), 100)
    assert np.array_equal(res[0][4], bins) and np.array_equal(res[1][4], bins)
    assert abs(res[0][5] - float((nl[:, :, :3].astype(np.float64) ** 2).sum())) < 1e-3 * res[0][5]


def test_oracle_pairwise_mlp_matches_torch_autograd():
    """The numpy restatement of the config-3 network (analytic du/dr) against torch autograd of the same network."""
    import oracle
    import htf
    from htf import synthetic
    pos, lo, hi = synthetic.lattice_fluid((6, 6, 6), 0.7, seed=1)
    nl, _, _ = oracle.nlist(pos, lo, hi, 2.5, 64)
    m = htf.models.PairwiseMLPModel(64, r_cut=2.5, seed=3)
    ref = m([torch.from_numpy(nl), None], True)[0].detach().numpy()
    out = oracle.pairwise_mlp(nl, m.raw_parameters().numpy(), 2.5)
    scale = np.abs(ref).max()
    assert np.abs(out - ref).max() < 1e-5 * scale          # fp32 tolerance of the north star
    assert np.abs(ref[:, :3]).max() > 0.1


def test_oracle_pairwise_mlp_force_is_minus_energy_gradient():
    """oracle.pairwise_mlp's analytic du/dr against a central finite difference of its own energies (float64 weights
    would hide nothing here: the check is on the derivative chain, tolerance set by the fp32 energy noise)."""
    import oracle
    rng = np.random.default_rng(0)
    raw = np.concatenate([(rng.standard_normal(64 * 32) / np.sqrt(32)), 0.1 * rng.standard_normal(64),
                          (rng.standard_normal(64 * 64) / 8), 0.1 * rng.standard_normal(64),
                          (rng.standard_normal(64 * 64) / 8), 0.1 * rng.standard_normal(64),
                          (rng.standard_normal(64) / 8), 0.1 * rng.standard_normal(1)]).astype(np.float32)
    nl = np.zeros((1, 4, 4), dtype=np.float32)
    nl[0, 0, :3] = [0.9, 0.3, -0.2]
    nl[0, 1, :3] = [-1.1, 0.8, 0.5]
    nl[0, 2, :3] = [0.2, -1.6, 0.9]                            # slot 3 stays padding
    out = oracle.pairwise_mlp(nl, raw, 2.5)
    h = 2e-3
    for slot in range(3):
        for ax in range(3):
            a, b = nl.copy(), nl.copy()
            a[0, slot, ax] += h; b[0, slot, ax] -= h
            de = (oracle.pairwise_mlp(a, raw, 2.5)[0, 3] - oracle.pairwise_mlp(b, raw, 2.5)[0, 3]) / (2 * h)
            # F_i = sum_j 2 dE_i/dd_ij (compute_nlist_forces): moving ONE slot's d changes e_i by de, its share of F is 2 de
            a1, b1 = nl.copy(), nl.copy()
            a1[0, :, :] = 0; b1[0, :, :] = 0
            a1[0, 0] = nl[0, slot]; b1[0, 0] = nl[0, slot]
            f_slot = oracle.pairwise_mlp(a1, raw, 2.5)[0, ax]   # force of a row holding only this slot
            assert abs(f_slot - 2 * de) < 2e-3 * max(1.0, abs(f_slot)), (slot, ax, f_slot, 2 * de)
    assert np.isfinite(out).all()


def test_slab_plan_and_roi_cover_the_halo():
    """host-side planning of the slab exchange: capacity covers both faces, the ROI covers the slab +- r_cut."""
    from htf import parallel, synthetic
    pos, lo, hi = synthetic.lattice_fluid((8, 8, 32), 0.7, seed=4)
    a, b = parallel.row_shard(pos.shape[0], 4, 1)
    shard = pos[a:b]
    lo_face, hi_face, width, cap = parallel.slab_plan(shard, 2, 2.5)
    z = shard[:, 2]
    assert lo_face == float(z.min()) and hi_face == float(z.max()) and width > 2.5
    assert cap >= max((z < lo_face + width).sum(), (z > hi_face - width).sum()) and cap % 256 == 0
    c, hw = parallel.roi_for_rows(shard, lo, hi, 2.5)
    assert hw[2] >= 0 and c[2] - hw[2] <= lo_face - 2.5 and c[2] + hw[2] >= hi_face + 2.5
    assert hw[0] < 0 and hw[1] < 0                               # x and y span the box: unrestricted
