"""Pins the CPU oracle against the reference's own known-answer-by-construction tests.

The reference stores no golden vectors for this path and neither hoomd nor tensorflow is
importable here, so every test below restates one of the reference's self-checking tests
without HOOMD: same lattice, same cutoff, same tolerance, an independent float64 O(N^2)
computation as the truth (what `hoomd.md.pair.lj` / the Python loops provide there).
"""
import numpy as np
import pytest

from htf import synthetic


def min_image64(d, L):
    return d - np.round(d / L) * L


def brute_pairs(pos, lo, hi, r_cut):
    """float64 O(N^2) neighbor vectors: dict row -> list of (j, d)."""
    xyz = pos[:, :3].astype(np.float64)
    L = np.asarray(hi, np.float64) - np.asarray(lo, np.float64)
    out = []
    for i in range(len(xyz)):
        d = min_image64(xyz - xyz[i], L)
        r = np.sqrt((d ** 2).sum(1))
        js = np.where((r <= r_cut) & (np.arange(len(xyz)) != i))[0]
        out.append((js, d[js]))
    return out


def lj_analytic(pos, lo, hi, r_cut):
    """hoomd.md.pair.lj(epsilon=1, sigma=1, r_cut) net force, per-particle energy and virial (xx,xy,...)."""
    n = pos.shape[0]
    F = np.zeros((n, 3)); E = np.zeros(n); V = np.zeros((n, 6))
    for i, (js, d) in enumerate(brute_pairs(pos, lo, hi, r_cut)):
        r2 = (d ** 2).sum(1)
        ir6 = 1.0 / r2 ** 3
        # F on i = -dU/dr_i ; d = r_j - r_i  ->  F_i = -(48 r^-14 - 24 r^-8) d
        fdivr = (48.0 * ir6 * ir6 - 24.0 * ir6) / r2
        F[i] = -(fdivr[:, None] * d).sum(0)
        E[i] = 0.5 * (4.0 * (ir6 * ir6 - ir6)).sum()
        pick = [(0, 0), (0, 1), (0, 2), (1, 1), (1, 2), (2, 2)]
        for c, (k, l) in enumerate(pick):
            V[i, c] = 0.5 * (fdivr * d[:, k] * d[:, l]).sum()      # HOOMD pair virial
    return F, E, V


@pytest.mark.parametrize("n,a", [(5, 3.0), (3, 4.0)])
def test_lj_forces_vs_analytic_lj(oracle_mod, n, a):
    """htf/test-py/test_tensorflow.py:335-382: LJModel(32) vs hoomd.md.pair.lj, r_cut=5, atol 1e-5."""
    pos, lo, hi = synthetic.square_lattice(n, a)
    pos = synthetic.perturb(pos, lo, hi, 0.15, seed=1)      # the reference runs 20 NVT steps first
    pos[:, 2] = 0.0
    nl, idx, cnt = oracle_mod.nlist(pos, lo, hi, 5.0, 32)
    fe, v9, v6 = oracle_mod.lj(nl)
    F, E, V = lj_analytic(pos, lo, hi, 5.0)
    np.testing.assert_allclose(fe[:, :3], F, atol=1e-5)
    assert np.all(np.sum(F ** 2, axis=1) > 1e-4 ** 2), "forces are too low to assess"
    np.testing.assert_allclose(fe[:, 3], E, atol=1e-5)       # :400-431 (w column = per-particle energy)


def test_lj_virial_vs_pair_virial(oracle_mod):
    """htf/test-py/test_tensorflow.py:619-671: virial xx, xy vs lj.forces[j].virial, atol 1e-5 (3x3, a=4)."""
    pos, lo, hi = synthetic.square_lattice(3, 4.0)
    pos = synthetic.perturb(pos, lo, hi, 0.1, seed=2)
    pos[:, 2] = 0.0
    nl, _, _ = oracle_mod.nlist(pos, lo, hi, 5.0, 32)
    fe, v9, v6 = oracle_mod.lj(nl)
    F, E, V = lj_analytic(pos, lo, hi, 5.0)
    np.testing.assert_allclose(v6[:, 0:2], V[:, 0:2], atol=1e-5)
    np.testing.assert_allclose(v6, V, atol=1e-5)             # all pairs are attractive here (r > 2^(1/6))
    # 3x3 -> 6 layout (htf/TensorflowCompute.cc:294-299)
    np.testing.assert_array_equal(v6, v9[:, [0, 1, 2, 4, 5, 8]])
    np.testing.assert_allclose(v9[:, 1], v9[:, 3], rtol=1e-6, atol=1e-12)   # (w*dx)*dy vs (w*dy)*dx


def test_inverse_r_forces_vs_python_loop(oracle_mod):
    """htf/test-py/test_tensorflow.py:20-35,81-104: SimplePotential (F = -sum d/|d|) vs the O(N^2) loop."""
    pos, lo, hi = synthetic.square_lattice(3, 4.0)
    pos = synthetic.perturb(pos, lo, hi, 0.2, seed=3)
    pos[:, 2] = 0.0
    N, rcut = 9, 5.0
    nl, _, _ = oracle_mod.nlist(pos, lo, hi, rcut, N - 1)
    d = nl[:, :, :3].astype(np.float64)
    r = np.sqrt((d ** 2).sum(-1, keepdims=True))
    with np.errstate(invalid="ignore", divide="ignore"):
        fr = np.where(r > 0, -d / r, 0.0)
    got = fr.sum(1)
    want = np.zeros((N, 3))
    xyz = pos[:, :3].astype(np.float64)
    L = np.asarray(hi, np.float64) - np.asarray(lo, np.float64)
    for i in range(N):
        for j in range(i + 1, N):
            rr = min_image64(xyz[j] - xyz[i], L)
            rd = np.sqrt((rr ** 2).sum())
            if rd <= rcut:
                f = -rr / rd
                want[i] += f
                want[j] -= f
    np.testing.assert_allclose(got, want, atol=1e-5)


def test_nlist_is_full_not_half(oracle_mod):
    """htf/test-py/test_tensorflow.py:559-579: 3x3 lattice a=4, r_cut=5 -> every row has 4 neighbors."""
    pos, lo, hi = synthetic.square_lattice(3, 4.0)
    nl, _, cnt = oracle_mod.nlist(pos, lo, hi, 5.0, 32)
    ncount = np.sum(np.sum(nl ** 2, axis=2) > 0.1, axis=1)
    assert np.min(ncount) == 4 and np.array_equal(ncount, cnt)


def test_nlist_vs_bruteforce_bcc(oracle_mod):
    """htf/test-py/test_utils.py:401-430: bcc 4x4x4 a=4.0, r_cut=5, NN=32; sorted per-row r to 5 decimals."""
    pos, lo, hi = synthetic.bcc_lattice(4, 4.0)
    pos = synthetic.perturb(pos, lo, hi, 0.1, seed=4)
    nl, idx, cnt = oracle_mod.nlist(pos, lo, hi, 5.0, 32)
    assert cnt.max() <= 32
    r = np.sqrt((nl[:, :, :3].astype(np.float64) ** 2).sum(-1))
    for i, (js, d) in enumerate(brute_pairs(pos, lo, hi, 5.0)):
        want = np.zeros(32); want[:len(js)] = np.sqrt((d ** 2).sum(1))
        np.testing.assert_array_almost_equal(np.sort(r[i]), np.sort(want), decimal=5)
        assert set(idx[i][idx[i] >= 0]) == set(js)


def test_cells_equal_bruteforce(oracle_mod):
    """the oracle's cell-list candidate path is bit-identical to its O(N^2) loop (incl. 2-D and tiny grids)."""
    cases = [synthetic.lattice_fluid((8, 8, 12), 0.7, seed=6) + (2.5, 64),
             synthetic.square_lattice(16, 2.0) + (3.0, 64),
             synthetic.lattice_fluid((4, 8, 8), 0.1, seed=1, two_types_p=0.3) + (3.0, 32)]
    for pos, lo, hi, rc, K in cases:
        a = oracle_mod.nlist(pos, lo, hi, rc, K, cells=False)
        b = oracle_mod.nlist(pos, lo, hi, rc, K, cells=True)
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
    pos, lo, hi, rc, K = cases[0]
    full = oracle_mod.nlist(pos, lo, hi, rc, K, cells=True)
    part = oracle_mod.nlist(pos, lo, hi, rc, K, 100, 300, cells=True)
    for x, y in zip(full, part):
        assert np.array_equal(x[100:300], y)          # batch_size / offset chunking (TensorflowCompute.cc:143-150)


def test_types_carried_in_w(oracle_mod):
    """htf/test-py/test_tensorflow.py:46-70: three types visible in nlist[..., 3] and positions[:, 3]."""
    pos, lo, hi = synthetic.lattice_fluid((5, 5, 5), 0.3, seed=8)
    pos[:, 3] = np.arange(pos.shape[0]) % 3
    nl, idx, cnt = oracle_mod.nlist(pos, lo, hi, 3.0, 32)
    assert len(np.unique(nl[:, :, 3].astype(int))) == 3
    v = idx >= 0
    assert np.array_equal(nl[:, :, 3][v], pos[idx[v], 3])


def test_overflow_wraps_modulo_k(oracle_mod):
    """htf/test-py/test_tensorflow.py:830-848 (K=4, r_cut=10, 8x8 lattice) + the % K rule of TensorflowCompute.cc:370."""
    pos, lo, hi = synthetic.square_lattice(8, 4.0)
    nl, idx, cnt = oracle_mod.nlist(pos, lo, hi, 10.0, 4)
    assert cnt.min() > 4                                   # 'Neighbor list is full!'
    big, idx_big, _ = oracle_mod.nlist(pos, lo, hi, 10.0, 64)
    for r in range(pos.shape[0]):
        c = cnt[r]
        for s in range(4):                                 # slot s holds the last neighbor q with q % 4 == s
            q = max(q for q in range(c) if q % 4 == s)
            assert idx[r, s] == idx_big[r, q]


def test_typed_rdf_symmetry(oracle_mod):
    """htf/test-py/test_tensorflow.py:450-485: rdf(A->B) == rdf(B->A), sum > 0."""
    pos, lo, hi = synthetic.typed_chains()
    pos = synthetic.perturb(pos, lo, hi, 0.05, seed=1)
    nl, _, cnt = oracle_mod.nlist(pos, lo, hi, 10.0, 256)
    assert cnt.max() <= 256
    ha = oracle_mod.rdf_hist(nl, (0, 10), 100, row_type=pos[:, 3], type_i=0, type_j=1)
    hb = oracle_mod.rdf_hist(nl, (0, 10), 100, row_type=pos[:, 3], type_i=1, type_j=0)
    rdfa, rs = oracle_mod.rdf_from_hist(ha, (0, 10))
    rdfb, _ = oracle_mod.rdf_from_hist(hb, (0, 10))
    assert rdfa.sum() > 0 and len(rdfa) == 100
    np.testing.assert_array_almost_equal(rdfa, rdfb)


def test_rdf_histogram_against_float64_binning(oracle_mod):
    """the recalled TF rule (UNPINNED) must at least agree with a float64 histogram away from bin edges."""
    pos, lo, hi = synthetic.lattice_fluid((8, 8, 8), 0.7, seed=12)
    nl, _, _ = oracle_mod.nlist(pos, lo, hi, 2.5, 64)
    h = oracle_mod.rdf_hist(nl, (0.0, 2.5), 100)
    assert h.sum() == nl.shape[0] * 64
    r = np.sqrt((nl[:, :, :3].astype(np.float64) ** 2).sum(-1)).ravel()
    b64 = np.minimum((r / (2.5 / 102)).astype(np.int64), 101)
    h64 = np.bincount(b64, minlength=102)
    assert np.abs(h - h64).sum() <= 4                      # only values within an ulp of an edge may move
    assert h[0] == (r == 0).sum()                          # padded zeros land in bin 0 ("remove 0s", simmodel.py:667)


def test_eds_layer_statistics(oracle_mod):
    """htf/layers.py:142-195: window = second half of each period, Adam step at n == period-1, alpha sign."""
    eds = oracle_mod.EDSLayer(set_point=4.0, period=10, learning_rate=0.5)
    rng = np.random.default_rng(0)
    alphas = [float(eds(5.0 + 0.1 * rng.standard_normal())) for _ in range(30)]
    assert all(a == 0.0 for a in alphas[:9])               # no update before the first period ends
    assert alphas[9] != 0.0 and alphas[9] == alphas[10] == alphas[18]
    # cv above the set point: gradient = -2 (mean - sp) ssd / period / 2 < 0 -> Adam moves alpha up
    assert alphas[9] > 0.0 and alphas[19] > alphas[9]
    assert eds.n == 0 and eds.t == 3


def test_oracle_reproduces_golden_fixtures():
    """tests/golden/*.npz (made by tests/golden/make_golden.py after an analytic float64 LJ check) pin the oracle."""
    import glob
    import os
    import oracle
    files = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))
    assert len(files) >= 3
    for f in files:
        g = np.load(f)
        K, r_cut = int(g["K"]), float(g["r_cut"])
        nl, idx, cnt = oracle.nlist(g["pos"], g["lo"], g["hi"], r_cut, K)
        key = np.where(idx < 0, np.iinfo(np.int32).max, idx)
        order = np.argsort(key, axis=1, kind="stable")
        assert np.array_equal(np.take_along_axis(idx, order, axis=1), g["idx_sorted"]), f
        assert np.array_equal(np.take_along_axis(nl, order[:, :, None], axis=1).view(np.uint32), g["nlist_sorted"].view(np.uint32)), f
        assert np.array_equal(cnt, g["count"])
        fe, _, v6 = oracle.lj(nl)
        assert np.array_equal(fe, g["force_energy"]) and np.array_equal(v6, g["virial6"]), f
        assert np.array_equal(oracle.rdf_hist(nl, (0.0, r_cut), 100), g["rdf_hist"]), f


def test_coordination_cv_against_float64_and_finite_differences(oracle_mod):
    """The smooth coordination CV of BASELINE config 5 has no reference implementation: the oracle is pinned
    against a float64 evaluation of cn_i = sum_j 1/(1 + (r/r0)^6) and a central difference of it for the gradient sums."""
    from htf import synthetic
    pos, lo, hi = synthetic.lattice_fluid((5, 5, 5), 0.8442, seed=5)
    nl, _, _ = oracle_mod.nlist(pos, lo, hi, 2.8, 96)
    r0 = 1.3
    cn, g = oracle_mod.coordination_cv(nl, r0)

    def cn64(nl_):
        a = nl_[..., :3].astype(np.float64) + 1e-7                      # nlist_rinv's safe norm offsets
        r = np.sqrt((a * a).sum(-1))
        s = np.where(r > 3e-6, 1.0 / (1.0 + (r / r0) ** 6), 0.0)
        return s.sum(1)
    ref = cn64(nl)
    assert np.abs(cn - ref).max() <= 1e-5 * np.abs(ref).max()
    h = 1e-3
    for ax in range(3):
        up, dn = nl.astype(np.float64), nl.astype(np.float64)          # perturb in float64: h is below fp32 resolution of d
        mask = (np.abs(nl[..., :3]).sum(-1) > 0)                          # move every real neighbor of every row along ax
        up[..., ax] += h * mask
        dn[..., ax] -= h * mask
        fd = (cn64(up) - cn64(dn)) / (2 * h)
        assert np.abs(g[:, ax] - fd).max() <= 2e-4 * max(1.0, np.abs(fd).max()), ax


def test_reference_fixture_comparator_and_any_committed_reference_fixtures():
    """oracle/make_ref_fixtures.py runs the unmodified reference (hoomd + tensorflow) on the golden systems wherever
    that stack exists and compares with the oracle-made fixtures.  Here: its comparator accepts the oracle's own
    output, rejects a one-count change of the RDF histogram, and any committed tests/golden/*_ref.npz must pass."""
    import glob
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("make_ref_fixtures", os.path.join(root, "oracle", "make_ref_fixtures.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    g = dict(np.load(os.path.join(root, "tests", "golden", "fluid216_rc25.npz")))
    r = {"nlist_sorted": m.sort_rows_by_value(g["nlist_sorted"]), "count": g["count"], "force_energy": g["force_energy"],
         "virial6": g["virial6"], "rdf_hist": g["rdf_hist"], "hoomd_double": np.bool_(False),
         "hoomd_version": np.str_("-"), "tf_version": np.str_("-")}
    assert m.compare("oracle vs itself", g, r)
    bad = dict(r)
    bad["rdf_hist"] = r["rdf_hist"].copy()
    bad["rdf_hist"][5] += 1
    assert not m.compare("one count off", g, bad)
    for f in sorted(glob.glob(os.path.join(root, "tests", "golden", "*_ref.npz"))):
        name = os.path.basename(f)[:-8]
        assert m.compare(name, dict(np.load(os.path.join(root, "tests", "golden", name + ".npz"))), dict(np.load(f)))


def _mlp_raw(seed=3):
    rng = np.random.default_rng(seed)
    parts = []
    for fo, fi in ((64, 32), (64, 64), (64, 64), (1, 64)):
        parts.append((rng.standard_normal((fo, fi)) / np.sqrt(fi)).ravel())
        parts.append(0.1 * rng.standard_normal(fo))
    return np.concatenate(parts)


def test_training_gradient_oracle_matches_torch_double_backward(oracle_mod):
    """The reverse sweep through the value / tangent chain (oracle.pairwise_mlp_train_grads, the checker of the
    tensor-core training kernel) against torch differentiating MSE(compute_nlist_forces output, labels) -- the double
    backward Keras performs in the reference's train_on_batch (htf/tensorflowcompute.py:346-370) -- in float64."""
    torch = pytest.importorskip("torch")
    from htf import synthetic
    pos, lo, hi = synthetic.lattice_fluid((5, 5, 5), 0.7, seed=3)
    nl, _, _ = oracle_mod.nlist(pos, lo, hi, 2.5, 64)
    raw = _mlp_raw()
    labels = np.random.default_rng(1).standard_normal((pos.shape[0], 4))
    loss, g, pred = oracle_mod.pairwise_mlp_train_grads(nl, raw, 2.5, labels)
    t = torch.tensor(raw, dtype=torch.float64, requires_grad=True)
    cuts = np.cumsum([0, 2048, 64, 4096, 64, 4096, 64, 64, 1])
    W1, b1, W2, b2, W3, b3, w4, b4 = (t[cuts[i]:cuts[i + 1]] for i in range(8))
    nlt = torch.tensor(nl, dtype=torch.float64, requires_grad=True)
    a = nlt[..., :3] + 1e-7
    r = torch.sqrt((a * a).sum(-1))
    mu = torch.linspace(0, 2.5, 32, dtype=torch.float64)
    h = torch.exp(-(r[..., None] - mu) ** 2 / (mu[1] - mu[0]))
    for W, b, fi in ((W1, b1, 32), (W2, b2, 64), (W3, b3, 64)):
        h = torch.tanh(h @ W.reshape(64, fi).T + b)
    u = torch.where(r > 3e-6, h @ w4 + b4[0], torch.zeros_like(r))
    e = 0.5 * u.sum(1)
    G = torch.autograd.grad(e.sum(), nlt, create_graph=True)[0]
    predt = torch.cat([(2 * G).sum(1)[:, :3], e[:, None]], 1)
    L = ((predt - torch.tensor(labels)) ** 2).mean()
    L.backward()
    assert abs(loss - float(L)) <= 1e-12 * abs(float(L))
    np.testing.assert_allclose(pred, predt.detach().numpy(), rtol=1e-10, atol=1e-12)
    gt = t.grad.numpy()
    assert np.abs(g - gt).max() <= 1e-10 * np.abs(gt).max()
    # float32 forward restatement agrees with the float64 predictions
    pred32 = oracle_mod.pairwise_mlp(nl, raw, 2.5)
    assert np.abs(pred32 - pred).max() <= 2e-4 * np.abs(pred).max()


def test_adam_oracle_matches_torch_adam_shape_and_keras_formula(oracle_mod):
    """oracle.adam_step is tf.keras.optimizers.Adam's update; the first steps coincide with torch.optim.Adam up to the
    placement of epsilon (sqrt(v) + eps vs sqrt(v_hat) + eps), i.e. to ~eps / |g| relative."""
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(0)
    p0 = rng.standard_normal(50).astype(np.float32)
    p, m, v = p0.copy(), np.zeros(50, np.float32), np.zeros(50, np.float32)
    tp = torch.tensor(p0.copy(), requires_grad=True)
    opt = torch.optim.Adam([tp], lr=1e-3, eps=1e-7)
    for t in range(1, 6):
        g = rng.standard_normal(50).astype(np.float32)
        p, m, v = oracle_mod.adam_step(p, g, m, v, t)
        tp.grad = torch.tensor(g)
        opt.step()
        np.testing.assert_allclose(p, tp.detach().numpy(), rtol=0, atol=2e-6)
    assert np.abs(p - p0).max() > 1e-3                        # it moved: five steps of about lr each
