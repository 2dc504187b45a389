"""GPU tests of the user-facing API: the reference's own test cases (htf/test-py/test_tensorflow.py,
test_utils.py) restated on htf.sim (the HOOMD stand-in) + tfcompute + SimModel."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def py_forces(system, rcut):
    """'1 / r^2 force' loop of htf/test-py/test_tensorflow.py:20-35."""
    snap = system.take_snapshot()
    position = snap.particles.position.astype(np.float64)
    N = len(position)
    forces = np.zeros((N, 3))
    for i in range(N):
        for j in range(i + 1, N):
            r = snap.box.min_image(position[j] - position[i])
            rd = np.sqrt(np.sum(r ** 2))
            if rd <= rcut:
                f = -r / rd
                forces[i, :] += f
                forces[j, :] -= f
    return forces


def lattice_system(n, a, kT=None, seed=2, dt=0.005):
    import htf
    system = htf.sim.create_lattice(htf.sim.sq(a), n=[n, n])
    system.integrator = htf.sim.NVE(dt)
    if kT is not None:
        system.randomize_velocities(kT=kT, seed=seed)
    return system


@pytest.mark.parametrize("batch_size", [None, 4])
def test_force_overwrite(batch_size):
    """test_tensorflow.py:81-129: SimplePotential forces == the O(N^2) python loop along an NVE run."""
    import htf
    N, rcut = 9, 5.0
    model = htf.models.SimplePotential(N - 1)
    tfc = htf.tfcompute(model)
    system = lattice_system(3, 4.0, kT=2)
    tfc.attach(htf.sim.nlist_cell(system), r_cut=rcut, batch_size=batch_size)
    system.run(2)
    for _ in range(3):
        want = py_forces(system, rcut)
        got = system.compute_net_force()[:, :3].cpu().numpy()
        np.testing.assert_allclose(got, want, atol=1e-5)
        system.run(100)


def test_lj_forces_vs_pair_lj_and_autograd():
    """test_tensorflow.py:335-382: LJModel(32) vs hoomd.md.pair.lj(r_cut=5) over a trajectory, atol 1e-5;
    the fused kernel and the literal autograd body agree."""
    import htf
    system = lattice_system(5, 3.0, kT=1, seed=1)
    fused = htf.tfcompute(htf.models.LJModel(32))
    fused.attach(htf.sim.nlist_cell(system), r_cut=5.0)
    lit = htf.tfcompute(htf.models.LJModelAutograd(32))
    lit.attach(htf.sim.nlist_cell(system), r_cut=5.0)
    system.forces.remove(lit)                       # only `fused` drives the dynamics
    lj = htf.sim.ReferenceLJ(system, r_cut=5.0)
    system.run(20)
    for t in range(10):
        system.run(1)
        f = fused.compute_forces(0).cpu().numpy()
        g = lit.compute_forces(0).cpu().numpy()
        ref = lj.compute_forces().cpu().numpy()
        np.testing.assert_allclose(f[:, :3], ref[:, :3], atol=1e-5)
        np.testing.assert_allclose(f[:, 3], ref[:, 3], atol=1e-5)
        np.testing.assert_allclose(g, f, atol=2e-6)
        assert np.all(np.sum(ref[:, :3] ** 2, axis=1) > 1e-4 ** 2), "Forces are too low to assess!"


def test_lj_energy_conservation():
    """test_tensorflow.py:532-557: NVE with the LJ model conserves PE + KE (atol 1e-3)."""
    import htf
    system = lattice_system(3, 4.0, kT=0.8, seed=1, dt=0.001)
    tfc = htf.tfcompute(htf.models.LJModel(32))
    tfc.attach(htf.sim.nlist_cell(system), r_cut=5.0)
    energy = []
    for i in range(6):
        system.run(250)
        pe = tfc.get_log_value()
        ke = 0.5 * float((system.velocities.double() ** 2).sum())
        energy.append(pe + ke)
        if i > 1:
            np.testing.assert_allclose(energy[-1], energy[-2], atol=1e-3)


def test_nlist_count_and_accessors():
    """test_tensorflow.py:559-579 (full, not half: 4 neighbors on the 3x3 lattice) and :46-70 (array getters)."""
    import htf
    system = lattice_system(3, 4.0, kT=0.8, seed=1, dt=0.001)
    tfc = htf.tfcompute(htf.models.LJVirialModel(32, virial=True))
    tfc.attach(htf.sim.nlist_cell(system), r_cut=5.0)
    system.run(1)
    nl = tfc.get_nlist_array()
    ncount = np.sum(np.sum(nl ** 2, axis=2) > 0.1, axis=1)
    assert np.min(ncount) == 4
    assert tfc.get_virial_array().shape == (9, 9) and tfc.get_forces_array().shape == (9, 4)
    assert tfc.get_positions_array().shape == (9, 4)


def test_virial_vs_pair_virial():
    """test_tensorflow.py:619-671: virial xx, xy vs the pair virial of hoomd.md.pair.lj, atol 1e-5."""
    import htf
    system = lattice_system(3, 4.0, kT=1, seed=1)
    tfc = htf.tfcompute(htf.models.LJVirialModel(32, virial=True))
    tfc.attach(htf.sim.nlist_cell(system), r_cut=5.0)
    system.run(9)
    v = tfc.get_virial_array()
    pos = system.positions[:, :3].double().cpu().numpy()
    L = system.box.L
    for i in range(9):
        d = pos - pos[i]
        d -= np.round(d / L) * L
        r2 = (d ** 2).sum(1)
        m = (r2 <= 25.0) & (np.arange(9) != i)
        ir6 = 1.0 / r2[m] ** 3
        fdivr = (48 * ir6 * ir6 - 24 * ir6) / r2[m]
        np.testing.assert_allclose(v[i][0], 0.5 * (fdivr * d[m, 0] * d[m, 0]).sum(), atol=1e-5)
        np.testing.assert_allclose(v[i][1], 0.5 * (fdivr * d[m, 0] * d[m, 1]).sum(), atol=1e-5)


def test_overflow_raises():
    """test_tensorflow.py:830-848: K=4, r_cut=10 with check_nlist=True -> 'Neighbor list is full!'."""
    import htf
    system = lattice_system(8, 4.0, kT=1, seed=1)
    tfc = htf.tfcompute(htf.models.LJModel(4, check_nlist=True))
    tfc.attach(htf.sim.nlist_cell(system), r_cut=10.0)
    with pytest.raises(RuntimeError, match="Neighbor list is full"):
        system.run(2)


def test_skew_fails():
    """test_tensorflow.py:321-333: a tilted box is rejected."""
    import htf
    system = lattice_system(3, 4.0)
    system.tilt = (0.5, 0.0, 0.0)
    tfc = htf.tfcompute(htf.models.WrapModel(0, output_forces=False))
    system.integrator = htf.sim.NVE(0.005)
    tfc.attach(system=system)
    with pytest.raises(htf._lib.HtfError, match="skewed"):
        system.run(1)


def test_rdf_models():
    """test_tensorflow.py:433-485: LJRDF running mean; typed RDF A->B == B->A."""
    import htf
    system = lattice_system(3, 4.0, kT=0.8, seed=1, dt=0.001)
    model = htf.models.LJRDF(32)
    tfc = htf.tfcompute(model)
    tfc.attach(htf.sim.nlist_cell(system), r_cut=5.0, batch_size=4)
    system.run(10)
    rdf = model.avg_rdf.result().cpu().numpy()
    assert len(rdf) > 5 and np.sum(rdf) > 0
    pos, lo, hi = htf.synthetic.typed_chains()
    system = htf.sim.System(pos, lo, hi)
    system.integrator = htf.sim.NVE(0.001)
    system.randomize_velocities(kT=0.8, seed=1)
    model = htf.models.LJTypedModel(256)
    tfc = htf.tfcompute(model)
    tfc.attach(htf.sim.nlist_cell(system), r_cut=10.0)
    system.run(10)
    rdfa, rdfb = model.avg_rdfa.result().cpu().numpy(), model.avg_rdfb.result().cpu().numpy()
    assert np.sum(rdfa) > 0
    np.testing.assert_array_almost_equal(rdfa, rdfb)


def test_mapped_nlist():
    """test_tensorflow.py:581-617: bead 0 = mean of the AA positions; AA and CG type sets only share 0."""
    import htf
    N, CGN = 9, 2
    model = htf.models.MappedNlist(N - 1, output_forces=False)
    tfc = htf.tfcompute(model)
    system = lattice_system(3, 4.0, dt=0.001)
    aa_group, mapped_group = tfc.enable_mapped_nlist(system, htf.models.MappedNlist.my_map)
    assert len(aa_group) == N and len(mapped_group) == 2 and len(system) == N + CGN
    system.randomize_velocities(kT=0.8, seed=1)
    system.velocities[N:] = 0
    tfc.attach(htf.sim.nlist_cell(system), r_cut=5.0, save_output_period=2)
    system.run(8)
    positions = tfc.outputs[0].reshape(-1, N + CGN, 4)
    np.testing.assert_allclose(positions[1:, N, :3], np.mean(positions[1:, :N, :3], axis=1), atol=1e-5)
    aa = set(np.unique(tfc.outputs[1][..., -1].astype(int)))
    cg = set(np.unique(tfc.outputs[2][..., -1].astype(int)))
    assert aa.intersection(cg) == set([0])


def test_training_force_matching():
    """test_tensorflow.py:400-431 + :155-271: label mode with reference forces; LJModel-vs-label error < 1e-5,
    labels land in get_forces_array(); a trainable model's weights move."""
    import htf
    Ne, rcut = 5, 3.0
    system = lattice_system(Ne, 2.0, kT=0.8, seed=1, dt=0.01)
    lj = htf.sim.ReferenceLJ(system, r_cut=rcut)
    system.forces.append(lj)
    model = htf.models.LJModel(32, output_forces=False)
    model.compile(loss="MeanSquaredError")
    tfc = htf.tfcompute(model)
    tfc.attach(htf.sim.nlist_cell(system), train=True, r_cut=rcut, period=100)
    tfc.set_reference_forces(lj)
    system.run(300)
    assert abs(float(model.metrics[0].result())) < 1e-5
    # the label buffer holds the summed reference forces of the last update (w column = per-particle energy)
    tfc.period = 1
    system.run(1)
    np.testing.assert_allclose(tfc.get_forces_array(), lj.forces.cpu().numpy(), rtol=1e-6)
    system.half_step_hooks.remove(tfc)
    tm = htf.models.TrainModel(16, output_forces=False, dim=8, top_neighs=5)
    tm.compile(loss=["MeanSquaredError", None])
    w0 = tm.dense1.weight.detach().clone()
    tfc2 = htf.tfcompute(tm)
    tfc2.attach(htf.sim.nlist_cell(system), train=True, r_cut=rcut, save_output_period=2)
    tfc2.set_reference_forces(lj)
    system.run(40)
    assert float((tm.dense1.weight.detach().cpu() - w0.cpu()).abs().max()) > 1e-4
    assert tfc2.outputs[0].shape[0] >= 10


def test_compute_nlist_known_answers():
    """htf/test-py/test_utils.py:187-270: 10 collinear particles."""
    import htf
    N = 10
    positions = torch.arange(N, dtype=torch.float32)[:, None].repeat(1, 3).cuda()
    box_size = [100.0, 100.0, 100.0]
    nlist = htf.compute_nlist(positions, 100.0, 9, box_size, return_types=False, sorted=True).cpu().numpy()
    np.testing.assert_array_almost_equal(nlist[0, 0, :], [1, 1, 1, 1])
    np.testing.assert_array_almost_equal(nlist[-1, -1, :], [-9, -9, -9, 0])
    ext = torch.cat([positions, torch.zeros((N, 1), device="cuda")], dim=1)
    nlist = htf.compute_nlist(ext, 100.0, 9, box_size, return_types=True, sorted=True).cpu().numpy()
    np.testing.assert_array_almost_equal(nlist[0, 0, :], [1, 1, 1, 0])
    em = np.zeros((N, N), dtype=bool)
    em[0, 1] = em[0, 2] = True
    nlist = htf.compute_nlist(positions, 100.0, 9, box_size, sorted=True, exclusion_matrix=em).cpu().numpy()
    np.testing.assert_array_almost_equal(nlist[0, 0, 3], 3)
    np.testing.assert_array_almost_equal(nlist[-1, -1, :], [-9, -9, -9, 0])
    nlist = htf.compute_nlist(positions, 5.5, 9, box_size, sorted=True).cpu().numpy()
    np.testing.assert_array_almost_equal(nlist[0, 0, :], [1, 1, 1, 1])
    np.testing.assert_array_almost_equal(nlist[-1, -1, :], [0, 0, 0, 0])
    with pytest.raises(ValueError):
        htf.compute_nlist(positions, 5.5, 9, box_size, return_types=True)


def test_nlist_compare_and_pairwise():
    """test_utils.py:401-430 (tfcompute nlist vs compute_nlist, sorted r to 5 decimals) and :432-437."""
    import htf
    system = htf.sim.create_lattice(htf.sim.bcc(4.0), n=4)
    system.integrator = htf.sim.NVE(0.001)
    system.randomize_velocities(kT=0.8, seed=1)
    tfc = htf.tfcompute(htf.models.LJModel(32))
    tfc.attach(htf.sim.nlist_cell(system), r_cut=5.0)
    system.run(50)
    nl = tfc.get_nlist_array()
    cn = htf.compute_nlist(system.positions, 5.0, 32, system.box.L).cpu().numpy()
    r = np.sort(np.sqrt((nl[:, :, :3] ** 2).sum(-1)), axis=1)
    cr = np.sort(np.sqrt((cn[:, :, :3].astype(np.float64) ** 2).sum(-1)), axis=1)
    np.testing.assert_array_almost_equal(r, cr, decimal=5)
    out = htf.compute_pairwise(htf.models.LJModel(4), np.linspace(0.5, 1.5, 5))
    assert out[0].shape[0] == 5
    e = out[0][:, 0, 3]
    rr = np.linspace(0.5, 1.5, 5)
    np.testing.assert_allclose(e, 2 * (rr ** -12 - rr ** -6), rtol=1e-3, atol=1e-4)


def test_eds_bias_converges():
    """test_utils.py:447-461: EDS drives the CV mean to within sqrt(0.5) of the set point 4."""
    import htf
    system = lattice_system(3, 4.0, kT=0.2, seed=2, dt=0.05)
    model = htf.models.EDSModel(0, set_point=4.0)
    tfc = htf.tfcompute(model)
    tfc.attach(system=system, save_output_period=10)
    system.run(1000)
    assert np.isfinite(np.mean(tfc.outputs[0]))
    assert (float(model.cv_avg.result()) - 4) ** 2 < 0.5


@pytest.mark.gpu
def test_device_integrators_match_torch_formulas_and_conserve_energy():
    """htf_integrate_half (velocity Verlet kick+drift+wrap / kick) against the same step written with torch ops, then
    energy conservation over a config-1 style LJ run that never leaves the GPU, then the Langevin thermostat."""
    import htf
    from htf import sim, synthetic
    pos, lo, hi = synthetic.lattice_fluid((4, 8, 8), 0.10, seed=1)
    def make(fused):
        s = sim.System(pos, lo, hi).randomize_velocities(2.0, seed=3)
        model = htf.models.LJModel(32)
        tfc = htf.tfcompute(model)
        tfc.attach(sim.nlist_cell(s), r_cut=3.0)
        s.forces.append(tfc)
        s.integrator = sim.NVE(0.005, fused=fused)
        return s
    a, b = make(True), make(False)
    a.run(5); b.run(5)
    assert torch.allclose(a.positions, b.positions, atol=2e-6) and torch.allclose(a.velocities, b.velocities, atol=2e-6)
    def energy(s):
        return float(0.5 * (s.velocities.double() ** 2).sum() + s.net_force[:, 3].double().sum())
    e0 = energy(a)
    a.run(400)
    # plain truncation at r_cut (no shift): every pair crossing the cutoff moves the total by 5.5e-3, so only drift matters
    assert abs(energy(a) - e0) < 0.01 * abs(e0), (e0, energy(a))
    assert bool(((a.positions[:, :3] >= torch.tensor(lo, device="cuda")) & (a.positions[:, :3] < torch.tensor(hi, device="cuda"))).all())
    # Langevin on an ideal gas: the velocity variance relaxes to kT
    class NoForce:
        def compute_forces(self, t):
            return torch.zeros((g.N, 4), device="cuda")
    p2, lo2, hi2 = synthetic.lattice_fluid((16, 16, 16), 0.5, seed=2)
    g = sim.System(p2, lo2, hi2)
    g.forces.append(NoForce())
    g.integrator = sim.Langevin(0.01, kT=1.5, seed=9, gamma=2.0)
    g.run(400)
    T = float((g.velocities.double() ** 2).mean())
    assert abs(T - 1.5) < 0.1, T


def test_wca_layer_model_runs_and_matches_closed_form():
    """htf/test-py/test_layers.py:9-23 (WCA(32) on the 3x3 a=4 lattice, r_cut 5, batch_size 4, 10 NVE steps) plus the
    closed form of the layer: energy = (sigma/r)^6 for r < 2^(1/3) sigma, clipped to [0, 10] (htf/layers.py:52-98)."""
    import htf
    model = htf.models.WCA(32)
    tfc = htf.tfcompute(model)
    system = lattice_system(3, 4.0, kT=0.8, seed=1, dt=0.001)
    tfc.attach(htf.sim.nlist_cell(system), r_cut=5.0, batch_size=4)
    system.run(10)
    f = tfc.get_forces_array()
    assert f.shape == (9, 4) and np.isfinite(f).all()
    # closed form on a hand-made neighbor tensor
    wca = htf.WCARepulsion(1.2).cuda()
    r = torch.tensor([0.5, 1.0, 1.4, 1.6, 3.0], device="cuda")
    nl = torch.zeros((1, 8, 4), device="cuda")
    nl[0, :5, 0] = r
    e = wca(nl)[0].detach().cpu().numpy()
    rr = r.cpu().numpy().astype(np.float64)
    want = np.where(rr < 1.2 * 2 ** (1 / 3), np.clip((1.2 / (rr + 3e-6)) ** 6, 0, 10), 0.0)
    np.testing.assert_allclose(e[:5], want, rtol=2e-5, atol=1e-6)
    assert np.all(e[5:] == 0.0)                              # padded slots carry no energy
    assert float(wca.regularization()) == pytest.approx(-1e-3 * 1.2)
    # the layer is trainable: d(sum energy)/d sigma is finite and non-zero
    wca(nl).sum().backward()
    assert wca.sigma.grad is not None and float(wca.sigma.grad) != 0.0


class _FakeGroup:
    def __init__(self, u, sel):
        self.u, self.sel = u, sel
        self.atoms = self

    @property
    def positions(self):
        return self.u.frames[self.u.cursor][self.sel]

    @property
    def types(self):
        return self.u.types[self.sel]

    def __len__(self):
        return int(self.sel.sum())


class _FakeTs:
    def __init__(self, frame, n_atoms, dimensions):
        self.frame, self.n_atoms, self.dimensions = frame, n_atoms, dimensions


class _FakeTrajectory:
    def __init__(self, u):
        self.u = u
        self.totaltime = float(len(u.frames))

    def __iter__(self):
        for i in range(len(self.u.frames)):
            self.u.cursor = i
            yield _FakeTs(i, self.u.frames[i].shape[0], self.u.dimensions)


class _FakeUniverse:
    """The slice of the MDAnalysis API that iter_from_trajectory touches (htf/utils.py:627-749)."""

    def __init__(self, frames, types, box):
        self.frames, self.types, self.cursor = frames, np.asarray(types), 0
        self.dimensions = np.array(list(box) + [90.0, 90.0, 90.0])
        self.trajectory = _FakeTrajectory(self)

    def select_atoms(self, selection):
        sel = np.ones(len(self.types), bool) if selection == "all" else (self.types == selection.split()[-1])
        return _FakeGroup(self, sel)


def test_iter_from_trajectory(oracle_mod):
    """htf/test-py/test_utils.py:599-636 on a synthetic 'universe': frames come out with per-frame neighbor lists that
    equal the O(N^2) rule of utils.compute_nlist (:75-161), positions carry the type index, the box is the hoomd
    [3,3] tensor, the period / selection arguments work, and the LJ model gives non-zero forces on every frame."""
    import htf
    rng = np.random.default_rng(5)
    L = np.array([12.0, 14.0, 16.0])
    n, NN, r_cut = 300, 48, 3.0
    frames = [(rng.random((n, 3)) * L).astype(np.float32) for _ in range(4)]
    types = np.array(["A", "B"])[rng.integers(0, 2, n)]
    u = _FakeUniverse(frames, types, L)
    model = htf.models.LJVirialModel(NN, virial=True)
    seen = []
    for inputs, ts in htf.iter_from_trajectory(NN, u, r_cut=r_cut, period=1):
        nlist, positions, box = inputs
        seen.append(ts.frame)
        assert nlist.shape == (n, NN, 4) and positions.shape == (n, 4) and box.shape == (3, 3)
        np.testing.assert_allclose(box.cpu().numpy()[1], L, rtol=1e-6)
        assert float(box[2].abs().sum()) < 1e-4                     # orthorhombic: no tilt
        assert set(np.unique(positions[:, 3].cpu().numpy())) <= {0.0, 1.0}
        # per-frame list == brute force on this frame (index in the last column, utils.compute_nlist semantics)
        x = frames[ts.frame].astype(np.float64)
        d = x[None, :, :] - x[:, None, :]
        d -= np.round(d / L) * L
        r = np.sqrt((d ** 2).sum(-1))
        want = [set(np.nonzero((r[i] <= r_cut) & (r[i] >= 5e-4))[0]) for i in range(n)]
        got_nl = nlist.cpu().numpy()
        for i in range(0, n, 7):
            valid = np.abs(got_nl[i, :, :3]).sum(-1) > 0
            got = set(got_nl[i, valid, 3].astype(int))
            borderline = {j for j in want[i] ^ got if abs(r[i, j] - r_cut) < 1e-4}
            assert (want[i] ^ got) <= borderline
        out = model(inputs)
        assert float(out[0].abs().sum()) != 0.0, "Forces not be computed correctly"
    assert seen == [0, 1, 2, 3]
    # period and selection
    picked = [ts.frame for _, ts in htf.iter_from_trajectory(NN, u, r_cut=r_cut, period=2)]
    assert picked == [0, 2]
    nsel = int((types == "B").sum())
    for inputs, ts in htf.iter_from_trajectory(32, u, selection="type B", r_cut=1.0, period=1):
        assert inputs[1].shape == (nsel, 4) and inputs[0].shape == (nsel, 32, 4)
    # static_nlist reproduces the reference's quirk: the list of the first frame is reused (htf/utils.py:717-721)
    lists = [inp[0] for inp, _ in htf.iter_from_trajectory(NN, u, r_cut=r_cut, static_nlist=True)]
    assert all(l is lists[0] for l in lists)


def test_fused_tfcompute_paths_match_model_call():
    """Built-in LJ models under tfcompute take the one-call library path (fused_rows): same forces / virial as calling
    the model on a separately built neighbor tensor, for whole-system and row-batched updates; pinned host mirrors
    receive the same numbers."""
    import htf
    from htf import synthetic
    pos, lo, hi = synthetic.lattice_fluid((12, 12, 12), 0.7, seed=9)
    n, K, r_cut = pos.shape[0], 64, 2.5
    ref_sys = htf.sim.System(pos, lo, hi)
    ctx = htf.HtfContext(n, K, r_cut)
    ctx.set_box(lo, hi)
    nl = ctx.build_nlist(ref_sys.positions)
    fe_ref, v6_ref = ctx.lj_forces(nl, virial=True, virial_components=6)
    for batch in (None, 500):
        system = htf.sim.System(pos, lo, hi)
        model = htf.models.LJVirialModel(K, virial=True)
        tfc = htf.tfcompute(model)
        tfc.attach(htf.sim.nlist_cell(system), r_cut=r_cut, batch_size=batch)
        h_f = torch.empty((n, 4)).pin_memory()
        h_v = torch.empty((n, 6)).pin_memory()
        tfc.set_host_outputs(h_f, h_v)
        for t in range(3):
            f = tfc.compute_forces(t)
        tfc.host_sync()
        torch.cuda.synchronize()
        assert torch.equal(f, fe_ref) and torch.equal(tfc.virial6(), v6_ref)
        assert torch.equal(h_f, fe_ref.cpu()) and torch.equal(h_v, v6_ref.cpu())
        v9 = tfc.get_virial_array()
        assert v9.shape == (n, 9) and np.array_equal(v9[:, [0, 1, 2, 4, 5, 8]], v6_ref.cpu().numpy().astype(np.float64))
        assert np.array_equal(v9[:, 3], v9[:, 1]) and np.array_equal(v9[:, 7], v9[:, 5])
        np.testing.assert_allclose(tfc.get_log_value(), float(fe_ref[:, 3].double().sum()), rtol=1e-6)
    # fused=False on the model keeps the generic path (model call on the built tensor)
    system = htf.sim.System(pos, lo, hi)
    model = htf.models.LJModel(K)
    model.fused = False
    tfc = htf.tfcompute(model)
    tfc.attach(htf.sim.nlist_cell(system), r_cut=r_cut)
    f = tfc.compute_forces(0)
    assert torch.equal(f, ctx.lj_forces(nl))


def test_models_without_output_forces_can_differentiate():
    """Round-1 advisor finding: a model built with output_forces=False that calls compute_nlist_forces at inference
    (force-matching models evaluated with train=False, utils.compute_pairwise) must work -- tf.gradients always does
    in the reference (htf/simmodel.py:526-555)."""
    import htf

    class Probe(htf.SimModel):
        def setup(self):
            self.scale = torch.nn.Parameter(torch.tensor(1.0))

        def compute(self, nlist, positions, box):
            rinv = htf.nlist_rinv(nlist)
            energy = self.scale * rinv.sum(dim=1)
            return htf.compute_nlist_forces(nlist, energy), htf.compute_positions_forces(positions, (positions[:, :3] ** 2).sum())

    model = Probe(8, output_forces=False).cuda()
    nl = torch.zeros((4, 8, 4), device="cuda")
    nl[:, 0, 0] = 1.5
    pos = torch.ones((4, 4), device="cuda")
    box = torch.tensor([[0.0, 0, 0], [10, 10, 10], [0, 0, 0]], device="cuda")
    f, fp = model([nl, pos, box], False)
    assert f.shape == (4, 4) and float(f[:, 0].abs().sum()) > 0 and torch.allclose(fp[:, :3], -2 * pos[:, :3])
    out = htf.compute_pairwise(model, np.linspace(0.5, 2.0, 4))
    assert out[0].shape[0] == 4 and np.isfinite(out[0]).all()


def test_unstuff4_converts_type_bits():
    """htf/TFArrayComm.cu:9-28: the w component of HOOMD's Scalar4 holds the type as int bits."""
    import htf
    rng = np.random.default_rng(0)
    n = 1000
    xyz = rng.standard_normal((n, 3)).astype(np.float32)
    types = rng.integers(0, 7, n).astype(np.int32)
    stuffed = np.concatenate([xyz, types.view(np.float32)[:, None]], axis=1)
    ctx = htf.HtfContext(n, 8, 1.0)
    d = torch.from_numpy(stuffed).cuda()
    out = ctx.unstuff4(d)
    got = out.cpu().numpy()
    assert np.array_equal(got[:, :3], xyz) and np.array_equal(got[:, 3], types.astype(np.float32))
    ctx.unstuff4(d, out=d)                                   # in place
    assert torch.equal(d, out)
