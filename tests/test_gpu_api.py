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
