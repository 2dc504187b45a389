"""Seeded synthetic systems (host side, numpy): the inputs of the parity tests and of bench.py.

Generators follow SURVEY.md section 8(d): simple-cubic sites with a uniform jitter of
+-0.15 a, box centred on the origin, positions wrapped into [-L/2, L/2), type stored as a
float value in column 3.  The lattice helpers reproduce the systems of the reference's own
tests (hoomd.lattice.sq / bcc via hoomd.init.create_lattice,
/root/reference htf/test-py/test_tensorflow.py:335-349, htf/test-py/test_utils.py:408-409).
"""
import numpy as np

# BASELINE.json configs -> (sites nx,ny,nz, density, r_cut, K, seed)
CONFIGS = {
    "cfg1": dict(sites=(4, 8, 8), rho=0.10, r_cut=3.0, K=32, seed=1),
    "cfg2": dict(sites=(32, 32, 64), rho=0.70, r_cut=2.5, K=64, seed=2),
    "cfg3": dict(sites=(64, 128, 128), rho=0.70, r_cut=2.5, K=64, seed=3),
    "cfg4": dict(sites=(64, 64, 64), rho=0.70, r_cut=2.5, K=64, seed=4),
    "cfg5": dict(sites=(128, 128, 256), rho=0.8442, r_cut=2.8, K=96, seed=5),
}


def lattice_fluid(sites, rho, seed, jitter=0.15, two_types_p=None, dtype=np.float32):
    """Jittered simple-cubic fluid.  Returns (pos[N,4] f32, lo[3], hi[3]).

    site (ix,iy,iz) -> ((i+0.5) a + U(-jitter a, jitter a)) mod L per axis in fp64, shifted by
    -L/2, cast to fp32.  Particle order is z-slowest / x-fastest, i.e. spatially coherent.
    """
    nx, ny, nz = sites
    a = (1.0 / rho) ** (1.0 / 3.0)
    L = np.array([nx * a, ny * a, nz * a], dtype=np.float64)
    rng = np.random.default_rng(seed)
    iz, iy, ix = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    base = np.stack([ix.ravel(), iy.ravel(), iz.ravel()], axis=1).astype(np.float64)
    n = base.shape[0]
    xyz = (base + 0.5) * a + rng.uniform(-jitter * a, jitter * a, size=(n, 3))
    xyz = np.mod(xyz, L) - 0.5 * L
    pos = np.zeros((n, 4), dtype=dtype)
    pos[:, :3] = xyz.astype(dtype)
    # fp32 rounding may put a coordinate exactly on +L/2: fold it back into [-L/2, L/2)
    lo = (-0.5 * L).astype(dtype)
    hi = (0.5 * L).astype(dtype)
    for ax in range(3):
        over = pos[:, ax] >= hi[ax]
        pos[over, ax] -= (hi[ax] - lo[ax])
    if two_types_p is not None:
        pos[:, 3] = (rng.random(n) < two_types_p).astype(dtype)
    return pos, lo, hi


def config(name, two_types_p=None):
    c = CONFIGS[name]
    pos, lo, hi = lattice_fluid(c["sites"], c["rho"], c["seed"], two_types_p=two_types_p)
    return pos, lo, hi, c["r_cut"], c["K"]


def square_lattice(n, a, lz=1.0):
    """hoomd.init.create_lattice(hoomd.lattice.sq(a), n=[n,n]): 2-D, box centred, thin in z."""
    nx, ny = (n, n) if np.isscalar(n) else n
    L = np.array([nx * a, ny * a, lz], dtype=np.float64)
    xs = (np.arange(nx) + 0.5) * a - 0.5 * L[0]
    ys = (np.arange(ny) + 0.5) * a - 0.5 * L[1]
    pos = np.zeros((nx * ny, 4), dtype=np.float32)
    pos[:, 0] = np.tile(xs, ny)
    pos[:, 1] = np.repeat(ys, nx)
    return pos, (-0.5 * L).astype(np.float32), (0.5 * L).astype(np.float32)


def bcc_lattice(n, a):
    """hoomd.lattice.bcc(a) replicated n^3 (htf/test-py/test_utils.py:408-409)."""
    L = np.array([n * a] * 3, dtype=np.float64)
    g = np.arange(n)
    iz, iy, ix = np.meshgrid(g, g, g, indexing="ij")
    corner = np.stack([ix.ravel(), iy.ravel(), iz.ravel()], axis=1).astype(np.float64) * a
    xyz = np.concatenate([corner + 0.25 * a, corner + 0.75 * a], axis=0) - 0.5 * L
    pos = np.zeros((xyz.shape[0], 4), dtype=np.float32)
    pos[:, :3] = xyz.astype(np.float32)
    return pos, (-0.5 * L).astype(np.float32), (0.5 * L).astype(np.float32)


def typed_chains():
    """Two-type 10-bead chains x27 of the typed-RDF test (htf/test-py/test_tensorflow.py:457-475)."""
    base = np.array([[-4.5 + i, 0.0, 0.0] for i in range(10)], dtype=np.float64)
    types = np.array([0] * 7 + [1] * 3, dtype=np.float64)
    out, tt = [], []
    for kz in range(3):
        for ky in range(3):
            for kx in range(3):
                out.append(base + np.array([10.0 * kx, 10.0 * ky, 10.0 * kz]))
                tt.append(types)
    xyz = np.concatenate(out) - np.array([10.0, 10.0, 10.0])
    L = np.array([30.0, 30.0, 30.0])
    xyz = np.mod(xyz + 0.5 * L, L) - 0.5 * L
    pos = np.zeros((xyz.shape[0], 4), dtype=np.float32)
    pos[:, :3] = xyz.astype(np.float32)
    pos[:, 3] = np.concatenate(tt).astype(np.float32)
    return pos, (-0.5 * L).astype(np.float32), (0.5 * L).astype(np.float32)


def perturb(pos, lo, hi, sigma, seed):
    """Gaussian displacement + wrap back into the box (a stand-in for a few MD steps)."""
    rng = np.random.default_rng(seed)
    L = (np.asarray(hi, dtype=np.float64) - np.asarray(lo, dtype=np.float64))
    xyz = pos[:, :3].astype(np.float64) + rng.normal(0.0, sigma, size=(pos.shape[0], 3))
    xyz = np.mod(xyz - np.asarray(lo, dtype=np.float64), L) + np.asarray(lo, dtype=np.float64)
    out = pos.copy()
    out[:, :3] = xyz.astype(np.float32)
    for ax in range(3):
        over = out[:, ax] >= hi[ax]
        out[over, ax] -= (hi[ax] - lo[ax])
    return out
