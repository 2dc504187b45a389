"""Randomised parity of the neighbor build against the oracle (GPU).

Seeded random boxes (non-cubic, from one cell to a dozen cells per dimension, so every staging / minimum-image path of
the tile kernel and the per-cell kernel is hit), random uniform gases with clusters (cells of very different
populations, some far above 32 rows), random cutoffs, K values and row shards.  Bars as everywhere: neighbor
(index, type) sets and (dx, dy, dz, type) values bit-exact per row after sorting by index, counts exact."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from test_gpu_parity import _ctx, gpu_nlist, sort_rows      # noqa: E402


def _random_case(seed):
    rng = np.random.default_rng(1000 + seed)
    r_cut = float(rng.choice([1.5, 2.0, 2.5, 3.0]))
    ncell = rng.integers(1, 13, size=3)                       # cells per dimension the box is sized for
    if seed % 5 == 0:
        ncell[2] = 1                                          # quasi 2-D
    L = (ncell * r_cut * rng.uniform(1.001, 1.6, size=3)).astype(np.float64)
    if seed % 7 == 3:
        L *= 0.9                                              # boxes a little under one cell: r_cut > L/2 images
    lo = -0.5 * L if seed % 3 else np.zeros(3)                # centred boxes and boxes starting at the origin
    hi = lo + L
    rho = float(rng.choice([0.05, 0.3, 0.7, 1.1]))
    n = int(np.clip(rho * np.prod(L), 20, 2400 if seed % 4 == 1 else 6000))
    xyz = rng.uniform(0.0, 1.0, size=(n, 3)) * L + lo
    if seed % 4 == 1:                                         # a dense blob: some cells far above the average
        k = n // 4
        c = rng.uniform(0.2, 0.8, size=3) * L + lo
        xyz[:k] = c + rng.normal(0.0, 0.35 * r_cut, size=(k, 3))
        xyz[:k] = np.mod(xyz[:k] - lo, L) + lo
    pos = np.zeros((n, 4), dtype=np.float32)
    pos[:, :3] = xyz.astype(np.float32)
    # fp32 rounding may land a coordinate exactly on hi: keep everything inside [lo, hi)
    lo32, hi32 = lo.astype(np.float32), hi.astype(np.float32)
    pos[:, :3] = np.minimum(np.maximum(pos[:, :3], lo32), np.nextafter(hi32, lo32))
    pos[:, 3] = rng.integers(0, 3, n)
    K = int(rng.choice([8, 32, 64, 96, 160]))
    return pos, lo32, hi32, r_cut, K


@pytest.mark.parametrize("seed", range(28))
def test_random_boxes_match_the_oracle(oracle_mod, seed):
    pos, lo, hi, r_cut, K = _random_case(seed)
    n = pos.shape[0]
    ctx = _ctx(n, K, r_cut, lo, hi)
    a, b = (0, n) if seed % 2 == 0 else (n // 5, n - n // 7)           # whole system / a row shard
    nl_g, idx_g, cnt_g = gpu_nlist(ctx, pos, a, b)
    nl_o, idx_o, cnt_o = oracle_mod.nlist(pos, lo, hi, r_cut, 1024, a, b, cells=False)   # every neighbor, brute force
    assert np.array_equal(cnt_g, cnt_o), "neighbor counts differ (grid %s)" % (ctx.cell_grid(),)
    ok = cnt_o <= K
    nls, ids = sort_rows(nl_g.cpu().numpy(), idx_g)
    want_i = np.full((b - a, K), -1, dtype=idx_o.dtype)
    want_v = np.zeros((b - a, K, 4), dtype=np.float32)
    m = min(K, idx_o.shape[1])
    nlo, ido = sort_rows(nl_o, idx_o)
    want_i[:, :m], want_v[:, :m] = ido[:, :m], nlo[:, :m]
    assert np.array_equal(ids[ok], want_i[ok]), "neighbor index sets differ"
    assert np.array_equal(nls[ok].view(np.uint32), want_v[ok].view(np.uint32)), "(dx,dy,dz,type) not bit-exact"
    for r in np.where(~ok)[0][:20]:                                     # overflowed rows: K distinct genuine neighbors
        assert len(set(idx_g[r])) == K and set(idx_g[r]).issubset(set(idx_o[r][idx_o[r] >= 0]))
    # the force pass with and without the builder's counts on the same tensor
    fe_a, v_a = ctx.lj_forces(nl_g, virial=True)
    fe_b, v_b = ctx.lj_forces(nl_g, virial=True, counts=torch.from_numpy(cnt_g).cuda())
    torch.cuda.synchronize()
    assert torch.equal(fe_a, fe_b) and torch.equal(v_a, v_b)
