"""Regenerates tests/golden/*.npz.

The reference stores no golden vectors for this path and cannot be imported here (no hoomd / tensorflow), so these
fixtures are produced by the CPU oracle on the reference's own known-answer systems (htf/test-py/test_tensorflow.py:
335-349 5x5 square lattice a=4.0 r_cut=5; test_utils.py:408-409 bcc 4x4x4 a=4.0) AFTER checking the oracle's LJ forces
against the analytic pair sum written out below in float64.  They pin (a) the oracle against silent changes and
(b) the CUDA path on the GPU box without needing anything but numpy there.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "hoomd-tf_b200"))
import oracle                      # noqa: E402
from htf import synthetic          # noqa: E402


def analytic_lj(pos, lo, hi, r_cut):
    """float64 minimum-image LJ forces (epsilon = sigma = 1), the reference tests' Python loop (test_tensorflow.py:20-35)."""
    L = np.asarray(hi, dtype=np.float64) - np.asarray(lo, dtype=np.float64)
    x = pos[:, :3].astype(np.float64)
    f = np.zeros_like(x)
    for i in range(len(x)):
        d = x - x[i]
        d -= np.round(d / L) * L
        r2 = (d * d).sum(1)
        m = (r2 > 0) & (r2 <= r_cut * r_cut)
        inv2 = 1.0 / r2[m]
        inv6 = inv2 ** 3
        f[i] = -((48.0 * inv6 * inv6 - 24.0 * inv6) * inv2)[:, None].__mul__(d[m]).sum(0)
    return f


def case(name, pos, lo, hi, r_cut, K):
    nl, idx, cnt = oracle.nlist(pos, lo, hi, r_cut, K)
    fe, v9, v6 = oracle.lj(nl)
    ref = analytic_lj(pos, lo, hi, r_cut)
    scale = max(np.abs(ref).max(), 1e-3)
    assert np.abs(fe[:, :3] - ref).max() <= 1e-4 * scale + 2e-5, name      # the 1e-7 / 3e-6 offsets of nlist_rinv included
    hist = oracle.rdf_hist(nl, (0.0, r_cut), 100)
    key = np.where(idx < 0, np.iinfo(np.int32).max, idx)
    order = np.argsort(key, axis=1, kind="stable")
    np.savez_compressed(os.path.join(HERE, name + ".npz"), pos=pos, lo=np.asarray(lo, np.float32), hi=np.asarray(hi, np.float32),
                        r_cut=np.float32(r_cut), K=np.int32(K), nlist_sorted=np.take_along_axis(nl, order[:, :, None], axis=1),
                        idx_sorted=np.take_along_axis(idx, order, axis=1), count=cnt, force_energy=fe, virial6=v6, rdf_hist=hist)
    print(name, "N", pos.shape[0], "K", K, "max count", cnt.max(), "max |F|", np.abs(fe[:, :3]).max())


if __name__ == "__main__":
    p, lo, hi = synthetic.square_lattice(5, 4.0)
    case("sq5x5_a4_rc5", synthetic.perturb(p, lo, hi, 0.05, seed=1), lo, hi, 5.0, 32)
    p, lo, hi = synthetic.bcc_lattice(4, 4.0)
    case("bcc4_a4_rc5", synthetic.perturb(p, lo, hi, 0.05, seed=2), lo, hi, 5.0, 32)
    p, lo, hi = synthetic.lattice_fluid((6, 6, 6), 0.7, seed=9)
    case("fluid216_rc25", p, lo, hi, 2.5, 64)
