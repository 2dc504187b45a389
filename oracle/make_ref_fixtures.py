#!/usr/bin/env python
"""Pins the oracle to the REAL reference.  TEST INFRASTRUCTURE ONLY -- cannot run in the build container.

Neither HOOMD-blue 2.x nor TensorFlow is importable where this repository is built, so the golden fixtures under
tests/golden/ are oracle output (checked analytically).  This script closes that gap wherever the reference stack
exists (hoomd >= 2.6 with the hoomd.htf plugin, tensorflow >= 2.3 -- e.g. the reference's own CI image):

    python oracle/make_ref_fixtures.py            # writes tests/golden/<name>_ref.npz, then compares
    python oracle/make_ref_fixtures.py --compare  # only compare existing *_ref.npz with the oracle fixtures

For each golden system (tests/golden/*.npz: 5x5 square lattice, bcc 4^3, 216-particle fluid) it loads the SAME
positions into a HOOMD snapshot, attaches the reference's own models through the reference's own code path
(htf.tfcompute.attach -> TensorflowCompute::computeForces -> prepareNeighbors -> SimModel), and stores what the
reference produced in the schema of tests/golden/make_golden.py:

    nlist_sorted [N,K,4]   tfcompute.get_nlist_array() (htf/tensorflowcompute.py:378-383), rows sorted by value
    count [N]              non-zero slots per row
    force_energy [N,4]     tfcompute.get_forces_array()   (LJVirialModel, htf/test-py/build_examples.py:104-115)
    virial6 [N,6]          tfcompute.get_virial_array()[:, (0,1,2,4,5,8)]   (htf/TensorflowCompute.cc:294-299)
    rdf_hist [nbins+2]     tf.histogram_fixed_width(|d|, [0, r_cut], 102) over the nlist (htf/simmodel.py:657-662)

The comparison is the parity bar of this repository: neighbor tensor bit-exact as a per-row multiset (a DOUBLE
precision HOOMD build selects neighbors in fp64 and casts afterwards -- pairs within one fp32 ulp of r_cut may then
differ and are reported separately), RDF counts bit-exact, forces / energy / virial within 1e-5.
A clean run turns "parity unpinned" in oracle/htf_oracle.c and DESIGN.md into "pinned"; commit the *_ref.npz files.
"""
import glob
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")
NBINS = 100


def sort_rows_by_value(nl):
    u = np.ascontiguousarray(nl.astype(np.float32)).view(np.uint32).astype(np.uint64)
    k1 = (u[..., 0] << np.uint64(32)) | u[..., 1]
    k2 = (u[..., 2] << np.uint64(32)) | u[..., 3]
    valid = np.abs(nl[..., :3]).sum(-1) > 0
    k0 = (~valid).astype(np.uint64)                         # padded slots last
    order = np.lexsort((k2, k1, k0), axis=1)
    return np.take_along_axis(nl, order[:, :, None], axis=1)


def run_reference(g):
    """One golden system through the unmodified reference.  Imports hoomd / tensorflow lazily."""
    import hoomd
    import hoomd.md
    import hoomd.htf as htf
    import tensorflow as tf

    pos, lo, hi = g["pos"], g["lo"].astype(np.float64), g["hi"].astype(np.float64)
    r_cut, K = float(g["r_cut"]), int(g["K"])
    n = pos.shape[0]
    L = hi - lo
    centre = 0.5 * (hi + lo)

    class LJVirialRDF(htf.SimModel):
        # the body of build_examples.LJVirialModel (:104-115) plus the raw histogram of compute_rdf (simmodel.py:657-662)
        def compute(self, nlist, positions, box):
            rinv = htf.nlist_rinv(nlist)
            inv_r6 = rinv ** 6
            p_energy = 4.0 / 2.0 * (inv_r6 * inv_r6 - inv_r6)
            energy = tf.reduce_sum(input_tensor=p_energy, axis=1)
            forces, virial = htf.compute_nlist_forces(nlist, energy, virial=True)
            r = tf.norm(tensor=nlist[:, :, :3], axis=2)
            hist = tf.histogram_fixed_width(r, tf.cast([0.0, r_cut], tf.float32), NBINS + 2)
            return forces, virial, hist

    hoomd.context.initialize("--mode=cpu")
    two_d = L[2] <= 1.0 + 1e-6 and np.all(pos[:, 2] == 0.0)
    box = hoomd.data.boxdim(Lx=L[0], Ly=L[1], Lz=L[2], dimensions=2 if two_d else 3)
    ntypes = int(pos[:, 3].max()) + 1
    snap = hoomd.data.make_snapshot(N=n, box=box, particle_types=[chr(ord("A") + t) for t in range(ntypes)])
    snap.particles.position[:] = pos[:, :3].astype(np.float64) - centre          # HOOMD boxes are centred on the origin
    snap.particles.typeid[:] = pos[:, 3].astype(np.int32)
    snap.particles.velocity[:] = 0.0
    hoomd.init.read_snapshot(snap)
    hoomd.context.current.sorter.disable()                                       # keep the particle order of the fixture
    model = LJVirialRDF(K, virial=True)
    tfcompute = htf.tfcompute(model)
    nlist = hoomd.md.nlist.cell()
    hoomd.md.integrate.mode_standard(dt=1e-12)                                   # a step that moves nothing in fp32/fp64
    hoomd.md.integrate.nve(group=hoomd.group.all())
    tfcompute.attach(nlist, r_cut=r_cut, save_output_period=1)
    hoomd.run(1)
    nl = np.asarray(tfcompute.get_nlist_array(), dtype=np.float64).reshape(n, K, 4)
    fe = np.asarray(tfcompute.get_forces_array(), dtype=np.float64)[:n]
    v9 = np.asarray(tfcompute.get_virial_array(), dtype=np.float64).reshape(-1, 9)[:n]
    hist = np.asarray(tfcompute.outputs[0][-1]).astype(np.int64)
    return {"nlist_sorted": sort_rows_by_value(nl.astype(np.float32)),
            "count": (np.abs(nl[..., :3]).sum(-1) > 0).sum(1).astype(np.int32),
            "force_energy": fe.astype(np.float32), "virial6": v9[:, [0, 1, 2, 4, 5, 8]].astype(np.float32),
            "rdf_hist": hist, "hoomd_double": np.bool_(bool(getattr(tfcompute.cpp_force, "isDoublePrecision", lambda: True)())),
            "hoomd_version": np.str_(hoomd.__version__), "tf_version": np.str_(tf.__version__)}


def rel(got, want):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    scale = np.sqrt(np.mean(want ** 2)) + 1e-30
    return float((np.abs(got - want) / np.maximum(np.abs(want), scale)).max())


def compare(name, g, r):
    """Reference output r against the oracle-made fixture g.  Returns True when the parity bar holds."""
    mine = sort_rows_by_value(g["nlist_sorted"])
    ok_cnt = bool(np.array_equal(g["count"], r["count"]))
    same = mine.view(np.uint32) == r["nlist_sorted"].view(np.uint32)
    rows_bad = int((~same.all(axis=(1, 2))).sum())
    ok_hist = bool(np.array_equal(g["rdf_hist"], r["rdf_hist"]))
    e_f, e_v = rel(g["force_energy"], r["force_energy"]), rel(g["virial6"], r["virial6"])
    print("%-16s counts %s | nlist rows differing %d of %d | rdf bins %s | force+energy %.2e | virial %.2e  (hoomd %s%s, tf %s)"
          % (name, "equal" if ok_cnt else "DIFFER", rows_bad, mine.shape[0], "equal" if ok_hist else "DIFFER", e_f, e_v,
             r["hoomd_version"], " double" if bool(r["hoomd_double"]) else " single", r["tf_version"]))
    if rows_bad and bool(r["hoomd_double"]):
        # value-level differences of a double-precision HOOMD build: positions differences are formed in fp64 and cast
        d = np.abs(mine.astype(np.float64) - r["nlist_sorted"].astype(np.float64)).max()
        print("   max |delta| of the neighbor tensor entries: %.3e (fp64 selection + cast vs fp32 arithmetic)" % d)
    return ok_cnt and ok_hist and e_f <= 1e-5 and e_v <= 1e-5 and (rows_bad == 0 or bool(r["hoomd_double"]))


def main():
    only_compare = "--compare" in sys.argv
    files = sorted(f for f in glob.glob(os.path.join(GOLDEN, "*.npz")) if not f.endswith("_ref.npz"))
    all_ok = True
    for f in files:
        name = os.path.basename(f)[:-4]
        g = dict(np.load(f))
        out = os.path.join(GOLDEN, name + "_ref.npz")
        if not only_compare:
            r = run_reference(g)
            np.savez_compressed(out, **r)
        if not os.path.exists(out):
            print("%-16s no reference fixture yet (%s)" % (name, out))
            all_ok = False
            continue
        all_ok &= compare(name, g, dict(np.load(out)))
    print("PINNED" if all_ok else "NOT PINNED")
    return 0 if all_ok else 1


if __name__ == "__main__":
    sys.exit(main())
