"""Dev helper: time the fused LJ + coordination-CV (+RDF) pass on a synthetic config."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hoomd-tf_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch, htf
from htf import synthetic
name = sys.argv[1] if len(sys.argv) > 1 else "cfg5"
pos, lo, hi, r_cut, K = synthetic.config(name)
n = pos.shape[0]
ctx = htf.HtfContext(n, K, r_cut); ctx.set_box(lo, hi)
d = torch.from_numpy(pos).cuda()
nl, cnt = ctx.build_nlist(d, want_count=True)
print('mean count %.1f of K=%d' % (float(cnt.float().mean()), K))
fe = torch.empty((n, 4), device="cuda"); vir = torch.empty((n, 6), device="cuda")
cv_row = torch.empty((n, 4), device="cuda"); cv_sum = torch.zeros(1, dtype=torch.float64, device="cuda")
bins = torch.zeros(102, dtype=torch.int64, device="cuda")
def timeit(fn, reps=10):
    for _ in range(3): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
gb = n * K * 16 / 1e6
for label, fn in (("lj", lambda: ctx.lj_forces(nl, out=fe)),
                  ("lj+cv", lambda: ctx.lj_cv_forces(nl, 1.3, cv_row, cv_sum)),
                  ("lj+cv+rdf", lambda: ctx.lj_cv_forces(nl, 1.3, cv_row, cv_sum, bins=bins, r_range=(0.0, r_cut), nbins=100)),
                  ("lj+rdf", lambda: ctx.lj_step_forces_only(nl, fe, vir, bins, (0.0, r_cut), 100)),
                  ("rdf only", lambda: ctx.rdf_hist(nl, (0.0, r_cut), 100, bins=bins)),
                  ("lj+vir      [counts]", lambda: ctx.lj_forces(nl, virial=True, out=fe, virial_out=vir, counts=cnt)),
                  ("lj+cv       [counts]", lambda: ctx.lj_cv_forces(nl, 1.3, cv_row, cv_sum, counts=cnt)),
                  ("lj+cv+rdf   [counts]", lambda: ctx.lj_cv_forces(nl, 1.3, cv_row, cv_sum, bins=bins, r_range=(0.0, r_cut), nbins=100, counts=cnt)),
                  ("lj+rdf      [counts]", lambda: ctx.lj_step_forces_only(nl, fe, vir, bins, (0.0, r_cut), 100, counts=cnt))):
    ms = timeit(fn)
    print("%-22s %.3f ms  %.0f GB/s" % (label, ms, gb / ms))
