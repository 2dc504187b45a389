"""Dev helper for ncu: a few lj_step passes on a synthetic config (no timing here)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hoomd-tf_b200")); sys.path.insert(0, ROOT)
import torch, htf
from htf import synthetic
name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
rdf = len(sys.argv) > 3 and sys.argv[3] == "rdf"
pos, lo, hi, r_cut, K = synthetic.config(name)
n = pos.shape[0]
ctx = htf.HtfContext(n, K, r_cut); ctx.set_box(lo, hi)
dpos = torch.from_numpy(pos).cuda()
nl = torch.empty((n, K, 4), device="cuda"); fe = torch.empty((n, 4), device="cuda"); vir = torch.empty((n, 6), device="cuda")
bins = torch.zeros(102, dtype=torch.int64, device="cuda") if rdf else None
for _ in range(steps):
    ctx.lj_step(dpos, nlist_out=nl, force_out=fe, virial_out=vir, bins=bins, r_range=(0, r_cut), nbins=100)
torch.cuda.synchronize()
print("done", ctx.launches)
