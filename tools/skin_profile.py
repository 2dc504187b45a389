"""Dev helper for ncu: the buffered-list path on a synthetic config."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hoomd-tf_b200")); sys.path.insert(0, ROOT)
import torch, htf
from htf import synthetic
pos, lo, hi, r_cut, K = synthetic.config(sys.argv[1] if len(sys.argv) > 1 else "cfg3")
n = pos.shape[0]
ctx = htf.HtfContext(n, K, r_cut); ctx.set_box(lo, hi); ctx.skin_configure(0.4)
d = torch.from_numpy(pos).cuda(); nl = torch.empty((n, K, 4), device="cuda")
ctx.skin_rebuild(d)
for _ in range(3): ctx.skin_nlist(d, out=nl)
torch.cuda.synchronize(); print("status", ctx.skin_status())
