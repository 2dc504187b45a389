"""Dev helper: time the halo selection (count + scan + scatter of the two slab faces) on one GPU, in a CUDA graph."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hoomd-tf_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch, htf
from htf import synthetic
pos, lo, hi, r_cut, K = synthetic.config("cfg3")
n = pos.shape[0]
ctx = htf.HtfContext(n, K, r_cut); ctx.set_box(lo, hi)
d = torch.from_numpy(pos).cuda()
cap = 131072
a = torch.empty((cap, 4), device="cuda"); b = torch.empty((cap, 4), device="cuda")
cnt = torch.zeros(2, dtype=torch.int32, device="cuda")
zlo, zhi = float(lo[2]) + r_cut, float(hi[2]) - r_cut
def fn(): ctx.pack_halo_pair(d, 2, zlo, zhi, a, b, counts=cnt)
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(3): fn()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g, stream=s):
    fn()
for _ in range(5): g.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(100): g.replay()
e1.record(); torch.cuda.synchronize()
print("pack_halo_pair (3 launches, graph): %.2f us; counts %s" % (e0.elapsed_time(e1) * 10.0, cnt.tolist()))
