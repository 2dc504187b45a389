"""Dev helper: device time of the halo packing (both slab faces of a 1 M-row slab) alone, replayed from a CUDA graph."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hoomd-tf_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
import htf
from htf import synthetic

pos, lo, hi, r_cut, K = synthetic.config(sys.argv[1] if len(sys.argv) > 1 else "cfg3")
n = pos.shape[0]
ctx = htf.HtfContext(n, K, r_cut); ctx.set_box(lo, hi)
lo_face, hi_face, width, cap = htf.parallel.slab_plan(pos, 2, r_cut)
d = torch.from_numpy(pos).cuda()
a = torch.empty((cap, 4), device="cuda"); b = torch.empty((cap, 4), device="cuda")
cnt = torch.zeros(2, dtype=torch.int32, device="cuda")
fn = lambda: ctx.pack_halo_pair(d, 2, lo_face + width, hi_face - width, a, b, counts=cnt)
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(3): fn()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g, stream=s):
    fn()
reps = 200
for _ in range(5): g.replay()
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps): g.replay()
e1.record(); torch.cuda.synchronize()
print("pack_halo_pair n=%d cap=%d counts=%s: %.2f us per call (graph replay, back to back)" % (n, cap, cnt.tolist(), e0.elapsed_time(e1) / reps * 1e3))
