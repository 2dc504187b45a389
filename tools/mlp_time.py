"""Time the tcgen05 pairwise-MLP kernel at the config-3 size (1M particles, K=64)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "hoomd-tf_b200"))
import torch, numpy as np
import htf
from htf import synthetic
from htf.context import HtfContext
n_side = int(sys.argv[1]) if len(sys.argv) > 1 else 64
pos, lo, hi = synthetic.lattice_fluid((n_side, n_side, n_side * 4), 0.7, seed=3) if n_side == 64 else synthetic.lattice_fluid((n_side,) * 3, 0.7, seed=3)
N, K, rc = pos.shape[0], 64, 2.5
ctx = HtfContext(N, K, rc); ctx.set_box(lo, hi)
nl = ctx.build_nlist(torch.from_numpy(pos).cuda())
m = htf.models.PairwiseMLPModel(K, r_cut=rc).cuda()
raw = m.raw_parameters(); packed = ctx.mlp_pack(raw)
out = torch.empty((N, 4), device="cuda")
for _ in range(3): ctx.mlp_forces(nl, packed, rc, out)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(10): ctx.mlp_forces(nl, packed, rc, out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
pairs = N * K
flops = pairs * 2.0 * (32 * 64 + 64 * 64 * 2 + 64 * 64 * 2 + 64 * 32)   # 3 fwd + 3 grad GEMMs
print("N=%d pairs=%.3g  %.3f ms  %.1f TFLOP/s (dense bf16)  %.2f Gpair/s  nlist read %.0f GB/s" % (N, pairs, ms, flops / ms / 1e9, pairs / ms / 1e6, pairs * 16 / ms / 1e6))
