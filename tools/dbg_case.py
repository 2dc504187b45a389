import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hoomd-tf_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch, htf
from htf import synthetic
for n, a, K, rc in [(3, 4.0, 8, 5.0), (5, 3.0, 32, 5.0), (16, 2.0, 64, 3.0)]:
    pos, lo, hi = synthetic.square_lattice(n, a)
    ctx = htf.HtfContext(pos.shape[0], K, rc); ctx.set_box(lo, hi)
    print(n, a, K, "grid", ctx.cell_grid(), flush=True)
    d = torch.from_numpy(pos).cuda()
    try:
        ctx.bin_particles(d); torch.cuda.synchronize(); print(" bin ok", flush=True)
        nl = ctx.build_nlist(d, rebin=False); torch.cuda.synchronize(); print(" build ok", float(nl.abs().sum()), flush=True)
    except Exception as e:
        print(" ERR", e, flush=True)
