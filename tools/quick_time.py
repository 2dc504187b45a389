"""Dev helper: per-phase device timings (CUDA events) of the path on a synthetic config."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hoomd-tf_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
import htf
from htf import synthetic

def timeit(fn, n=20, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(np.min(ts))

name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
pos, lo, hi, r_cut, K = synthetic.config(name)
n = pos.shape[0]
ctx = htf.HtfContext(n, K, r_cut); ctx.set_box(lo, hi)
print(name, "N", n, "K", K, "grid", ctx.cell_grid())
dpos = torch.from_numpy(pos).cuda()
nl = torch.empty((n, K, 4), device="cuda"); fe = torch.empty((n, 4), device="cuda"); vir = torch.empty((n, 6), device="cuda")
bins = torch.zeros(102, dtype=torch.int64, device="cuda")
ctx.bin_particles(dpos)
t_bin = timeit(lambda: ctx.bin_particles(dpos))
cnt = torch.empty((n,), dtype=torch.int32, device="cuda")
t_build_nc = timeit(lambda: ctx.build_nlist(dpos, out=nl, rebin=False))
t_build = timeit(lambda: ctx.build_nlist(dpos, out=nl, rebin=False, count_out=cnt))
t_lj_full = timeit(lambda: ctx.lj_forces(nl, virial=True, out=fe, virial_out=vir))
t_lj = timeit(lambda: ctx.lj_forces(nl, virial=True, out=fe, virial_out=vir, counts=cnt))
t_ljnv = timeit(lambda: ctx.lj_forces(nl, virial=False, out=fe))
t_rdf = timeit(lambda: ctx.rdf_hist(nl, (0, r_cut), 100, bins=bins))
t_step = timeit(lambda: ctx.lj_step(dpos, nlist_out=nl, force_out=fe, virial_out=vir))
t_step_rdf = timeit(lambda: ctx.lj_step(dpos, nlist_out=nl, force_out=fe, virial_out=vir, bins=bins, r_range=(0, r_cut), nbins=100))
print("overflow", ctx.overflow())
gb = lambda b, ms: b / ms / 1e6
print("bin      med %.3f ms min %.3f" % t_bin)
print("build    med %.3f ms min %.3f  -> %.0f GB/s" % (t_build_nc + (gb(n * (16 * K + 16), t_build_nc[0]),)))
print("build+cnt med %.3f ms min %.3f" % t_build)
print("lj+vir(all slots) med %.3f ms min %.3f  -> %.0f GB/s" % (t_lj_full + (gb(n * (16 * K + 40), t_lj_full[0]),)))
print("lj+vir   med %.3f ms min %.3f  -> %.0f GB/s (algorithmic bytes)" % (t_lj + (gb(n * (16 * K + 40), t_lj[0]),)))
print("lj       med %.3f ms min %.3f" % t_ljnv)
print("rdf      med %.3f ms min %.3f" % t_rdf)
print("step     med %.3f ms min %.3f  -> %.0f GB/s, %.3e particle-steps/s" % (t_step + (gb(n * (32 * K + 56), t_step[0]), n / t_step[0] * 1e3)))
print("step+rdf med %.3f ms min %.3f" % t_step_rdf)

# ---- the same step replayed from a CUDA graph (launch-bound small systems) ----
try:
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            ctx.lj_step(dpos, nlist_out=nl, force_out=fe, virial_out=vir, bins=bins, r_range=(0, r_cut), nbins=100)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        ctx.lj_step(dpos, nlist_out=nl, force_out=fe, virial_out=vir, bins=bins, r_range=(0, r_cut), nbins=100)
    t_graph = timeit(lambda: g.replay())
    print("step+rdf (CUDA graph replay) med %.3f ms min %.3f  -> %.3e particle-steps/s" % (t_graph + (n / t_graph[0] * 1e3,)))
except Exception as ex:
    print("graph capture failed:", repr(ex)[:300])
