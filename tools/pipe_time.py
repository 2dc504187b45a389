"""Dev helper: the pipelined step (htf_lj_step) against the back-to-back sequence, per slab count.

  python tools/pipe_time.py [cfg3|cfg5] [--shuffle]
Prints per-phase times of the serial sequence, then the whole step (CUDA-graph replay) for several slab counts,
and checks that the pipelined outputs are bit-identical to the serial ones.
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hoomd-tf_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
import htf
from htf import synthetic


def timeit(fn, n=30, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n


name = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else "cfg3"
pos, lo, hi, r_cut, K = synthetic.config(name)
if "--shuffle" in sys.argv:
    pos = pos[np.random.default_rng(0).permutation(pos.shape[0])]
n = pos.shape[0]
ctx = htf.HtfContext(n, K, r_cut); ctx.set_box(lo, hi)
print(name, "N", n, "K", K, "grid", ctx.cell_grid(), "shuffled" if "--shuffle" in sys.argv else "lattice order")
dpos = torch.from_numpy(pos).cuda()
nl = torch.empty((n, K, 4), device="cuda"); fe = torch.empty((n, 4), device="cuda"); vir = torch.empty((n, 6), device="cuda")
ctx.bin_particles(dpos)
t_bin = timeit(lambda: ctx.bin_particles(dpos))
t_build = timeit(lambda: ctx.build_nlist(dpos, out=nl, rebin=False))
t_lj = timeit(lambda: ctx.lj_forces(nl, virial=True, out=fe, virial_out=vir))
print("serial phases: bin %.3f  build %.3f (%.0f GB/s)  lj+vir %.3f (%.0f GB/s)  sum %.3f ms"
      % (t_bin, t_build, n * (16 * K + 16) / t_build / 1e6, t_lj, n * (16 * K + 40) / t_lj / 1e6, t_bin + t_build + t_lj))
fe0, vir0, nl0 = fe.clone(), vir.clone(), nl.clone()
assert ctx.overflow() == 0

side = torch.cuda.Stream()
combos = [(0, 2, 1)] + [(sl, bps, bs) for bs in (1, 2) for bps in (1, 2, 3) for sl in (2, 4, 8)]
if len(sys.argv) > 2 and sys.argv[-1].startswith("combos="):
    combos = [tuple(int(x) for x in c.split(",")) for c in sys.argv[-1][7:].split(";")]
for slabs, bps, bstreams in combos:
    os.environ["HTF_PIPE_PASS_BPS"] = str(bps); os.environ["HTF_PIPE_BUILD_STREAMS"] = str(bstreams)
    ctx = htf.HtfContext(n, K, r_cut); ctx.set_box(lo, hi)
    ctx.set_pipeline(slabs)
    fe.zero_(); vir.zero_(); nl.zero_()
    torch.cuda.synchronize()
    with torch.cuda.stream(side):
        for _ in range(3):
            ctx.lj_step(dpos, nlist_out=nl, force_out=fe, virial_out=vir)
    torch.cuda.synchronize()
    same = bool(torch.equal(fe, fe0) and torch.equal(vir, vir0) and torch.equal(nl, nl0))
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        ctx.lj_step(dpos, nlist_out=nl, force_out=fe, virial_out=vir)
    t_graph = timeit(lambda: g.replay(), n=100)
    print("slabs %2d pass-blocks/SM %d build-streams %d: graph %.3f ms  -> %.3e particle-steps/s, path %.0f GB/s   identical: %s"
          % (slabs, bps, bstreams, t_graph, n / t_graph * 1e3, n * (32 * K + 56) / t_graph / 1e6, same))
    del g, ctx
