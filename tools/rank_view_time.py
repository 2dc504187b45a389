"""Dev helper: what ONE rank of a 2-rank weak-scaling run computes, timed on one GPU without any exchange: the rank's
own 1 M rows + the two received faces (sentinel padded to the exchange capacity), region-of-interest binning, build of
the own rows, force pass.  Compared with the single-GPU step it separates the cost of the rank layout (more particles
binned, halo cell layers) from the cost of the exchange itself and of the ranks waiting for each other."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hoomd-tf_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
import htf
from htf import synthetic

world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
c = dict(synthetic.CONFIGS["cfg3"]); sites = list(c["sites"]); sites[2] *= world
pos, lo, hi = synthetic.lattice_fluid(tuple(sites), c["rho"], c["seed"])
r_cut, K = c["r_cut"], c["K"]
n = pos.shape[0]; per = n // world
rank = 0
a, b = rank * per, (rank + 1) * per

def graph_time(fn, reps=200):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s): fn()
    for _ in range(5): g.replay()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3

def run(label, d_all, rows, roi):
    ctx = htf.HtfContext(d_all.shape[0], K, r_cut); ctx.set_box(lo, hi)
    if roi is not None: ctx.set_roi(*roi)
    nl = torch.empty((rows, K, 4), device="cuda"); fe = torch.empty((rows, 4), device="cuda"); vir = torch.empty((rows, 6), device="cuda")
    cnt = torch.empty((rows,), dtype=torch.int32, device="cuda")
    ctx.bin_particles(d_all); ctx.build_nlist(d_all, 0, rows, out=nl, rebin=False, count_out=cnt)
    assert ctx.overflow() == 0
    t_bin = graph_time(lambda: ctx.bin_particles(d_all))
    t_build = graph_time(lambda: ctx.build_nlist(d_all, 0, rows, out=nl, rebin=False, count_out=cnt))
    t_force = graph_time(lambda: ctx.lj_forces(nl, virial=True, virial_components=6, out=fe, virial_out=vir, counts=cnt))
    print("%-34s entries %8d grid %s: bin %.1f build %.1f force %.1f us (sum %.1f)" % (label, d_all.shape[0], ctx.cell_grid(), t_bin, t_build, t_force, t_bin + t_build + t_force))

# single-GPU reference: the config itself
p1, lo1, hi1 = synthetic.lattice_fluid(tuple(c["sites"]), c["rho"], c["seed"])
lo_keep, hi_keep = lo, hi
lo, hi = lo1, hi1
run("single GPU (1 M, periodic)", torch.from_numpy(p1).cuda(), p1.shape[0], None)
lo, hi = lo_keep, hi_keep
# rank view
own = pos[a:b]
lo_face, hi_face, width, cap = htf.parallel.slab_plan(own, 2, r_cut)
nxt, prv = (rank + 1) % world, (rank - 1) % world
def face(r, low):
    q = pos[r * per:(r + 1) * per]; z = q[:, 2]
    f = q[z < z.min() + width] if low else q[z > z.max() - width]
    out = np.full((cap, 4), 1e30, dtype=np.float32); out[:, 3] = 0.0; out[:len(f)] = f
    return out
local = np.concatenate([own, face(nxt, True), face(prv, False)], axis=0)
roi = htf.parallel.roi_for_rows(own, lo, hi, r_cut)
run("rank 0 of %d (own + 2 faces, ROI)" % world, torch.from_numpy(local).cuda(), per, roi)
