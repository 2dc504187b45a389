"""Dev helper: what the graph boundaries and the event records between the phases cost per step (1 M x 64, one GPU):
(a) three graphs with an event record after each (what bench.py's timed region does), (b) three graphs, no events,
(c) ONE graph for the whole step, (d) one graph + the five event records around it."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hoomd-tf_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
import htf
from htf import synthetic

pos, lo, hi, r_cut, K = synthetic.config("cfg3")
n = pos.shape[0]
ctx = htf.HtfContext(n, K, r_cut); ctx.set_box(lo, hi)
d = torch.from_numpy(pos).cuda()
nl = torch.empty((n, K, 4), device="cuda"); fe = torch.empty((n, 4), device="cuda"); vir = torch.empty((n, 6), device="cuda")
cnt = torch.empty((n,), dtype=torch.int32, device="cuda")
f_bin = lambda: ctx.bin_particles(d)
f_build = lambda: ctx.build_nlist(d, 0, n, out=nl, rebin=False, count_out=cnt)
f_force = lambda: ctx.lj_forces(nl, virial=True, virial_components=6, out=fe, virial_out=vir, counts=cnt)
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(3): f_bin(); f_build(); f_force()
torch.cuda.synchronize()
def cap(*fns):
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        for f in fns: f()
    return g
g3 = [cap(f_bin), cap(f_build), cap(f_force)]
g1 = cap(f_bin, f_build, f_force)
ev = lambda: torch.cuda.Event(enable_timing=True)
def run(step, reps=300):
    for _ in range(10): step(None)
    torch.cuda.synchronize()
    marks = [[ev() for _ in range(5)] for _ in range(reps)]
    e0, e1 = ev(), ev()
    e0.record()
    for i in range(reps): step(marks[i])
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
def a(m):
    if m: m[0].record(); m[1].record()
    g3[0].replay()
    if m: m[2].record()
    g3[1].replay()
    if m: m[3].record()
    g3[2].replay()
    if m: m[4].record()
def b(m):
    g3[0].replay(); g3[1].replay(); g3[2].replay()
def c(m):
    g1.replay()
def dd(m):
    if m: m[0].record(); m[1].record(); m[2].record(); m[3].record()
    g1.replay()
    if m: m[4].record()
for name, f in (("a: 3 graphs + events", a), ("b: 3 graphs", b), ("c: 1 graph", c), ("d: 1 graph + 5 events", dd), ("a again", a)):
    print("%-24s %.1f us/step" % (name, run(f)))
