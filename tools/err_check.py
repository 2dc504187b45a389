"""Dev helper: error of the GPU LJ pass vs the oracle (metric of tests/test_gpu_parity.py) on cfg2."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "hoomd-tf_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch, htf, oracle
from htf import synthetic
def err(got, want):
    got = np.asarray(got, np.float64); want = np.asarray(want, np.float64)
    scale = np.sqrt(np.mean(want ** 2)); e = np.abs(got - want) / np.maximum(np.abs(want), scale)
    return e.max(), scale
for name in ["cfg2"]:
    pos, lo, hi, rc, K = synthetic.config(name)
    ctx = htf.HtfContext(pos.shape[0], K, rc); ctx.set_box(lo, hi)
    nl = ctx.build_nlist(torch.from_numpy(pos).cuda())
    nl_o, _, _ = oracle.nlist(pos, lo, hi, rc, K, cells=True)
    fe_o, v9_o, v6_o = oracle.lj(nl_o)
    fe, v6 = ctx.lj_forces(nl, virial=True)
    fe2 = ctx.lj_forces(torch.from_numpy(nl_o).cuda())
    # float64 truth from the oracle tensor
    d = nl_o[:, :, :3].astype(np.float64) + 1e-7; rt = np.sqrt((d**2).sum(-1)); s = np.where(rt > 3e-6, 1/(rt+3e-6), 0)
    coef = np.where(rt > 3e-6, (24*s**7 - 48*s**13)/np.where(rt>0, rt, 1), 0)
    f64 = (coef[:, :, None]*d).sum(1)
    print(name, "forces gpu-order vs oracle   max rel err %.3e (rms scale %.3g)" % err(fe.cpu().numpy()[:, :3], fe_o[:, :3]))
    print(name, "forces oracle-order vs oracle max rel err %.3e" % err(fe2.cpu().numpy()[:, :3], fe_o[:, :3])[0])
    print(name, "oracle fp32 vs float64 truth  max rel err %.3e" % err(fe_o[:, :3], f64)[0])
    print(name, "gpu    fp32 vs float64 truth  max rel err %.3e" % err(fe.cpu().numpy()[:, :3], f64)[0])
    print(name, "energy %.3e  virial %.3e" % (err(fe.cpu().numpy()[:, 3], fe_o[:, 3])[0], err(v6.cpu().numpy(), v6_o)[0]))
