import sys, os
sys.path.insert(0, "/root/repo/hoomd-tf_b200"); sys.path.insert(0, "/root/repo")
import numpy as np, torch, htf, oracle
from htf import synthetic
pos, lo, hi = synthetic.lattice_fluid((8, 8, 8), 0.5, seed=9)
pos[-40:, 3] = 2.0 + (np.arange(40) % 2)
K, r_cut = 64, 3.0
ctx = htf.HtfContext(pos.shape[0], K, r_cut); ctx.set_box(lo, hi); ctx.set_mapped_nlist(2)
dpos = torch.from_numpy(pos).cuda()
nl, idx, cnt = ctx.build_nlist(dpos, want_idx=True, want_count=True)
nl_o, idx_o, cnt_o = oracle.nlist(pos, lo, hi, r_cut, K, map_type_start=2)
idx = idx.cpu().numpy(); nl = nl.cpu().numpy()
print("cnt equal", np.array_equal(cnt.cpu().numpy(), cnt_o), cnt_o.max())
key = np.where(idx < 0, 2**31-1, idx); order = np.argsort(key, 1, kind="stable")
ids = np.take_along_axis(idx, order, 1); nls = np.take_along_axis(nl, order[:, :, None], 1)
bad = np.where((ids != idx_o).any(1))[0]; print("bad idx rows", bad[:10], len(bad))
bad2 = np.where((nls.view(np.uint32) != nl_o.view(np.uint32)).any((1,2)))[0]; print("bad nl rows", bad2[:10], len(bad2))
if len(bad2):
    r = bad2[0]; print(ids[r][:12], idx_o[r][:12]); print(nls[r][:4]); print(nl_o[r][:4])
# also without idx
nl2 = ctx.build_nlist(dpos).cpu().numpy()
print("noidx variant equals idx variant:", np.array_equal(nl2, nl))
