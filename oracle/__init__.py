"""CPU oracle for the hoomd-tf nlist -> forces+virial path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package, and only as the checker or the CPU
baseline.  The product (``hoomd-tf_b200/``) never imports it.

The arithmetic lives in ``htf_oracle.c`` (fp32, no FMA, fixed operation order); this
module is the numpy/ctypes front end plus the scalar EDS layer restatement
(/root/reference htf/layers.py:142-195).  See the header of ``htf_oracle.c`` for the
reference file:line each function follows and for the pinning status
("parity unpinned" for the RDF bin rule and the EDS/Adam update).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libhtf_oracle.so")
_lib = None


def build(force=False):
    """Compile libhtf_oracle.so with the committed Makefile (gcc, -ffp-contract=off)."""
    src = os.path.join(_HERE, "htf_oracle.c")
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= os.path.getmtime(src)):
        return _LIB_PATH
    subprocess.check_call(["make", "-s", "-C", _HERE, "libhtf_oracle.so"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = ctypes.CDLL(_LIB_PATH)
        fp = ctypes.POINTER(ctypes.c_float)
        ip = ctypes.POINTER(ctypes.c_int32)
        lp = ctypes.POINTER(ctypes.c_int64)
        i64, i32, f32 = ctypes.c_int64, ctypes.c_int, ctypes.c_float
        for name in ("htf_oracle_nlist", "htf_oracle_nlist_cells"):
            fn = getattr(L, name)
            fn.restype = ctypes.c_int
            fn.argtypes = [fp, i64, fp, fp, f32, i32, i64, i64, i32, fp, ip, ip]
        L.htf_oracle_lj.restype = ctypes.c_int
        L.htf_oracle_lj.argtypes = [fp, i64, i32, fp, fp, fp]
        L.htf_oracle_rdf_hist.restype = ctypes.c_int
        L.htf_oracle_rdf_hist.argtypes = [fp, i64, i32, fp, f32, f32, i32, i32, i32, lp]
        L.htf_oracle_cv.restype = ctypes.c_int
        L.htf_oracle_cv.argtypes = [fp, i64, i32, f32, fp, fp]
        L.htf_oracle_num_threads.restype = ctypes.c_int
        L.htf_oracle_set_threads.restype = ctypes.c_int
        L.htf_oracle_set_threads.argtypes = [ctypes.c_int]
        _lib = L
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct)) if a is not None else None


def num_threads():
    return int(lib().htf_oracle_num_threads())


def set_threads(n=0):
    """OpenMP threads for the following calls (0 = every online processor); returns the count in effect.
    bench.py calls this because torchrun exports OMP_NUM_THREADS=1."""
    return int(lib().htf_oracle_set_threads(int(n)))


def nlist(pos, box_lo, box_hi, r_cut, K, row_lo=0, row_hi=None, cells=None, map_type_start=-1,
          want_idx=True):
    """Neighbor tensor of rows [row_lo,row_hi) (htf/TensorflowCompute.cc:304-374).

    Returns (nlist[rows,K,4] f32, idx[rows,K] i32 (-1 padded), count[rows] i32).
    ``cells=None`` picks the O(N^2) loop for n <= 4096 and the cell-list candidates above.
    """
    pos = _f32(pos)
    n = pos.shape[0]
    row_hi = n if row_hi is None else row_hi
    rows = row_hi - row_lo
    lo, hi = _f32(box_lo), _f32(box_hi)
    out = np.empty((rows, K, 4), dtype=np.float32)
    idx = np.empty((rows, K), dtype=np.int32) if want_idx else None
    cnt = np.empty((rows,), dtype=np.int32)
    if cells is None:
        cells = n > 4096
    fn = lib().htf_oracle_nlist_cells if cells else lib().htf_oracle_nlist
    rc = fn(_ptr(pos, ctypes.c_float), n, _ptr(lo, ctypes.c_float), _ptr(hi, ctypes.c_float),
            float(r_cut), int(K), int(row_lo), int(row_hi), int(map_type_start),
            _ptr(out, ctypes.c_float), _ptr(idx, ctypes.c_int32), _ptr(cnt, ctypes.c_int32))
    if rc != 0:
        raise RuntimeError("oracle nlist failed: %d" % rc)
    return out, idx, cnt


def lj(nl, virial=True):
    """LJModel forces+energy [rows,4] and virial ([rows,9], [rows,6]) from a neighbor tensor."""
    nl = _f32(nl)
    rows, K = nl.shape[0], nl.shape[1]
    fe = np.empty((rows, 4), dtype=np.float32)
    v9 = np.empty((rows, 9), dtype=np.float32) if virial else None
    v6 = np.empty((rows, 6), dtype=np.float32) if virial else None
    rc = lib().htf_oracle_lj(_ptr(nl, ctypes.c_float), rows, K, _ptr(fe, ctypes.c_float),
                             _ptr(v9, ctypes.c_float), _ptr(v6, ctypes.c_float))
    if rc != 0:
        raise RuntimeError("oracle lj failed: %d" % rc)
    return fe, v9, v6


def rdf_hist(nl, r_range, nbins=100, row_type=None, type_i=None, type_j=None):
    """compute_rdf's integer histogram over nbins+2 bins (htf/simmodel.py:662)."""
    nl = _f32(nl)
    rows, K = nl.shape[0], nl.shape[1]
    hist = np.zeros((nbins + 2,), dtype=np.int64)
    rt = _f32(row_type) if row_type is not None else None
    if type_i is not None and rt is None:
        raise ValueError("type_i needs row_type")
    rc = lib().htf_oracle_rdf_hist(_ptr(nl, ctypes.c_float), rows, K, _ptr(rt, ctypes.c_float),
                                   float(np.float32(r_range[0])), float(np.float32(r_range[1])), int(nbins),
                                   -1 if type_i is None else int(type_i),
                                   -1 if type_j is None else int(type_j),
                                   _ptr(hist, ctypes.c_int64))
    if rc != 0:
        raise RuntimeError("oracle rdf failed: %d" % rc)
    return hist


def coordination_cv(nl, r0):
    """Smooth coordination numbers cn[rows] and their gradient sums grad[rows,3] (BASELINE config 5 CV)."""
    nl = _f32(nl)
    rows, K = nl.shape[0], nl.shape[1]
    cn = np.empty((rows,), dtype=np.float32)
    g = np.empty((rows, 3), dtype=np.float32)
    rc = lib().htf_oracle_cv(_ptr(nl, ctypes.c_float), rows, K, float(r0), _ptr(cn, ctypes.c_float), _ptr(g, ctypes.c_float))
    if rc != 0:
        raise RuntimeError("oracle cv failed: %d" % rc)
    return cn, g


def rdf_from_hist(hist, r_range, nbins=100):
    """hist[nbins+2] -> (rdf[nbins], bin centres[nbins]) as htf/simmodel.py:663-669 (fp32)."""
    lo, hi = np.float32(r_range[0]), np.float32(r_range[1])
    shell = np.linspace(lo, hi, nbins + 1, dtype=np.float32)
    vis = ((shell[1:] + shell[:-1]) * np.float32(0.5)).astype(np.float32)
    vols = (shell[1:] ** 3 - shell[:-1] ** 3).astype(np.float32)
    return (hist[1:-1].astype(np.float32) / vols).astype(np.float32), vis


def pairwise_mlp(nl, raw, rbf_high):
    """fp32 numpy restatement of the pairwise-MLP force field (BASELINE config 3): RBFExpansion(0, rbf_high, 32)
    (htf/layers.py:27-49) -> 3 x Dense(64, tanh) -> Dense(1), e_i = 1/2 sum_j u(r_ij) over the non-padded slots,
    forces by compute_nlist_forces' convention (htf/simmodel.py:542-550) with the analytic du/dr.
    ``raw`` is the parameter blob of include/htf_b200.h (torch.nn.Linear layout).  Returns [rows,4] (F, e).
    PARITY UNPINNED against the reference (it has no such model with fixed weights); it is pinned against the
    torch autograd evaluation of the same network in tests/."""
    f = np.float32
    nl = _f32(nl)
    raw = _f32(raw)
    o = 0
    def take(n, shape):
        nonlocal o
        a = raw[o:o + n].reshape(shape); o += n
        return a
    W1, b1 = take(64 * 32, (64, 32)), take(64, (64,))
    W2, b2 = take(64 * 64, (64, 64)), take(64, (64,))
    W3, b3 = take(64 * 64, (64, 64)), take(64, (64,))
    w4, b4 = take(64, (64,)), take(1, (1,))
    d = nl[..., :3] + f(1e-7)
    r = np.sqrt((d * d).sum(-1, dtype=f))
    mu = np.linspace(0.0, rbf_high, 32, dtype=f)
    gap = f(mu[1] - mu[0])
    u_c = r[..., None] - mu
    phi = np.exp(-(u_c * u_c) / gap).astype(f)
    dphi = (f(-2.0) * u_c / gap * phi).astype(f)
    h, hp = phi, dphi
    for W, b in ((W1, b1), (W2, b2), (W3, b3)):
        z, zp = h @ W.T + b, hp @ W.T
        h = np.tanh(z).astype(f)
        hp = ((f(1.0) - h * h) * zp).astype(f)
    u, du = h @ w4 + b4[0], hp @ w4
    mask = r > f(3e-6)
    coef = np.where(mask, du / r, f(0)).astype(f)
    out = np.empty(nl.shape[:1] + (4,), dtype=f)
    out[:, :3] = (coef[..., None] * d).sum(1, dtype=f)
    out[:, 3] = (f(0.5) * np.where(mask, u, f(0))).sum(1, dtype=f)
    return out


class EDSLayer:
    """Scalar restatement of htf/layers.py:142-195 (EDSLayer.call) with tf.compat.v1 Adam
    (beta1 .9, beta2 .999, eps 1e-8, lr_t = lr*sqrt(1-b2^t)/(1-b1^t)); fp32 state.
    PARITY UNPINNED: the reference only tests a convergence band (test_utils.py:447-461)."""

    def __init__(self, set_point, period, learning_rate=1e-2, cv_scale=1.0):
        f = np.float32
        self.set_point, self.period = f(set_point), int(period)
        self.lr, self.cv_scale = f(learning_rate), f(cv_scale)
        self.mean = f(0); self.ssd = f(0); self.n = 0; self.alpha = f(0)
        self.m = f(0); self.v = f(0); self.t = 0

    def __call__(self, cv):
        f = np.float32
        cv = f(cv)
        h = self.period // 2
        if self.n == 0:                                   # layers.py:161-165
            self.mean = f(0); self.ssd = f(0)
        if self.n > h:                                    # :169-178
            delta = f(cv - self.mean)
            self.mean = f(self.mean + f(delta / f(self.n - h)))
            self.ssd = f(self.ssd + f(delta * f(cv - self.mean)))
        if self.n == self.period - 1:                     # :181-190
            g = f(f(f(f(f(-2) * f(self.mean - self.set_point)) * self.ssd) / f(self.period)) / f(2))
            g = f(g / self.cv_scale)
            self.t += 1
            self.m = f(f(0.9) * self.m + f(0.1) * g)
            self.v = f(f(0.999) * self.v + f(0.001) * f(g * g))
            lr_t = f(self.lr * f(np.sqrt(f(1) - f(0.999) ** self.t)) / f(f(1) - f(0.9) ** self.t))
            self.alpha = f(self.alpha - f(lr_t * self.m) / f(f(np.sqrt(self.v)) + f(1e-8)))
        self.n = (self.n + 1) % self.period               # :193
        return self.alpha


def pairwise_mlp_train_grads(nl, raw, rbf_high, labels):
    """Gradient of the force-matching loss of the pairwise MLP with respect to its parameters -- numpy restatement of
    what Keras ``train_on_batch`` differentiates in the reference's label/training mode
    (/root/reference htf/tensorflowcompute.py:346-370: MSE between the model's first output [N,4] = forces + energy and
    the label forces [N,4]), for the config-4 model (RBF(32) -> 3 x Dense(64, tanh) -> Dense(1), e_i = 1/2 sum_j u).

    L = mean over the N x 4 entries of (pred - labels)^2.  With r, u(r), u'(r) per pair, F_i = sum_j u' a / r and
    e_i = 1/2 sum_j u (valid pairs), dL/dtheta = sum_p [g_p du'_p/dtheta + h_p du_p/dtheta] with
    g_p = (dF_i . a_p / r_p) / (2N) and h_p = de_i / (4N): the reverse sweep through the value/tangent chain below.
    Returns (loss, grads in the raw-blob layout of include/htf_b200.h, pred[N,4]).  float64 throughout: this is the
    checker for the tensor-core training kernel (PARITY UNPINNED against the reference, which has no fixed-weight
    instance of such a model; pinned against torch autograd in tests/)."""
    f = np.float64
    nl = np.asarray(nl, dtype=f)
    raw = np.asarray(raw, dtype=f)
    labels = np.asarray(labels, dtype=f)
    o = 0
    def take(n, shape):
        nonlocal o
        a = raw[o:o + n].reshape(shape); o += n
        return a
    W1, b1 = take(64 * 32, (64, 32)), take(64, (64,))
    W2, b2 = take(64 * 64, (64, 64)), take(64, (64,))
    W3, b3 = take(64 * 64, (64, 64)), take(64, (64,))
    w4, b4 = take(64, (64,)), take(1, (1,))
    N, K = nl.shape[0], nl.shape[1]
    a = nl[..., :3] + 1e-7
    r = np.sqrt((a * a).sum(-1))
    mask = r > 3e-6
    mu = np.linspace(0.0, rbf_high, 32)
    gap = mu[1] - mu[0]
    uc = r[..., None] - mu
    phi = np.exp(-(uc * uc) / gap)
    dphi = -2.0 * uc / gap * phi
    hs, hps = [phi], [dphi]
    for W, b in ((W1, b1), (W2, b2), (W3, b3)):
        z, zp = hs[-1] @ W.T + b, hps[-1] @ W.T
        h = np.tanh(z)
        hs.append(h); hps.append((1.0 - h * h) * zp)
    u, du = hs[-1] @ w4 + b4[0], hps[-1] @ w4
    coef = np.where(mask, du / np.where(mask, r, 1.0), 0.0)
    pred = np.empty((N, 4))
    pred[:, :3] = (coef[..., None] * a).sum(1)
    pred[:, 3] = (0.5 * np.where(mask, u, 0.0)).sum(1)
    diff = pred - labels
    loss = float((diff ** 2).mean())
    g = np.where(mask, (diff[:, None, :3] * a).sum(-1) / np.where(mask, r, 1.0), 0.0) / (2.0 * N)     # weight of du'/dtheta
    hh = np.where(mask, diff[:, 3:4] / (4.0 * N), 0.0) * np.ones((1, K))                             # weight of du/dtheta
    # reverse sweep
    gw4 = (hh[..., None] * hs[3] + g[..., None] * hps[3]).sum((0, 1))
    gb4 = np.array([hh.sum()])
    hb, hpb = hh[..., None] * w4, g[..., None] * w4            # adjoints of h3, h3'
    grads = []
    for l, W in ((3, W3), (2, W2), (1, W1)):
        h, hp = hs[l], hps[l]
        s = 1.0 - h * h
        zb = s * hb - 2.0 * h * hp * hpb
        zpb = s * hpb
        gW = np.einsum("nko,nki->oi", zb, hs[l - 1]) + np.einsum("nko,nki->oi", zpb, hps[l - 1])
        gb = zb.sum((0, 1))
        grads.append((gW, gb))
        hb, hpb = zb @ W, zpb @ W
    (gW3, gb3), (gW2, gb2), (gW1, gb1) = grads
    flat = np.concatenate([gW1.ravel(), gb1, gW2.ravel(), gb2, gW3.ravel(), gb3, gw4, gb4])
    return loss, flat, pred


def adam_step(params, grads, m, v, t, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-7):
    """One Keras Adam step (tf.keras.optimizers.Adam, the optimizer behind the reference's ``train_on_batch``;
    defaults lr 1e-3, epsilon 1e-7): t is the step count AFTER the increment.  fp32, returns (params, m, v)."""
    f = np.float32
    params, grads, m, v = (np.asarray(x, dtype=f) for x in (params, grads, m, v))
    # Keras holds beta_1 / beta_2 as float32 tensors: 1 - beta is the float32 difference (1 - 0.999f = 0.00100004673)
    m = (f(beta1) * m + (f(1.0) - f(beta1)) * grads).astype(f)
    v = (f(beta2) * v + (f(1.0) - f(beta2)) * grads * grads).astype(f)
    lr_t = f(lr * np.sqrt(1.0 - beta2 ** t) / (1.0 - beta1 ** t))
    params = (params - lr_t * m / (np.sqrt(v) + f(eps))).astype(f)
    return params, m, v
