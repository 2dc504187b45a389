"""Thin torch-tensor front end over the C ABI (one ``HtfContext`` per device).

Plays the role of the reference's ``_htf.TensorflowComputeGPU`` object plus its
``TFArrayComm`` buffers (/root/reference htf/TensorflowCompute.cc:422-486,
htf/TFArrayComm.h:86-231): torch only provides device memory and the current stream.
"""
import ctypes

import torch

from . import _lib


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _check_dev_f32(t, name, last=None):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise ValueError("%s must be a CUDA tensor (the hot path has no CPU fallback)" % name)
    if t.dtype != torch.float32 or not t.is_contiguous():
        raise ValueError("%s must be contiguous float32" % name)
    if last is not None and t.shape[-1] != last:
        raise ValueError("%s must have last dimension %d" % (name, last))


def _check_counts(c, rows):
    if c is None:
        return
    if not isinstance(c, torch.Tensor) or not c.is_cuda or c.dtype != torch.int32 or not c.is_contiguous() \
            or c.numel() != rows:
        raise ValueError("counts must be a contiguous int32 CUDA tensor with one entry per row")


class HtfContext:
    def __init__(self, n_max, nneighbor_cutoff, r_cut, device=None, deterministic=True):
        if not torch.cuda.is_available():
            raise RuntimeError("htf_b200 needs a CUDA device: the hot path has no CPU fallback")
        self.lib = _lib.load()
        idx = None if device is None else torch.device(device).index
        # an index-less "cuda" means the process's current device (one process per GPU sets it to its rank)
        self.device = torch.device("cuda", torch.cuda.current_device() if idx is None else idx)
        self.K = int(nneighbor_cutoff)
        self.r_cut = float(r_cut)
        self._h = ctypes.c_void_p()
        flags = _lib.FLAG_DETERMINISTIC if deterministic else 0
        rc = self.lib.htf_create(ctypes.byref(self._h), self.device.index, int(n_max), self.K, self.r_cut, flags)
        if rc != 0:
            raise _lib.HtfError(rc, self.lib.htf_last_error(None).decode())
        self._overflow = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.box = None

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.htf_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- helpers --
    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _ck(self, rc):
        return _lib.check(self._h, rc)

    # -- configuration --
    def set_box(self, lo, hi, tilt=(0.0, 0.0, 0.0)):
        """updateBox (htf/TensorflowCompute.cc:272-282) + skew check (htf/simmodel.py:195)."""
        a3 = ctypes.c_float * 3
        self._ck(self.lib.htf_set_box(self._h, a3(*[float(x) for x in lo]), a3(*[float(x) for x in hi]),
                                      a3(*[float(x) for x in tilt])))
        self.box = ([float(x) for x in lo], [float(x) for x in hi], [float(x) for x in tilt])

    def set_cutoff(self, r_cut, nneighbor_cutoff):
        self._ck(self.lib.htf_set_cutoff(self._h, float(r_cut), int(nneighbor_cutoff)))
        self.r_cut, self.K = float(r_cut), int(nneighbor_cutoff)

    def set_roi(self, center=None, half_width=None):
        """Bin only particles within ``half_width`` (per axis, minimum image) of ``center``; None switches it off."""
        if center is None:
            self._ck(self.lib.htf_set_roi(self._h, None, None))
            return
        a3 = ctypes.c_float * 3
        self._ck(self.lib.htf_set_roi(self._h, a3(*[float(x) for x in center]), a3(*[float(x) for x in half_width])))

    def pack_halo(self, pos, axis, threshold, below, out, count=None):
        """Stable copy of the particles beyond a plane into the fixed-capacity buffer ``out`` (sentinel padded)."""
        _check_dev_f32(pos, "positions", 4)
        _check_dev_f32(out, "halo buffer", 4)
        self._ck(self.lib.htf_pack_halo(self._h, _ptr(pos), pos.shape[0], int(axis), float(threshold), int(bool(below)),
                                        _ptr(out), out.shape[0], _ptr(count), _ptr(self._overflow), self._stream()))
        return out

    def pack_halo_pair(self, pos, axis, threshold_lo, threshold_hi, out_lo, out_hi, counts=None):
        """Both slab faces in one pass: ``pos[axis] < threshold_lo`` -> out_lo, ``> threshold_hi`` -> out_hi."""
        _check_dev_f32(pos, "positions", 4)
        _check_dev_f32(out_lo, "halo buffer", 4)
        _check_dev_f32(out_hi, "halo buffer", 4)
        if out_lo.shape[0] != out_hi.shape[0]:
            raise ValueError("the two halo buffers must have the same capacity")
        self._ck(self.lib.htf_pack_halo_pair(self._h, _ptr(pos), pos.shape[0], int(axis), float(threshold_lo),
                                             float(threshold_hi), _ptr(out_lo), _ptr(out_hi), out_lo.shape[0],
                                             _ptr(counts), _ptr(self._overflow), self._stream()))
        return out_lo, out_hi

    def eds_step(self, cv, layer):
        """One EDSLayer update on the device (``layer``: htf.layers.EDSLayer with CUDA float32 state)."""
        self._ck(self.lib.htf_eds_step(self._h, _ptr(cv), _ptr(layer.set_point), _ptr(layer.mean), _ptr(layer.ssd),
                                       _ptr(layer.n), _ptr(layer.alpha), _ptr(layer.adam_m), _ptr(layer.adam_v),
                                       _ptr(layer.adam_t), int(layer.period), float(layer.learning_rate),
                                       float(layer.cv_scale), self._stream()))

    def integrate_half(self, half, pos, vel, force, dt, gamma=0.0, kT=0.0, flat=False, seed=0, timestep=0):
        """Velocity-Verlet half step on the device (half 0: kick + drift + wrap, half 1: kick); see include/htf_b200.h."""
        _check_dev_f32(pos, "positions", 4)
        _check_dev_f32(force, "forces", 4)
        if not (vel.is_cuda and vel.dtype == torch.float32 and vel.is_contiguous() and vel.shape == (pos.shape[0], 3)):
            raise ValueError("velocities must be a contiguous float32 CUDA tensor [N,3]")
        self._ck(self.lib.htf_integrate_half(self._h, int(half), _ptr(pos), _ptr(vel), _ptr(force), pos.shape[0], float(dt),
                                             float(gamma), float(kT), int(bool(flat)), int(seed), int(timestep),
                                             self._stream()))

    # ---- buffered ("skin") lists: search every few steps, distance filter every step ----
    def skin_configure(self, skin, k_candidates=0):
        self._ck(self.lib.htf_skin_configure(self._h, float(skin), int(k_candidates)))

    def skin_rebuild(self, pos, row_lo=0, row_hi=None):
        _check_dev_f32(pos, "positions", 4)
        n = pos.shape[0]
        self._ck(self.lib.htf_skin_rebuild(self._h, _ptr(pos), n, int(row_lo), n if row_hi is None else int(row_hi),
                                           self._stream()))

    def skin_nlist(self, pos, row_lo=0, row_hi=None, out=None, want_idx=False, want_count=False, count_out=None):
        """The per-step pass: neighbor tensor of the rows from the candidate lists of the last ``skin_rebuild``."""
        _check_dev_f32(pos, "positions", 4)
        n = pos.shape[0]
        row_hi = n if row_hi is None else int(row_hi)
        rows = row_hi - int(row_lo)
        if out is None:
            out = torch.empty((rows, self.K, 4), dtype=torch.float32, device=self.device)
        idx = torch.empty((rows, self.K), dtype=torch.int32, device=self.device) if want_idx else None
        cnt = torch.empty((rows,), dtype=torch.int32, device=self.device) if want_count else None
        if count_out is not None:                       # caller-owned counts (not returned)
            _check_counts(count_out, rows)
            cnt = count_out
        self._ck(self.lib.htf_skin_nlist(self._h, _ptr(pos), n, int(row_lo), row_hi, _ptr(out), _ptr(idx), _ptr(cnt),
                                         _ptr(self._overflow), self._stream()))
        if want_idx or want_count:
            return out, idx, cnt
        return out

    def skin_status(self, reset=True):
        """(rows used with a particle displaced by more than skin/2, rows whose candidate list overflowed); syncs."""
        h = (ctypes.c_int32 * 2)()
        self._ck(self.lib.htf_skin_status(self._h, h, int(bool(reset)), self._stream()))
        return int(h[0]), int(h[1])

    def set_mapped_nlist(self, map_type_start):
        self._ck(self.lib.htf_set_mapped_nlist(self._h, -1 if map_type_start is None else int(map_type_start)))

    def cell_grid(self):
        n = (ctypes.c_int * 3)()
        self._ck(self.lib.htf_get_cell_grid(self._h, n))
        return tuple(n)

    @property
    def launches(self):
        return int(self.lib.htf_launch_count(self._h))

    # -- the path --
    def bin_particles(self, pos):
        _check_dev_f32(pos, "positions", 4)
        self._ck(self.lib.htf_bin_particles(self._h, _ptr(pos), pos.shape[0], self._stream()))

    def build_nlist(self, pos, row_lo=0, row_hi=None, out=None, want_idx=False, want_count=False, rebin=True,
                    count_out=None):
        """positions [N,4] -> nlist [rows,K,4] (and optionally idx [rows,K], count [rows])."""
        _check_dev_f32(pos, "positions", 4)
        n = pos.shape[0]
        row_hi = n if row_hi is None else int(row_hi)
        rows = row_hi - int(row_lo)
        if rebin:
            self.bin_particles(pos)
        if out is None:
            out = torch.empty((rows, self.K, 4), dtype=torch.float32, device=self.device)
        else:
            _check_dev_f32(out, "nlist out", 4)
        idx = torch.empty((rows, self.K), dtype=torch.int32, device=self.device) if want_idx else None
        cnt = torch.empty((rows,), dtype=torch.int32, device=self.device) if want_count else None
        if count_out is not None:                       # caller-owned counts (not returned)
            _check_counts(count_out, rows)
            cnt = count_out
        self._ck(self.lib.htf_build_nlist(self._h, _ptr(pos), n, int(row_lo), row_hi, _ptr(out), _ptr(idx),
                                          _ptr(cnt), _ptr(self._overflow), self._stream()))
        res = (out,)
        if want_idx:
            res += (idx,)
        if want_count:
            res += (cnt,)
        return res if len(res) > 1 else out

    def overflow(self, reset=True):
        """max neighbor count over rows that reached K since the last reset (0 = no row is full).
        Synchronises the stream (D2H read of one int)."""
        v = int(self._overflow.item())
        if reset and v:
            self._overflow.zero_()
        return v

    def lj_forces(self, nlist, virial=False, virial_components=6, out=None, virial_out=None, counts=None):
        """LJ forces + energy (+ virial) of a neighbor tensor.  ``counts`` int32[rows] (``build_nlist(want_count=True)``)
        lets the pass skip every row's zero padding instead of reading it."""
        _check_dev_f32(nlist, "nlist", 4)
        _check_counts(counts, nlist.shape[0])
        if nlist.dim() != 3:
            raise ValueError("nlist must be [rows, K, 4]")
        rows, k = nlist.shape[0], nlist.shape[1]
        fe = out if out is not None else torch.empty((rows, 4), dtype=torch.float32, device=self.device)
        vir = None
        if virial:
            vir = virial_out if virial_out is not None else \
                torch.empty((rows, virial_components), dtype=torch.float32, device=self.device)
        self._ck(self.lib.htf_lj_forces(self._h, _ptr(nlist), rows, int(k), _ptr(counts), _ptr(fe), _ptr(vir),
                                        int(virial_components), self._stream()))
        return (fe, vir) if virial else fe

    def lj_step_forces_only(self, nlist, force_out, virial_out, bins, r_range, nbins=100, counts=None):
        """LJ forces+virial with the compute_rdf histogram fused into the same pass over ``nlist``."""
        _check_dev_f32(nlist, "nlist", 4)
        _check_counts(counts, nlist.shape[0])
        vc = virial_out.shape[1] if virial_out is not None else 6
        self._ck(self.lib.htf_lj_forces_rdf(self._h, _ptr(nlist), nlist.shape[0], int(nlist.shape[1]),
                                            _ptr(counts), _ptr(force_out), _ptr(virial_out), int(vc), _ptr(bins),
                                            float(r_range[0]), float(r_range[1]), int(nbins), self._stream()))
        return force_out

    def lj_cv_forces(self, nlist, r0, cv_row, cv_sum, force_out=None, virial_out=None, bins=None,
                     r_range=(0.0, 1.0), nbins=100, counts=None):
        """LJ forces(+virial) + smooth coordination CV (+ RDF) in one pass (BASELINE config 5 model).
        ``cv_row`` float32[rows,4] = (dCV-sum/dd x,y,z, cn_i); ``cv_sum`` float64[1] is incremented."""
        _check_dev_f32(nlist, "nlist", 4)
        _check_dev_f32(cv_row, "cv_row", 4)
        if cv_sum.dtype != torch.float64 or not cv_sum.is_cuda:
            raise ValueError("cv_sum must be a float64 CUDA tensor")
        rows, k = nlist.shape[0], nlist.shape[1]
        if force_out is None:
            force_out = torch.empty((rows, 4), dtype=torch.float32, device=self.device)
        vc = virial_out.shape[1] if virial_out is not None else 6
        _check_counts(counts, rows)
        self._ck(self.lib.htf_lj_cv_forces(self._h, _ptr(nlist), rows, int(k), _ptr(counts), float(r0), _ptr(force_out),
                                           _ptr(virial_out), int(vc), _ptr(cv_row), _ptr(cv_sum), _ptr(bins),
                                           float(r_range[0]), float(r_range[1]), int(nbins), self._stream()))
        return force_out

    def mlp_param_sizes(self):
        """(number of fp32 values in the raw parameter blob, bytes of the packed blob)."""
        a, b = ctypes.c_int(), ctypes.c_int()
        self.lib.htf_mlp_param_sizes(ctypes.byref(a), ctypes.byref(b))
        return a.value, b.value

    def mlp_pack(self, raw, packed=None):
        """fp32 parameter blob (torch.nn.Linear layout, see include/htf_b200.h) -> packed bf16 operand layouts."""
        n_raw, n_packed = self.mlp_param_sizes()
        if not (raw.is_cuda and raw.dtype == torch.float32 and raw.is_contiguous() and raw.numel() == n_raw):
            raise ValueError("raw MLP parameters must be a contiguous float32 CUDA tensor of %d values" % n_raw)
        if packed is None:
            packed = torch.empty(n_packed, dtype=torch.uint8, device=self.device)
        self._ck(self.lib.htf_mlp_pack(self._h, _ptr(raw), _ptr(packed), self._stream()))
        return packed

    def mlp_forces(self, nlist, packed, rbf_high, out=None, counts=None):
        """Pairwise-MLP forces+energy [rows,4] on the tensor cores (tcgen05).  ``counts`` int32[rows] (the builder's
        neighbors per row) makes the compaction pre-pass of large tensors read only the valid slots."""
        _check_dev_f32(nlist, "nlist", 4)
        rows, k = nlist.shape[0], nlist.shape[1]
        _check_counts(counts, rows)
        if out is None:
            out = torch.empty((rows, 4), dtype=torch.float32, device=self.device)
        self._ck(self.lib.htf_mlp_forces(self._h, _ptr(nlist), rows, int(k), _ptr(counts), _ptr(packed), float(rbf_high),
                                         _ptr(out), self._stream()))
        return out

    def mlp_train_grads(self, nlist, raw, rbf_high, labels, n_total=None, grads=None, pred=None, loss=None):
        """Gradient of the force-matching MSE of the pairwise MLP w.r.t. its raw parameters (see include/htf_b200.h).
        Returns (grads float32[10497], pred float32[rows,4], loss float32[1])."""
        _check_dev_f32(nlist, "nlist", 4)
        _check_dev_f32(labels, "labels", 4)
        rows, k = nlist.shape[0], nlist.shape[1]
        n_raw, _ = self.mlp_param_sizes()
        if not (raw.is_cuda and raw.dtype == torch.float32 and raw.is_contiguous() and raw.numel() == n_raw):
            raise ValueError("raw MLP parameters must be a contiguous float32 CUDA tensor of %d values" % n_raw)
        if labels.shape[0] != rows:
            raise ValueError("labels must have one row per nlist row")
        if grads is None:
            grads = torch.empty(n_raw, dtype=torch.float32, device=self.device)
        if pred is None:
            pred = torch.empty((rows, 4), dtype=torch.float32, device=self.device)
        if loss is None:
            loss = torch.empty(1, dtype=torch.float32, device=self.device)
        self._ck(self.lib.htf_mlp_train_grads(self._h, _ptr(nlist), rows, int(k), _ptr(raw), float(rbf_high), _ptr(labels),
                                              int(rows if n_total is None else n_total), _ptr(pred), _ptr(grads), _ptr(loss),
                                              self._stream()))
        return grads, pred, loss

    def adam_step(self, params, grads, m, v, t, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-7):
        """Fused Keras-Adam update in place; ``t`` is a float32[1] CUDA step counter that the call increments."""
        for x in (params, grads, m, v, t):
            if not (x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()):
                raise ValueError("adam_step takes contiguous float32 CUDA tensors")
        self._ck(self.lib.htf_adam_step(self._h, _ptr(params), _ptr(grads), _ptr(m), _ptr(v), _ptr(t), int(params.numel()),
                                        float(lr), float(beta1), float(beta2), float(eps), self._stream()))
        return params

    def rdf_hist(self, nlist, r_range, nbins=100, row_pos=None, type_i=None, type_j=None, bins=None,
                 type_tensor=None):
        """compute_rdf's integer histogram: int64[nbins+2], accumulated into ``bins`` if given.

        Row types come either from ``row_pos`` ([rows,4] positions, column 3) or from
        ``type_tensor`` (any 1-D float32 view, e.g. ``positions[:, 3]``; its stride is honoured)."""
        _check_dev_f32(nlist, "nlist", 4)
        rows, k = nlist.shape[0], nlist.shape[1]
        if bins is None:
            bins = torch.zeros((nbins + 2,), dtype=torch.int64, device=self.device)
        tptr, tstride = None, 0
        if row_pos is not None:
            _check_dev_f32(row_pos, "row positions", 4)
            tptr, tstride = ctypes.c_void_p(row_pos.data_ptr() + 12), 4
        elif type_tensor is not None:
            t = type_tensor
            if not (t.is_cuda and t.dtype == torch.float32 and t.dim() == 1):
                raise ValueError("type_tensor must be a 1-D float32 CUDA tensor")
            tptr, tstride = ctypes.c_void_p(t.data_ptr()), (t.stride(0) if t.shape[0] > 1 else 1)
        if tptr is not None and (row_pos if row_pos is not None else type_tensor).shape[0] != rows:
            raise ValueError("row types must have one entry per nlist row")
        self._ck(self.lib.htf_rdf_hist(self._h, _ptr(nlist), rows, int(k), tptr, int(tstride),
                                       float(r_range[0]), float(r_range[1]), int(nbins),
                                       -1 if type_i is None else int(type_i),
                                       -1 if type_j is None else int(type_j), _ptr(bins), self._stream()))
        return bins

    def lj_step(self, pos, row_lo=0, row_hi=None, nlist_out=None, force_out=None, virial_out=None,
                virial_components=6, bins=None, r_range=(0.0, 1.0), nbins=100):
        """One computeForces pass of the built-in LJ model (htf/TensorflowCompute.cc:130-216)."""
        _check_dev_f32(pos, "positions", 4)
        n = pos.shape[0]
        row_hi = n if row_hi is None else int(row_hi)
        rows = row_hi - int(row_lo)
        if force_out is None:
            force_out = torch.empty((rows, 4), dtype=torch.float32, device=self.device)
        self._ck(self.lib.htf_lj_step(self._h, _ptr(pos), n, int(row_lo), row_hi, _ptr(nlist_out), _ptr(force_out),
                                      _ptr(virial_out), int(virial_components), _ptr(self._overflow), _ptr(bins),
                                      float(r_range[0]), float(r_range[1]), int(nbins), self._stream()))
        return force_out

    def unstuff4(self, pos_hoomd, out=None):
        """HOOMD Scalar4 positions (type as int bits in .w) -> positions with the type as a float value."""
        _check_dev_f32(pos_hoomd, "positions", 4)
        if out is None:
            out = torch.empty_like(pos_hoomd)
        self._ck(self.lib.htf_unstuff4(self._h, _ptr(pos_hoomd), _ptr(out), pos_hoomd.shape[0], self._stream()))
        return out

    def lj_rows(self, n_all, row_lo, row_hi, nlist_out=None, force_out=None, virial_out=None, virial_components=6,
                bins=None, r_range=(0.0, 1.0), nbins=100):
        """Row batch of the LJ step on the particles of the last ``bin_particles`` (no re-binning)."""
        rows = int(row_hi) - int(row_lo)
        if force_out is None:
            force_out = torch.empty((rows, 4), dtype=torch.float32, device=self.device)
        if virial_out is not None:
            virial_components = virial_out.shape[1]
        self._ck(self.lib.htf_lj_rows(self._h, int(n_all), int(row_lo), int(row_hi), _ptr(nlist_out), _ptr(force_out),
                                      _ptr(virial_out), int(virial_components), _ptr(self._overflow), _ptr(bins),
                                      float(r_range[0]), float(r_range[1]), int(nbins), self._stream()))
        return force_out

    def lj_cv_step(self, pos, r0, cv_row, cv_sum, row_lo=0, row_hi=None, nlist_out=None, force_out=None,
                   virial_out=None, bins=None, r_range=(0.0, 1.0), nbins=100):
        """One step of the EDS-biased config-5 model: bin, build and the fused LJ + CV (+RDF) pass, pipelined."""
        if isinstance(pos, int):
            n, pos = pos, None                      # already binned: ``pos`` is the particle count
        else:
            _check_dev_f32(pos, "positions", 4)
            n = pos.shape[0]
        row_hi = n if row_hi is None else int(row_hi)
        rows = row_hi - int(row_lo)
        if force_out is None:
            force_out = torch.empty((rows, 4), dtype=torch.float32, device=self.device)
        vc = virial_out.shape[1] if virial_out is not None else 6
        self._ck(self.lib.htf_lj_cv_step(self._h, _ptr(pos), n, int(row_lo), row_hi, _ptr(nlist_out), float(r0),
                                         _ptr(force_out), _ptr(virial_out), int(vc), _ptr(cv_row), _ptr(cv_sum),
                                         _ptr(self._overflow), _ptr(bins), float(r_range[0]), float(r_range[1]),
                                         int(nbins), self._stream()))
        return force_out

    def set_pipeline(self, slabs):
        """Slabs of the pipelined step (build of slab i+1 overlaps the pair pass of slab i); <= 1 switches it off."""
        self._ck(self.lib.htf_set_pipeline(self._h, int(slabs)))

    # ---- peer-memory exchange (one process per GPU of one node; see include/htf_b200.h) ----
    def comm_create(self, rank, world, halo_capacity):
        """Allocate this rank's exchange window; returns its 64-byte IPC handle (bytes)."""
        buf = ctypes.create_string_buffer(64)
        self._ck(self.lib.htf_comm_create(self._h, int(rank), int(world), int(halo_capacity), buf))
        return buf.raw

    def comm_connect(self, handles):
        """``handles``: the 64-byte handles of all ranks in rank order."""
        blob = b"".join(bytes(h) for h in handles)
        self._ck(self.lib.htf_comm_connect(self._h, blob))

    def comm_exchange_halo(self, local, n_own, axis, threshold_lo, threshold_hi):
        _check_dev_f32(local, "local positions", 4)
        self._ck(self.lib.htf_comm_exchange_halo(self._h, _ptr(local), int(n_own), int(axis), float(threshold_lo),
                                                 float(threshold_hi), _ptr(self._overflow), self._stream()))
        return local

    def comm_allreduce(self, values):
        """In-place sum over all ranks of a small int64 / float64 / float32 CUDA vector (<= 16384 values)."""
        fns = {torch.int64: self.lib.htf_comm_allreduce_i64, torch.float64: self.lib.htf_comm_allreduce_f64,
               torch.float32: self.lib.htf_comm_allreduce_f32}
        if not values.is_cuda or not values.is_contiguous() or values.dtype not in fns:
            raise ValueError("comm_allreduce takes a contiguous int64, float64 or float32 CUDA tensor")
        fn = fns[values.dtype]
        self._ck(fn(self._h, _ptr(values), int(values.numel()), self._stream()))
        return values

    def comm_status(self):
        v = ctypes.c_int32(0)
        self._ck(self.lib.htf_comm_status(self._h, ctypes.byref(v), self._stream()))
        return int(v.value)

    def comm_destroy(self):
        self._ck(self.lib.htf_comm_destroy(self._h))
