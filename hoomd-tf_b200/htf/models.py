"""Built-in models of the hot path.

Each class is the drop-in for one of the reference's example models
(/root/reference htf/test-py/build_examples.py, htf/test-py/benchmark.py): same name, same
``compute`` signature, same outputs -- but the body calls the fused sm_100a kernels
(``ops.lj_forces``, ``compute_rdf``) instead of building a differentiable graph.  The
``*Autograd`` variants keep the reference's literal bodies on torch autograd; they are the
semantic cross-check for the fused ones and the template for user models.
"""
import torch

from . import ops
from .layers import EDSLayer, RBFExpansion
from .simmodel import (MeanTensor, Mean, SimModel, compute_nlist_forces, compute_positions_forces, compute_rdf,
                       nlist_rinv, safe_norm, wrap_vector)


class LJModel(SimModel):
    """build_examples.py:67-77 / benchmark.py:12-23, fused.

    ``compute`` is the drop-in body for a given neighbor tensor; under ``tfcompute`` the whole row batch is ONE
    library call instead (``fused_rows``: build + LJ pass, pipelined slab by slab, straight into the force rows)."""

    def compute(self, nlist, positions, box):
        return ops.lj_forces(nlist)

    def fused_rows(self, tfc, n, off, hi):
        tfc.ctx.lj_rows(n, off, hi, nlist_out=tfc._nlist_buf, force_out=tfc._forces[off:hi])


class LJVirialModel(SimModel):
    """build_examples.py:104-115 (construct with ``virial=True``), fused."""

    def compute(self, nlist, positions, box):
        return ops.lj_forces(nlist, virial=True)

    def fused_rows(self, tfc, n, off, hi):
        tfc.ctx.lj_rows(n, off, hi, nlist_out=tfc._nlist_buf, force_out=tfc._forces[off:hi],
                        virial_out=tfc.virial6_rows()[off:hi] if self.virial else None)


class LJRDF(SimModel):
    """build_examples.py:297-314: LJ forces + running mean of the RDF over [3, 5]."""

    def setup(self, r_range=(3.0, 5.0), nbins=100):
        self.avg_rdf = MeanTensor()
        self.r_range, self.nbins = tuple(r_range), nbins

    def compute(self, nlist, positions, box):
        rdf, rs = compute_rdf(nlist, self.r_range, positions[:, 3], nbins=self.nbins)
        self.avg_rdf.update_state(rdf)
        return ops.lj_forces(nlist)


class LJTypedModel(SimModel):
    """build_examples.py:80-101: typed RDFs A->B and B->A next to (scaled-down) LJ forces."""

    def setup(self):
        self.avg_rdfa = MeanTensor()
        self.avg_rdfb = MeanTensor()

    def compute(self, nlist, positions, box):
        forces = ops.lj_forces(nlist) * (1e-10 / 2.0)
        rdfa, rs = compute_rdf(nlist, [0, 10], positions[:, 3], type_i=0, type_j=1)
        rdfb, rs = compute_rdf(nlist, [0, 10], positions[:, 3], type_i=1, type_j=0)
        self.avg_rdfa.update_state(rdfa)
        self.avg_rdfb.update_state(rdfb)
        return forces


# ---- literal restatements on autograd (user-model style) ----
class LJModelAutograd(SimModel):
    def compute(self, nlist, positions, box):
        rinv = nlist_rinv(nlist)
        inv_r6 = rinv ** 6
        p_energy = 4.0 / 2.0 * (inv_r6 * inv_r6 - inv_r6)
        energy = p_energy.sum(dim=1)
        return compute_nlist_forces(nlist, energy)


class LJVirialModelAutograd(SimModel):
    def compute(self, nlist, positions, box):
        rinv = nlist_rinv(nlist)
        inv_r6 = rinv ** 6
        p_energy = 4.0 / 2.0 * (inv_r6 * inv_r6 - inv_r6)
        energy = p_energy.sum(dim=1)
        return compute_nlist_forces(nlist, energy, virial=True)


class SimplePotential(SimModel):
    """build_examples.py:9-24: F = -sum_j d/|d| (no energy)."""

    def compute(self, nlist, positions):
        nlist = nlist[:, :, :3]
        neighs_rs = torch.linalg.norm(nlist, dim=2, keepdim=True)
        fr = -1.0 * (1.0 / neighs_rs) * nlist
        real_fr = torch.where(torch.isfinite(fr), fr, torch.zeros_like(nlist))
        return real_fr.sum(dim=1)


class BenchmarkPotential(SimModel):
    """build_examples.py:27-32."""

    def compute(self, nlist):
        rinv = nlist_rinv(nlist)
        return compute_nlist_forces(nlist, rinv)


class NoForceModel(SimModel):
    """build_examples.py:35-43."""

    def compute(self, nlist, positions):
        neighs_rs = torch.linalg.norm(nlist[:, :, :3], dim=2)
        energy = torch.where(neighs_rs == 0, torch.zeros_like(neighs_rs), 1.0 / neighs_rs)
        pos_norm = torch.linalg.norm(positions, dim=1)
        return energy, pos_norm


class WrapModel(SimModel):
    """build_examples.py:52-59."""

    def compute(self, nlist, positions, box):
        return wrap_vector(positions[0, :3] - positions[-1, :3], box)


class BenchmarkNonlistModel(SimModel):
    """build_examples.py:62-66."""

    def compute(self, nlist, positions, box):
        ps = torch.linalg.norm(positions, dim=1)
        energy = torch.where(ps == 0, torch.zeros_like(ps), 1.0 / ps)
        return compute_positions_forces(positions, energy)


class EDSModel(SimModel):
    """build_examples.py:118-135: EDS bias on the distance of particle 0 from the origin."""

    def setup(self, set_point):
        self.cv_avg = Mean()
        self.eds_bias = EDSLayer(float(set_point), 5, 1 / 5)

    def compute(self, nlist, positions, box):
        rvec = wrap_vector(positions[0, :3], box)
        cv = torch.linalg.norm(rvec)
        self.cv_avg.update_state(cv)
        alpha = self.eds_bias(cv)
        energy = (cv - 5) ** 2 + cv * alpha
        forces = compute_positions_forces(positions, energy)
        return forces, alpha


class MappedNlist(SimModel):
    """build_examples.py:183-196."""

    @staticmethod
    def my_map(pos, box):
        x = pos[:, :3].mean(dim=0, keepdim=True)
        cg1 = torch.cat((x, torch.zeros((1, 1), dtype=x.dtype, device=x.device)), -1)
        cg2 = torch.tensor([[0, 0, 0.1, 1]], dtype=x.dtype, device=x.device)
        return torch.cat((cg1, cg2), dim=0)

    def compute(self, nlist, positions, box):
        nlist, cnlist = self.mapped_nlist(nlist)
        return positions, nlist, cnlist


class NlistNN(SimModel):
    """build_examples.py:199-218 / examples/08: sorted 1/r of the nearest neighbors -> 3 Dense layers."""

    def setup(self, dim, top_neighs):
        self.dense1 = torch.nn.Linear(top_neighs, dim)
        self.dense2 = torch.nn.Linear(dim, dim)
        self.last = torch.nn.Linear(dim, 1)
        self.top_neighs = top_neighs

    def compute(self, nlist, positions, box):
        rinv = nlist_rinv(nlist)
        top_n = torch.sort(rinv, dim=1, descending=True).values[:, :self.top_neighs]
        x = self.dense1(top_n.reshape(-1, self.top_neighs))
        x = self.dense2(x)
        energy = self.last(x)
        return compute_nlist_forces(nlist, energy)


class TrainModel(SimModel):
    """build_examples.py:246-268."""

    def setup(self, dim, top_neighs):
        self.dense1 = torch.nn.Linear(top_neighs, dim)
        self.dense2 = torch.nn.Linear(dim, dim)
        self.last = torch.nn.Linear(dim, 1)
        self.top_neighs = top_neighs
        self.output_zero = False

    def compute(self, nlist, positions, training):
        rinv = nlist_rinv(nlist)
        top_n = torch.sort(rinv, dim=1, descending=True).values[:, :self.top_neighs]
        x = self.dense2(self.dense1(top_n))
        energy = self.last(x)
        if training:
            energy = energy * 2
        forces = compute_nlist_forces(nlist, energy)
        if self.output_zero:
            energy = energy * 0.0
        return forces, energy.sum()


class WCA(SimModel):
    """build_examples.py:221-228: trainable WCA repulsion layer -> nlist forces."""

    def setup(self):
        from .layers import WCARepulsion
        self.wca = WCARepulsion(0.5)

    def compute(self, nlist):
        energy = self.wca(nlist)
        return compute_nlist_forces(nlist, energy)


class RBF(SimModel):
    """build_examples.py:231-241: per-pair radial basis features -> Dense(1)."""

    def setup(self, low, high, count):
        self.rbf = RBFExpansion(low, high, count)
        self.dense = torch.nn.Linear(count, 1)

    def compute(self, nlist):
        r = safe_norm(nlist[:, :, :3], axis=2)
        energy = self.dense(self.rbf(r)).sum()
        return compute_nlist_forces(nlist, energy)


class EDSCoordinationModel(SimModel):
    """BASELINE config 5: LJ fluid with an EDS bias on the smooth coordination number
    ``CV = mean_i sum_j 1/(1 + (r_ij/r0)^6)`` and a running RDF, all from ONE pass over the neighbor tensor.

    Bias energy ``alpha * CV`` (htf/layers.py EDSLayer, examples/03): force on row i is
    ``F_LJ,i + 2 alpha / N * sum_j ds/dd_ij`` (compute_nlist_forces convention, htf/simmodel.py:542-550); the
    energy column carries ``e_i + alpha * cn_i / N`` so that the total is ``E_LJ + alpha * CV``.  With row shards
    (``group`` given) the CV sum, the particle count and the histogram are all-reduced.
    """

    def setup(self, set_point, period=25, learning_rate=5.0, r0=1.3, rdf_range=None, nbins=100, group=None,
              cv_scale=1.0):
        self.eds_bias = EDSLayer(float(set_point), period, learning_rate, cv_scale)
        self.r0, self.rdf_range, self.nbins, self.group = float(r0), rdf_range, nbins, group
        self.cv_avg = Mean()
        self.avg_rdf = MeanTensor()
        self.last_bins = None

    fused_whole_shard_only = True      # the CV is a global quantity: row batches cannot be biased one by one

    def compute(self, nlist, positions, box):
        # row_counts: optional int32[rows] from the builder of THIS nlist (set by the caller that built it)
        fe, _, cv_row, cv_sum, bins = ops.lj_cv_forces(nlist, self.r0, rdf_range=self.rdf_range, nbins=self.nbins,
                                                       counts=getattr(self, "row_counts", None))
        return self._finish(nlist.shape[0], fe, cv_row, cv_sum, bins)

    def fused_rows(self, tfc, n, off, hi):
        """Whole shard in one library call: build + fused LJ/CV/RDF pass, pipelined (htf_lj_cv_step)."""
        rows, dev = hi - off, tfc._forces.device
        st = getattr(self, "_fused_state", None)
        if st is None or st[0].shape[0] != rows or st[0].device != dev:
            st = (torch.empty((rows, 4), dtype=torch.float32, device=dev),
                  torch.zeros(1, dtype=torch.float64, device=dev),
                  torch.zeros(self.nbins + 2, dtype=torch.int64, device=dev) if self.rdf_range is not None else None,
                  torch.empty((rows, 4), dtype=torch.float32, device=dev))
            self._fused_state = st
        cv_row, cv_sum, bins, fe = st
        cv_sum.zero_()
        if bins is not None:
            bins.zero_()
        tfc.ctx.lj_cv_step(n, self.r0, cv_row, cv_sum, off, hi, nlist_out=tfc._nlist_buf, force_out=fe, bins=bins,
                           r_range=self.rdf_range if self.rdf_range is not None else (0.0, 1.0), nbins=self.nbins)
        forces, alpha, cv = self._finish(rows, fe, cv_row, cv_sum, bins, out=tfc._forces[off:hi])
        self.last_outputs = (alpha, cv)

    def _finish(self, nrows, fe, cv_row, cv_sum, bins, out=None):
        import torch.distributed as dist
        from .simmodel import rdf_from_hist
        key = (int(nrows), fe.device)
        if getattr(self, "_n_key", None) != key:                # the row count as a device scalar, made once per shape
            self._n_key, self._n_dev = key, torch.tensor([float(nrows)], dtype=torch.float64, device=fe.device)
        n = self._n_dev
        if self.group is not None or (dist.is_available() and dist.is_initialized() and self.group is not False):
            g = self.group if self.group not in (None, False) else None
            if dist.is_initialized() and dist.get_world_size(g) > 1:
                packed = torch.cat([cv_sum, n])
                dist.all_reduce(packed, group=g)
                cv_sum, n = packed[:1], packed[1:]
                if bins is not None:
                    bins = bins.clone()                 # keep the rank-local histogram buffer reusable
                    dist.all_reduce(bins, group=g)
        cv = (cv_sum / n).to(torch.float32)[0]
        self.cv_avg.update_state(cv)
        alpha = self.eds_bias(cv)
        if getattr(self, "_coef4", None) is None or self._coef4.device != fe.device:
            self._coef4 = torch.tensor([2.0, 2.0, 2.0, 1.0], dtype=torch.float32, device=fe.device)
        # one contiguous pass: (F, e) += (2 s, 2 s, 2 s, s) * (grad sums, cn),  s = alpha / N
        forces = torch.addcmul(fe, cv_row, self._coef4 * (alpha / n.to(torch.float32))[0], out=out)
        if bins is not None:
            self.last_bins = bins
            rdf, _ = rdf_from_hist(bins, self.rdf_range, self.nbins)
            self.avg_rdf.update_state(rdf)
        return forces, alpha, cv


class PairwiseMLPModel(SimModel):
    """BASELINE config 3: per-pair neural force field, 32 radial basis features -> 3 x Dense(64, tanh) -> Dense(1),
    ``e_i = 1/2 sum_j u(r_ij)`` over the non-padded slots (the per-pair form of the reference's `RBF` / `NlistNN`
    models, htf/test-py/build_examples.py:199-241).  Inference runs the fused tcgen05 kernel (bf16 operands, fp32
    accumulation); ``fused=False`` or ``training=True`` evaluates the same network with torch autograd in fp32 --
    the reference the tensor-core result is validated against, and the path online training uses."""

    def setup(self, r_cut, seed=3, fused=True):
        self.r_cut, self.fused = float(r_cut), fused
        self.rbf = RBFExpansion(0.0, float(r_cut), 32)
        self.dense1 = torch.nn.Linear(32, 64)
        self.dense2 = torch.nn.Linear(64, 64)
        self.dense3 = torch.nn.Linear(64, 64)
        self.last = torch.nn.Linear(64, 1)
        g = torch.Generator().manual_seed(int(seed))
        with torch.no_grad():
            for lin in (self.dense1, self.dense2, self.dense3, self.last):
                lin.weight.copy_(torch.randn(lin.weight.shape, generator=g) / lin.in_features ** 0.5)   # N(0, 1/fan_in)
                lin.bias.copy_(0.1 * torch.randn(lin.bias.shape, generator=g))

    def raw_parameters(self):
        """fp32 blob in the layout of include/htf_b200.h (htf_mlp_pack)."""
        return torch.cat([self.dense1.weight.reshape(-1), self.dense1.bias, self.dense2.weight.reshape(-1),
                          self.dense2.bias, self.dense3.weight.reshape(-1), self.dense3.bias,
                          self.last.weight.reshape(-1), self.last.bias]).detach().to(torch.float32).contiguous()

    def pair_energy(self, nlist):
        r = safe_norm(nlist[:, :, :3], axis=2)
        x = torch.tanh(self.dense1(self.rbf(r)))
        x = torch.tanh(self.dense2(x))
        x = torch.tanh(self.dense3(x))
        u = self.last(x)[..., 0]
        return torch.where(r > 3e-6, u, torch.zeros_like(u))

    def compute(self, nlist, positions, training):
        if self.fused and not training and nlist.is_cuda:
            return ops.mlp_forces(nlist, self.raw_parameters(), self.r_cut)
        energy = 0.5 * self.pair_energy(nlist).sum(dim=1)
        return compute_nlist_forces(nlist, energy)

    # ---- online force matching (BASELINE config 4): the Keras train_on_batch of the reference as one library call ----
    def _linears(self):
        return (self.dense1, self.dense2, self.dense3, self.last)

    def load_raw_parameters(self, raw):
        """Inverse of ``raw_parameters``: write a flat fp32 blob back into the Linear layers."""
        o = 0
        with torch.no_grad():
            for lin in self._linears():
                n = lin.weight.numel()
                lin.weight.copy_(raw[o:o + n].view_as(lin.weight)); o += n
                n = lin.bias.numel()
                lin.bias.copy_(raw[o:o + n]); o += n

    def train_on_batch(self, x, y, reset_metrics=False, n_total=None, group=None):
        """One optimizer step on MSE(forces+energy [N,4], labels [N,4]) (htf/tensorflowcompute.py:367-370).

        With the default Adam optimizer on a CUDA device this is the fused path: ``htf_mlp_train_grads`` (inference pass
        + hand-written reverse sweep through the force gradient on the tensor cores), a sum of the gradients over the
        ranks when ``torch.distributed`` is initialised (row shards; ``n_total`` = rows of all ranks), and the fused
        ``htf_adam_step``.  Any other optimizer / loss, CPU tensors or ``fused=False`` use torch autograd."""
        nlist = x[0]
        mse = self.loss is not None and len(self.loss) >= 1 and isinstance(self.loss[0], str) and \
            self.loss[0].lower() in ("meansquarederror", "mse", "mean_squared_error") and all(l is None for l in self.loss[1:])
        adam = isinstance(self.optimizer, torch.optim.Adam) and len(self.optimizer.param_groups) == 1
        if not (self.fused and mse and adam and torch.is_tensor(nlist) and nlist.is_cuda):
            return super().train_on_batch(x, y, reset_metrics=reset_metrics)
        import torch.distributed as dist
        if reset_metrics:
            for m_ in self.metrics:
                m_.reset()
        ctx = ops.default_context(nlist.device)
        nl = nlist.detach()
        if nl.dtype != torch.float32 or not nl.is_contiguous():
            nl = nl.to(torch.float32).contiguous()
        labels = y.detach().to(torch.float32).contiguous()
        raw = self.raw_parameters()
        st = getattr(self, "_fused_adam", None)
        if st is None or st[0].device != raw.device:
            st = (torch.zeros_like(raw), torch.zeros_like(raw), torch.zeros(1, dtype=torch.float32, device=raw.device))
            self._fused_adam = st
        world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        if n_total is None:
            key = (int(nl.shape[0]), world)
            if getattr(self, "_n_total_key", None) != key:          # one collective + host read per shard shape, then cached
                n_total = nl.shape[0]
                if world > 1:
                    t_ = torch.tensor([n_total], dtype=torch.int64, device=raw.device)
                    dist.all_reduce(t_, group=group)
                    n_total = int(t_.item())
                self._n_total_key, self._n_total = key, n_total
            n_total = self._n_total
        grads, pred, loss = ctx.mlp_train_grads(nl, raw, self.r_cut, labels, n_total=n_total)
        if world > 1:
            packed = torch.cat([grads, loss])
            dist.all_reduce(packed, group=group)                    # ~10.5k floats: the weight-gradient all-reduce
            grads, loss = packed[:-1].contiguous(), packed[-1:]
        pg = self.optimizer.param_groups[0]
        ctx.adam_step(raw, grads, st[0], st[1], st[2], lr=pg["lr"], beta1=pg["betas"][0], beta2=pg["betas"][1], eps=pg["eps"])
        self.load_raw_parameters(raw)
        self.last_grads, self.last_pred, self.last_loss = grads, pred, loss
        self.metrics[0].update_state(loss[0])
        return loss[0]
