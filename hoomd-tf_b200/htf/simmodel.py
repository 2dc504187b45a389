"""SimModel and the nlist math of hoomd-tf, on torch device tensors.

Mirrors /root/reference htf/simmodel.py for the nlist -> forces+virial path:
``SimModel`` (:8-339), ``compute_nlist_forces`` (:526-555), ``compute_positions_forces``
(:492-506), ``nlist_rinv`` (:618-635), ``safe_norm`` (:581-594), ``box_size`` (:597-603),
``wrap_vector`` (:606-615), ``compute_rdf`` (:638-669), ``masked_nlist`` (:672-693).

User ``compute()`` bodies receive CUDA tensors ``nlist[N,K,4]``, ``positions[N,4]``,
``box[3,3]`` produced by libhtf_b200; arbitrary bodies differentiate with torch autograd
(the role TensorFlow plays in the reference), the built-in models in ``htf.models`` call
the fused sm_100a kernels (``lj_forces``, ``compute_rdf``) instead.
"""
import torch

from . import ops

_STATE = {"training": False}


class SimModel(torch.nn.Module):
    """The main way a model interacts with a simulation (htf/simmodel.py:8-74)."""

    def __init__(self, nneighbor_cutoff, output_forces=True, virial=False, check_nlist=False,
                 dtype=torch.float32, name="htf-model", **kwargs):
        super().__init__()
        self.nneighbor_cutoff = nneighbor_cutoff
        self.output_forces = output_forces
        self.virial = virial
        self.check_nlist = check_nlist
        self.model_dtype = dtype
        self.name = name
        self._map_nlist = False
        if SimModel.compute == self.__class__.compute:
            raise AttributeError("You must implement compute method in subclass")
        try:
            code = self.compute.__code__
            self._arg_count = code.co_argcount - 1                     # - 1 for self
            self._pass_training = "training" == code.co_varnames[self._arg_count]
            if self._pass_training:
                self._arg_count -= 1
        except AttributeError:
            raise AttributeError("SimModel child class must implement compute method, and should not implement call")
        self.batch_steps = 0
        self._running_means = []
        self.loss = None
        self.optimizer = None
        self.metrics = []
        self.setup(**kwargs)

    def get_config(self):
        return {"nneighbor_cutoff": self.nneighbor_cutoff, "output_forces": self.output_forces,
                "virial": self.virial, "check_nlist": self.check_nlist, "name": self.name,
                "dtype": self.model_dtype}

    def compute(self, nlist, positions, box, training=True):
        raise AttributeError("You must implement compute in your subclass")

    def setup(self, **kwargs):
        pass

    def retrace_compute(self):
        """No tracing compiler here (htf/simmodel.py:147-163 re-traces a tf.function): nothing to do."""
        return None

    # -- Keras ``model(inputs, training)`` --
    def forward(self, inputs, training=False):
        inputs = list(inputs)
        # forces come from gradients w.r.t. the nlist / positions (htf/simmodel.py:505,542); tf.gradients works for
        # any model there (output_forces or not, training or not), so the two inputs are always differentiable leaves
        for i in (0, 1):
            if i < len(inputs) and torch.is_tensor(inputs[i]) and inputs[i].is_floating_point():
                inputs[i] = inputs[i].detach().requires_grad_(True)
        prev = _STATE["training"]
        _STATE["training"] = bool(training)
        try:
            with torch.enable_grad():
                if self._pass_training:
                    out = self.compute(*inputs[:self._arg_count], training)
                else:
                    out = self.compute(*inputs[:self._arg_count])
        finally:
            _STATE["training"] = prev
        if torch.is_tensor(out):
            out = (out,)
        return tuple(out)

    call = forward

    # -- mapped nlist (htf/simmodel.py:257-287) --
    def mapped_nlist(self, nlist):
        if not self._map_nlist:
            raise ValueError("You must call tfcompute.enable_mapped_nlist before using mapped_nlist")
        return nlist[:self._map_i], nlist[self._map_i:]

    def mapped_positions(self, positions):
        if not self._map_nlist:
            raise ValueError("You must call tfcompute.enable_mapped_nlist before using mapped_nlist")
        return positions[:self._map_i], positions[self._map_i:]

    # -- the slice of the Keras training API that tfcompute uses (htf/tensorflowcompute.py:88-95,367-370) --
    def compile(self, optimizer=None, loss=None, lr=1e-3):
        """``loss``: a name / callable or a list with one entry per model output (``None`` = not trained)."""
        if optimizer is None or optimizer == "Adam" or optimizer == "adam":
            params = [p for p in self.parameters() if p.requires_grad]
            optimizer = torch.optim.Adam(params, lr=lr, eps=1e-7) if params else None    # Keras defaults
        self.optimizer = optimizer
        self.loss = list(loss) if isinstance(loss, (list, tuple)) else [loss]
        self.metrics = [_MeanMetric()]
        return self

    def train_on_batch(self, x, y, reset_metrics=False):
        """One optimizer step on loss(model(x)[0], y) -- Keras MSE = mean over all elements."""
        if self.loss is None:
            raise ValueError("SimModel has not been compiled")
        if reset_metrics:
            for m in self.metrics:
                m.reset()
        out = self.forward(x, training=True)
        total = None
        for o, l in zip(out, self.loss):
            if l is None:
                continue
            fn = _resolve_loss(l)
            val = fn(o, y.to(o.dtype))
            total = val if total is None else total + val
        if total is None:
            raise ValueError("no trainable output (all losses are None)")
        if self.optimizer is not None and total.requires_grad:
            self.optimizer.zero_grad(set_to_none=True)
            total.backward()
            self.optimizer.step()
        self.metrics[0].update_state(total.detach())
        return total.detach()


class _MeanMetric:
    """keras.metrics.Mean stand-in that stays on the device (no per-step D2H)."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.total, self.count = None, 0

    def update_state(self, v):
        v = v.detach() if torch.is_tensor(v) else torch.as_tensor(v)
        self.total = v.clone() if self.total is None else self.total + v
        self.count += 1

    def result(self):
        if self.total is None:
            return torch.tensor(0.0)
        return self.total / self.count


class MeanTensor(_MeanMetric):
    """keras.metrics.MeanTensor stand-in (running mean of a tensor, e.g. an RDF)."""


Mean = _MeanMetric


def _resolve_loss(l):
    if callable(l):
        return l
    if isinstance(l, str) and l.lower() in ("meansquarederror", "mse", "mean_squared_error"):
        return lambda pred, target: torch.mean((pred - target) ** 2)
    raise ValueError("unknown loss %r" % (l,))


# ---------------------------------------------------------------------------------------
# force math
# ---------------------------------------------------------------------------------------
def _add_energy(forces, energy):
    """htf/simmodel.py:558-578."""
    if energy.dim() > 1:
        energy = energy.reshape(energy.shape[0], -1).sum(dim=1, keepdim=True)
        return torch.cat([forces[:, :3], energy.to(forces.dtype)], dim=-1)
    if energy.dim() == 0:
        return torch.cat([forces[:, :3], energy.reshape(1, 1).expand(forces.shape[0], 1).to(forces.dtype)], dim=-1)
    return torch.cat([forces[:, :3], energy.reshape(forces.shape[0], 1).to(forces.dtype)], dim=-1)


def _compute_virial(nlist, nlist_forces):
    """htf/simmodel.py:509-523: -sum_j |F_ij| / (2 r_ij) r_ij (x) r_ij  ->  [N,3,3]."""
    nlist3 = nlist[:, :, :3]
    rij_outer = torch.einsum("ijk,ijl->ijkl", nlist3, nlist3)
    r_mag = torch.linalg.norm(nlist3, dim=2)
    f_mag = torch.linalg.norm(nlist_forces, dim=2)
    f_rs = torch.where(r_mag == 0, torch.zeros_like(f_mag), f_mag / (2.0 * r_mag))     # divide_no_nan
    return -1.0 * torch.einsum("ij,ijkl->ikl", f_rs, rij_outer)


def compute_nlist_forces(nlist, energy, virial=False):
    """Pairwise forces [N,4] (xyz + per-particle energy) from an energy that depends on the nlist
    (htf/simmodel.py:526-555): F_i = sum_j 2 dE/d nlist[i,j,:]."""
    if not (torch.is_tensor(energy) and energy.requires_grad):
        raise ValueError("Could not find dependence between energy and nlist. Did you put them in wrong order?")
    grads = torch.autograd.grad(energy.sum(), nlist, create_graph=_STATE["training"], allow_unused=True)[0]
    if grads is None:
        raise ValueError("Could not find dependence between energy and nlist. Did you put them in wrong order?")
    nlist_forces = grads * 2.0
    nlist_reduce = nlist_forces.sum(dim=1)
    if virial:
        return _add_energy(nlist_reduce, energy), _compute_virial(nlist, nlist_forces)
    return _add_energy(nlist_reduce, energy)


def compute_positions_forces(positions, energy):
    """F = -dE/d positions, [N,4] with the energy column (htf/simmodel.py:492-506)."""
    grads = torch.autograd.grad(energy.sum(), positions, create_graph=_STATE["training"], allow_unused=True)[0]
    if grads is None:
        raise ValueError("Could not find dependence between energy and positions")
    return _add_energy(-grads, energy)


def safe_norm(tensor, delta=1e-7, **kwargs):
    """tf.norm(tensor + delta) (htf/simmodel.py:581-594); ``axis`` is accepted like in TF."""
    dim = kwargs.pop("axis", kwargs.pop("dim", None))
    return torch.linalg.norm(tensor + delta, dim=dim, **kwargs)


def box_size(box):
    """hi - lo of the [3,3] box tensor (htf/simmodel.py:597-603)."""
    return box[1, :] - box[0, :]


def wrap_vector(r, box):
    """Minimum image of r: r - round(r / L) L, round half to even like TF (htf/simmodel.py:606-615)."""
    bs = box_size(box)
    return r - torch.round(r / bs) * bs


def nlist_rinv(nlist):
    """N x NN tensor of 1/r, zero for empty neighbors, with the reference's exact offsets
    (htf/simmodel.py:618-635): r~ = ||d + 1e-7||, 1/(r~ + 3e-6) where r~ > 3e-6."""
    delta = 3e-6
    r = safe_norm(nlist[:, :, :3], axis=2, delta=delta / 3 / 10)
    return torch.where(r > delta, 1.0 / (r + delta), torch.zeros_like(r))


def masked_nlist(nlist, type_tensor, type_i=None, type_j=None):
    """htf/simmodel.py:672-693: keep rows of type_i (boolean mask), zero entries whose neighbor type != type_j."""
    if type_i is not None:
        nlist = nlist[type_tensor == type_i]
    if type_j is not None:
        mask = (nlist[:, :, 3] == type_j).to(nlist.dtype)
        nlist = nlist * mask[:, :, None]
    return nlist


def compute_rdf(nlist, r_range, type_tensor=None, nbins=100, type_i=None, type_j=None):
    """Un-normalised pairwise RDF (htf/simmodel.py:638-669): returns (rdf[nbins], bin centres[nbins]).

    The histogram over nbins+2 bins runs in the sm_100a kernel (bit-exact with the reference's
    tf.histogram_fixed_width rule); the first and last bin are dropped (:668).
    """
    hist = ops.rdf_hist(nlist, r_range, nbins=nbins, type_tensor=type_tensor, type_i=type_i, type_j=type_j)
    return rdf_from_hist(hist, r_range, nbins)


_RDF_SHELLS = {}


def rdf_from_hist(hist, r_range, nbins=100):
    key = (float(r_range[0]), float(r_range[1]), int(nbins), hist.device)
    cached = _RDF_SHELLS.get(key)
    if cached is None:                      # shell volumes / bin centres of a range are made once per device
        shell_rs = torch.linspace(key[0], key[1], nbins + 1, dtype=torch.float32, device=hist.device)
        vis_rs = (shell_rs[1:] + shell_rs[:-1]) * 0.5
        vols = shell_rs[1:] ** 3 - shell_rs[:-1] ** 3
        cached = _RDF_SHELLS[key] = (vols, vis_rs)
    vols, vis_rs = cached
    return hist[1:-1].to(torch.float32) / vols, vis_rs
