"""Model layers of hoomd-tf on torch (/root/reference htf/layers.py): RBFExpansion (:7-49),
WCARepulsion (:52-98), EDSLayer (:101-195).  All state lives on the device; no call
synchronises the host."""
import torch

from .simmodel import nlist_rinv


class RBFExpansion(torch.nn.Module):
    """exp(-(d - mu)^2 / gap), mu = linspace(low, high, count) (htf/layers.py:27-49)."""

    def __init__(self, low, high, count):
        super().__init__()
        self.low, self.high, self.count = low, high, count
        centers = torch.linspace(float(low), float(high), count, dtype=torch.float32)
        self.register_buffer("centers", centers)
        self.gap = float(centers[1] - centers[0]) if count > 1 else 1.0

    def get_config(self):
        return {"low": self.low, "high": self.high, "count": self.count}

    def forward(self, inputs):
        return torch.exp(-(inputs[..., None] - self.centers.to(inputs.device)) ** 2 / self.gap)


class WCARepulsion(torch.nn.Module):
    """Trainable repulsion (sigma/r)^6 for r < 2^(1/3) sigma, clipped to [0, 10] (htf/layers.py:52-98)."""

    def __init__(self, sigma, regularization_strength=1e-3):
        super().__init__()
        self.sigma = torch.nn.Parameter(torch.tensor(float(sigma)))
        self.regularization_strength = regularization_strength

    def get_config(self):
        return {"sigma": float(self.sigma.detach())}

    def regularization(self):
        return -self.regularization_strength * self.sigma

    def forward(self, nlist):
        rinv = nlist_rinv(nlist)
        rp = (self.sigma * rinv) ** 6
        r = torch.linalg.norm(nlist[:, :, :3], dim=2)
        r_pair_energy = (r < self.sigma * 2 ** (1 / 3)).to(rp.dtype) * rp
        return torch.clamp(r_pair_energy, 0, 10)


class EDSLayer(torch.nn.Module):
    """EDS coupling constant alpha for a scalar collective variable (htf/layers.py:101-195).

    Statistics (Welford mean / ssd) are gathered over the second half of each ``period``; at
    ``n == period - 1`` one tf.compat.v1 Adam step (beta1 .9, beta2 .999, eps 1e-8,
    lr_t = lr sqrt(1 - b2^t) / (1 - b1^t)) moves alpha along
    ``-2 (mean - set_point) ssd / period / 2 / cv_scale``.  Masked arithmetic keeps every step on
    the device exactly like the reference's ``tf.function``.
    """

    def __init__(self, set_point, period, learning_rate=1e-2, cv_scale=1.0, name="eds-layer"):
        super().__init__()
        if isinstance(set_point, int) or (torch.is_tensor(set_point) and not set_point.is_floating_point()):
            raise ValueError("EDS only works with floats, not dtype " + str(type(set_point)))
        self.name = name
        self.fused = True               # CUDA tensors: htf_eds_step; False keeps the torch formulas (the CPU/test reference)
        self.period = int(period)
        self.cv_scale = float(cv_scale)
        self.learning_rate = float(learning_rate)
        z = lambda: torch.zeros((), dtype=torch.float32)
        self.register_buffer("set_point", torch.as_tensor(set_point, dtype=torch.float32))
        self.register_buffer("mean", z())
        self.register_buffer("ssd", z())
        self.register_buffer("n", torch.zeros((), dtype=torch.int32))
        self.register_buffer("alpha", z())
        self.register_buffer("adam_m", z())
        self.register_buffer("adam_v", z())
        self.register_buffer("adam_t", z())

    def get_config(self):
        return {"set_point": float(self.set_point), "period": self.period, "cv_scale": self.cv_scale,
                "learning_rate": self.learning_rate}

    @torch.no_grad()
    def forward(self, cv):
        cv = cv.detach().to(torch.float32)
        for b in (self.set_point, self.mean, self.ssd, self.n, self.alpha, self.adam_m, self.adam_v, self.adam_t):
            if b.device != cv.device:
                self.to(cv.device)
                break
        if cv.is_cuda and self.fused:
            # one libhtf_b200 launch on the device-resident state instead of ~50 element-wise ones
            from . import ops
            ops.default_context(cv.device).eds_step(cv.reshape(()).contiguous(), self)
            return self.alpha.clone()
        h = self.period // 2
        reset_mask = (self.n != 0).to(torch.float32)
        self.mean.mul_(reset_mask)
        self.ssd.mul_(reset_mask)
        update_mask = (self.n > h).to(torch.float32)
        delta = (cv - self.mean) * update_mask
        denom = (self.n - h).to(torch.float32)
        self.mean.add_(torch.where(denom == 0, torch.zeros_like(delta), delta / denom))     # divide_no_nan
        self.ssd.add_(delta * (cv - self.mean))
        upd = (self.n == self.period - 1).to(torch.float32)
        grad = upd * -2.0 * (self.mean - self.set_point) * self.ssd / float(self.period) / 2.0 / self.cv_scale
        # tf.compat.v1.train.AdamOptimizer.apply_gradients, executed only when upd == 1
        t_new = self.adam_t + upd
        m_new = torch.where(upd > 0, 0.9 * self.adam_m + 0.1 * grad, self.adam_m)
        v_new = torch.where(upd > 0, 0.999 * self.adam_v + 0.001 * grad * grad, self.adam_v)
        tt = torch.clamp(t_new, min=1.0)
        lr_t = self.learning_rate * torch.sqrt(1.0 - 0.999 ** tt) / (1.0 - 0.9 ** tt)
        step = lr_t * m_new / (torch.sqrt(v_new) + 1e-8)
        self.alpha.sub_(upd * step)
        self.adam_t.copy_(t_new)
        self.adam_m.copy_(m_new)
        self.adam_v.copy_(v_new)
        self.n.copy_((self.n + 1) % self.period)
        return self.alpha.clone()
