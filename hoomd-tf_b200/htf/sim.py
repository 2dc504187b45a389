"""Minimal stand-in for the HOOMD-blue side of hoomd-tf (HOOMD is not a dependency here).

The reference plugs into HOOMD's integrator as a ForceCompute
(/root/reference htf/tensorflowcompute.py:180-188, htf/TensorflowCompute.cc:130).  This module
provides just enough of that host: a particle system resident on the GPU, lattice
initialisers matching ``hoomd.init.create_lattice`` for the unit cells the reference tests
use, velocity-Verlet NVE and Langevin integrators, and ``run(steps)`` which calls every
attached force compute each step (and the half-step hook of label/training mode,
htf/TensorflowCompute.h:53-71).  It is host plumbing around the hot path, not part of it.
"""
import math

import numpy as np
import torch

from . import synthetic


class Box:
    def __init__(self, lo, hi):
        self.lo = np.asarray(lo, dtype=np.float32)
        self.hi = np.asarray(hi, dtype=np.float32)

    @property
    def L(self):
        return (self.hi.astype(np.float64) - self.lo.astype(np.float64))

    @property
    def Lx(self): return float(self.L[0])

    @property
    def Ly(self): return float(self.L[1])

    @property
    def Lz(self): return float(self.L[2])

    def min_image(self, r):
        r = np.asarray(r, dtype=np.float64)
        return r - np.round(r / self.L) * self.L

    def tensor(self, device, tilt=(0.0, 0.0, 0.0)):
        """[3,3] = lo; hi; (xy, xz, yz)  (htf/TensorflowCompute.cc:272-282)."""
        return torch.tensor(np.stack([self.lo, self.hi, np.asarray(tilt, np.float32)]), dtype=torch.float32,
                            device=device)


class Snapshot:
    def __init__(self, system):
        p = system.positions.detach().cpu().numpy()
        self.box = system.box
        self.particles = type("P", (), {})()
        self.particles.N = p.shape[0]
        self.particles.position = p[:, :3].copy()
        self.particles.typeid = p[:, 3].astype(np.int64)
        self.particles.velocity = system.velocities.detach().cpu().numpy().copy()


class System:
    """Particles on one GPU: positions [N,4] (x,y,z,type as float), velocities [N,3], unit mass."""

    def __init__(self, pos, lo, hi, device="cuda", dim=3, tilt=(0.0, 0.0, 0.0)):
        self.device = torch.device(device)
        self.positions = torch.as_tensor(np.ascontiguousarray(pos, dtype=np.float32)).to(self.device)
        self.velocities = torch.zeros((self.positions.shape[0], 3), dtype=torch.float32, device=self.device)
        self.box = Box(lo, hi)
        self.tilt = tuple(float(t) for t in tilt)
        self.dim = dim
        self.forces = []            # force computes: objects with compute_forces(timestep) -> [N,4]
        self.half_step_hooks = []   # label/training mode computes (HalfStepHook)
        self.net_force = torch.zeros((self.positions.shape[0], 4), dtype=torch.float32, device=self.device)
        self.timestep = 0
        self.integrator = None
        self.map_types = set()

    @property
    def N(self):
        return self.positions.shape[0]

    def __len__(self):
        return self.N

    def take_snapshot(self):
        return Snapshot(self)

    def box_tensor(self):
        return self.box.tensor(self.device, self.tilt)

    def randomize_velocities(self, kT, seed):
        g = torch.Generator(device="cpu").manual_seed(int(seed))
        v = torch.randn((self.N, 3), generator=g, dtype=torch.float32) * math.sqrt(kT)
        if self.dim == 2:
            v[:, 2] = 0.0
        v -= v.mean(dim=0, keepdim=True)
        self.velocities = v.to(self.device)
        return self

    def compute_net_force(self):
        """Sum of all attached force computes at the current positions ([N,4]: xyz + energy)."""
        total = torch.zeros((self.N, 4), dtype=torch.float32, device=self.device)
        for f in self.forces:
            total = total + f.compute_forces(self.timestep)
        self.net_force = total
        return total

    def wrap(self):
        lo = torch.tensor(self.box.lo, device=self.device)
        L = torch.tensor(self.box.L.astype(np.float32), device=self.device)
        xyz = self.positions[:, :3]
        xyz = xyz - torch.floor((xyz - lo) / L) * L
        xyz = torch.where(xyz >= lo + L, xyz - L, xyz)          # fp32 rounding can land exactly on hi
        self.positions[:, :3] = xyz

    def run(self, steps):
        if self.integrator is None:
            raise ValueError("Must have integrator set to receive forces")
        for _ in range(int(steps)):
            self.integrator.step(self)
            self.timestep += 1


def create_lattice(unitcell, n, a=None, device="cuda"):
    """hoomd.init.create_lattice for 'sq' (2-D), 'sc', 'bcc'; ``unitcell`` = (name, a) or a name + ``a``."""
    if isinstance(unitcell, (tuple, list)):
        name, a = unitcell
    else:
        name = unitcell
    if name == "sq":
        pos, lo, hi = synthetic.square_lattice(n if np.isscalar(n) else tuple(n), a)
        return System(pos, lo, hi, device=device, dim=2)
    if name == "bcc":
        pos, lo, hi = synthetic.bcc_lattice(n if np.isscalar(n) else n[0], a)
        return System(pos, lo, hi, device=device)
    if name == "sc":
        nn = (n, n, n) if np.isscalar(n) else tuple(n)
        pos, lo, hi = synthetic.lattice_fluid(nn, 1.0 / a ** 3, seed=0, jitter=0.0)
        return System(pos, lo, hi, device=device)
    raise ValueError("unknown unit cell %r" % (name,))


def sq(a):
    return ("sq", a)


def bcc(a):
    return ("bcc", a)


def sc(a):
    return ("sc", a)


class NList:
    """Stand-in for ``hoomd.md.nlist.cell()``: the neighbor search itself is libhtf_b200's cell list;
    this object only carries the system it belongs to and the subscribed cutoffs."""

    def __init__(self, system, check_period=1):
        self.system = system
        self.check_period = check_period
        self._rcut_subscribers = []

    def subscribe(self, fn):
        self._rcut_subscribers.append(fn)

    def update_rcut(self):
        self.r_cut = max([float(f()) for f in self._rcut_subscribers] + [0.0])


def nlist_cell(system, check_period=1):
    return NList(system, check_period)


class _Integrator:
    """Velocity Verlet around the force computes.  On the GPU both half steps are libhtf_b200 kernels
    (htf_integrate_half: kick + drift + wrap, kick) so a trajectory never leaves the device; ``fused=False`` (and
    CPU tensors) use the same formulas written with torch ops -- the reference the kernels are tested against."""

    def __init__(self, dt, fused=True):
        self.dt = float(dt)
        self.fused = bool(fused)
        self._have_force = False
        self._ctx = None

    def _forces(self, system):
        return system.compute_net_force()[:, :3]

    def _kernel_ctx(self, system):
        if not (self.fused and system.positions.is_cuda):
            return None
        if self._ctx is None:
            from .context import HtfContext
            self._ctx = HtfContext(system.N, 1, 1.0, device=system.device)
            self._ctx.set_box(system.box.lo, system.box.hi)
        return self._ctx

    def _fused_step(self, system, ctx, gamma=0.0, kT=0.0, seed=0):
        if not self._have_force:
            self._f4 = system.compute_net_force().contiguous()
            self._have_force = True
        if not system.velocities.is_contiguous():
            system.velocities = system.velocities.contiguous()
        ctx.integrate_half(0, system.positions, system.velocities, self._f4, self.dt, gamma, kT, system.dim == 2, seed,
                           system.timestep)
        for h in system.half_step_hooks:
            h.half_step(system.timestep)
        self._f4 = system.compute_net_force().contiguous()
        ctx.integrate_half(1, system.positions, system.velocities, self._f4, self.dt, gamma, kT, system.dim == 2, seed,
                           system.timestep)


class NVE(_Integrator):
    """Velocity Verlet (hoomd.md.integrate.nve with mode_standard(dt))."""

    def step(self, system):
        dt = self.dt
        ctx = self._kernel_ctx(system)
        if ctx is not None:
            return self._fused_step(system, ctx)
        if not self._have_force:
            self._f = self._forces(system)
            self._have_force = True
        system.velocities = system.velocities + 0.5 * dt * self._f
        system.positions[:, :3] += dt * system.velocities
        if system.dim == 2:
            system.positions[:, 2] = 0.0
        system.wrap()
        for h in system.half_step_hooks:
            h.half_step(system.timestep)
        self._f = self._forces(system)
        system.velocities = system.velocities + 0.5 * dt * self._f


class Langevin(_Integrator):
    """Simple Langevin step (hoomd.md.integrate.langevin): velocity Verlet whose two half kicks each carry friction
    and an independent random impulse sized for dt/2, so that <v^2> relaxes to kT."""

    def __init__(self, dt, kT, seed, gamma=1.0, fused=True):
        super().__init__(dt, fused)
        self.kT, self.gamma = float(kT), float(gamma)
        self._gen = None
        self._seed = int(seed)

    def step(self, system):
        dt = self.dt
        ctx = self._kernel_ctx(system)
        if ctx is not None:
            return self._fused_step(system, ctx, self.gamma, self.kT, self._seed)
        if self._gen is None:
            self._gen = torch.Generator(device=system.device).manual_seed(self._seed)
        if not self._have_force:
            self._f = self._forces(system)
            self._have_force = True
        sigma = math.sqrt(4.0 * self.gamma * self.kT / dt)       # each half kick lasts dt/2 and has its own impulse

        def noise():
            z = torch.randn((system.N, 3), generator=self._gen, device=system.device, dtype=torch.float32)
            if system.dim == 2:
                z[:, 2] = 0.0
            return sigma * z
        fr = self._f - self.gamma * system.velocities + noise()
        system.velocities = system.velocities + 0.5 * dt * fr
        system.positions[:, :3] += dt * system.velocities
        if system.dim == 2:
            system.positions[:, 2] = 0.0
        system.wrap()
        for h in system.half_step_hooks:
            h.half_step(system.timestep)
        self._f = self._forces(system)
        system.velocities = system.velocities + 0.5 * dt * (self._f - self.gamma * system.velocities + noise())


class ReferenceLJ:
    """hoomd.md.pair.lj(epsilon, sigma, r_cut) stand-in used as label force (set_reference_forces) and as
    the analytic comparison of the reference tests.  Plain torch O(N^2): host-side test/label helper for
    small systems only -- not the product path."""

    def __init__(self, system, r_cut, epsilon=1.0, sigma=1.0):
        self.system, self.r_cut, self.eps, self.sigma = system, float(r_cut), float(epsilon), float(sigma)
        self.name = "lj"

    def compute_forces(self, timestep=0):
        s = self.system
        xyz = s.positions[:, :3].double()
        L = torch.tensor(s.box.L, device=s.device, dtype=torch.float64)
        d = xyz[None, :, :] - xyz[:, None, :]
        d = d - torch.round(d / L) * L
        r2 = (d ** 2).sum(-1)
        n = xyz.shape[0]
        mask = (r2 <= self.r_cut ** 2) & ~torch.eye(n, dtype=torch.bool, device=s.device)
        r2s = torch.where(mask, r2, torch.ones_like(r2))
        sr6 = (self.sigma ** 2 / r2s) ** 3
        fdivr = torch.where(mask, self.eps * (48.0 * sr6 * sr6 - 24.0 * sr6) / r2s, torch.zeros_like(r2))
        f = -(fdivr[:, :, None] * d).sum(1)
        e = 0.5 * torch.where(mask, 4.0 * self.eps * (sr6 * sr6 - sr6), torch.zeros_like(r2)).sum(1)
        self.forces = torch.cat([f, e[:, None]], dim=1).float()
        return self.forces


class PairLJ:
    """hoomd.md.pair.lj(epsilon = sigma = 1, r_cut) stand-in for LARGE systems: the library's own nlist + LJ kernels
    (HtfContext.lj_step), optionally scaled.  Used as the label force of online force matching (set_reference_forces,
    BASELINE config 4) where the O(N^2) ``ReferenceLJ`` is out of reach."""

    def __init__(self, system, r_cut, nneighbor_cutoff=64, scale=1.0):
        from .context import HtfContext
        self.system, self.r_cut, self.K, self.scale = system, float(r_cut), int(nneighbor_cutoff), float(scale)
        self.ctx = HtfContext(max(system.N, 1), self.K, self.r_cut, device=system.device)
        self.ctx.set_box(system.box.lo, system.box.hi)
        self.name = "lj"
        self.forces = None

    def compute_forces(self, timestep=0):
        s = self.system
        if self.forces is None or self.forces.shape[0] != s.N:
            self.forces = torch.empty((s.N, 4), dtype=torch.float32, device=s.device)
        self.ctx.lj_step(s.positions, force_out=self.forces)
        if self.scale != 1.0:
            self.forces.mul_(self.scale)
        return self.forces
