"""tfcompute: applies a SimModel to a running simulation.

Host-side mirror of /root/reference htf/tensorflowcompute.py (class tfcompute) fused with
the orchestration of its C++ half, ``TensorflowCompute<M>::computeForces``
(htf/TensorflowCompute.cc:130-216): period gate, row batches, neighbor tensor build,
model call, force/virial hand-over, output capture.  The device work goes through
libhtf_b200 (``HtfContext``); there is no TensorFlow, no C++ -> Python callback and no
per-batch device synchronisation.
"""
import numpy as np
import torch

from . import sim as _sim
from .context import HtfContext


class tfcompute:
    def __init__(self, model):
        self.model = model
        self.cpp_force = None
        self._nlist = None
        self.map_types = set()
        self.ctx = None
        self.system = None
        self.shard = None           # (row_lo, row_hi): this rank's particle rows when sharded over GPUs
        self._ref_forces = []
        self._map_typeid_start = None

    # ------------------------------------------------------------------ attach
    def attach(self, nlist=None, r_cut=0, period=1, batch_size=None, train=False, save_output_period=None,
               system=None):
        """htf/tensorflowcompute.py:38-188.  ``nlist`` is a ``sim.NList`` (the stand-in for
        ``hoomd.md.nlist.cell()``); ``system`` is only needed when the model takes no neighbor list."""
        self.enabled = True
        self.force_name = "tfcompute"
        r_cut = float(r_cut)
        self.r_cut = r_cut
        self.batch_size = 0 if batch_size is None else int(batch_size)
        self.save_output_period = save_output_period
        self.outputs = None
        self._calls = 0
        self.period = int(period)
        self._output_offset = 0
        if self.model.output_forces:
            self._output_offset = 1
        if self.model.virial:
            self._output_offset = 2
        if train:
            try:
                i = 0
                for i, l in enumerate(self.model.loss):
                    if l is None:
                        break
                else:
                    i = len(self.model.loss)
                self._output_offset = i
            except (AttributeError, TypeError):
                raise ValueError("SimModel has not been compiled")
        self.train = train
        self.nneighbor_cutoff = self.model.nneighbor_cutoff
        if nlist is not None:
            nlist.subscribe(self.rcut)
            nlist.update_rcut()
            self._nlist = nlist
            system = nlist.system
        elif self.nneighbor_cutoff != 0:
            raise ValueError("Must provide an nlist if you have nneighbor_cutoff > 0")
        if system is None:
            raise RuntimeError("Must initialize a system first")
        self.system = system
        self.model.to(system.device)          # parameters / layer state live next to the particles
        self.force_mode ="tf2hoomd" if self.model.output_forces else "hoomd2tf"
        n = system.N
        self.ctx = HtfContext(max(n, 1), max(1, self.nneighbor_cutoff), r_cut if r_cut > 0 else 1.0,
                              device=system.device)
        self.ctx.set_box(system.box.lo, system.box.hi)          # a skewed box is rejected per step, as in the reference
        if self.model._map_nlist:
            self.ctx.set_mapped_nlist(self._map_typeid_start)
        self.dtype = torch.float32
        self._forces = torch.zeros((n, 4), dtype=torch.float32, device=system.device)
        self._virial = torch.zeros((n, 9), dtype=torch.float32, device=system.device)
        self._virial6 = None        # [n,6] written directly by fused built-in models (what HOOMD keeps of the 3x3)
        self._nlist_buf = None
        self._nlist_store = None    # reused [batch rows, K, 4] buffer of the fused models
        self._positions_buf = None
        self._host = None           # pinned host mirrors (forces, virial6) filled batch by batch on a copy stream
        if self.force_mode == "tf2hoomd":
            system.forces.append(self)
        else:
            if system.integrator is None:
                raise ValueError("Must have integrator set to receive forces")
            system.half_step_hooks.append(self)
        self.cpp_force = self          # the reference exposes the C++ object here; everything lives on self

    def enable_mapped_nlist(self, system, mapping_fxn):
        """Append CG beads to the particle array so that they get neighbor lists of their own
        (htf/tensorflowcompute.py:198-263).  ``mapping_fxn(positions[N,4], [Lx,Ly,Lz]) -> [M,4]``."""
        aa = system.positions
        cg = mapping_fxn(aa, [system.box.Lx, system.box.Ly, system.box.Lz])
        cg = torch.as_tensor(cg, dtype=torch.float32, device=system.device)
        M, AAN = cg.shape[0], aa.shape[0]
        start = int(aa[:, 3].max().item()) + 1
        cg = cg.clone()
        cg[:, 3] += start
        for t in cg[:, 3].tolist():
            self.map_types.add(int(t))
        system.positions = torch.cat([aa, cg], dim=0).contiguous()
        system.velocities = torch.cat([system.velocities, torch.zeros((M, 3), device=system.device)], dim=0)
        system.net_force = torch.zeros((AAN + M, 4), device=system.device)
        system.map_types = set(self.map_types)
        self.model._map_nlist = True
        self.model._map_fxn = mapping_fxn
        self.model._map_i = AAN
        self._map_typeid_start = start
        if self.ctx is not None:
            self.ctx.set_mapped_nlist(start)
        return list(range(AAN)), list(range(AAN, AAN + M))

    def set_reference_forces(self, *forces):
        """Label forces for training (htf/tensorflowcompute.py:265-282, TensorflowCompute.cc:251-269)."""
        if self.force_mode == "tf2hoomd":
            raise ValueError("Only valid to set reference forces if mode is hoomd2tf")
        for f in forces:
            if not hasattr(f, "compute_forces"):
                raise ValueError("given force does not seem like a force compute")
            self._ref_forces.append(f)

    def rcut(self):
        return self.r_cut

    # ------------------------------------------------------------------ the step
    def compute_forces(self, timestep):
        """ForceCompute::compute -> TensorflowCompute::computeForces (htf/TensorflowCompute.cc:130-216)."""
        if timestep % self.period == 0:
            self._update(timestep)
        return self._forces

    def half_step(self, timestep):
        """HalfStepHookWrapper::update (htf/TensorflowCompute.h:61-64): label / training mode."""
        if timestep % self.period == 0:
            self._update(timestep)

    def _update(self, timestep):
        s = self.system
        if sum(abs(t) for t in s.tilt) >= 1e-4:
            self.ctx.set_box(s.box.lo, s.box.hi, s.tilt)        # raises: "box is skewed"
        if self.ctx.box is None or list(self.ctx.box[0]) != [float(x) for x in s.box.lo] or \
                list(self.ctx.box[1]) != [float(x) for x in s.box.hi]:
            self.ctx.set_box(s.box.lo, s.box.hi)                # the box changed since attach (updateBox runs per step)
        n = s.N
        if self._forces.shape[0] != n:
            self._forces = torch.zeros((n, 4), dtype=torch.float32, device=s.device)
            self._virial = torch.zeros((n, 9), dtype=torch.float32, device=s.device)
            self._virial6 = None
        if self.batch_size == 0 and self.model._map_nlist:
            self._start_update()
        pos = s.positions
        if self.force_mode == "hoomd2tf":
            # labels: net force, or the sum of the chosen reference forces (TensorflowCompute.cc:177-187)
            if self._ref_forces:
                lab = None
                for f in self._ref_forces:
                    v = f.compute_forces(timestep)
                    lab = v if lab is None else lab + v
                self._forces = lab.to(torch.float32)
            else:
                self._forces = s.net_force.clone()
        bs = self.batch_size if self.batch_size > 0 else max(n, 1)
        K = self.nneighbor_cutoff
        if K > 0:
            self.ctx.bin_particles(pos)                          # once per update, like m_nlist->compute (:162-163)
        batch_index = 0
        r0, r1 = self.shard if self.shard is not None else (0, n)
        bs = min(bs, max(r1 - r0, 1))
        # built-in models whose whole batch is one library call (build + pair pass, pipelined): no model call, the
        # kernels write straight into the force / virial rows
        fused = getattr(self.model, "fused_rows", None)
        saving = self.save_output_period and (self._calls + 1) % self.save_output_period == 0
        if not (fused is not None and K > 0 and self.force_mode == "tf2hoomd" and not self.train and not saving
                and getattr(self.model, "fused", True)):
            fused = None
        if fused is not None and getattr(self.model, "fused_whole_shard_only", False) and bs < r1 - r0:
            fused = None
        if fused is not None and (self._nlist_store is None or self._nlist_store.shape[0] < min(bs, r1 - r0)
                                  or self._nlist_store.shape[1] != K):
            self._nlist_store = torch.empty((max(min(bs, r1 - r0), 1), K, 4), dtype=torch.float32, device=s.device)
        if self._host is not None:
            torch.cuda.current_stream(s.device).wait_event(self._host["done"])      # last step's copies have left
        for off in range(r0, max(r1, r0 + 1), bs):
            hi = min(r1, off + bs)
            if fused is not None:
                if batch_index == 0:
                    self._calls += 1
                self._nlist_buf = self._nlist_store[:hi - off]
                self._positions_buf = pos[off:hi]
                fused(self, n, off, hi)
                if self.model.check_nlist and self.ctx.overflow() >= K:
                    raise RuntimeError("Neighbor list is full!")
            else:
                if K > 0:
                    self._nlist_buf = self.ctx.build_nlist(pos, off, hi, rebin=False)
                else:
                    self._nlist_buf = torch.zeros((1, 1, 4), dtype=torch.float32, device=s.device)
                self._positions_buf = pos[off:hi]
                self._finish_update(batch_index, off, hi)
            if self._host is not None:
                self._mirror(off, hi, r0)
            batch_index += 1
        if self._host is not None:
            self._host["done"].record(self._host["stream"])

    # ------------------------------------------------------------------ host mirrors
    def set_host_outputs(self, forces=None, virial6=None):
        """Pinned host tensors that receive this rank's force rows [rows,4] (and virial rows [rows,6]) after every
        update, batch by batch on a copy stream: the device->host copy of batch b runs under the kernels of batch
        b+1.  This is the reference's CPU-HOOMD mode, where TfToHoomd copies the force tensor into host arrays
        (htf/tf2hoomd_op/tf2hoomd.cc:48-59), without its per-batch device synchronisation.  ``host_sync()`` waits."""
        dev = self.system.device
        for t in (forces, virial6):
            if t is not None and not t.is_pinned():
                raise ValueError("host outputs must be pinned tensors")
        self._host = {"f": forces, "v": virial6, "stream": torch.cuda.Stream(device=dev),
                      "done": torch.cuda.Event(), "ready": torch.cuda.Event()}
        self._host["done"].record(torch.cuda.current_stream(dev))

    def _mirror(self, off, hi, r0):
        h = self._host
        main = torch.cuda.current_stream(self.system.device)
        h["ready"].record(main)
        h["stream"].wait_event(h["ready"])
        with torch.cuda.stream(h["stream"]):
            if h["f"] is not None:
                h["f"][off - r0:hi - r0].copy_(self._forces[off:hi], non_blocking=True)
            if h["v"] is not None:
                h["v"][off - r0:hi - r0].copy_(self.virial6((off, hi)), non_blocking=True)

    def host_sync(self):
        if self._host is not None:
            self._host["stream"].synchronize()

    def _start_update(self):
        """precompute: apply the mapping function and write the bead positions back
        (htf/simmodel.py:289-339, htf/TFArrayComm.cu:31-41 copy3)."""
        s = self.system
        i = self.model._map_i
        bs = s.box_tensor()[1] - s.box_tensor()[0]
        cg = self.model._map_fxn(s.positions[:i], bs)
        s.positions[i:, :3] = torch.as_tensor(cg, dtype=torch.float32, device=s.device)[:, :3]

    def _finish_update(self, batch_index, off=0, hi=None):
        """htf/tensorflowcompute.py:313-370."""
        if batch_index == 0:
            self._calls += 1
        s = self.system
        hi = s.N if hi is None else hi
        box = s.box_tensor()
        if self.model.check_nlist and self.nneighbor_cutoff > 0:
            # the reference counts x>0 entries (htf/simmodel.py:216-224); the kernel knows the exact count
            if self.ctx.overflow() >= self.nneighbor_cutoff:
                raise RuntimeError("Neighbor list is full!")
        inputs = [self._nlist_buf, self._positions_buf, box]
        save = self.save_output_period and self._calls % self.save_output_period == 0
        if not self.train:
            output = self.model(inputs, self.train)
            if save:
                self._save(output[self._output_offset:])
            if self.force_mode == "tf2hoomd":
                self._compute_outputs(off, hi, *output[:self._output_offset])
        else:
            labels = self._forces[off:hi]
            if save:
                output = self.model(inputs, self.train)
                self._save(output[self._output_offset:])
            self.model.train_on_batch(x=inputs, y=labels, reset_metrics=False)

    def _save(self, outs):
        vals = [o.detach().cpu().numpy()[np.newaxis, ...] for o in outs]
        if self.outputs is None:
            self.outputs = vals
        else:
            self.outputs = [np.append(o1, o2, axis=0) for o1, o2 in zip(self.outputs, vals)]

    def _compute_outputs(self, off, hi, forces, virial=None):
        """SimModel.compute_outputs + TfToHoomd (htf/simmodel.py:240-255): forces -> rows [off,hi)."""
        forces = forces.detach()
        if forces.shape[1] == 3:
            forces = torch.cat([forces, torch.zeros((forces.shape[0], 1), dtype=forces.dtype, device=forces.device)], 1)
        self._forces[off:hi] = forces.to(torch.float32)
        if virial is not None:
            self._virial[off:hi] = virial.detach().reshape(-1, 9).to(torch.float32)

    # ------------------------------------------------------------------ accessors (htf/tensorflowcompute.py:372-392)
    def get_positions_array(self):
        return self._positions_buf.detach().cpu().numpy().astype(np.float64)

    def get_nlist_array(self):
        return self._nlist_buf.detach().cpu().numpy().astype(np.float64).reshape(-1, self.nneighbor_cutoff, 4)

    def get_forces_array(self):
        return self._forces.detach().cpu().numpy().astype(np.float64)

    def get_virial_array(self):
        if self._virial6 is not None:
            return self._virial6[:, [0, 1, 2, 1, 3, 4, 2, 4, 5]].detach().cpu().numpy().astype(np.float64)
        return self._virial.detach().cpu().numpy().astype(np.float64).reshape((-1, 9))

    def virial6_rows(self):
        """The [n,6] buffer fused models write their virial into (allocated on first use)."""
        n = self._forces.shape[0]
        if self._virial6 is None or self._virial6.shape[0] != n:
            self._virial6 = torch.zeros((n, 6), dtype=torch.float32, device=self._forces.device)
        return self._virial6

    def virial6(self, rows=None):
        """Device tensor [rows, 6] = xx, xy, xz, yy, yz, zz: the six components HOOMD keeps of the 3x3 virial
        (receiveVirial, htf/TensorflowCompute.cc:285-301)."""
        if self._virial6 is not None:
            return self._virial6 if rows is None else self._virial6[rows[0]:rows[1]]
        v = self._virial if rows is None else self._virial[rows[0]:rows[1]]
        return v[:, [0, 1, 2, 4, 5, 8]]

    def get_log_value(self):
        """HOOMD log quantity "tensorflow": the potential energy sum (htf/TensorflowCompute.cc:377-395)."""
        return float(self._forces[:, 3].sum().item())

    def update_coeffs(self):
        pass
