#!/usr/bin/env python
"""bench.py -- particle-timesteps/s of the nlist -> forces+virial hot path (BASELINE.json metric).

One "step" = one pass of the path over one synthetic LJ fluid: [halo exchange,] cell binning, padded neighbor
tensor [N,K,4], LJ forces + per-particle energy + 6-component virial.  Default workload:
1,048,576 particles, K=64, r_cut=2.5, rho=0.7 (BASELINE.json configs[2] geometry with the LJ
model, the size the north-star target is quoted on); the 1 GiB neighbor tensor is larger than
L2, so no flush is needed between steps.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg3|cfg2|cfg1|cfg5] [--model lj|mlp|eds] [--rdf]
  python bench.py --impl reference ...      # the CPU restatement (oracle/) on the host cores

N > 1 is launched by torchrun (one rank per GPU, NCCL): rows are sharded by particle index (z-slabs), the two slab
faces are exchanged every step, the RDF histogram is all-reduced when --rdf is on.  The headline record is weak
scaling (N x the config's particles); the same line carries a strong-scaling record of the named size under
"strong".  Every record carries a "parity" block: each rank checks a slice of its rows against the CPU oracle
outside the timed region and the ranks all-reduce the size-independent sums.  Rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "hoomd-tf_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "particle-timesteps/s (nlist->forces+virial)"
UNIT = "particle-timesteps/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=["cfg1", "cfg2", "cfg3", "cfg4", "cfg5", "refbench"],
                    help="cfg1..cfg5: BASELINE.json configs (default cfg3 = 1M particles, K=64, the size the target is "
                         "quoted on); refbench: the reference's own pytest-benchmark (htf/test-py/benchmark.py: N=256 2-D "
                         "lattice, NN=64, r_cut=3, Langevin, 4000 + 1000 steps, whole simulation) through the public API")
    ap.add_argument("--model", default="lj", choices=["lj", "mlp", "eds", "mlp-train"],
                    help="lj: closed-form LJ + virial (the headline path); mlp: BASELINE config 3's pairwise-MLP force "
                         "field on the tensor cores (forces + energy); eds: BASELINE config 5's EDS bias on the smooth "
                         "coordination-number CV + 100-bin RDF + LJ in one fused pass (use with --workload cfg5); mlp-train: "
                         "BASELINE config 4's online force matching -- every step trains the pairwise MLP on LJ label forces: "
                         "inference pass, hand-written reverse sweep through the force gradient, gradient all-reduce, fused "
                         "Adam (use with --workload cfg4)")
    ap.add_argument("--skin", type=float, default=0.0,
                    help="> 0: buffered neighbor lists like HOOMD's r_buff (the reference's own split): search with "
                         "r_cut + skin every --rebuild-every steps, distance filter every step; the particles then "
                         "move ballistically each step (|v| dt = --step-length) so that rebuilds are really needed")
    ap.add_argument("--rebuild-every", type=int, default=10)
    ap.add_argument("--step-length", type=float, default=0.0087,
                    help="displacement per step in the --skin workload (LJ liquid at T*=1, dt=0.005: 0.005*sqrt(3))")
    ap.add_argument("--no-counts", action="store_true",
                    help="the pair pass reads all K slots of every row (default: the builder hands it per-row neighbor "
                         "counts and it reads only the valid slots)")
    ap.add_argument("--no-graph", action="store_true",
                    help="launch every kernel from the host instead of replaying the three CUDA graphs "
                         "(binning | build | forces) the step is captured into")
    ap.add_argument("--phase-every", type=int, default=10,
                    help="every n-th step of the timed region runs phase by phase (one graph per phase, CUDA events between: "
                         "the per-kernel times of the roofline block); the other steps are one graph replay. 1 = every step")
    ap.add_argument("--rdf", action="store_true", help="fuse the 100-bin compute_rdf histogram into the force pass")
    ap.add_argument("--exchange", default="halo", choices=["halo", "allgather"],
                    help="N>1: slab halo exchange (send/recv of the two faces) or all-gather of every position")
    ap.add_argument("--scaling", default="both", choices=["both", "weak", "strong"],
                    help="N>1: the headline record is weak scaling (N x the config's particles); 'both' adds a strong-scaling "
                         "record (the named size split over the N GPUs) under the key 'strong'")
    ap.add_argument("--shuffle", action="store_true",
                    help="single GPU: random particle order instead of the generator's spatially coherent lattice order")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle-checked parity leg (it is on by default)")
    ap.add_argument("--transport", default="auto", choices=["auto", "p2p", "nccl"],
                    help="halo exchange transport: p2p = the library's peer-memory windows (fused pack + send over NVLink, "
                         "flags, no NCCL on the data path), nccl = pack kernels + NCCL send/recv, auto = p2p when the "
                         "peers can be mapped")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-batches", type=int, default=4,
                    help="row batches per step in the end-to-end leg (device->host copy of batch b overlaps batch b+1)")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md, MEASURED_PEAKS.json absent)"


MLP_FLOP_PER_PAIR = 2 * 2 * (32 * 64 + 64 * 64 + 64 * 64 + 64)     # value and tangent chains, multiply+add
# training step, per valid pair: the forward chains (once), the input-gradient chains back to h1 (two 64x64 layers, value
# and tangent adjoints) and the weight gradients of all three layers (both adjoints) + the last layer
MLP_TRAIN_FLOP_PER_PAIR = MLP_FLOP_PER_PAIR + 2 * 2 * (2 * 64 * 64) + 2 * 2 * (32 * 64 + 64 * 64 + 64 * 64) + 2 * 2 * 64


def tensor_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["bf16_tflops_sustained"]), "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"
    return 1500.0, "fallback (B200_PROFILING.md, MEASURED_PEAKS.json absent)"


def ncu_traffic(args, rows, K):
    """dram bytes per launch of the dominant kernel from the committed ncu capture (profiles/) -- only when that capture
    was taken on THIS configuration (same rows per launch and K), else None."""
    path = os.path.join(ROOT, "profiles", "nlist_build_traffic.json")
    if os.path.exists(path):
        with open(path) as f:
            t = json.load(f)
        if int(t.get("rows", -1)) == int(rows) and int(t.get("K", -1)) == int(K) and args.model != "mlp":
            return t
    return None


def ncu_force_traffic(args, rows, K, counts):
    """Same for the LJ forces+virial pass (profiles/pair_pass_traffic.json: taken with the per-row counts)."""
    path = os.path.join(ROOT, "profiles", "pair_pass_traffic.json")
    if os.path.exists(path) and args.model == "lj" and not args.rdf:
        with open(path) as f:
            t = json.load(f)
        if int(t.get("rows", -1)) == int(rows) and int(t.get("K", -1)) == int(K) and bool(t.get("counts")) == bool(counts):
            return t
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 20 ms while the benchmark runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(suffix=".csv")
        self.f = open(self.path, "w")
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "20"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None
        t0 = time.time()
        while self.p is not None and self.lines() == 0 and time.time() - t0 < 5.0:
            time.sleep(0.02)                 # nvidia-smi needs a moment before its first sample

    def lines(self):
        try:
            with open(self.path) as f:
                return sum(1 for _ in f)
        except OSError:
            return 0

    def stop(self, first_line=0, last_line=None):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        time.sleep(0.05)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.close()
        rows = []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        with open(self.path) as f:
            for line in f.read().splitlines():
                c = [x.strip() for x in line.split(",")]
                if len(c) < 9:
                    continue
                try:
                    rows.append((float(c[1]), float(c[2]), float(c[3]),
                                 [nme for nme, v in zip(names, c[5:9]) if v.lower().startswith("active")]))
                except ValueError:
                    continue
        os.unlink(self.path)
        window = rows[first_line:last_line] if last_line is not None else rows[first_line:]
        label = "timed region"
        if len(window) < 3:                  # very short timed region: use every sample taken under load
            window, label = rows, "warm-up + timed region"
        if window:
            sm = sorted(r[0] for r in window)
            out = {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(r[1] for r in window),
                   "power_w_max": max(r[2] for r in window),
                   "reasons": sorted({x for r in window for x in r[3]}), "samples": len(window), "window": label}
        return out


def workload(args, world, scaling="weak"):
    """Synthetic system of the named config.  weak: one config-size slab per GPU (the box grows along z);
    strong: the named size itself, split over the GPUs."""
    from htf import synthetic
    c = dict(synthetic.CONFIGS[args.workload])
    sites = list(c["sites"])
    if scaling == "weak":
        sites[2] *= world
    pos, lo, hi = synthetic.lattice_fluid(tuple(sites), c["rho"], c["seed"])
    if args.shuffle:
        import numpy as np
        pos = pos[np.random.default_rng(0).permutation(pos.shape[0])]
    return pos, lo, hi, c["r_cut"], c["K"]


# --------------------------------------------------------------------------------------------
def run_reference(args):
    """The reference arm: the CPU restatement of the reference's path (oracle/), all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import numpy as np
    import oracle
    oracle.build()
    threads = oracle.set_threads(0)        # torchrun exports OMP_NUM_THREADS=1: take every host core explicitly
    world = max(1, args.gpus)
    pos_g, lo, hi, r_cut, K = workload(args, world)          # the b200 arm's system at this N (weak scaling)
    n_global = pos_g.shape[0]
    rows = min(n_global, 2048 if args.model == "mlp-train" else 32768 if args.model == "mlp" else 262144)   # bounded sample
    g0 = (n_global // world - rows) // 2
    pos, a0 = slab_with_halo(pos_g, g0, g0 + rows, r_cut)    # what one rank of a row-sharded CPU run would hold
    n = pos.shape[0]
    if args.model in ("mlp", "mlp-train"):
        raw = mlp_raw_parameters()

    def step():
        nl, _, _ = oracle.nlist(pos, lo, hi, r_cut, K, a0, a0 + rows, cells=True, want_idx=False)
        if args.model == "mlp-train":
            return cpu_train_step(oracle, nl, raw, r_cut)
        if args.model == "mlp":
            return oracle.pairwise_mlp(nl, raw, r_cut)
        fe, _, v6 = oracle.lj(nl, virial=True)
        if args.rdf:
            oracle.rdf_hist(nl, (0.0, r_cut), 100)
        return fe

    warm = max(args.warmup, 3)
    for _ in range(warm):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = rows * args.steps / dt
    sample = ("%d-row slab of the %d-particle system per step (cell binning of the slab and its r_cut halo, %d particles, "
              "included)" % (rows, n_global, n))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": warm, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfg_dict(args, n_global, K, r_cut, world),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "host_cores": os.cpu_count()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "CPU restatement of TensorflowCompute::prepareNeighbors + closed-form LJModel (oracle/), "
                "not HOOMD+TensorFlow: the reference stack cannot be built or imported here",
    }
    print(json.dumps(line))
    return 0


def slab_with_halo(pos, a0, b0, r_cut, skin=0.4):
    """Rows [a0, b0) of a z-sorted system plus every particle within r_cut + skin of them in z (periodic images are
    not needed for an interior slab).  Returns (subset positions, index of row a0 inside the subset)."""
    import numpy as np
    z = pos[:, 2]
    zlo, zhi = float(z[a0:b0].min()) - r_cut - skin, float(z[a0:b0].max()) + r_cut + skin
    keep = np.nonzero((z >= zlo) & (z <= zhi))[0]
    if keep.size == 0 or keep[0] > a0 or keep[-1] < b0 - 1 or args_shuffled(pos, keep, a0, b0):
        return pos, a0                                            # not a contiguous slab (e.g. shuffled order): whole system
    return np.ascontiguousarray(pos[keep]), int(np.searchsorted(keep, a0))


def args_shuffled(pos, keep, a0, b0):
    import numpy as np
    i0 = int(np.searchsorted(keep, a0))
    return not np.array_equal(keep[i0:i0 + (b0 - a0)], np.arange(a0, b0))


def mlp_raw_parameters(seed=3):
    """Random-init weights of the config-3 architecture (N(0, 1/fan_in), biases 0.1 N(0,1)) as the raw blob."""
    import numpy as np
    rng = np.random.default_rng(seed)
    parts = []
    for fan_out, fan_in in ((64, 32), (64, 64), (64, 64), (1, 64)):
        parts.append((rng.standard_normal((fan_out, fan_in)) / np.sqrt(fan_in)).astype(np.float32).ravel())
        parts.append((0.1 * rng.standard_normal(fan_out)).astype(np.float32))
    return np.concatenate(parts)


def cfg_dict(args, n, K, r_cut, world):
    return {"workload": "lj_fluid_%s%s%s" % (args.workload, "+rdf100" if args.rdf else "", "+pairwise_mlp" if args.model == "mlp" else "+pairwise_mlp_training" if args.model == "mlp-train" else ""),
            "particles": n, "particles_per_gpu": n // world, "nneighbor_cutoff": K, "r_cut": r_cut,
            "model": ("pairwise MLP: RBF(32) -> 3 x Dense(64, tanh) -> Dense(1), bf16 operands / fp32 accumulation (tcgen05)"
                      if args.model == "mlp" else
                      "online force matching: pairwise MLP (RBF(32) -> 3 x Dense(64, tanh) -> Dense(1)) trained every step on "
                      "0.05 x LJ label forces, MSE over [N,4] + Adam(1e-3): inference pass (tcgen05) + reverse sweep through the "
                      "force gradient (warp MMAs, bf16 / fp32 accumulation) + fused Adam"
                      if args.model == "mlp-train" else
                      "LJ + EDS bias (period 25, lr 5.0) on the smooth coordination CV (r0 1.3) + 100-bin RDF, one fused pass"
                      if args.model == "eds" else "LJ (nlist_rinv closed form) + 6-component virial"),
            "sharding": ("particle rows (z-slabs), %s per step" % ("halo exchange of the two slab faces"
                         if args.exchange == "halo" else "all-gather of all positions")) if world > 1 else "single GPU",
            "l2": "per-GPU working set %.0f MiB/step > 126 MB L2, no flush needed" % (n // world * K * 16 / 2 ** 20)}


# --------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(local) if world > 1 else None
    globals()["_NUMA_NOTE"] = numa
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    env = dict(world=world, rank=rank, local=local, dev=dev, dist=dist, numa=numa)

    line = measure(args, env, "weak", full=True)           # the headline record: one config-size slab per GPU
    if world > 1 and args.scaling in ("both", "strong") and args.skin <= 0.0:
        st = measure(args, env, "strong", full=False)      # the named size itself, split over the GPUs
        if rank == 0:
            line["strong"] = {k: st[k] for k in ("value", "ms_per_step", "exchange_ms", "per_rank", "exchange", "parity", "config") if k in st}
            line["strong"]["efficiency_inputs"] = {
                "particles": st["config"]["particles"], "n_gpus": world, "ms_per_step": st["ms_per_step"],
                "note": "strong-scaling efficiency = value(N) / (N x value(1)) with value(1) from the --gpus 1 line"}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def bind_to_gpu_numa_node(gpu_index):
    """Pin this rank to the CPUs NVML reports as local to its GPU, so that the pinned host buffers of the end-to-end leg
    (allocated afterwards, first touch) and the copy submissions sit on the GPU's own NUMA node / PCIe root.  torchrun
    does not bind ranks.  Returns a short description, or None when NVML is not available."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(gpu_index))
        pynvml.nvmlDeviceSetCpuAffinity(h)
        cpus = sorted(os.sched_getaffinity(0))
        return "NVML ideal CPU affinity: %d CPUs (%d..%d)" % (len(cpus), cpus[0], cpus[-1])
    except Exception as ex:                                  # noqa: BLE001 -- optional tuning, reported in the record
        return "not bound (%s)" % type(ex).__name__


def measure(args, env, scaling, full):
    """One timed case.  Returns the JSON record (complete on rank 0)."""
    import numpy as np
    import torch
    import htf
    world, rank, local, dev, dist = env["world"], env["rank"], env["local"], env["dev"], env["dist"]

    pos, lo, hi, r_cut, K = workload(args, world, scaling)
    n = pos.shape[0]
    per = n // world
    row_lo, row_hi = rank * per, (rank + 1) * per if rank < world - 1 else n
    rows = row_hi - row_lo
    g_lo, g_hi = row_lo, row_hi                      # this rank's rows in the global numbering

    ctx = htf.HtfContext(n, K, r_cut, device=dev)
    ctx.set_box(lo, hi)
    if world > 1:
        # bin only what can matter for this rank's rows (slab +- (r_cut + skin)), like HOOMD's ghost layer
        ctx.set_roi(*htf.parallel.roi_for_rows(pos[row_lo:row_hi], lo, hi, r_cut))
    halo = world > 1 and args.exchange == "halo"
    xch = None
    if halo:
        # rows are z-slabs (the generator's particle order is z-slowest): exchange only the two faces
        lo_face, hi_face, width, cap_h = htf.parallel.slab_plan(pos[row_lo:row_hi], 2, r_cut)
        t = torch.tensor([cap_h], dtype=torch.int64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        xch = htf.parallel.SlabExchange(ctx, rows, 2, lo_face, hi_face, width, int(t.item()), transport=args.transport)
        xch.own.copy_(torch.from_numpy(pos[row_lo:row_hi]).to(dev))
        d_pos_all, d_shard = xch.local, xch.own
        row_lo, row_hi = 0, rows                     # local indexing: own rows come first
    else:
        d_pos_all = torch.from_numpy(pos).to(dev)
        d_shard = d_pos_all[row_lo:row_hi].clone()
    p2p = xch is not None and xch.transport == "p2p"
    nl = torch.empty((rows, K, 4), dtype=torch.float32, device=dev)
    fe = torch.empty((rows, 4), dtype=torch.float32, device=dev)
    vir = torch.empty((rows, 6), dtype=torch.float32, device=dev)
    bins = torch.zeros(102, dtype=torch.int64, device=dev) if args.rdf else None
    packed = None
    if args.model == "mlp":
        packed = ctx.mlp_pack(torch.from_numpy(mlp_raw_parameters()).to(dev))
    train = None
    if args.model == "mlp-train":
        raw0 = torch.from_numpy(mlp_raw_parameters()).to(dev)
        train = {"raw": raw0.clone(), "m": torch.zeros_like(raw0), "v": torch.zeros_like(raw0),
                 "t": torch.zeros(1, dtype=torch.float32, device=dev), "g": torch.empty_like(raw0),
                 "loss": torch.zeros(1, dtype=torch.float32, device=dev), "labels": None, "raw0": raw0}
    eds_model = None
    if args.model == "eds":
        # EDS period 25, learning rate 5.0 as examples/03; the set point becomes the initial CV + 5 % below
        eds_model = htf.models.EDSCoordinationModel(K, set_point=1.0, period=25, learning_rate=5.0, r0=1.3,
                                                    rdf_range=(0.0, r_cut), nbins=100)
    ev = lambda: torch.cuda.Event(enable_timing=True)

    skin = args.skin > 0.0
    if skin:
        if world > 1:
            raise SystemExit("--skin is single-GPU (the halo layout changes every step)")
        ctx.skin_configure(args.skin)
        g = torch.Generator(device="cpu").manual_seed(7)
        v = torch.randn((n, 4), generator=g)
        v[:, 3] = 0.0
        v[:, :3] *= args.step_length / v[:, :3].norm(dim=1, keepdim=True)
        d_vel = v.to(dev)
        d_lo = torch.tensor(list(lo) + [0.0], dtype=torch.float32, device=dev)
        d_L = torch.tensor([hi[a] - lo[a] for a in range(3)] + [1.0], dtype=torch.float32, device=dev)
        skin_state = {"t": 0, "rebuilds": 0}

    # the builder's neighbors-per-row ride along to the pair pass, which then reads only each row's valid slots
    # (--model mlp: its compaction pre-pass then reads 4 bytes per row instead of the tensor twice)
    cnt = None if (skin or args.no_counts) else torch.empty((rows,), dtype=torch.int32, device=dev)
    if eds_model is not None:
        eds_model.row_counts = cnt

    def phase_exchange():
        if halo:
            xch.exchange()                                            # the path's one exchange step (2 faces)
        elif world > 1:
            dist.all_gather_into_tensor(d_pos_all, d_shard)           # ... or every position

    def phase_bin():
        ctx.bin_particles(d_pos_all)

    def phase_build():
        if skin:
            # the particles move (ballistic, |v| dt = --step-length); every --rebuild-every steps they are wrapped back
            # into the box and the candidate lists (r_cut + skin) are rebuilt; the distance filter runs every step
            d_pos_all.add_(d_vel)
            if skin_state["t"] % args.rebuild_every == 0:
                wrapped = torch.remainder(d_pos_all - d_lo, d_L) + d_lo
                wrapped[:, 3] = d_pos_all[:, 3]
                d_pos_all.copy_(wrapped)
                ctx.skin_rebuild(d_pos_all)
                skin_state["rebuilds"] += 1
            skin_state["t"] += 1
            ctx.skin_nlist(d_pos_all, out=nl)
        else:
            ctx.build_nlist(d_pos_all, row_lo, row_hi, out=nl, rebin=False, count_out=cnt)

    def phase_force():
        if train is not None:
            # one online force-matching step: gradient of MSE(model forces+energy, labels), summed over ranks, Adam
            ctx.mlp_train_grads(nl, train["raw"], r_cut, train["labels"], n_total=n, grads=train["g"], pred=fe, loss=train["loss"])
            if world > 1:
                if p2p:
                    ctx.comm_allreduce(train["g"])                    # ~10.5k weight gradients through the peer mailboxes
                else:
                    dist.all_reduce(train["g"])
            ctx.adam_step(train["raw"], train["g"], train["m"], train["v"], train["t"], lr=1e-3)
        elif eds_model is not None:
            eds_model.compute(nl, None, None)                         # fused LJ + CV + RDF pass, all-reduces, EDS update, bias
        elif packed is not None:
            ctx.mlp_forces(nl, packed, r_cut, out=fe, counts=cnt)
        elif bins is not None:
            bins.zero_()
            ctx.lj_step_forces_only(nl, fe, vir, bins, (0.0, r_cut), 100, counts=cnt)
            if world > 1:
                if p2p:
                    ctx.comm_allreduce(bins)                          # mailboxes in peer memory, graph-capturable
                else:
                    dist.all_reduce(bins)
        else:
            ctx.lj_forces(nl, virial=True, virial_components=6, out=fe, virial_out=vir, counts=cnt)

    def step(marks=None):
        if marks is not None:
            marks[0].record()
        phase_exchange()
        if marks is not None:
            marks[1].record()
        if not skin:
            phase_bin()
        if marks is not None:
            marks[2].record()
        phase_build()
        if marks is not None:
            marks[3].record()
        phase_force()
        if marks is not None:
            marks[4].record()

    if train is not None:
        # labels: 0.05 x the LJ forces of the same configuration (the library's own LJ pass, outside every timed region)
        phase_exchange(); phase_bin(); phase_build()
        train["labels"] = ctx.lj_forces(nl).mul_(0.05)
        torch.cuda.synchronize()
    if eds_model is not None:
        step()                                                        # one pass to measure the initial CV
        eds_model.eds_bias.set_point.fill_(float(eds_model.cv_avg.result()) * 1.05)
        eds_model.cv_avg.reset()

    # ---- the step's phases captured into CUDA graphs (same kernels, same stream order; the events between the
    #      phases stay live).  Binning alone is six dependent launches of a few microseconds.  NCCL stays eager. ----
    graphs = None
    step_graph = None
    launches_per_step = None
    eager_force = eds_model is not None or (bins is not None and world > 1 and not p2p) or \
        (train is not None and world > 1 and not p2p)              # host-side collectives / metric updates
    if not skin and not args.no_graph and (world == 1 or halo):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                step()                                          # warm-up on the capture stream (allocations, func attributes)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graphs = []
        l0 = ctx.launches
        pack_graph = None
        if halo:                                                # the device half of the exchange
            pack_graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(pack_graph, stream=side):
                xch.pack()
        for fn in (phase_bin, phase_build) + (() if eager_force else (phase_force,)):
            g_ = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_, stream=side):
                fn()
            graphs.append(g_)
        launches_per_step = ctx.launches - l0
        # ... and the whole step as ONE graph (same kernels, same order) when nothing in it needs the host: the five
        # event records and the graph boundaries of the phase-by-phase form cost 14 + 4 us per step (tools/graph_gap_time.py),
        # so only every --phase-every-th step of the timed region runs phase by phase with the events between
        if not eager_force and (not halo or p2p) and args.phase_every != 1:
            step_graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(step_graph, stream=side):
                if halo:
                    xch.pack()
                phase_bin(); phase_build(); phase_force()

        def step(marks=None):                                   # noqa: F811 -- the graph replay of the same step
            if marks is None and step_graph is not None:
                step_graph.replay()
                return
            if marks is not None:
                marks[0].record()
            if pack_graph is not None:
                pack_graph.replay()
                xch.swap()
            if marks is not None:
                marks[1].record()
            graphs[0].replay()
            if marks is not None:
                marks[2].record()
            graphs[1].replay()
            if marks is not None:
                marks[3].record()
            if eager_force:
                phase_force()
            else:
                graphs[2].replay()
            if marks is not None:
                marks[4].record()

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    sampler = ClockSampler(local) if (rank == 0 and full) else None
    warm = max(args.warmup, 3)
    for _ in range(warm):
        step()
    sync_all()
    assert ctx.overflow() == 0, "neighbor list overflowed K: the run is void"

    # ---- timed region: exactly --steps steps, device timed, clocks sampled meanwhile ----
    every = max(1, args.phase_every) if step_graph is not None else 1
    marks = [[ev() for _ in range(5)] if i % every == 0 else None for i in range(args.steps)]
    skin_t0 = skin_state["t"] if skin else 0
    line0 = sampler.lines() if sampler else 0
    launches0 = ctx.launches
    e0, e1 = ev(), ev()
    sync_all()
    e0.record()
    for i in range(args.steps):
        step(marks[i])
    e1.record()
    sync_all()
    ms = e0.elapsed_time(e1)
    launches = ctx.launches - launches0 if graphs is None else launches_per_step * args.steps
    assert ctx.overflow() == 0, "neighbor list overflowed K: the run is void"
    if skin:
        assert ctx.skin_status() == (0, 0), "a buffered list was used past skin/2 or overflowed: the run is void"
    clocks = sampler.stop(line0) if sampler else None
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    phase = lambda a, b: float(np.mean([m[a].elapsed_time(m[b]) for m in marks if m is not None]))
    exchange_ms, bin_ms, build_ms, force_ms = phase(0, 1), phase(1, 2), phase(2, 3), phase(3, 4)
    per_rank = None
    if world > 1:
        # every rank's own phase times: a rank waits for both neighbours inside its exchange phase, so the rank with the
        # longest compute phases shows the shortest exchange and sets the pace of all
        mine = {"exchange_ms": round(exchange_ms, 5), "compute_ms": round(bin_ms + build_ms + force_ms, 5)}
        allr = [None] * world
        dist.all_gather_object(allr, mine)
        per_rank = {"exchange_ms": [r_["exchange_ms"] for r_ in allr], "compute_ms": [r_["compute_ms"] for r_ in allr],
                    "what": "per rank: exchange phase (pack + signal + wait for both neighbours + gather) and "
                            "binning + build + forces, CUDA events, mean over the timed steps"}
    value = n * args.steps / (ms * 1e-3)

    # ---- parity leg (outside the timed region; the oracle is the checker, never the thing measured) ----
    parity = None
    if not args.no_parity and not skin:
        parity = parity_leg(args, env, htf, ctx, pos, lo, hi, r_cut, K, g_lo, rows, row_lo, row_hi, d_pos_all, nl, fe, vir,
                            phase_exchange, packed)

    # ---- end to end through the public API with host buffers (tfcompute + built-in LJ virial model) ----
    e2e = None
    if full and not args.no_e2e:
        e2e = run_e2e(args, htf, torch, dist, world, rank, dev, pos, lo, hi, r_cut, K, g_lo, g_hi)

    line = None
    if rank == 0:
        peak, peak_src = peaks()
        alg_build = rows * (16 * K + 16)
        achieved = alg_build / (build_ms * 1e-3) / 1e9
        traffic = ncu_traffic(args, rows, K) if (full and not skin) else None
        step_ms = ms / args.steps
        ftraffic = ncu_force_traffic(args, rows, K, cnt is not None) if (full and not skin) else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": step_ms, "higher_is_better": True,
            "scaling": scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg_dict(args, n, K, r_cut, world),
            "exchange": ({"transport": xch.transport, "note": xch.transport_note or None,
                          "what": "fused pack + send into the neighbours' peer-memory windows, flags, gather (htf_comm_exchange_halo)"
                                  if p2p else "pack kernels + NCCL send/recv"} if xch is not None else None),
            "roofline": {"bound": "hbm", "kernel": "nlist_tile2_kernel (+ per-cell fallback pass)", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_build, "kernel_ms": build_ms,
                         "traffic": traffic["dram_bytes_per_launch"] if traffic else None,
                         "traffic_source": traffic["source"] if traffic else
                         "no committed ncu capture for this configuration (profiles/nlist_build_traffic.json is keyed by rows and K)",
                         "path_frac": (rows * (32 * K + 56)) / (step_ms * 1e-3) / 1e9 / peak,
                         "path_bytes_per_row": {"nlist written and re-read (32K+56)": 32 * K + 56,
                                                "nlist written once (16K+56)": 16 * K + 56},
                         "path_frac_single_write": (rows * (16 * K + 56)) / (step_ms * 1e-3) / 1e9 / peak,
                         "binning_ms": bin_ms, "force_kernel_ms": force_ms,
                         "force_kernel_frac": rows * (16 * K + 40) / (force_ms * 1e-3) / 1e9 / peak,
                         "force_kernel_traffic": ftraffic["dram_bytes_per_launch"] if ftraffic else None,
                         "force_kernel_frac_dram": (ftraffic["dram_bytes_per_launch"] / (force_ms * 1e-3) / 1e9 / peak
                                                    if ftraffic else None),
                         "force_kernel_note": ("algorithmic bytes (16K+40 per row) over the measured time; the pass gets the "
                                               "builder's per-row counts and reads only the valid slots of each row, so its "
                                               "DRAM traffic is below the algorithmic figure and force_kernel_frac can exceed 1; "
                                               "force_kernel_frac_dram is the committed ncu capture's DRAM bytes over the "
                                               "measured time (profiles/pair_pass_traffic.json)"
                                               if cnt is not None else "reads all K slots of every row")},
            "exchange_ms": exchange_ms if world > 1 else 0.0,
            "per_rank": per_rank,
            "clocks": clocks, "gpu_launches": launches,
            "launch": (("1 CUDA graph per step (%sbinning, build, forces: %d kernels); every %d-th step of the timed region "
                        "runs as %d graphs with CUDA event records between the phases (the phase times of this record)"
                        % ("halo exchange, " if halo else "", launches_per_step, every, len(graphs) + (1 if halo else 0)))
                       if step_graph is not None else
                       "%d CUDA graphs per step (%sbinning | build%s), %d kernels"
                       % (len(graphs) + (1 if halo else 0), "halo packing | " if halo else "",
                          "" if eager_force else " | forces", launches_per_step)
                       if graphs is not None else "stream launches from the host"),
            "parity": parity,
            "e2e": e2e,
        }
        if args.shuffle:
            line["config"]["particle_order"] = "shuffled (random permutation of the lattice order)"
        if skin:
            cfg = line["config"]
            cfg["workload"] += "+skin"
            cfg["skin"] = {"r_buff": args.skin, "rebuild_every": args.rebuild_every, "step_length": args.step_length,
                           "rebuilds_in_timed_region": sum(1 for i in range(args.steps)
                                                           if (skin_t0 + i) % args.rebuild_every == 0),
                           "note": "HOOMD-style buffered list: search (r_cut + r_buff) amortised, prepareNeighbors-style "
                                   "distance filter every step; particles move ballistically every step"}
            line["roofline"]["kernel"] = "nlist_filter_kernel (per-step pass; the search is amortised inside build_ms)"
        if args.model == "mlp":
            # the dominant kernel of this workload is the tensor-core MLP: report it against the measured bf16 peak,
            # counting only the VALID (non-padded) pairs as useful work
            tpeak, tsrc = tensor_peak()
            valid_pairs = int((nl[:, :, :3].abs().sum(-1) > 0).sum().item())
            flop = valid_pairs * MLP_FLOP_PER_PAIR
            tf = flop / (force_ms * 1e-3) / 1e12
            line["dtype"] = "bf16"
            line["roofline"] = {"bound": "tensor", "kernel": "mlp_force_kernel (tcgen05, operands in TMEM)", "achieved": tf,
                                "peak": tpeak, "unit": "TFLOP/s", "frac": tf / tpeak, "peak_source": tsrc,
                                "algorithmic_flop_per_launch": flop, "kernel_ms": force_ms, "traffic": None,
                                "valid_pairs": valid_pairs, "slots": rows * K,
                                "frac_counting_padded_slots": rows * K * MLP_FLOP_PER_PAIR / (force_ms * 1e-3) / 1e12 / tpeak,
                                "note": "41,216 useful flop per VALID pair (value + tangent through 32-64-64-64-1); the kernel's own "
                                        "bound is the MUFU pipe (200 MUFU per pair at 16 lanes/clk/SM), see DESIGN.md",
                                "nlist_build_ms": build_ms, "nlist_build_frac": achieved / peak}
        if train is not None:
            tpeak, tsrc = tensor_peak()
            valid_pairs = int((nl[:, :, :3].abs().sum(-1) > 0).sum().item())
            flop = valid_pairs * MLP_TRAIN_FLOP_PER_PAIR
            tf = flop / (force_ms * 1e-3) / 1e12
            line["dtype"] = "bf16"
            line["roofline"] = {"bound": "tensor", "kernel": "training step: mlp_force_kernel (tcgen05) + mlp_train_kernel (warp MMAs) + "
                                "gradient reduce + Adam", "achieved": tf, "peak": tpeak, "unit": "TFLOP/s", "frac": tf / tpeak,
                                "peak_source": tsrc, "algorithmic_flop_per_launch": flop, "kernel_ms": force_ms, "traffic": None,
                                "valid_pairs": valid_pairs, "slots": rows * K,
                                "note": "%d useful flop per VALID pair: forward value+tangent chains once, input-gradient chains, "
                                        "weight gradients (the kernel's own forward recompute and the padded slots are not counted)"
                                        % MLP_TRAIN_FLOP_PER_PAIR,
                                "nlist_build_ms": build_ms, "nlist_build_frac": achieved / peak,
                                "loss_after_timed_region": float(train["loss"].item()), "adam_steps": int(train["t"].item())}
        if world == 1 and full and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args, pos, lo, hi, r_cut, K)
    del ctx, nl, fe, vir, d_pos_all, d_shard, xch, graphs
    torch.cuda.empty_cache()
    return line


def _sorted_rows(a):
    """Rows of a [rows, K, 4] float32 tensor (numpy), each sorted by the bit patterns of (dx, dy, dz, type): slot order
    is unspecified, the multiset of entries is what must match bit for bit."""
    import numpy as np
    u = np.ascontiguousarray(a).view(np.uint32).astype(np.uint64)
    key1 = (u[..., 0] << np.uint64(32)) | u[..., 1]
    key2 = (u[..., 2] << np.uint64(32)) | u[..., 3]
    order = np.lexsort((key2, key1), axis=1)
    return np.take_along_axis(a, order[:, :, None], axis=1)


def parity_leg(args, env, htf, ctx, pos, lo, hi, r_cut, K, g_lo, rows, row_lo, row_hi, d_pos_all, nl, fe, vir,
               phase_exchange, packed):
    """Every rank checks a slice of its rows against the CPU oracle (neighbor tensor bit-exact as a multiset, RDF bins of
    the slice bit-exact, forces / virial 1e-5) and the ranks all-reduce the size-independent sums.  Mirrors the intent
    of the reference's MPI test (/root/reference htf/test-py/test_mpi_tensorflow.py:59-80): a domain-decomposed run
    must reproduce the single-domain numbers."""
    import numpy as np
    import torch
    import oracle
    world, rank, dev, dist = env["world"], env["rank"], env["dev"], env["dist"]
    oracle.build()
    oracle.set_threads(max(1, (os.cpu_count() or 1) // world))
    m = min(rows, 2048)
    a0 = (rows - m) // 2                                     # slice in the middle of this rank's rows (local numbering)
    # one untimed LJ step with the RDF histogram through the one-call entry point, on freshly exchanged positions
    phase_exchange()
    bins = torch.zeros(102, dtype=torch.int64, device=dev)
    vir_c = torch.empty((rows, 6), dtype=torch.float32, device=dev)
    fe_c = ctx.lj_step(d_pos_all, row_lo, row_hi, nlist_out=nl, virial_out=vir_c, bins=bins, r_range=(0.0, r_cut), nbins=100)
    bins_slice = ctx.rdf_hist(nl[a0:a0 + m], (0.0, r_cut), 100)
    torch.cuda.synchronize()
    sub, s0 = slab_with_halo(pos, g_lo + a0, g_lo + a0 + m, r_cut) if world * rows == pos.shape[0] and not args.shuffle \
        else (pos, g_lo + a0)
    nl_o, _, cnt_o = oracle.nlist(sub, lo, hi, r_cut, K, s0, s0 + m, cells=True, want_idx=False)
    fe_o, _, v6_o = oracle.lj(nl_o)
    h_o = oracle.rdf_hist(nl_o, (0.0, r_cut), 100)
    nl_g = nl[a0:a0 + m].cpu().numpy()
    nlist_ok = bool(np.array_equal(_sorted_rows(nl_g).view(np.uint32), _sorted_rows(nl_o).view(np.uint32)))
    rdf_ok = bool(np.array_equal(bins_slice.cpu().numpy(), h_o))

    def rel(got, want):
        got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
        scale = np.sqrt(np.mean(want ** 2)) + 1e-30
        return float((np.abs(got - want) / np.maximum(np.abs(want), scale)).max())

    f_err = rel(fe_c[a0:a0 + m].cpu().numpy(), fe_o)
    v_err = rel(vir_c[a0:a0 + m].cpu().numpy(), v6_o)
    train_err = None
    if args.model == "mlp-train":
        mm = min(m, 256)
        raw = torch.from_numpy(mlp_raw_parameters()).to(dev)
        lab = (0.05 * fe_o[:mm]).astype(np.float32)
        g_k, _, _ = ctx.mlp_train_grads(nl[a0:a0 + mm].contiguous(), raw, r_cut, torch.from_numpy(lab).to(dev))
        _, g_ref, _ = oracle.pairwise_mlp_train_grads(nl_o[:mm], mlp_raw_parameters(), r_cut, lab)
        g_k = g_k.cpu().numpy().astype(np.float64)
        cuts = np.cumsum([0, 2048, 64, 4096, 64, 4096, 64, 64, 1])
        train_err = max(float(np.abs(g_k[cuts[i]:cuts[i + 1]] - g_ref[cuts[i]:cuts[i + 1]]).max()
                              / (np.sqrt(np.mean(g_ref[cuts[i]:cuts[i + 1]] ** 2)) + 1e-30)) for i in range(8))
    mlp_err = None
    if packed is not None:
        mm = min(m, 256)
        f_mlp = ctx.mlp_forces(nl[a0:a0 + mm].contiguous(), packed, r_cut).cpu().numpy().astype(np.float64)
        f_ref = oracle.pairwise_mlp(nl_o[:mm], mlp_raw_parameters(), r_cut).astype(np.float64)
        mlp_err = float(np.abs(f_mlp[:, :3] - f_ref[:, :3]).max() / (np.sqrt(np.mean(f_ref[:, :3] ** 2)) + 1e-30))
    # size-independent sums over ALL rows of all ranks
    sums = torch.cat([fe_c[:, :3].double().sum(0), fe_c[:, :3].double().abs().sum().reshape(1),
                      fe_c[:, 3].double().sum().reshape(1), vir_c.double().sum(0)])
    flags = torch.tensor([int(nlist_ok), int(rdf_ok)], dtype=torch.int64, device=dev)
    errs = torch.tensor([f_err, v_err], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(sums)
        dist.all_reduce(bins)
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
        dist.all_reduce(errs, op=dist.ReduceOp.MAX)
    sums = sums.cpu().numpy()
    n_global = pos.shape[0]
    out = {
        "checker": "oracle/ (CPU restatement), %d rows per rank from the middle of its shard" % m,
        "nlist_multiset_bit_exact": bool(flags[0].item()), "rdf_slice_bit_exact": bool(flags[1].item()),
        "force_energy_max_rel_err": float(errs[0].item()), "virial_max_rel_err": float(errs[1].item()), "tolerance": 1e-5,
        "rdf_total": int(bins.sum().item()), "rdf_total_expected": int(n_global) * K,
        "sum_force_over_sum_abs_force": float(np.abs(sums[:3]).max() / max(sums[3], 1e-30)),
        "sum_energy": float(sums[4]), "sum_virial6": [float(x) for x in sums[5:11]],
        "rdf_bins_sha": __import__("hashlib").sha1(bins.cpu().numpy().tobytes()).hexdigest()[:16],
    }
    if mlp_err is not None:
        out["mlp_force_max_err_over_rms"] = mlp_err
        out["mlp_tolerance"] = 1e-1
    if train_err is not None:
        t_ = torch.tensor([train_err], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_, op=dist.ReduceOp.MAX)
        out["train_grad_max_err_over_block_rms"] = float(t_.item())
        out["train_grad_tolerance"] = 1e-1
        mlp_err = float(t_.item()) if mlp_err is None else max(mlp_err, float(t_.item()))
    out["ok"] = bool(out["nlist_multiset_bit_exact"] and out["rdf_slice_bit_exact"] and out["force_energy_max_rel_err"] <= 1e-5
                     and out["virial_max_rel_err"] <= 1e-5 and out["rdf_total"] == out["rdf_total_expected"]
                     and out["sum_force_over_sum_abs_force"] <= 1e-5 and (mlp_err is None or mlp_err < 1e-1))
    return out


def run_e2e(args, htf, torch, dist, world, rank, dev, pos, lo, hi, r_cut, K, row_lo, row_hi):
    """Same metric through tfcompute (the user-facing call) with HOST buffers: every step copies the rank's positions
    from pinned host memory and reads forces+virial back into pinned host memory.  The step runs in row batches
    (tfcompute's batch_size, the reference's own chunking): the device->host copy of batch b leaves on a copy stream
    while batch b+1 is built and evaluated, so only the first build and the last copy are exposed."""
    import numpy as np
    n = pos.shape[0]
    rows = row_hi - row_lo
    halo = world > 1 and args.exchange == "halo"
    if args.model == "mlp-train":
        return run_e2e_train(args, htf, torch, dist, world, rank, dev, pos, lo, hi, r_cut, K, row_lo, row_hi)
    if args.model == "mlp":
        model = htf.models.PairwiseMLPModel(K, r_cut=r_cut).to(dev)
    elif args.model == "eds":
        model = htf.models.EDSCoordinationModel(K, set_point=30.0, period=25, learning_rate=5.0, r0=1.3,
                                                rdf_range=(0.0, r_cut), nbins=100)
    else:
        model = htf.models.LJVirialModel(K, virial=True)
    tfc = htf.tfcompute(model)
    nbatch = args.e2e_batches if args.model == "lj" else 1
    batch = None if nbatch <= 1 else (rows + nbatch - 1) // nbatch
    h_pos = torch.from_numpy(pos[row_lo:row_hi].copy()).pin_memory()
    h_f = torch.empty((rows, 4), dtype=torch.float32).pin_memory()
    h_v = torch.empty((rows, 6), dtype=torch.float32).pin_memory() if args.model == "lj" else None
    if halo:
        # like the device-resident leg: the rank's system is [own rows | halo from below | halo from above]
        lo_face, hi_face, width, cap_h = htf.parallel.slab_plan(pos[row_lo:row_hi], 2, r_cut)
        t_ = torch.tensor([cap_h], dtype=torch.int64, device=dev)
        dist.all_reduce(t_, op=dist.ReduceOp.MAX)
        cap_h = int(t_.item())
        local0 = np.concatenate([pos[row_lo:row_hi], np.full((2 * cap_h, 4), 1e30, dtype=np.float32)])
        system = htf.sim.System(local0, lo, hi, device=dev)
        tfc.attach(htf.sim.nlist_cell(system), r_cut=r_cut, batch_size=batch)
        xch = htf.parallel.SlabExchange(tfc.ctx, rows, 2, lo_face, hi_face, width, cap_h, transport=args.transport)
        system.positions = xch.local
        d_shard = xch.own
        out_lo, out_hi = 0, rows
    else:
        system = htf.sim.System(pos, lo, hi, device=dev)
        tfc.attach(htf.sim.nlist_cell(system), r_cut=r_cut, batch_size=batch)
        d_shard = system.positions[row_lo:row_hi] if world == 1 else torch.empty((rows, 4), dtype=torch.float32, device=dev)
        out_lo, out_hi = row_lo, row_hi
    tfc.shard = (out_lo, out_hi)
    tfc.set_host_outputs(h_f, h_v)
    if world > 1:
        tfc.ctx.set_roi(*htf.parallel.roi_for_rows(pos[row_lo:row_hi], lo, hi, r_cut))

    def step(t):
        d_shard.copy_(h_pos, non_blocking=True)
        if halo:
            xch.exchange()
        elif world > 1:
            dist.all_gather_into_tensor(system.positions, d_shard)
        tfc.compute_forces(t)                       # forces (+virial) of every batch are mirrored into h_f / h_v

    steps = max(3, min(args.steps, 10))
    for t in range(3):
        step(t)
    tfc.host_sync()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for t in range(steps):
        step(t)
    torch.cuda.current_stream().wait_event(tfc._host["done"])      # the last device->host copies belong to the step
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        tt = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    # the mirrored host result is the device result
    ok = bool(torch.equal(h_f, tfc._forces[out_lo:out_hi].cpu()))
    return {"value": n * steps / (ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h_pos.numel() * 4),
            "d2h_bytes_per_step": int(h_f.numel() * 4 + (h_v.numel() * 4 if h_v is not None else 0)), "steps": steps,
            "ms_per_step": ms / steps, "row_batches": max(nbatch, 1), "host_result_matches_device": ok, "cpu_binding": globals().get("_NUMA_NOTE"),
            "api": "htf.tfcompute(%s).compute_forces, pinned host positions in, forces%s mirrored to pinned host memory "
                   "batch by batch (set_host_outputs)"
                   % (("PairwiseMLPModel", "+energy") if args.model == "mlp" else
                      ("EDSCoordinationModel", "+energy") if args.model == "eds" else ("LJVirialModel", "+virial"))}


def run_e2e_train(args, htf, torch, dist, world, rank, dev, pos, lo, hi, r_cut, K, row_lo, row_hi):
    """Config 4 end to end: tfcompute in the reference's label / training mode (attach(train=True) +
    set_reference_forces, htf/tensorflowcompute.py:265-282,346-370).  Every step the positions come from pinned host
    memory, the half-step hook builds the neighbor tensor, takes the label forces and runs one fused train_on_batch;
    the loss is read back into pinned host memory."""
    n = pos.shape[0]
    if world > 1:
        return {"value": None, "unit": UNIT, "note": "the training e2e leg is single-GPU; the multi-GPU gradient all-reduce is in the "
                                                      "device-resident leg", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    model = htf.models.PairwiseMLPModel(K, r_cut=r_cut, output_forces=False).to(dev)
    model.compile("Adam", "MeanSquaredError")
    system = htf.sim.System(pos, lo, hi, device=dev)
    system.integrator = htf.sim.NVE(0.005)
    tfc = htf.tfcompute(model)
    tfc.attach(htf.sim.nlist_cell(system), r_cut=r_cut, train=True)
    tfc.set_reference_forces(htf.sim.PairLJ(system, r_cut, K, scale=0.05))
    h_pos = torch.from_numpy(pos.copy()).pin_memory()
    h_loss = torch.zeros(1, dtype=torch.float32).pin_memory()

    def step(t):
        system.positions.copy_(h_pos, non_blocking=True)
        tfc.half_step(t)
        h_loss.copy_(model.last_loss.reshape(1), non_blocking=True)

    steps = max(3, min(args.steps, 10))
    for t in range(3):
        step(t)
    torch.cuda.synchronize()
    l_first = float(h_loss[0])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(steps):
        step(t)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    return {"value": n * steps / (ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h_pos.numel() * 4), "d2h_bytes_per_step": 4,
            "steps": steps, "ms_per_step": ms / steps, "loss_first": l_first, "loss_last": float(h_loss[0]),
            "api": "htf.tfcompute(PairwiseMLPModel).attach(train=True) + set_reference_forces(PairLJ x 0.05): half_step -> "
                   "build, labels, fused train_on_batch; pinned host positions in, loss out"}


def cpu_baseline(args, pos, lo, hi, r_cut, K):
    """The CPU restatement (oracle/) timed on this box's host cores on a bounded sample of the workload."""
    import oracle
    oracle.build()
    oracle.set_threads(0)
    n = pos.shape[0]
    rows = min(n, 2048 if args.model == "mlp-train" else 32768 if args.model == "mlp" else 262144)
    a0 = (n - rows) // 2
    raw = mlp_raw_parameters() if args.model in ("mlp", "mlp-train") else None
    reps, t_tot = 0, 0.0
    while t_tot < 8.0 and reps < 6:
        t0 = time.perf_counter()
        nl, _, _ = oracle.nlist(pos, lo, hi, r_cut, K, a0, a0 + rows, cells=True, want_idx=False)
        if args.model == "mlp-train":
            cpu_train_step(oracle, nl, raw, r_cut)
        elif raw is not None:
            oracle.pairwise_mlp(nl, raw, r_cut)
        else:
            oracle.lj(nl, virial=True)
        if args.rdf:
            oracle.rdf_hist(nl, (0.0, r_cut), 100)
        t_tot += time.perf_counter() - t0
        reps += 1
    return {"value": rows * reps / t_tot, "unit": UNIT, "cores": oracle.num_threads(), "kind": "port",
            "sample": "%d x a %d-row slab of the %d-particle system (cell binning of all particles included)" % (reps, rows, n),
            "host_cores": os.cpu_count()}


def cpu_train_step(oracle, nl, raw, r_cut):
    """One force-matching step on the CPU: labels 0.05 x LJ, float64 gradient sweep, Adam (oracle/)."""
    import numpy as np
    fe, _, _ = oracle.lj(nl, virial=False)
    loss, g, _ = oracle.pairwise_mlp_train_grads(nl, raw, r_cut, 0.05 * fe)
    z = np.zeros_like(raw, dtype=np.float32)
    oracle.adam_step(raw, g, z, z, 1)
    return loss


def run_refbench(args):
    """The reference's one published benchmark, end to end through the kept API (/root/reference htf/test-py/benchmark.py:
    25-48): 256 particles on a 2-D square lattice a=2.0, HOOMD-side LJ pair force + the LJ SimModel through tfcompute
    (NN=64, r_cut=3.0), Langevin kT=1 dt=0.005, 4000 equilibration steps, then 1000 timed steps, 5 rounds (median).
    Published: median 2.007 s per 1000 steps = 1.28e5 particle-timesteps/s on a Xeon Gold 6130."""
    import numpy as np
    import torch
    import htf
    from htf import sim
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU: the hot path has no CPU fallback")
    published = 256 * 1000 / 2.007
    system = sim.create_lattice(sim.sq(2.0), [16, 16])
    nlist = sim.nlist_cell(system, check_period=1)
    system.forces.append(sim.ReferenceLJ(system, 3.0))            # hoomd.md.pair.lj stays attached in the reference run
    system.integrator = sim.Langevin(0.005, kT=1.0, seed=42)
    system.run(4000)
    tfc = htf.tfcompute(htf.models.LJModel(64))
    tfc.attach(nlist, r_cut=3.0)
    system.forces.append(tfc)
    system.run(50)
    torch.cuda.synchronize()
    rounds = []
    l0 = tfc.ctx.launches
    for _ in range(5):
        t0 = time.perf_counter()
        system.run(1000)
        torch.cuda.synchronize()
        rounds.append(time.perf_counter() - t0)
    med = float(np.median(rounds))
    value = 256 * 1000 / med
    print(json.dumps({
        "metric": "particle-timesteps/s (whole simulation, the reference's test_lj_benchmark)", "value": value,
        "unit": UNIT, "n_gpus": 1, "steps": 1000, "warmup": 4050, "ms_per_step": med, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": value / published, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "refbench: N=256 2-D sq lattice a=2.0, NN=64, r_cut=3.0, Langevin kT=1 dt=0.005, "
                               "1000 steps x 5 rounds (median), wall clock incl. the Python driver loop",
                   "published": {"value": published, "where": "Xeon Gold 6130, HOOMD+TensorFlow, BASELINE.md section 1"}},
        "rounds_s": rounds, "gpu_launches": tfc.ctx.launches - l0}))
    return 0


if __name__ == "__main__":
    a = parse()
    if a.workload == "refbench" and a.impl != "reference":
        sys.exit(run_refbench(a))
    sys.exit(run_reference(a) if a.impl == "reference" else run_b200(a))
