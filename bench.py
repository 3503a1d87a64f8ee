#!/usr/bin/env python
"""bench.py -- FNO spectral-conv fwd+adjoint grid-points/sec (BASELINE.json's metric).

    python bench.py --gpus 1 --steps K --warmup W            # our arm, config C3
    torchrun ... bench.py --gpus N ...                        # our arm, config C4 ladder (weak scaling)
    python bench.py --impl reference ...                      # restated reference on host cores

A "step" is one pass of the hot path over one batch of synthetic input: the forward apply
``y = S(θ) * x`` (taped), the adjoint apply ``x~ = S(θ)' * y2`` and the reverse-mode pullback
(parameter gradients + input cotangent) of the spectral convolution
``S = dft_all' * (ParDiagonal ⊗ ParMatrix) * dft_all`` of examples/fno.jl.  ``value`` counts
each grid point (spatial-temporal location x batch; channels are in the bytes) once per step.

Prints ONE JSON line (rank 0).  Keys follow the driver contract; ``roofline`` is for the
dominant kernel timed live with CUDA events on the launching stream, ``cpu_baseline`` is the
oracle (a restatement of the reference's operation sequence on NumPy/pocketfft -- the
reference itself is Julia and cannot run here) on a bounded sample, ``e2e`` is the same
metric through the host-buffer C-ABI call (pinned H2D + apply + D2H inside the timed region).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np

C3_SHAPE = [64, 64, 64, 32]          # [nx, ny, nz, nt] (examples/fno.jl order: time last = slowest array dim)
CHANNELS = 20
MODES = 8
# weak-scaling ladder, 2^24 grid points per GPU, [nx, ny, nz, nt] (x = fastest spatial dim, z = distributed dim).
# Growing x, then y, then z keeps the reference's own order heuristic (largest Domain-Range first, src/ParKron.jl:47-50)
# transforming the distributed z dim LAST, so the repartition moves the already-truncated tensor.
C4_LADDER = {1: [64, 64, 64, 64], 2: [128, 64, 64, 64], 4: [128, 128, 64, 64], 8: [128, 128, 128, 64]}
CPU_SAMPLE_SHAPE = [64, 64, 32, 32]  # cpu_baseline sample: 1/4 of C3's grid points
REF_SAMPLE_SHAPE = [32, 32, 32, 32]  # --impl reference sample per step: 1/8 of C3's grid points (named in config.workload)
os.environ.setdefault("PO_PEER_TIMEOUT_MS", "60000")  # a stuck peer must fail the bench loudly, not hang the box


def seeded_normal(n, seed, rank=0):
    """i.i.d. standard normal from numpy.random.default_rng (PCG64), generated in Float64 and
    cast; x uses seed 0, cotangents seed 1, parameters seed 2.  With several ranks every rank draws its own shard
    from the stream keyed (seed, rank)."""
    rng = np.random.default_rng(seed if rank == 0 else [seed, rank])
    return rng.standard_normal(n).astype(np.float32)


def space_modes(n):
    return [(1, MODES), (n - MODES + 1, n)]


def time_modes():
    return [(1, MODES)]


def modes_for(shape):
    """8 low + 8 high modes per spatial dim, 8 low modes for the real (time) dim.
    examples/fno.jl:17 zips the Kron factors with reverse(modes), which pairs spatial dim i with
    modes[nspace-1-i]; emit the list in that order so unequal extents get in-range ranges."""
    sp = shape[:-1]
    m = [[(1, MODES), (s - MODES + 1, s)] for s in reversed(sp)]
    m.append([(1, MODES)])
    return m


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            f = [v.strip() for v in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# -------------------------------------------------------------------------------------------
# CPU arm: the oracle in reference order
# -------------------------------------------------------------------------------------------
def cpu_pair_time(shape, reps=1, workers=None):
    """Seconds for one fwd+adjoint pair of the restated reference on `shape` (best of reps)."""
    import scipy.fft
    from oracle import parops_oracle as O
    S, _ = O.spectral_convolution(O.F32, shape, modes_for(shape), CHANNELS)
    theta = O.init(S, seed=2)
    rng = np.random.default_rng(0)
    x = rng.standard_normal((S.Domain, 1)).astype(np.float32)
    y2 = np.random.default_rng(1).standard_normal((S.Range, 1)).astype(np.float32)
    Sadj = S.adjoint()
    workers = workers or os.cpu_count()
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        with scipy.fft.set_workers(workers):
            O.pullback(S, x, y2, theta)   # forward apply + reverse mode (xbar + parameter gradients)
            Sadj(y2, theta)               # adjoint apply
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return best, workers


def run_reference(args):
    """The restated reference (NumPy/pocketfft oracle in the reference's operation order; Julia/FFTW are absent) on the
    host cores.  One step = the same layer on a BOUNDED SAMPLE grid (REF_SAMPLE_SHAPE, 1/8 of C3's points): the full C3 grid
    takes ~30 s per step on 16 cores, so K driver steps of it would not end within minutes.  The sample is named in
    config.workload -- the value is a per-point rate on a smaller, more cache-friendly problem, NOT a like-for-like C3 number."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    shape = REF_SAMPLE_SHAPE
    pts = int(np.prod(shape))
    for _ in range(args.warmup):
        cpu_pair_time(shape)
    t0 = time.perf_counter()
    workers = os.cpu_count()
    for _ in range(args.steps):
        _, workers = cpu_pair_time(shape)
    total = time.perf_counter() - t0
    ms = total / args.steps * 1e3
    val = pts / (ms / 1e3)
    sec1, _ = cpu_pair_time(shape, workers=1)
    sample = "same layer (c=%d, %d modes/dim, Float32, batch 1) on a %s grid = 1/%d of C3's grid points per step" % (
        CHANNELS, MODES, "x".join(map(str, shape)), int(np.prod(C3_SHAPE)) // pts)
    cfg = workload_config(args.gpus)
    cfg["workload"] = "SAMPLED %s -- reference arm: %s" % (cfg["workload"], sample)
    cfg["shape"] = shape
    cfg["same_config"] = False
    line = {
        "impl": "reference", "metric": "FNO spectral-conv fwd+adjoint grid-points/sec", "value": val, "unit": "grid-points/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": val, "unit": "grid-points/s", "cores": workers, "kind": "port", "sample": sample,
                         "value_1_thread": pts / sec1},
        "e2e": {"value": val, "unit": "grid-points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "restated reference (NumPy/pocketfft in the reference's operation order); Julia/FFTW are absent.  Per-point rate on the "
                "sampled grid named in config.workload; a ratio against the GPU arm's C3/C4 line is an extrapolation, not same-config",
    }
    emit(line)
    return 0


def workload_config(n_gpus):
    if n_gpus == 1:
        return {"workload": "C3: examples/fno.jl 4-D spectral-conv layer 64x64x64x32, 20 channels, 8 modes/dim, Float32, batch 1, "
                            "fwd apply + adjoint apply + parameter gradient (pullback)", "shape": C3_SHAPE, "channels": CHANNELS, "modes_per_dim": MODES,
                "batch": 1, "l2": "inputs (671 MB) larger than the 126 MB L2", "parallelism": "single GPU"}
    shape = C4_LADDER[n_gpus]
    return {"workload": "C4: distributed 4-D spectral-conv layer, global %s, 20 channels, 8 modes/dim, Float32, batch 1, z-slabs "
                        "[1,1,1,P,1], ParRepartition all-to-all over NCCL, fwd + adjoint + parameter gradient" % "x".join(map(str, shape)),
            "shape": shape, "channels": CHANNELS, "modes_per_dim": MODES, "batch": 1,
            "l2": "per-rank inputs (1.34 GB) larger than the 126 MB L2", "parallelism": "domain decomposition, %d z-slabs" % n_gpus}


# -------------------------------------------------------------------------------------------
# our arm
# -------------------------------------------------------------------------------------------
def build_single(parops, shape):
    S = parops.spectral_convolution(parops.Float32, shape, modes_for(shape), CHANNELS)
    theta = parops.gpu(parops.init(S, rng=2))
    St = S(theta)
    return St, St.adjoint(), S.Domain, S.Range


def build_distributed(parops, shape, P, rank):
    """C4: Kron(RF_t, RF_z, RF_y, RF_x, I_c) distributed over z-slabs, local frequency weights,
    and the adjoint pipeline back (src/ParKron.jl:244-309)."""
    T, CT = parops.Float32, parops.ComplexF32
    nx, ny, nz, nt = shape

    def rf(Tin, n, rng):
        F = parops.ParDFT(Tin, n)
        return parops.compose(parops.ParRestriction(CT, F.Range, rng), F)

    # Kron factor 1 = slowest dim: t, z, y, x, c   (array dims fastest first: c, x, y, z, t)
    K = parops.ParKron(rf(T, nt, time_modes()), rf(CT, nz, space_modes(nz)), rf(CT, ny, space_modes(ny)),
                       rf(CT, nx, space_modes(nx)), parops.ParIdentity(T, CHANNELS))
    # input: z-slabs [1,1,1,P,1]; output of the transform: slabs of the kept t-modes [1,1,1,1,P] -- the layout
    # distribute() already has after the z-transform, so `dims_out` saves the repartition back (and its mirror
    # in the inverse): 2 all-to-alls per apply instead of 4.  The frequency weights live on the t-mode slabs.
    dims = [1, 1, 1, P, 1]
    dims_out = [1, 1, 1, 1, P]
    Tdist = parops.distribute(K, dims, dims_out, rank=rank)
    kt_local = parops.local_size(MODES, rank, P)
    n_modes_local = 2 * MODES * 2 * MODES * 2 * MODES * kt_local
    ew, mix = parops.ParDiagonal(CT, n_modes_local), parops.ParMatrix(CT, CHANNELS, CHANNELS)
    W = parops.kron(ew, mix)
    S = parops.compose(Tdist.adjoint(), W, Tdist)
    # the channel-mixing matrix is REPLICATED: it gets a random stream of its own, so that every rank draws the same matrix whatever the
    # length of its shard of the per-mode diagonal (uneven mode slabs; the reference gets there with ParBroadcasted, src/ParMatrix.jl:21)
    theta = parops.init(mix, rng=2)
    theta.update(parops.init(ew, rng=[2, rank]))
    theta = parops.gpu(theta)
    St = S(theta)
    return St, St.adjoint(), S.Domain, S.Range


def distributed_check(parops, eng, world, rank, shape=None):
    """One-off correctness check beside the perf number: the distributed layer (same kernel families as the timed
    C4 shapes: row ring / staged inverse, pruned y-DFT, NVLink repartition, fused z-mix-z) on a reduced global grid
    against the SERIAL oracle on rank 0.  Returns {"shape", "rel_l2_fwd", "rel_l2_adj"} on rank 0."""
    import torch
    import torch.distributed as dist
    from oracle import parops_oracle as O
    nz = 32 if world <= 4 else 8 * world
    shape = shape or [64, 128, nz, 16]  # nx, ny, nz, nt
    St, Sadj, dom, ran = build_distributed(parops, shape, world, rank)
    nx, ny, nzz, nt = shape
    gs = (CHANNELS, nx, ny, nzz, nt)
    dims = [1, 1, 1, world, 1]
    nzl = [parops.local_size(nzz, r, world) for r in range(world)]
    z0 = sum(nzl[:rank])
    # every rank draws the same global arrays and keeps its z-slab
    xg = seeded_normal(int(np.prod(gs)), 0).reshape(gs, order="F")
    yg = seeded_normal(int(np.prod(gs)), 1).reshape(gs, order="F")
    xl = eng.to_device(np.asfortranarray(xg[:, :, :, z0:z0 + nzl[rank], :]).reshape(-1, 1, order="F"))
    yl = eng.to_device(np.asfortranarray(yg[:, :, :, z0:z0 + nzl[rank], :]).reshape(-1, 1, order="F"))
    outs = []

    def gather_to_root(o, sizes):
        """dist.gather needs equal sizes: uneven shards (grids that do not divide by the rank count, the CPU tests) are padded to the largest"""
        if world == 1:
            return [o]
        if len(set(sizes)) == 1:
            parts = [torch.empty(n, device=o.device, dtype=o.dtype) for n in sizes] if rank == 0 else None
            dist.gather(o, parts, dst=0)
            return parts
        big = max(sizes)
        padded = torch.zeros(big, device=o.device, dtype=o.dtype)
        padded[:o.numel()] = o
        parts = [torch.empty(big, device=o.device, dtype=o.dtype) for _ in sizes] if rank == 0 else None
        dist.gather(padded, parts, dst=0)
        return [p[:n] for p, n in zip(parts, sizes)] if rank == 0 else None

    for op, v in ((St, xl), (Sadj, yl)):
        o = (op * v).reshape(-1).contiguous()
        parts = gather_to_root(o, [CHANNELS * nx * ny * n * nt for n in nzl])
        if rank == 0:
            outs.append(np.concatenate([p.cpu().numpy().reshape((CHANNELS, nx, ny, n, nt), order="F") for p, n in zip(parts, nzl)], axis=3))
    # parameters: the channel-mixing matrix is replicated (same seed on every rank); the per-mode diagonal is sharded
    # over the kept t-modes (slowest mode dim), so the serial diagonal is the concatenation of the ranks' shards
    leaves = {type(k).__name__: v for k, v in St_params(St).items()}
    dloc = torch.view_as_real(leaves["ParDiagonal"].reshape(-1).contiguous()).reshape(-1)  # (re, im) pairs: gather has no complex types
    kts = [parops.local_size(MODES, r, world) for r in range(world)]
    dparts = gather_to_root(dloc, [2 * (2 * MODES) ** 3 * k for k in kts])
    res = None
    if rank == 0:
        # examples/fno.jl:12-13 builds Kron(F_t, F_shape[1], F_shape[2], ...): its FIRST spatial entry is the SLOWEST array dim,
        # while build_distributed's [nx, ny, nz, nt] names x the fastest -- hand the serial operator the reversed spatial list
        oshape = [nzz, ny, nx, nt]
        So, parts_o = O.spectral_convolution(O.F32, oshape, modes_for(oshape), CHANNELS)
        theta_o = {parts_o["mix_channels"]: np.asfortranarray(leaves["ParMatrix"].cpu().numpy()),
                   parts_o["elementwise_weighting"]: np.concatenate([d.cpu().numpy().view(np.complex64) for d in dparts])}
        y_ref = So(xg.reshape(-1, 1, order="F"), theta_o).reshape(gs, order="F")
        x_ref = So.adjoint()(yg.reshape(-1, 1, order="F"), theta_o).reshape(gs, order="F")
        res = {"shape": shape, "rel_l2_fwd": float(np.linalg.norm(outs[0] - y_ref) / np.linalg.norm(y_ref)),
               "rel_l2_adj": float(np.linalg.norm(outs[1] - x_ref) / np.linalg.norm(x_ref)), "against": "serial oracle on rank 0",
               "tolerance": 1e-5}
        if not (res["rel_l2_fwd"] <= 1e-5 and res["rel_l2_adj"] <= 1e-5):
            raise SystemExit("distributed layer disagrees with the serial oracle: %r" % res)
    if world > 1:
        dist.barrier()
    return res


def St_params(St):
    """{leaf operator: parameter tensor} of a parameterised tree."""
    out = {}

    def walk(op):
        if hasattr(op, "leaf_and_tensor"):
            leaf, _, t = op.leaf_and_tensor()
            out[leaf] = t
            return
        for c in op.children():
            walk(c)
    walk(St)
    return out


def run_ours(args):
    import torch
    import parops

    n = args.gpus
    distributed = "RANK" in os.environ and int(os.environ.get("WORLD_SIZE", "1")) > 1
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if n > 1 and not distributed:
        raise SystemExit("--gpus %d needs torchrun (one process per GPU)" % n)
    torch.cuda.set_device(local_rank)
    eng = parops.default_engine()
    if distributed:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        parops.init_comm(eng)
        shape = C4_LADDER[world]
        dist_check = None if args.no_check else distributed_check(parops, eng, world, rank)
        St, Sadj, dom, ran = build_distributed(parops, shape, world, rank)
    else:
        dist_check = None
        shape = C3_SHAPE if args.workload == "c3" else C4_LADDER[1]
        if args.shape:  # experiments only: nx,ny,nz,nt
            shape = [int(v) for v in args.shape.split(",")]
        St, Sadj, dom, ran = build_single(parops, shape)
    pts_global = int(np.prod(shape))

    x = torch.from_numpy(seeded_normal(dom, 0, rank)).to("cuda").reshape(1, dom).t()
    y2 = torch.from_numpy(seeded_normal(ran, 1, rank)).to("cuda").reshape(1, ran).t()

    plan_f, _ = parops.get_plan(St, 1, eng)
    plan_a, _ = parops.get_plan(Sadj, 1, eng)

    def step():
        yf, back = parops.rrule(St, x)        # forward apply (taped)
        xa = Sadj * y2                        # adjoint apply  S(θ)' * y2
        grads, xbar = back(y2)                # parameter gradient + input cotangent (Zygote pullback)
        return yf, xa, grads, xbar

    def barrier():
        if distributed:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    plan_f.set_profiling(True)
    plan_a.set_profiling(True)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = eng.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    launches = eng.launches - l0
    prof = plan_f.step_profile() + plan_a.step_profile() + [r for r in plan_f.pullback_profile() if r[2] > 0]
    plan_f.set_profiling(False)
    plan_a.set_profiling(False)
    if distributed:
        import torch.distributed as dist
        t = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = pts_global / (ms_step / 1e3)

    # ---- roofline of the dominant kernel (largest share of the step)
    peak, peak_src = peaks()
    dom_k = max(prof, key=lambda r: r[1]) if prof else None
    roof = None
    if dom_k and dom_k[2] > 0:
        avg_ms = dom_k[1] / dom_k[2]
        achieved = dom_k[3] / (avg_ms / 1e3) / 1e9
        # DRAM traffic cannot be measured outside a profiler: null here; the ncu --set full capture of this kernel
        # (dram__bytes_read.sum + dram__bytes_write.sum per launch) is committed under profiles/ and named in DESIGN.md
        traffic = None
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "kernel": dom_k[0], "avg_ms": avg_ms, "algorithmic_bytes": dom_k[3], "peak_source": peak_src,
                "share_of_step": dom_k[1] / max(1e-9, sum(r[1] for r in prof)),
                "step_algorithmic_bytes": 2 * plan_f.algorithmic_bytes() + plan_a.algorithmic_bytes(),
                "step_frac": (2 * plan_f.algorithmic_bytes() + plan_a.algorithmic_bytes()) / (ms_step / 1e3) / 1e9 / peak,
                "kernels": [{"kernel": r[0], "avg_ms": r[1] / max(1, r[2]), "GBps": (r[3] / (r[1] / max(1, r[2]) / 1e3) / 1e9) if r[1] > 0 else None} for r in prof]}

    # ---- e2e: the same step with HOST (pinned) buffers: every step copies its inputs host->device and
    # reads every result (y, x~, xbar, parameter gradients) back device->host inside the timed region.
    # The forward and adjoint applies go through the C-ABI host call (po_plan_apply_host); the taped
    # forward + pullback need device residency for the tape, so they use explicit pinned copies.
    e2e = None
    if not args.no_e2e:
        xh = torch.empty(dom, dtype=torch.float32).pin_memory()
        yh2 = torch.empty(ran, dtype=torch.float32).pin_memory()
        xh.copy_(x.reshape(-1))
        yh2.copy_(y2.reshape(-1))
        y2n = yh2.numpy().reshape(ran, 1, order="F")
        out_y = torch.empty(ran, dtype=torch.float32).pin_memory()
        out_xb = torch.empty(dom, dtype=torch.float32).pin_memory()
        out_xa = torch.empty(dom, dtype=torch.float32).pin_memory()
        xan = out_xa.numpy().reshape(dom, 1, order="F")                        # host result buffer of the adjoint apply

        # Three independent host<->device streams: the forward (taped) apply, the adjoint apply through the C-ABI host
        # call, and the upload of the cotangent; PCIe is full duplex, so the D2H of one overlaps the H2D of the next.
        # (single-GPU only: with a communicator every exchange stays on one stream, in one order on all ranks)
        if distributed:
            s_fwd = s_adj = s_up = torch.cuda.current_stream()
        else:
            s_fwd, s_adj, s_up = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()

        gpin = {}  # pinned host buffers of the parameter gradients (allocated on first use)

        def e2e_step():
            """Schedule (PCIe is full duplex; each apply is a GLOBAL transform, so its D2H cannot start before its H2D has ended):
            H2D queue  x | ybar | y2          D2H queue  . | y | xbar | x~
            Everything asynchronous is enqueued BEFORE the one blocking call (the C-ABI host apply of the adjoint, which waits for
            its own stream): with whole-tensor transfers the floor is one H2D + three D2H = 4 x 671 MB at the link rate."""
            cur = torch.cuda.current_stream()
            for st_ in (s_fwd, s_adj, s_up):
                st_.wait_stream(cur)
            with torch.cuda.stream(s_fwd):
                xd = xh.to("cuda", non_blocking=True).reshape(1, dom).t()      # H2D x
                yf, back = parops.rrule(St, xd)
                out_y.copy_(yf.reshape(-1), non_blocking=True)                 # D2H y
            with torch.cuda.stream(s_up):
                ybd = yh2.to("cuda", non_blocking=True).reshape(1, ran).t()    # H2D cotangent
                ev_up = torch.cuda.Event()
                ev_up.record(s_up)
            with torch.cuda.stream(s_fwd):
                s_fwd.wait_event(ev_up)
                ybd.record_stream(s_fwd)
                grads, xbar = back(ybd)
                out_xb.copy_(xbar.reshape(-1), non_blocking=True)              # D2H xbar
                for k, v in grads.items():                                     # D2H parameter gradients
                    if k not in gpin:
                        gpin[k] = torch.empty(v.shape, dtype=v.dtype).pin_memory()
                    gpin[k].copy_(v, non_blocking=True)
            with torch.cuda.stream(s_adj):
                xa = parops.apply_operator(Sadj, y2n, out=xan)                 # H2D y2, adjoint apply, D2H x~ (po_plan_apply_host; waits for its stream)
            cur.wait_stream(s_fwd)
            cur.wait_stream(s_adj)
            torch.cuda.synchronize()
            return xa, gpin

        def e2e_pipeline(k_steps):
            """The same per-step work as e2e_step, software-pipelined ACROSS steps (single GPU): nothing blocks the host inside a
            step (the adjoint goes through po_plan_apply_host_async), the upload of the next step's x is enqueued behind this step's
            cotangent uploads, and the host only waits for step k-1 before it enqueues step k+1 (two steps in flight).
            H2D queue  ybar_k | y2_k | x_k+1        D2H queue  y_k | xbar_k | x~_k
            so both directions of the link carry three 671 MB tensors per step at the same time: floor = 3 x 671 MB at the link rate
            (+ the first upload, which nothing overlaps).  Every step still uploads all its inputs from pinned host memory and reads
            all its results back inside the timed region."""
            cur = torch.cuda.current_stream()
            for st_ in (s_fwd, s_adj, s_up):
                st_.wait_stream(cur)

            def upload_x():
                with torch.cuda.stream(s_up):
                    xd_ = xh.to("cuda", non_blocking=True)
                    ev_ = torch.cuda.Event()
                    ev_.record(s_up)
                return xd_, ev_

            nxt = upload_x()
            done = []
            for k in range(k_steps):
                xd, ev_x = nxt
                with torch.cuda.stream(s_fwd):
                    s_fwd.wait_event(ev_x)
                    xd.record_stream(s_fwd)
                    yf, back = parops.rrule(St, xd.reshape(1, dom).t())            # taped forward apply
                    out_y.copy_(yf.reshape(-1), non_blocking=True)                 # D2H y
                with torch.cuda.stream(s_up):
                    ybd = yh2.to("cuda", non_blocking=True).reshape(1, ran).t()    # H2D cotangent
                    ev_up = torch.cuda.Event()
                    ev_up.record(s_up)
                with torch.cuda.stream(s_adj):
                    parops.apply_operator(Sadj, y2n, out=xan, wait=False)          # H2D y2, adjoint apply, D2H x~: enqueued only
                    ev_adj = torch.cuda.Event()
                    ev_adj.record(s_adj)
                if k + 1 < k_steps:
                    nxt = upload_x()                                               # H2D x of the NEXT step
                with torch.cuda.stream(s_fwd):
                    s_fwd.wait_event(ev_up)
                    ybd.record_stream(s_fwd)
                    grads, xbar = back(ybd)
                    out_xb.copy_(xbar.reshape(-1), non_blocking=True)              # D2H xbar
                    for kk, v in grads.items():                                    # D2H parameter gradients
                        if kk not in gpin:
                            gpin[kk] = torch.empty(v.shape, dtype=v.dtype).pin_memory()
                        gpin[kk].copy_(v, non_blocking=True)
                    ev_fwd = torch.cuda.Event()
                    ev_fwd.record(s_fwd)
                done.append((ev_fwd, ev_adj))
                if k >= 1:                                                         # at most two steps in flight
                    done[k - 1][0].synchronize()
                    done[k - 1][1].synchronize()
            cur.wait_stream(s_fwd)
            cur.wait_stream(s_adj)
            cur.wait_stream(s_up)
            torch.cuda.synchronize()
            eng.raise_if_failed()
            return xan, gpin

        # one-off check that the overlapped host path returns what the device-resident path returns
        yf_ref, back_ref = parops.rrule(St, x)
        _, xbar_ref = back_ref(y2)
        xa_ref = Sadj * y2
        torch.cuda.synchronize()
        refs = (xa_ref.reshape(-1).cpu(), yf_ref.reshape(-1).cpu(), xbar_ref.reshape(-1).cpu())
        del yf_ref, back_ref, xbar_ref, xa_ref

        def host_path_errors(xa_host):
            got = (torch.from_numpy(np.ascontiguousarray(xa_host.reshape(-1))), out_y, out_xb)
            return {name: float((g - r).norm() / r.norm()) for g, r, name in zip(got, refs, ("x~", "y", "xbar"))}

        def clear_results():
            out_y.zero_(); out_xb.zero_(); out_xa.zero_()

        xa_chk, _ = e2e_step()
        errs = host_path_errors(xa_chk)
        if not all(v <= 1e-5 for v in errs.values()):
            raise SystemExit("e2e host path disagrees with the device path: %r" % errs)
        pipelined, pipe_note = False, None
        if not distributed and not args.no_e2e_pipeline:
            try:
                clear_results()
                xa_chk, _ = e2e_pipeline(3)
                errs_p = host_path_errors(xa_chk)
                pipelined = all(v <= 1e-5 for v in errs_p.values())
                if not pipelined:
                    pipe_note = "pipelined host path failed its check (%r): per-step schedule timed instead" % errs_p
            except Exception as e:  # noqa: BLE001 -- fall back to the per-step schedule, never lose the line
                pipe_note = "pipelined host path raised %r: per-step schedule timed instead" % (e,)
                torch.cuda.synchronize()

        ksteps = max(1, min(args.steps, 20 if pipelined else 5))
        if pipelined:
            barrier()
            t0 = time.perf_counter()
            e2e_pipeline(ksteps)
            dt = time.perf_counter() - t0
            # the per-step (unpipelined) schedule beside it, for the record
            t1 = time.perf_counter()
            for _ in range(2):
                e2e_step()
            torch.cuda.synchronize()
            dt_unpiped = (time.perf_counter() - t1) / 2
        else:
            e2e_step()
            barrier()
            t0 = time.perf_counter()
            for _ in range(ksteps):
                e2e_step()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            dt_unpiped = None
        if distributed:
            import torch.distributed as dist
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        h2d, d2h = int((dom + 2 * ran) * 4), int((ran + 2 * dom) * 4 + 2 * 8 * (CHANNELS * CHANNELS + 32768))
        e2e = {"value": pts_global / (dt / ksteps), "unit": "grid-points/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "steps": ksteps, "ms_per_step": dt / ksteps * 1e3,
               "link_GBps_per_direction": max(h2d, d2h) / (dt / ksteps) / 1e9,
               "schedule": "pipelined across steps" if pipelined else "per step",
               "note": ("pinned host buffers: H2D x, y2, ybar; taped forward + adjoint apply (po_plan_apply_host_async) + pullback; D2H y, x~, xbar, "
                        "parameter gradients.  Steps are software-pipelined over three streams (H2D queue ybar_k|y2_k|x_k+1 against D2H queue "
                        "y_k|xbar_k|x~_k, two steps in flight, the first upload and the last download inside the timed region): floor = 3 x 671 MB "
                        "per direction per step at the link rate; results checked once against the device-resident path") if pipelined else
                       ("pinned host buffers: H2D x, y2, ybar; taped forward + adjoint apply (po_plan_apply_host) + pullback; "
                        "D2H y, x~, xbar, parameter gradients; three streams, all asynchronous work enqueued before the one blocking host "
                        "apply (H2D queue x|ybar|y2 against D2H queue y|xbar|x~ on the full-duplex link: floor = 4 x 671 MB of whole-tensor transfers, "
                        "each apply being a global transform); results checked once against the device-resident path")}
        if dt_unpiped is not None:
            e2e["per_step_schedule_ms"] = dt_unpiped * 1e3
        if pipe_note:
            e2e["pipeline_note"] = pipe_note

    # ---- weak-scaling base: the C4 ladder's own N = 1 rung (64^4, 2^24 points per GPU like every rung above it), so that
    # per-GPU efficiency at N GPUs can be computed against the SAME per-GPU work instead of against C3 (half the points)
    weak_base = None
    if not distributed and args.workload == "c3" and not args.shape and not args.no_weak_base:
        del x, y2
        torch.cuda.empty_cache()
        bshape = C4_LADDER[1]
        Stb, Sadjb, domb, ranb = build_single(parops, bshape)
        xb_ = torch.from_numpy(seeded_normal(domb, 0)).to("cuda").reshape(1, domb).t()
        yb_ = torch.from_numpy(seeded_normal(ranb, 1)).to("cuda").reshape(1, ranb).t()

        def bstep():
            yf, back = parops.rrule(Stb, xb_)
            xa = Sadjb * yb_
            return yf, xa, back(yb_)

        for _ in range(3):
            bstep()
        torch.cuda.synchronize()
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0.record()
        for _ in range(args.steps):
            bstep()
        b1.record()
        torch.cuda.synchronize()
        bms = b0.elapsed_time(b1) / args.steps
        weak_base = {"shape": bshape, "grid_points": int(np.prod(bshape)), "ms_per_step": bms, "value": int(np.prod(bshape)) / (bms / 1e3),
                     "unit": "grid-points/s", "note": "N = 1 rung of the C4 weak-scaling ladder (same 2^24 points per GPU as N = 2, 4, 8): "
                     "weak-scaling efficiency at N = value(N) / (N * this value)"}
        del xb_, yb_, Stb, Sadjb

    # ---- secondary BASELINE.json configs (C1 latency, C2, C5 arms, C3(ii) ParTensor, block epilogue) in the driver-visible line
    secondary = None
    if not distributed and args.workload == "c3" and not args.shape and not args.no_secondary:
        try:  # in a process of its own: it captures CUDA graphs and switches kernel arms through the environment
            torch.cuda.empty_cache()
            out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "bench_configs.py")], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                                 timeout=300, text=True)
            if out.returncode != 0:
                raise RuntimeError("scripts/bench_configs.py exited %d: %s" % (out.returncode, out.stderr[-300:]))
            secondary = json.loads(out.stdout[out.stdout.index("{"):])
            for v in secondary.values():  # keep the line readable: per-kernel lists stay in profiles/
                for k in [k for k in v if k.startswith("kernels") or k.endswith("_kernels")]:
                    del v[k]
        except Exception as e:  # noqa: BLE001 -- a secondary config must never cost the headline line
            secondary = {"error": repr(e)[:400]}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        sec, workers = cpu_pair_time(CPU_SAMPLE_SHAPE)
        spts = int(np.prod(CPU_SAMPLE_SHAPE))
        cpu = {"value": spts / sec, "unit": "grid-points/s", "cores": workers, "kind": "port",
               "sample": "same layer (c=20, 8 modes/dim, Float32, batch 1) on a %s sub-grid (1/%d of the grid points), one fwd+adjoint "
                         "pair, %.1f s" % ("x".join(map(str, CPU_SAMPLE_SHAPE)), pts_global // spts, sec)}

    if rank == 0:
        line = {
            "metric": "FNO spectral-conv fwd+adjoint grid-points/sec", "value": value, "unit": "grid-points/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(world),
            "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "inputs": "numpy default_rng PCG64 standard normal, Float64 cast to Float32: x seed 0, cotangent seed 1, parameters seed 2 (BASELINE.md 4)",
        }
        if weak_base is not None:
            line["weak_base"] = weak_base
        if secondary is not None:
            line["secondary"] = secondary
        if dist_check is not None:
            line["dist_check"] = dist_check
        emit(line)
    if distributed:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    return 0


def _claim_stdout():
    """The driver reads ONE JSON line from stdout; NCCL / torchrun print banners there.  Send
    everything else to stderr and keep the real stdout for the result line."""
    real = os.dup(1)
    sys.stdout.flush()
    os.dup2(2, 1)
    return os.fdopen(real, "w")


RESULT_OUT = None


def emit(line: dict):
    RESULT_OUT.write(json.dumps(line) + "\n")
    RESULT_OUT.flush()


def main():
    global RESULT_OUT
    RESULT_OUT = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=["c3", "c4"])
    ap.add_argument("--shape", default="", help="experiments only: override the single-GPU grid nx,ny,nz,nt")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-e2e-pipeline", action="store_true", help="time the per-step host schedule instead of the cross-step pipeline")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-check", action="store_true", help="skip the one-off distributed-vs-serial-oracle check (N > 1)")
    ap.add_argument("--no-secondary", action="store_true", help="skip the secondary BASELINE.json configs (N = 1 only)")
    ap.add_argument("--no-weak-base", action="store_true", help="skip timing the C4 ladder's N = 1 rung (N = 1 only)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
