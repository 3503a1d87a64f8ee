"""Secondary configs of BASELINE.json (parity-test cases, measured here for the record):
C1 (2-D FNO 64x64, latency), C2 (truncated-DFT Kron on 128^3 x 16), C5 (FP64 Kron of dense 256^3).
Usage (GPU box): python scripts/bench_configs.py > gpurun_out/configs.json"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import parops
from parops import ComplexF32, ComplexF64, Float32, Float64


def timeit(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def timeit_graph(fn, reps=10, warm=3):
    """Device time per call with the host out of the picture: `reps` calls captured into ONE CUDA graph on a side stream and the
    graph replayed (the library enqueues everything on the caller's stream and never synchronises, so its launches -- PDL edges
    included -- capture as they are).  None if capture is not possible."""
    try:
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        g = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            keep = []
            with torch.cuda.graph(g, stream=side):
                for _ in range(reps):
                    keep.append(fn())
        torch.cuda.synchronize()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / (3 * reps)
        del g, keep
        return ms
    except Exception as e:  # noqa: BLE001
        sys.stderr.write("[bench_configs] CUDA-graph timing unavailable: %r\n" % (e,))
        try:
            torch.cuda.synchronize()
        except Exception:  # noqa: BLE001
            pass
        return None


def profile(op, x, eng):
    plan, _ = parops.get_plan(op, 1 if x.dim() == 1 else x.shape[1], eng)
    plan.set_profiling(True)
    for _ in range(5):
        op * x
    rows = plan.step_profile()
    plan.set_profiling(False)
    return [{"kernel": r[0], "avg_ms": r[1] / max(1, r[2]), "GBps": (r[3] / (r[1] / max(1, r[2]) / 1e3) / 1e9) if r[1] > 0 else None} for r in rows]


def bench_c2(eng, g):
    """C2 with the fused factor-pair kernels (po_c2.cuh) and, for the record, with PO_C2_FUSE=0 (three kernels per direction).
    `*_ms` is device time per apply from a CUDA-graph replay of 10 applies (no host in the loop); `*_call_ms` the Python-call loop."""
    K = parops.ParKron(parops.compose(parops.ParRestriction(ComplexF32, 65, [(1, 16)]), parops.ParDFT(Float32, 128)),
                       parops.ParDFT(ComplexF32, 128), parops.ParDFT(ComplexF32, 128))
    x = torch.randn((16, K.Domain), generator=g, device="cuda", dtype=torch.float32).t()
    y = K * x
    Ka = K.adjoint()
    pts = 128 ** 3 * 16
    alg = 134.2e6 + 33.6e6
    res = {"workload": "(R[1:16] o rDFT128) (x) DFT128 (x) DFT128, 128^3 x batch 16, Float32"}
    prev = os.environ.get("PO_C2_FUSE")
    for fuse, tag in (("1", ""), ("0", "unfused_")):
        os.environ["PO_C2_FUSE"] = fuse
        call_f, call_a = timeit(lambda: K * x), timeit(lambda: Ka * y)
        ms_f, ms_a = timeit_graph(lambda: K * x), timeit_graph(lambda: Ka * y)
        how = "cuda graph replay"
        if ms_f is None or ms_a is None:
            ms_f, ms_a, how = call_f, call_a, "python call loop"
        res.update({tag + "fwd_ms": ms_f, tag + "adj_ms": ms_a, tag + "fwd_call_ms": call_f, tag + "adj_call_ms": call_a, tag + "timing": how,
                    tag + "fwd_frac_hbm": alg / (ms_f / 1e3) / 1e9 / 6553.0, tag + "adj_frac_hbm": alg / (ms_a / 1e3) / 1e9 / 6553.0,
                    tag + "fwd_kernels": profile(K, x, eng), tag + "adj_kernels": profile(Ka, y, eng)})
        if not tag:
            res.update({"fwd_pts_per_s": pts / (ms_f / 1e3), "adj_pts_per_s": pts / (ms_a / 1e3)})
    if prev is None:
        os.environ.pop("PO_C2_FUSE", None)
    else:
        os.environ["PO_C2_FUSE"] = prev
    # A/B of the fused kernels' instantiations (0: run-time row pitch, 1: compile-time pitch, 3: + forward capped at 96 registers); the library default is in po_c2.cuh
    prevv = os.environ.get("PO_C2_VARIANT")
    ab = {}
    for v in (0, 1, 3):
        os.environ["PO_C2_VARIANT"] = str(v)
        ab["variant_%d" % v] = {"fwd_ms": timeit_graph(lambda: K * x), "adj_ms": timeit_graph(lambda: Ka * y)}
    if prevv is None:
        os.environ.pop("PO_C2_VARIANT", None)
    else:
        os.environ["PO_C2_VARIANT"] = prevv
    res["fused_kernel_variants"] = ab
    return res


def _partial(out):
    """BENCH_CONFIGS_PARTIAL=<path>: rewrite the results so far after every section (a run cut short still leaves its numbers)."""
    path = os.environ.get("BENCH_CONFIGS_PARTIAL")
    if path:
        with open(path, "w") as f:
            json.dump(out, f, indent=1)


def _sec_C2(out, eng, g):
    """C2"""
    out["C2"] = bench_c2(eng, g)


def _sec_C1(out, eng, g):
    """C1 (latency)"""
    S = parops.spectral_convolution(Float32, [64, 64], [[(1, 12), (53, 64)], [(1, 12)]], 20)
    th = parops.gpu(parops.init(S, rng=2))
    St = S(th)
    for b in (1, 16):
        x = torch.randn((b, S.Domain), generator=g, device="cuda", dtype=torch.float32).t()
        ms = timeit(lambda: St * x, reps=50)
        ms_g = timeit_graph(lambda: St * x, reps=20)
        out["C1_b%d" % b] = {"workload": "2-D FNO spectral conv 64x64, c=20, 12 modes, batch %d" % b, "apply_ms": ms, "pts_per_s": 4096 * b / (ms / 1e3),
                             "apply_ms_graph_replay": ms_g, "note": "apply_ms = Python call loop (host-bound: 5 launches + ctypes per apply); "
                             "apply_ms_graph_replay = the same 5-kernel chain replayed from a CUDA graph", "kernels": profile(St, x, eng)}


def _sec_C4_local(out, eng, g):
    """the 8-GPU rung's LOCAL shape on one GPU (128 x 128 x 16 x 64 per rank, rows of c * nx = 10 KB): the CTA-pair (cluster + DSMEM) kernels"""
    shape = [16, 128, 128, 64]  # examples/fno.jl order: first spatial entry = slowest array dim (z, the local slab), then y, x (fastest), time last
    modes = [[(1, 8), (sh - 7, sh)] for sh in reversed(shape[:-1])] + [[(1, 8)]]
    Sl = parops.spectral_convolution(Float32, shape, modes, 20)
    thl = parops.gpu(parops.init(Sl, rng=2))
    Slt = Sl(thl)
    xl = torch.randn((1, Sl.Domain), generator=g, device="cuda", dtype=torch.float32).t()
    ms_l = timeit(lambda: Slt * xl, reps=5)
    kl = profile(Slt, xl, eng)
    big = [k for k in kl if k["kernel"].startswith("fused-")]
    out["C4_local_shape_8gpu_rung"] = {"workload": "spectral-conv apply at the per-rank shape of the 8-GPU rung, 128x128x16x64, c = 20, Float32 (single GPU, no exchange)",
                                       "apply_ms": ms_l, "apply_frac_hbm": 2 * Sl.Domain * 4 / (ms_l / 1e3) / 1e9 / 6553.0,
                                       "fused_fwd_tx_ms": big[0]["avg_ms"] if big else None, "fused_fwd_tx_frac_hbm": (big[0]["GBps"] / 6553.0) if big else None,
                                       "fused_inv_xt_ms": big[-1]["avg_ms"] if big else None, "fused_inv_xt_frac_hbm": (big[-1]["GBps"] / 6553.0) if big else None,
                                       "kernels": kl}
    del xl, Slt, thl


def _sec_ParTensor(out, eng, g):
    """C3(ii): per-mode ParTensor weights: ic = oc = 20, M = 32768 modes, 104.9 MB ComplexF32 -- weight-streaming kernels"""
    ws = (20, 20, 16, 16, 16, 8)
    Tn = parops.ParTensor(ComplexF32, tuple(range(1, 7)), ws, tuple(range(2, 7)), (20, 16, 16, 16, 8), (1,) + tuple(range(3, 7)), (20, 16, 16, 16, 8))
    tht = parops.gpu(parops.init(Tn, rng=2))
    Tt = Tn(tht)
    xh = torch.view_as_complex(torch.randn((1, Tn.Domain, 2), generator=g, device="cuda", dtype=torch.float32)).t()
    ms = timeit(lambda: Tt * xh, reps=10)
    wbytes = 20 * 20 * 32768 * 8
    yb = torch.view_as_complex(torch.randn((1, Tn.Range, 2), generator=g, device="cuda", dtype=torch.float32)).t()

    def tpb():
        _, back = parops.rrule(Tt, xh)
        return back(yb)
    ms_pb = timeit(tpb, reps=10)
    kern = profile(Tt, xh, eng)
    kms = kern[0]["avg_ms"]
    gms = timeit_graph(lambda: Tt * xh, reps=10)
    out["C3ii_ParTensor"] = {"workload": "ParTensor rank 6 (oc, ic, 16, 16, 16, 8), ic = oc = 20, ComplexF32, batch 1: y = W x per mode",
                             "kernel_ms_cuda_events": kms, "weight_stream_GBps": wbytes / (kms / 1e3) / 1e9, "frac_hbm": wbytes / (kms / 1e3) / 1e9 / 6553.0,
                             "apply_ms_graph_replay": gms, "frac_hbm_graph_replay": (wbytes / (gms / 1e3) / 1e9 / 6553.0) if gms else None,
                             "python_call_ms": ms, "python_call_plus_pullback_ms": ms_pb,
                             "note": "one ~20 us kernel: the CUDA-event bracket of a single launch adds ~5 us and the Python call ~40 us of host time; "
                                     "ncu device time 20.3 us = 5.2 TB/s = 0.80 of the measured HBM peak (profiles/r2h_partensor_ncu_launches_chunk18KB.csv). "
                                     "pullback = taped apply + xbar (reads W once more) + Wbar (writes 105 MB)",
                             "kernels": kern}


def _sec_block_epilogue(out, eng, g):
    """FNO block epilogue at C3 size: gelu.(S x .+ W x), pointwise branch fused (po_block_epilogue) vs unfused"""
    c, npts = 20, 64 * 64 * 64 * 32
    Wc = parops.channel_mixing(Float32, [64, 64, 64, 32], c)
    thw = parops.gpu(parops.init(Wc, rng=3))
    Wt = Wc(thw)
    Wm = list(thw.values())[0]
    xb = torch.randn((1, c * npts), generator=g, device="cuda", dtype=torch.float32).t()
    sb = torch.randn((1, c * npts), generator=g, device="cuda", dtype=torch.float32).t()
    ms_fused = timeit(lambda: parops.block_epilogue(Wm, xb, sb), reps=5)
    ms_unf = timeit(lambda: parops.engine.fused_add_gelu(sb, Wt * xb), reps=5)
    gb = 3 * c * npts * 4 / 1e9
    out["block_epilogue_C3"] = {"workload": "gelu.(s .+ (I (x) W) x), c=20, 64x64x64x32 grid, Float32", "fused_ms": ms_fused, "unfused_ms": ms_unf,
                                "fused_GBps": gb / (ms_fused / 1e3), "fused_frac_hbm": gb / (ms_fused / 1e3) / 6553.0,
                                "kernels": profile(Wt, xb, eng)}
    del xb, sb


def _sec_C5(out, eng, g):
    """C5: ParDiagonal o (M (x) M (x) M), 256 x 256 Float64 factors -- the sweep BASELINE.json names: SIMT arm vs DMMA (legacy tensor path) arm"""
    # vs the generic tile kernel; FP64 peaks measured on this part: DFMA 34.0 TFLOP/s, DMMA 37.1 TFLOP/s (scripts/microbench/mb_fp64.cu)
    Ms = [parops.ParMatrix(Float64, 256, 256) for _ in range(3)]
    D = parops.ParDiagonal(Float64, 256 ** 3)
    K5 = parops.compose(D, parops.ParKron(*Ms))
    th = parops.gpu(parops.init(K5, rng=2))
    K5t = K5(th)
    K5a = K5t.adjoint()
    x = torch.randn(256 ** 3, generator=g, device="cuda", dtype=torch.float64)
    flops = 3 * 2 * 256 * 256 ** 3
    c5 = {"workload": "ParDiagonal(256^3) o (M(x)M(x)M), M 256x256, Float64, matvec and adjoint", "fp64_peaks_measured_TFLOPs": {"DFMA": 34.0, "DMMA": 37.1}}
    prev = os.environ.get("PO_DGEMM")
    for arm, name in ((1, "simt_128x128_prefetch"), (3, "simt_128x64_prefetch_2cta"), (2, "dmma_m8n8k4_128x128_prefetch"), (0, "generic_64x64_tile")):
        os.environ["PO_DGEMM"] = str(arm)
        ms_f = timeit(lambda: K5t * x, reps=5)
        ms_a = timeit(lambda: K5a * x, reps=5)
        c5[name] = {"matvec_ms": ms_f, "adjoint_ms": ms_a, "matvec_TFLOPs": flops / (ms_f / 1e3) / 1e12, "adjoint_TFLOPs": flops / (ms_a / 1e3) / 1e12,
                    "matvec_frac_of_DFMA_peak": flops / (ms_f / 1e3) / 1e12 / 34.0}
        c5["kernels_" + name] = profile(K5t, x, eng)
    if prev is None:
        os.environ.pop("PO_DGEMM", None)
    else:
        os.environ["PO_DGEMM"] = prev
    out["C5"] = c5
    del x


def _sec_few_modes(out, eng, g, shape=(64, 64, 64, 32), device="cuda"):
    """kept-mode counts below the fused kernels' 8 (+ 8): the widening pass (po_widen) against the 8-mode layer and against the generic kernels"""
    shape = list(shape)
    x = torch.randn((1, 20 * int(np.prod(shape))), generator=g, device=device, dtype=torch.float32).t()
    res = {"workload": "C3-size spectral-conv APPLY (64x64x64x32, c = 20, Float32, batch 1) with k modes per dim; 8 = the benchmark's layer"}
    for k in (8, 4, 2):
        modes = [[(1, k), (sh - k + 1, sh)] for sh in reversed(shape[:-1])] + [[(1, k)]]
        S = parops.spectral_convolution(Float32, shape, modes, 20)
        St = S(parops.gpu(parops.init(S, rng=2)))
        ms = timeit(lambda: St * x, reps=5)
        res["modes_%d_apply_ms" % k] = ms
        res["modes_%d_frac_hbm" % k] = 2 * x.numel() * 4 / (ms / 1e3) / 1e9 / 6553.0
        if k == 4:
            res["modes_4_kernels"] = profile(St, x, eng)
            prev = os.environ.get("PO_WIDEN")
            os.environ["PO_WIDEN"] = "0"  # read per plan; PO_PLAN_NO_PIPELINE gives the generic plan its own cache key
            try:
                plan, _ = parops.get_plan(St, 1, eng, parops._lib.PO_PLAN_NO_PIPELINE)
                msg = timeit(lambda: parops.apply_operator(St, x, eng, parops._lib.PO_PLAN_NO_PIPELINE), reps=3)
                res["modes_4_generic_kernels_apply_ms"] = msg
            finally:
                if prev is None:
                    os.environ.pop("PO_WIDEN", None)
                else:
                    os.environ["PO_WIDEN"] = prev
    out["few_kept_modes"] = res


SECTIONS = [("C2", _sec_C2), ("C1", _sec_C1), ("C4_local", _sec_C4_local), ("ParTensor", _sec_ParTensor), ("block_epilogue", _sec_block_epilogue),
            ("C5", _sec_C5), ("few_modes", _sec_few_modes)]


def run_all(eng=None, quick=False, only=None):
    """Measure the secondary BASELINE.json configs; returns a dict (bench.py embeds it as `secondary` in its JSON line).  Every section runs
    on its own: a failure is recorded under "<section>_error" and the others still report."""
    eng = eng or parops.default_engine()
    out = {}
    g = torch.Generator(device="cuda").manual_seed(0)
    for name, fn in SECTIONS:
        if only and name != only:
            continue
        try:
            fn(out, eng, g)
        except Exception as e:  # noqa: BLE001
            out[name + "_error"] = {"error": repr(e)[:300]}
            try:
                torch.cuda.synchronize()
            except Exception:  # noqa: BLE001
                pass
        _partial(out)
    return out


def main():
    out = run_all(only=(sys.argv[1] if len(sys.argv) > 1 else None))
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
