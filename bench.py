#!/usr/bin/env python
"""Headline benchmark: EI candidate evaluations / second (value + gradient) at N=2048 observations, D=16,
plus Gram + Cholesky milliseconds at the same size (BASELINE.json `metric`, config 4 of `configs`).

  python bench.py --gpus N --steps K --warmup W            our arm (one process per GPU; torchrun for N > 1)
  python bench.py --impl reference --steps K --warmup W    the reference's own CPU code (oracle/_ref) on host cores

A step = one sweep of `--candidates` candidate points per GPU through the hot path (k*, mu, sigma^2, EI and their
gradients for every candidate, then the arg-max; for N > 1 an NCCL all-gather of the per-rank (value, index) pair).
`value` times the sweep with candidates and results resident in HBM; `e2e` times the same sweep through the
host-buffer C-ABI call (slsgp_acq_batch) with pinned host candidates in and values + gradients back.
The headline sweep arithmetic is --mode (default "tensor": tcgen05 split-fp16, parity 1e-3); `modes` in the JSON line
holds a short run of every mode with its measured error against the FP64 sweep.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tools")]
import synth  # noqa: E402

N_OBS, DIM = 2048, 16
KERNEL_SE, ACQ_EI = 0, 0
NOISE = 0.005
METRIC = "ei_candidate_evals_per_sec"
UNIT = "evals/s"
WORKLOAD = "config4: EI acquisition sweep (value + gradient), N=2048 obs, D=16, ARD-SE kernel, U[0,1]^16 candidates"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["graft", "reference"], default="graft")
    ap.add_argument("--mode", choices=["tensor", "tensor_x2", "tensor_x1", "fp64"], default="tensor",
                    help="sweep arithmetic of the headline number (include/slsgp.h slsgp_sweep_mode)")
    ap.add_argument("--candidates", type=int, default=0, help="candidates per GPU per step (default 2^20; 2^18 for fp64)")
    ap.add_argument("--e2e-candidates", type=int, default=0)
    ap.add_argument("--no-mode-table", action="store_true", help="skip the short runs of the other sweep modes")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def fixed_model():
    X = synth.make_X(N_OBS, DIM, "uniform", seed=1)
    theta = synth.make_theta(DIM, "default")
    y = synth.make_y(X, seed=3)
    return X, theta, y


# ----------------------------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md "clocks DURING the timed region")
# ----------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.idx, self.rows, self.proc = device_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                smax = float(r[2])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------------
# CPU legs (the only places that execute oracle/): bounded samples of the same workload on the host cores
# ----------------------------------------------------------------------------------------------------------------
_W = {}


def _ref_worker_init():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import support as S
    X, theta, y = fixed_model()
    ref = S.Ref()
    h = ref.gpr_create(KERNEL_SE, X, y, theta, NOISE)  # GaussianProcessRegressor(X, y, theta, b), unmodified reference
    _W.update(ref=ref, h=h, reg=ref.gpr_regressor(h))
    # a small model of the same kind for the warm-up steps: one full-size evaluation costs 11-22 s of CPU time, and a CPU
    # code has no clocks or caches worth 3 x 22 s of warming; the TIMED steps are always full size
    Xs, ys = X[:, :256].copy(order="F"), y[:256].copy()
    hs = ref.gpr_create(KERNEL_SE, Xs, ys, theta, NOISE)
    _W.update(h_small=hs, reg_small=ref.gpr_regressor(hs))


def _ref_worker_warm(seed):
    x = synth.make_queries(1, DIM, seed=500 + seed)[:, 0]
    return float(_W["ref"].acq(_W["reg_small"], ACQ_EI, 1.0, x)[0])


def _ref_worker_eval(seed):
    """One EI candidate evaluation as the reference does it: CalcAcquisitionValue + CalcAcquisitionValueDerivative
    (src/acquisition-function.cpp:170-230), each recomputing f_best through N PredictMu calls."""
    x = synth.make_queries(1, DIM, seed=1000 + seed)[:, 0]
    t0 = time.perf_counter()
    v, g = _W["ref"].acq(_W["reg"], ACQ_EI, 1.0, x)
    return time.perf_counter() - t0, float(v)


def reference_available():
    return os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libsls_ref_probe.so"))


def run_reference_pool(steps, warmup, workers):
    """Each step: every worker process performs ONE reference evaluation (value + gradient) at N=2048, D=16."""
    from concurrent.futures import ProcessPoolExecutor
    import multiprocessing as mp
    with ProcessPoolExecutor(max_workers=workers, mp_context=mp.get_context("fork"), initializer=_ref_worker_init) as ex:
        for w in range(max(warmup, 1)):  # forces the initialisers (model build is not timed); warm-up on the small model
            list(ex.map(_ref_worker_warm, range(workers)))
        t0 = time.perf_counter()
        for s in range(steps):
            list(ex.map(_ref_worker_eval, range(100 * s, 100 * s + workers)))
        dt = time.perf_counter() - t0
    return dt


def port_baseline(seconds=12.0):
    """Fallback when oracle/_ref is absent: the plain-C port with alpha / f_best cached (algorithm-equivalent CPU)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import support as S
    X, theta, y = fixed_model()
    o = S.Oracle()
    m = o.model(KERNEL_SE, X, theta, NOISE, y)
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        o.acq_batch(m, ACQ_EI, 1.0, 0.9, synth.make_queries(64, DIM, seed=n))
        n += 64
    return n / (time.perf_counter() - t0), n


def cpu_baseline_block():
    if reference_available():
        _ref_worker_init()
        dt, _ = _ref_worker_eval(0)
        return {"value": 1.0 / dt, "unit": UNIT, "cores": 1, "kind": "reference",
                "sample": "1 EI value+gradient evaluation (CalcAcquisitionValue + ...Derivative of the unmodified "
                          "reference, GaussianProcessRegressor at N=2048, D=16; O(N^3) per evaluation as written), "
                          f"{dt:.1f} s; linked against include/eigen-lite because Eigen is not installed"}
    v, n = port_baseline()
    return {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"{n} candidates through oracle/slsgp_oracle.c with alpha and f_best cached"}


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if not reference_available():
        v, n = port_baseline(20.0)
        kind, cores, sample, dt, steps = "port", 1, f"{n} candidates through the plain-C port", n / v, args.steps
        value = v
    else:
        cores = min(os.cpu_count() or 1, 64)
        dt = run_reference_pool(args.steps, args.warmup, cores)
        value = args.steps * cores / dt
        kind = "reference"
        sample = (f"each step = {cores} worker processes x 1 EI value+gradient evaluation of the unmodified reference "
                  "(oracle/_ref, GaussianProcessRegressor N=2048 D=16); warm-up steps use a 256-point model of the same kind")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------------
def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def main_graft(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; libslsgp has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        # keep stdout to the one JSON line: NCCL prints its version banner there when NCCL_DEBUG=VERSION
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    pkg = importlib.import_module("sequential-line-search_b200")
    ctx = pkg.Context(local)
    stream = torch.cuda.Stream(device=local)
    ctx.set_stream(stream.cuda_stream)

    X, theta, y = fixed_model()
    MODES = {"fp64": pkg.SWEEP_FP64, "tensor": pkg.SWEEP_TENSOR, "tensor_x2": pkg.SWEEP_TENSOR_X2, "tensor_x1": pkg.SWEEP_TENSOR_X1}
    PASSES = {"fp64": 1, "tensor": 3, "tensor_x2": 2, "tensor_x1": 1}
    M = args.candidates or ((1 << 18) if args.mode == "fp64" else (1 << 20))
    Me = args.e2e_candidates or M // 2
    if args.mode == "fp64":
        gemm_key, kernels = "sweep_gemm", ("sweep_gemm", "sweep_kstar", "sweep_reduce", "sweep_grad_gemm", "sweep_finish")
    else:
        gemm_key, kernels = "tc_gemm", ("tc_gemm", "tc_kstar", "sweep_finish")

    # ---- model build: Gram + Cholesky (+ inverse, alpha), timed per phase with CUDA events (median of 7 after 3)
    ctx.set_data(X)
    phases = {k: [] for k in ("gram", "factor", "inverse", "alpha")}
    for it in range(10):
        ctx.gram(KERNEL_SE, theta, NOISE, want=False)
        ctx.factor()
        ctx.inverse(want=False)
        ctx.solve_alpha(y)
        if it >= 3:
            for k in phases:
                phases[k].append(ctx.phase_ms(k))
    aux = {k + "_ms": float(np.median(v)) for k, v in phases.items()}
    aux["gram_chol_ms"] = aux["gram_ms"] + aux["factor_ms"]

    # ---- device-resident candidates: 4 distinct batches per rank, rotated, so no step re-reads the previous inputs
    nb = 4
    gen = torch.Generator(device="cuda").manual_seed(1234 + rank)
    Xq = [torch.rand((M, DIM), dtype=torch.float64, device="cuda", generator=gen) for _ in range(nb)]
    val = torch.empty(M, dtype=torch.float64, device="cuda")
    grad = torch.empty((M, DIM), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()

    def step(i, m=M):
        q = Xq[i % nb]
        ctx.acq_batch_device(ACQ_EI, 1.0, q.data_ptr(), m, d_val=val.data_ptr(), d_grad=grad.data_ptr())
        v, idx = ctx.argmax_device(val.data_ptr(), m, index0=rank * M)  # synchronises this rank's stream
        if world > 1:  # the only collective on the path: (value, index) of every rank's winner
            return pkg.sharding.all_gather_winner(v, idx, device="cuda")
        return v, idx

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- the other sweep modes on a short run (rank 0's table; not the headline): throughput + error vs FP64
    mode_table = {}
    if not args.no_mode_table:
        Mt = min(M, 1 << 17)
        ctx.set_sweep_mode(pkg.SWEEP_FP64)
        step(0, Mt)
        ref_val, ref_grad = val[:Mt].clone(), grad[:Mt].clone()
        for name, mode in MODES.items():
            ctx.set_sweep_mode(mode)
            step(0, Mt)
            ev_v = float((val[:Mt] - ref_val).abs().max() / ref_val.abs().max())
            ev_g = float((grad[:Mt] - ref_grad).abs().max() / ref_grad.abs().max())
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            t0.record(stream)
            reps = 2 if name == "fp64" else 6
            for r in range(reps):
                step(r + 1, Mt)
            t1.record(stream)
            torch.cuda.synchronize()
            mode_table[name] = {"evals_per_s": reps * Mt / (t0.elapsed_time(t1) * 1e-3), "mma_passes": PASSES[name],
                                "ei_max_err_vs_fp64": ev_v, "grad_max_err_vs_fp64": ev_g}

    ctx.set_sweep_mode(MODES[args.mode])
    for i in range(args.warmup):
        step(i)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ctx.profile_enable(True)
    for k in kernels:
        ctx.profile_read(k)
    launches0 = ctx.launch_count()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(args.steps):
        best = step(args.warmup + i)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = ctx.launch_count() - launches0
    prof = {k: ctx.profile_read(k) for k in kernels}
    ctx.profile_enable(False)
    clocks = sampler.stop() if rank == 0 else None

    # ---- end to end through the host-buffer C-ABI call: pinned candidates in, values + gradients out
    hq = [torch.rand((Me, DIM), dtype=torch.float64, generator=torch.Generator().manual_seed(77 + rank + 10 * b)).pin_memory()
          for b in range(2)]
    hval = torch.empty(Me, dtype=torch.float64).pin_memory()
    hgrad = torch.empty((Me, DIM), dtype=torch.float64).pin_memory()
    lib, h = ctx.lib, ctx.h
    import ctypes as C
    dpt = C.POINTER(C.c_double)

    def e2e_step(i):
        st = lib.slsgp_acq_batch(h, ACQ_EI, 1.0, C.cast(hq[i % 2].data_ptr(), dpt), Me, C.cast(hval.data_ptr(), dpt),
                                 C.cast(hgrad.data_ptr(), dpt))
        assert st == 0, lib.slsgp_last_error(h)
        return float(hval.max())  # the host reads the step's result

    for i in range(max(args.warmup, 3)):
        e2e_step(i)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record(stream)
    for i in range(args.steps):
        e2e_step(i)
    f1.record(stream)
    barrier()
    ms_e2e = f0.elapsed_time(f1)

    # ---- max over ranks
    if world > 1:
        t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])

    if rank == 0:
        peaks, peak_src = measured_peaks()
        gemm_ms, gemm_n = prof[gemm_key]
        shard = min(M, 16384 if args.mode == "fp64" else 148 * 128 * 2)
        # algorithmic work per candidate = 2 N^2 (beta = K^-1 k*) + 7 N D + 2 N  FLOP
        flop_per_cand = 2.0 * N_OBS * N_OBS + 7.0 * N_OBS * DIM + 2.0 * N_OBS
        cands_per_launch = M * args.steps / max(gemm_n, 1)
        achieved = flop_per_cand * cands_per_launch / (gemm_ms / max(gemm_n, 1) * 1e-3) / 1e12 if gemm_n else None
        peak = peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops"))
        sweep_total = sum(v[0] for v in prof.values())
        if args.mode == "fp64":
            kname = "gemm64_dmma_kernel<NN> (beta = K^-1 k*, mma.sync.m8n8k4.f64)"
            peak, peak_src = 35.46, "measured cuBLAS DGEMM n=8192 on this pool (profiles/r01j_fp64_library_points.txt; datasheet 37; no FP64 figure in MEASURED_PEAKS.json),"
            note = "the kernel runs on the FP64 tensor pipe (DMMA)"
        else:
            kname = (f"tc_sweep_gemm_kernel<20, 2> (tcgen05.mma cta_group::2 kind::f16, 256x256x16, {PASSES[args.mode]} split-fp16 "
                     "pass(es) per pipeline stage, fused epilogue)")
            note = (f"achieved counts ALGORITHMIC flops; the tensor pipe executes {PASSES[args.mode]}x the 2N^2 term "
                    f"(executed ~{PASSES[args.mode] * achieved:.0f} TFLOP/s)") if achieved else ""
        traffic = None
        try:  # dram bytes per launch of the dominant kernel, from the committed ncu --set full capture
            tr = json.load(open(os.path.join(ROOT, "profiles", "r01_roofline_traffic.json")))["tc_sweep_gemm_kernel"]
            if args.mode == "tensor" and cands_per_launch:
                traffic = tr["dram_bytes_per_launch"] * cands_per_launch / tr["candidates_per_launch"]
        except Exception:
            traffic = None
        line = {
            "metric": METRIC, "value": world * M * args.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64" if args.mode == "fp64" else "f16x2-split/f32-acc",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "candidates_per_gpu_per_step": M, "n_obs": N_OBS, "dim": DIM,
                       "sweep_mode": args.mode, "shard_candidates": shard,
                       "collective": "all_gather(value,index) per step" if world > 1 else "none",
                       "l2": "4 rotating candidate batches; per-step working set (candidates, values, gradients, per-shard "
                             "k* operands) exceeds the 126 MB L2; no explicit flush"},
            "gram_chol_ms": aux["gram_chol_ms"], "aux": aux, "modes": mode_table,
            "e2e": {"value": world * Me * args.steps / (ms_e2e * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": Me * DIM * 8, "d2h_bytes_per_step": Me * (DIM + 1) * 8,
                    "candidates_per_gpu_per_step": Me, "api": "slsgp_acq_batch (host buffers)"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": kname,  # fp64 mode: the FP64 tensor pipe
                         "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                         "executed_tflops": (PASSES[args.mode] * achieved) if achieved else None,
                         "executed_frac": (PASSES[args.mode] * achieved / peak) if achieved else None,
                         "peak_source": (f"{peak_src} TFLOP/s" if args.mode == "fp64" else f"{peak_src} bf16 sustained (MEASURED_PEAKS.json)"),
                         "note": note,
                         "launches_timed": gemm_n, "avg_launch_ms": gemm_ms / max(gemm_n, 1),
                         "share_of_sweep": gemm_ms / sweep_total if sweep_total else None,
                         "kernel_ms": {k: v[0] for k, v in prof.items()}},
            "best": {"value": best[0], "index": best[1]},
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_block()
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        main_reference(a)
    else:
        main_graft(a)
