#!/usr/bin/env python
"""Headline benchmark (BASELINE.json `metric`): EI candidate evaluations / second (value + gradient) at N = 2048 observations,
D = 16, plus Gram + Cholesky milliseconds at the same size - config 4 of BASELINE.json `configs`:
"EI acquisition sweep: N=2048, D=16, 16M candidate points + gradients, sharded 1/2/4/8 GPU".

  python bench.py --gpus N --steps K --warmup W            our arm (one process per GPU; torchrun for N > 1)
  python bench.py --impl reference --steps K --warmup W    the reference's own CPU code (oracle/_ref) on the host cores

A step (our arm) = what one acquisition search costs once the data is on the device: fit the model (Gram, Cholesky, inverse,
alpha, f_best: what a regressor constructor does after its MAP fit; replicated on every rank), evaluate EI and its gradient for
this rank's share of the `--candidates` candidate points (default 2^24 in total = config 4's M; "scaling": "strong": the total is
fixed and split over the ranks; --scaling weak gives every rank the full count), arg-max, and for N > 1 the NCCL all-gather of the
per-rank (value, index) winner. Candidates are resident in HBM when the timed region starts.

One JSON line on rank 0. `value` is the headline arithmetic (--mode, default "tensor": tcgen05 split-fp16 with fp32 accumulation,
north_star's 1e-3 class); `value_fp64` is the same step in IEEE double (north_star's 1e-5 class) with its own `roofline_fp64`;
`e2e*` are the same metric through the host-buffer C ABI (pinned) and through the C++ host layer's CalcAcquisitionValues
(pageable Eigen storage); `rooflines` covers the Gram and Cholesky kernels; `cpu_baseline` holds the reference (EI evaluation and
CalcLargeKY + LLT) and an algorithm-equivalent all-core CPU implementation timed on this box.
"""
from __future__ import annotations

import argparse
import ctypes as C
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
WORKLOAD = "config4: EI acquisition sweep (value + gradient), N=2048 obs, D=16, ARD-SE kernel, 2^24 U[0,1]^16 candidates"
FLOP_PER_CAND = 2.0 * N_OBS * N_OBS + 7.0 * N_OBS * DIM + 2.0 * N_OBS  # 2 N^2 + 7 N D + 2 N = 8.62 MFLOP
BYTES_PER_CAND = 8.0 * (DIM + 1 + DIM)                                  # D doubles in, 1 + D doubles out = 264 B


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["graft", "reference"], default="graft")
    ap.add_argument("--mode", choices=["tensor", "tensor_x2", "tensor_x1", "fp64"], default="tensor",
                    help="sweep arithmetic of the headline number (include/slsgp.h slsgp_sweep_mode)")
    ap.add_argument("--scaling", choices=["strong", "weak"], default="strong")
    ap.add_argument("--candidates", type=int, default=1 << 24, help="candidates per step: in total (strong) or per GPU (weak)")
    ap.add_argument("--fp64-candidates", type=int, default=1 << 20, help="same for the IEEE-double arm")
    ap.add_argument("--e2e-candidates", type=int, default=1 << 22, help="same for the host-buffer arms")
    ap.add_argument("--no-fp64", action="store_true")
    ap.add_argument("--no-mode-table", action="store_true", help="skip the short runs of the other sweep modes")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-pageable", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip configs 1, 2, 3, 5 and the N grid")
    return ap.parse_args()


def fixed_model():
    X = synth.make_X(N_OBS, DIM, "uniform", seed=1)
    theta = synth.make_theta(DIM, "default")
    y = synth.make_y(X, seed=3)
    return X, theta, y


def config_block(total_candidates, world, scaling):
    """The one description of the workload both arms print (the reference arm evaluates a bounded sample of it)."""
    return {"workload": WORKLOAD, "n_obs": N_OBS, "dim": DIM, "kernel": "ARD squared exponential", "acquisition": "expected improvement",
            "candidates_per_step": total_candidates, "outputs": "EI value + gradient per candidate, arg-max",
            "l2": "inputs larger than L2: every step reads its own candidate batch (>= 128 MB per rank) and writes as much; 2 rotating batches"}


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
        sm, smax, power, reasons = [], None, [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                smax = float(r[2])
                power.append(float(r[3]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm), "power_w_median": float(np.median(power)) if power else None}


# ----------------------------------------------------------------------------------------------------------------
# CPU legs (the only places that execute oracle/): bounded samples of the same workload on the host cores.
# The reference is oracle/_ref/libsls_ref_loop_fma.so when present - the unmodified sources compiled -O3 -mavx2 -mfma, the
# fastest of the reference builds here - else libsls_ref_probe.so (-O2). Eigen is not installed in this image: both are linked
# against include/eigen-lite (stated in every `sample`).
# ----------------------------------------------------------------------------------------------------------------
_W = {}
_REF_CANDIDATES = [os.path.join(ROOT, "oracle", "_ref", n) for n in ("libsls_ref_loop_fma.so", "libsls_ref_probe.so")]


def reference_path():
    for p in _REF_CANDIDATES:
        if os.path.exists(p):
            return p
    return None


def reference_build():
    p = reference_path()
    return "-O3 -mavx2 -mfma" if p and p.endswith("_fma.so") else "-O2"


def _ref_lib():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import support as S
    ref = S.Ref.__new__(S.Ref)
    ref.lib = lib = C.CDLL(reference_path())
    for name in ("ref_gpr_create", "ref_gpr_regressor"):
        getattr(lib, name).restype = C.c_void_p
    for name in ("ref_predict_mu", "ref_predict_sigma", "ref_acq_value", "ref_btl", "ref_gram_chol_seconds"):
        getattr(lib, name).restype = C.c_double
    return ref, S


def _ref_worker_init():
    ref, S = _ref_lib()
    X, theta, y = fixed_model()
    h = ref.gpr_create(KERNEL_SE, X, y, theta, NOISE)  # GaussianProcessRegressor(X, y, theta, b), unmodified reference
    _W.update(ref=ref, h=h, reg=ref.gpr_regressor(h))
    # a small model of the same kind for the warm-up steps: one full-size evaluation costs ~10-20 s of CPU time, and a CPU
    # code has no clocks or caches worth 5 x 20 s of warming; the TIMED steps are always full size
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


def reference_gram_chol_ms(repeats=3):
    """CalcLargeKY + Eigen::LLT as PreferenceRegressor's constructor runs them (src/regressor.cpp:61-89,
    src/preference-regressor.cpp:289-290), timed inside the reference library; median of `repeats`."""
    ref, S = _ref_lib()
    X, theta, _ = fixed_model()
    th = S.f64(theta)
    t = [ref.lib.ref_gram_chol_seconds(KERNEL_SE, DIM, N_OBS, S._p(X), S._p(th), C.c_double(NOISE), None) for _ in range(repeats)]
    return float(np.median(t)) * 1e3


def algorithm_equivalent_cpu(seconds=6.0):
    """SURVEY.md 8(d) baseline (B): the SAME algorithm the GPU path runs - K^-1, alpha and f_best cached once per model, candidates
    in batches through BLAS-3 - in numpy / OpenBLAS on all host cores. Returns (EI evaluations / s, candidates done, Gram+Chol ms)."""
    X, theta, y = fixed_model()
    a, il = theta[0], 1.0 / theta[1:]
    Xs = X * il[:, None]
    gram_chol_ms = None
    for _ in range(3):  # best of 3 (the first pass pays for the BLAS thread pool)
        t0 = time.perf_counter()
        sq = (Xs * Xs).sum(0)
        K = a * np.exp(-0.5 * np.maximum(sq[:, None] + sq[None, :] - 2.0 * Xs.T @ Xs, 0.0)) + NOISE * np.eye(N_OBS)
        L = np.linalg.cholesky(K)
        dt = (time.perf_counter() - t0) * 1e3
        gram_chol_ms = dt if gram_chol_ms is None else min(gram_chol_ms, dt)
    Li = np.linalg.solve(L, np.eye(N_OBS))
    Kinv = Li.T @ Li
    alpha = Kinv @ y
    f_best = float(np.max((K - NOISE * np.eye(N_OBS)) @ alpha))
    from math import erf, pi, sqrt
    verf = np.vectorize(erf)
    done, B, t0 = 0, 4096, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        Q = synth.make_queries(B, DIM, seed=done)
        Qs = Q * il[:, None]
        k = a * np.exp(-0.5 * np.maximum((Qs * Qs).sum(0)[None, :] + sq[:, None] - 2.0 * Xs.T @ Qs, 0.0))  # N x B
        u = Kinv @ k                                                                                        # the 2 N^2 term
        mu, s2 = alpha @ k, np.maximum(a - (k * u).sum(0), 0.0)
        sg = np.sqrt(s2)
        w1, w2 = k * alpha[:, None], k * u
        # d k_i / d x = -c k_i (x - X_i) / l^2  (c = 2: the reference's SE x-derivative)
        dmu = -2.0 * (Q * w1.sum(0)[None, :] - X @ w1) * (il * il)[:, None]
        dsg = 2.0 * (Q * w2.sum(0)[None, :] - X @ w2) * (il * il)[:, None] / sg[None, :]
        z = (mu - f_best) / sg
        pdf, cdf = np.exp(-0.5 * z * z) / sqrt(2 * pi), 0.5 * (1.0 + verf(z / sqrt(2.0)))
        ei = (mu - f_best) * cdf + sg * pdf
        dz = (dmu - z[None, :] * dsg) / sg[None, :]
        gei = dmu * cdf + ((mu - f_best) * pdf)[None, :] * dz + dsg * pdf + (sg * (-z * pdf))[None, :] * dz
        done += B
        _ = float(ei.max()) + float(gei[0, 0])
    return done / (time.perf_counter() - t0), done, gram_chol_ms


def cpu_baseline_block():
    cores = os.cpu_count() or 1
    out = {}
    if reference_path():
        _ref_worker_init()
        dt, _ = _ref_worker_eval(0)
        out.update({"value": 1.0 / dt, "unit": UNIT, "cores": 1, "kind": "reference",
                    "sample": "1 EI value+gradient evaluation (CalcAcquisitionValue + ...Derivative of the unmodified reference, "
                              f"GaussianProcessRegressor at N=2048, D=16; O(N^3) per evaluation as written), {dt:.1f} s; built {reference_build()}, "
                              "linked against include/eigen-lite because Eigen is not installed in this image",
                    "gram_chol_ms": reference_gram_chol_ms(),
                    "gram_chol_sample": "CalcLargeKY + Eigen::LLT of the unmodified reference (src/regressor.cpp:61-89, src/preference-regressor.cpp:289-290), "
                                        "median of 3, 1 thread, eigen-lite's LLT"})
    v, n, gc = algorithm_equivalent_cpu()
    out["algorithm_equivalent"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "gram_chol_ms": gc,
                                   "sample": f"{n} candidates in batches of 4096 through numpy/OpenBLAS on all {cores} host threads with K^-1, alpha and f_best "
                                             "cached per model (the algorithm the GPU path runs; SURVEY.md 8(d) baseline B); Gram+Chol = vectorised Gram + LAPACK potrf"}
    if "value" not in out:
        out.update({"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": out["algorithm_equivalent"]["sample"]})
    return out


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    if not reference_path():
        v, n, _ = algorithm_equivalent_cpu(20.0)
        kind, cores, sample, dt, value = "port", os.cpu_count() or 1, f"{n} candidates through the numpy port (oracle/_ref absent)", n / v, v
    else:
        cores = min(os.cpu_count() or 1, 64)
        dt = run_reference_pool(args.steps, args.warmup, cores)
        value = args.steps * cores / dt
        kind = "reference"
        sample = (f"each step = {cores} worker processes x 1 EI value+gradient evaluation of the unmodified reference (oracle/_ref, built "
                  f"{reference_build()} against include/eigen-lite; GaussianProcessRegressor N=2048 D=16): a bounded sample of the 2^24 candidates of the "
                  "workload; warm-up steps use a 256-point model of the same kind")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_block(args.candidates, world, args.scaling),
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
        return json.load(open(p)), "MEASURED_PEAKS.json"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "B200_PROFILING.md fallback"


def traffic_from_capture(kernel, signature):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture, only when it was taken in the configuration
    that is running now (profiles/r02_roofline_traffic.json records the signature); else None."""
    try:
        rec = json.load(open(os.path.join(ROOT, "profiles", "r02_roofline_traffic.json")))[kernel]
    except Exception:
        return None, "no committed capture for this kernel"
    if any(rec.get("signature", {}).get(k) != v for k, v in signature.items()):
        return None, f"committed capture was taken in another configuration ({rec.get('signature')})"
    return rec["dram_bytes_per_launch"], f"from committed ncu capture {rec.get('source')}"


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
    strong = args.scaling == "strong"

    def per_rank(total):  # this rank's share of a step's candidates
        if not strong:
            return total
        base, rem = divmod(total, world)
        return base + (1 if rank < rem else 0)

    M, M64, Me = per_rank(args.candidates), per_rank(args.fp64_candidates), per_rank(args.e2e_candidates)
    Mtot, M64tot, Metot = (args.candidates, args.fp64_candidates, args.e2e_candidates) if strong else (world * args.candidates, world * args.fp64_candidates, world * args.e2e_candidates)
    first_index = rank * (args.candidates // world) + min(rank, args.candidates % world) if strong else rank * args.candidates

    def fit():
        """What a regressor constructor does after its MAP fit (src/preference-regressor.cpp:289-290 + the cached quantities)."""
        ctx.invalidate()  # a context does not recompute matrices it holds: every timed fit starts from the data alone
        ctx.gram(KERNEL_SE, theta, NOISE, want=False)
        ctx.factor()
        ctx.inverse(want=False)
        ctx.solve_alpha(y)

    # ---- model build phases: Gram, Cholesky, inverse, alpha, timed per phase with CUDA events (median of 9 after 3)
    ctx.set_data(X)
    phases = {k: [] for k in ("gram", "factor", "inverse", "alpha")}
    for it in range(12):
        fit()
        if it >= 3:
            for k in phases:
                phases[k].append(ctx.phase_ms(k))
    aux = {k + "_ms": float(np.median(v)) for k, v in phases.items()}
    aux["gram_chol_ms"] = aux["gram_ms"] + aux["factor_ms"]
    if rank == 0:
        aux["cusolver_potrf_ms"] = cusolver_potrf_ms(torch, ctx.gram(KERNEL_SE, theta, NOISE, want=True))

    # ---- device-resident candidates: 2 distinct batches per rank, rotated (each far larger than the 126 MB L2)
    nb = 2
    gen = torch.Generator(device="cuda").manual_seed(1234 + rank)
    Mmax = max(M, M64)
    Xq = [torch.rand((Mmax, DIM), dtype=torch.float64, device="cuda", generator=gen) for _ in range(nb)]
    val = torch.empty(Mmax, dtype=torch.float64, device="cuda")
    grad = torch.empty((Mmax, DIM), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()

    def step(i, m, refit=True):
        if refit:
            fit()
        q = Xq[i % nb]
        ctx.acq_batch_device(ACQ_EI, 1.0, q.data_ptr(), m, d_val=val.data_ptr(), d_grad=grad.data_ptr())
        v, idx = ctx.argmax_device(val.data_ptr(), m, index0=first_index)  # synchronises this rank's stream
        if world > 1:  # the only collective on the path: (value, index) of every rank's winner
            return pkg.sharding.all_gather_winner(v, idx, device="cuda")
        return v, idx

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        """W warm-up calls, then K timed calls bracketed by barrier + synchronize; CUDA events on the library's stream; max over ranks."""
        for i in range(warmup):
            fn(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        out = None
        for i in range(steps):
            out = fn(warmup + i)
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t[0])
        return ms, out

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
                step(r + 1, Mt, refit=False)
            t1.record(stream)
            torch.cuda.synchronize()
            mode_table[name] = {"evals_per_s": reps * Mt / (t0.elapsed_time(t1) * 1e-3), "mma_passes": PASSES[name],
                                "ei_max_err_vs_fp64": ev_v, "grad_max_err_vs_fp64": ev_g, "candidates": Mt}

    # ---- headline arm
    def run_arm(mode_name, m, kernels):
        ctx.set_sweep_mode(MODES[mode_name])
        ctx.profile_enable(True)
        for i in range(args.warmup):
            step(i, m)
        for k in kernels:  # discard the per-kernel records of the warm-up steps
            ctx.profile_read(k)
        launches0 = ctx.launch_count()
        ms, best = timed(lambda i: step(i, m), args.steps, 0)
        launches = ctx.launch_count() - launches0
        prof = {k: ctx.profile_read(k) for k in kernels}
        ctx.profile_enable(False)
        return ms, best, launches, prof

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    tensor_kernels = ("tc_gemm", "tc_kstar", "sweep_finish", "gram", "chol_step")
    fp64_kernels = ("sweep_gemm", "sweep_kstar", "sweep_reduce", "sweep_grad_gemm", "sweep_finish", "gram", "chol_step")
    head_kernels = fp64_kernels if args.mode == "fp64" else tensor_kernels
    ms, best, launches, prof = run_arm(args.mode, M, head_kernels)
    clocks = sampler.stop() if rank == 0 else None

    fp64 = None
    if not args.no_fp64 and args.mode != "fp64":
        sampler64 = ClockSampler(local)
        if rank == 0:
            sampler64.start()
        ms64, best64, launches64, prof64 = run_arm("fp64", M64, fp64_kernels)
        fp64 = {"ms": ms64, "launches": launches64, "prof": prof64, "clocks": sampler64.stop() if rank == 0 else None}
    ctx.set_sweep_mode(MODES[args.mode])

    # ---- end to end through the host-buffer C-ABI call: pinned candidates in, values + gradients out
    hq = [torch.rand((Me, DIM), dtype=torch.float64, generator=torch.Generator().manual_seed(77 + rank + 10 * b)).pin_memory()
          for b in range(2)]
    hval = torch.empty(Me, dtype=torch.float64).pin_memory()
    hgrad = torch.empty((Me, DIM), dtype=torch.float64).pin_memory()
    lib, h = ctx.lib, ctx.h
    dpt = C.POINTER(C.c_double)

    def e2e_step(i, m=Me):
        fit()
        st = lib.slsgp_acq_batch(h, ACQ_EI, 1.0, C.cast(hq[i % 2].data_ptr(), dpt), m, C.cast(hval.data_ptr(), dpt),
                                 C.cast(hgrad.data_ptr(), dpt))
        assert st == 0, lib.slsgp_last_error(h)
        return float(hval[:m].max())  # the host reads the step's result

    ms_e2e, _ = timed(e2e_step, args.steps, max(args.warmup, 3))
    e2e64 = None
    if fp64 is not None:
        Me64 = min(Me, M64)
        ctx.set_sweep_mode(pkg.SWEEP_FP64)
        ms_e2e64, _ = timed(lambda i: e2e_step(i, Me64), args.steps, max(args.warmup, 3))
        e2e64 = {"ms": ms_e2e64, "m": Me64}
        ctx.set_sweep_mode(MODES[args.mode])

    # ---- end to end through the C++ host layer with PAGEABLE memory: acquisition_func::CalcAcquisitionValues on a
    # GaussianProcessRegressor (Eigen matrices in / out, the call a C++ user of the drop-in makes); rank 0 only
    pageable = None
    if rank == 0 and not args.no_pageable:
        try:
            host = pkg.hostlib.Host()
            hreg = host.gpr_create(KERNEL_SE, X, y, theta, NOISE)
            reg = host.gpr_regressor(hreg)
            Mp = 1 << 20
            pageable = {}
            for name, mode in (("fp64", pkg.SWEEP_FP64), ("tensor", pkg.SWEEP_TENSOR)):
                host.regressor_set_sweep_mode(reg, mode)
                reps = 2 if name == "fp64" else max(3, min(args.steps, 6))
                sec = host.time_acq_values(reg, DIM, Mp, ACQ_EI, 1.0, reps)
                pageable[name] = {"value": Mp / sec, "unit": UNIT, "candidates_per_call": Mp, "calls": reps,
                                  "h2d_bytes_per_call": Mp * DIM * 8, "d2h_bytes_per_call": Mp * (DIM + 1) * 8,
                                  "api": "acquisition_func::CalcAcquisitionValues(GaussianProcessRegressor, Eigen::MatrixXd, .., &gradients) of libsls_b200_host.so, "
                                         "timed inside the C++ layer: pageable Eigen storage in and out, result allocations included"}
            host.regressor_set_sweep_mode(reg, pkg.SWEEP_FP64)
            host.gpr_destroy(hreg)
        except Exception as e:  # the host layer is optional for the metric; say why it is missing
            pageable = {"unavailable": repr(e)}

    if rank == 0:
        peaks, peak_src = measured_peaks()
        # FP64 tensor-pipe peak: cuBLAS DGEMM measured now on this GPU (no FP64 figure in MEASURED_PEAKS.json)
        A = torch.rand((8192, 8192), dtype=torch.float64, device="cuda")
        torch.matmul(A, A)
        best_ms = 1e9
        for _ in range(4):
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            torch.matmul(A, A)
            t1.record()
            torch.cuda.synchronize()
            best_ms = min(best_ms, t0.elapsed_time(t1))
        dgemm_tflops = 2.0 * 8192 ** 3 / (best_ms * 1e-3) / 1e12
        del A

        def gemm_roofline(prof_d, key, m_rank, steps, passes, peak, peak_source, kname, note):
            g_ms, g_n = prof_d[key]
            if not g_n:
                return None
            per_launch = m_rank * steps / g_n
            achieved = FLOP_PER_CAND * per_launch / (g_ms / g_n * 1e-3) / 1e12
            total = sum(v[0] for k, v in prof_d.items())
            return {"bound": "tensor", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                    "executed_tflops": passes * achieved, "executed_frac": passes * achieved / peak, "peak_source": peak_source, "note": note,
                    "launches_timed": g_n, "avg_launch_ms": g_ms / g_n, "candidates_per_launch": per_launch,
                    "algorithmic_flop_per_candidate": FLOP_PER_CAND, "share_of_step_kernel_time": g_ms / total if total else None,
                    "kernel_ms": {k: v[0] for k, v in prof_d.items()}, "kernel_launches": {k: v[1] for k, v in prof_d.items()}}

        shard = 148 * 128 * 2
        if args.mode == "fp64":
            roof = gemm_roofline(prof, "sweep_gemm", M, args.steps, 1, dgemm_tflops, "cuBLAS DGEMM 8192^3 measured in this run",
                                 "gemm64_dmma_kernel<NN> (beta = K^-1 k*, mma.sync.m8n8k4.f64)", "FP64 tensor pipe (DMMA)")
            traffic, traffic_note = None, "no capture"
        else:
            kname = (f"tc_sweep_gemm_kernel<{ctx_xp(DIM)}, 2, false> (tcgen05.mma cta_group::2 kind::f16, 256x256x16, {PASSES[args.mode]} split-fp16 "
                     "pass(es) per pipeline stage, fused epilogue)")
            roof = gemm_roofline(prof, "tc_gemm", M, args.steps, PASSES[args.mode], peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops")),
                                 f"bf16 sustained ({peak_src})", kname,
                                 f"achieved counts ALGORITHMIC flops (8.62 MFLOP per candidate); the tensor pipe executes {PASSES[args.mode]}x the 2N^2 term")
            traffic, traffic_note = traffic_from_capture("tc_sweep_gemm_kernel", {"n_obs": N_OBS, "dim": DIM, "passes": PASSES[args.mode],
                                                                               "shard_candidates": shard, "split": int(os.environ.get("SLSGP_TC_SPLIT", "2"))})
        if roof is not None:
            roof["traffic"], roof["traffic_note"] = traffic, traffic_note
        # Gram (HBM) and Cholesky (FP64 pipe + launch chain): kernel time from the per-kernel CUDA events of the timed steps
        gram_ms, gram_n = prof["gram"]
        chol_ms, chol_n = prof["chol_step"]
        gram_bytes = 8.0 * (N_OBS * DIM + N_OBS * N_OBS)
        rooflines = {}
        if gram_n:
            ach = gram_bytes / (gram_ms / gram_n * 1e-3) / 1e9
            rooflines["gram"] = {"bound": "hbm", "kernel": "gram_sym_kernel<0> (lower 64 x 64 tiles, mirrored through shared memory)", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                 "frac": ach / peaks["hbm_gbs"], "algorithmic_bytes_per_launch": gram_bytes, "avg_launch_ms": gram_ms / gram_n,
                                 "launches_timed": gram_n, "peak_source": peak_src,
                                 "traffic": traffic_from_capture("gram_tile_kernel", {"n_obs": N_OBS, "dim": DIM})[0]}
        if chol_n:
            nb64 = N_OBS // 64
            # the whole factorisation as timed alone (aux.factor_ms): the per-launch events of the profiled steps sit between
            # programmatic dependent launches and serialise them (measured inside the steps: chol_ms / chol_n * nb64)
            per_factor_ms = aux["factor_ms"]
            ach = N_OBS ** 3 / 3.0 / (per_factor_ms * 1e-3) / 1e12
            rooflines["cholesky"] = {"bound": "tensor", "kernel": f"chol_step_kernel<pair> x {nb64} dependent launches (FP64 DMMA trailing update, two pivots per barrier in the diagonal tile)", "achieved": ach,
                                     "peak": dgemm_tflops, "unit": "TFLOP/s", "frac": ach / dgemm_tflops, "algorithmic_flop": N_OBS ** 3 / 3.0,
                                     "ms_per_factorisation": per_factor_ms, "launches_timed": chol_n,
                                     "cusolver_potrf_ms": aux.get("cusolver_potrf_ms"),
                                     "note": "latency chain of N / 64 dependent steps, not a throughput kernel at this size; cusolver_potrf_ms = torch.linalg.cholesky on the same matrix in this run",
                                     "peak_source": "cuBLAS DGEMM 8192^3 measured in this run", "traffic": None}
        line = {
            "metric": METRIC, "value": Mtot * args.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64" if args.mode == "fp64" else "f16x2-split/f32-acc",
            "data": "synthetic",
            "config": dict(config_block(args.candidates, world, args.scaling), candidates_per_gpu_per_step=M, sweep_mode=args.mode,
                           shard_candidates=shard, step="fit (gram, cholesky, inverse, alpha) + sweep + arg-max" + (" + all_gather(value,index)" if world > 1 else ""),
                           collective="all_gather(value,index) per step" if world > 1 else "none"),
            "gram_chol_ms": aux["gram_chol_ms"], "aux": aux, "modes": mode_table,
            "e2e": {"value": Metot * args.steps / (ms_e2e * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": Me * DIM * 8, "d2h_bytes_per_step": Me * (DIM + 1) * 8,
                    "candidates_per_gpu_per_step": Me, "api": "slsgp_acq_batch (host buffers, pinned)"},
            "e2e_pageable": pageable,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roof, "rooflines": rooflines,
            "fp64_peak_tflops_measured": dgemm_tflops,
            "best": {"value": best[0], "index": best[1]},
        }
        if fp64 is not None:
            line["value_fp64"] = M64tot * args.steps / (fp64["ms"] * 1e-3)
            line["fp64"] = {"value": line["value_fp64"], "unit": UNIT, "dtype": "f64", "ms_per_step": fp64["ms"] / args.steps,
                            "candidates_per_step": M64tot, "gpu_launches": int(fp64["launches"]), "clocks": fp64["clocks"]}
            line["roofline_fp64"] = gemm_roofline(fp64["prof"], "sweep_gemm", M64, args.steps, 1, dgemm_tflops, "cuBLAS DGEMM 8192^3 measured in this run",
                                                  "gemm64_dmma_kernel<NN> (beta = K^-1 k*, mma.sync.m8n8k4.f64, 64x64 CTA tiles)", "FP64 tensor pipe (DMMA)")
            if e2e64:
                line["e2e_fp64"] = {"value": e2e64["m"] * world * args.steps / (e2e64["ms"] * 1e-3), "unit": UNIT,
                                    "h2d_bytes_per_step": e2e64["m"] * DIM * 8, "d2h_bytes_per_step": e2e64["m"] * (DIM + 1) * 8,
                                    "candidates_per_gpu_per_step": e2e64["m"], "api": "slsgp_acq_batch (host buffers, pinned), SLSGP_SWEEP_FP64"}
        if world == 1 and not args.no_configs:
            line["configs"] = extra_configs(pkg, ctx, torch)
            # the Gram kernel where it IS bound by HBM: at N = 2048 the 33.5 MB matrix never leaves the L2
            for row in line["configs"].get("grid_d16", []):
                if row["n_obs"] == 8192 and row.get("gram_ms"):
                    gb = 8.0 * (8192 * DIM + 8192.0 * 8192)
                    ach = gb / (row["gram_ms"] * 1e-3) / 1e9
                    line["rooflines"]["gram_n8192"] = {"bound": "hbm", "kernel": "gram_sym_kernel<0>", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                                       "frac": ach / peaks["hbm_gbs"], "algorithmic_bytes_per_launch": gb, "avg_launch_ms": row["gram_ms"],
                                                       "peak_source": peak_src, "traffic": None,
                                                       "note": "N = 8192, D = 16 (537 MB written); phase time of the last of three builds"}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_block()
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def cusolver_potrf_ms(torch, K_host, reps=7):
    """cuSOLVER potrf through torch.linalg.cholesky on the same Gram matrix, CUDA events, median: the library point the
    factorisation is read against (not on the product path; torch allocates the output inside the timed call)."""
    K = torch.from_numpy(np.ascontiguousarray(K_host)).cuda()
    for _ in range(3):
        torch.linalg.cholesky(K)
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        torch.linalg.cholesky(K)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    del K
    return float(np.median(ts))


def extra_configs(pkg, ctx, torch):
    """BASELINE.json configs 1, 2, 3, 5 and an N-grid on ONE GPU, so that they appear in the driver-run line (rank 0, N = 1 only).
    Config 4 is the headline. Every entry says what was timed."""
    out = {}
    KERNEL_MATERN = 1

    def device_ms(fn, reps):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps * 1e3

    # ---- config 2: GP regressor, N = 512, D = 8: build K + Cholesky (+ inverse, alpha), then 2^20 posterior queries
    N2, D2, M2 = 512, 8, 1 << 20
    X2, th2 = synth.make_X(N2, D2, "uniform", seed=1), synth.make_theta(D2, "default")
    y2 = synth.make_y(X2, seed=3)
    ctx.set_data(X2)

    def fit2():
        ctx.invalidate()
        ctx.gram(KERNEL_SE, th2, NOISE, want=False)
        ctx.factor()
        ctx.inverse(want=False)
        ctx.solve_alpha(y2)

    fit_ms = device_ms(fit2, 10)
    phases = {k: ctx.phase_ms(k) for k in ("gram", "factor", "inverse", "alpha")}
    q2 = torch.rand((M2, D2), dtype=torch.float64, device="cuda")
    mu2, sg2 = torch.empty(M2, dtype=torch.float64, device="cuda"), torch.empty(M2, dtype=torch.float64, device="cuda")
    c2 = {"n_obs": N2, "dim": D2, "queries": M2, "fit_ms_host_clock": fit_ms, "phase_ms": phases,
          "what": "slsgp_gram + factor + inverse + solve_alpha, then mu and sigma for 2^20 device-resident query points"}
    for name, mode in (("fp64", pkg.SWEEP_FP64), ("tensor", pkg.SWEEP_TENSOR)):
        ctx.set_sweep_mode(mode)
        ms = device_ms(lambda: (ctx.acq_batch_device(ACQ_EI, 1.0, q2.data_ptr(), M2, d_mu=mu2.data_ptr(), d_sigma=sg2.data_ptr()), ctx.synchronize()), 3)
        c2[f"posterior_queries_per_s_{name}"] = M2 / (ms * 1e-3)
    out["config2"] = c2
    del q2, mu2, sg2

    # ---- config 3: PreferenceRegressor MAP objective + gradient, N = 2048, D = 16, 683 triplets, 2066 variables
    X3, _, _ = fixed_model()
    off3, idx3 = synth.make_tuples(X3)
    ctx.set_sweep_mode(pkg.SWEEP_FP64)
    ctx.set_data(X3)
    ctx.set_preferences(off3, idx3)
    rng = np.random.default_rng(21)
    x3 = np.concatenate([0.05 * rng.standard_normal(N_OBS), [0.5, 0.005], np.full(DIM, 0.5)])
    t = []
    for it in range(53):
        xp = x3 * (1.0 + 0.01 * rng.standard_normal(len(x3)))
        t0 = time.perf_counter()
        ctx.map_objective_pref(KERNEL_SE, xp, True, 0.5, 0.5, 0.005, 0.25, 0.01)
        if it >= 3:
            t.append((time.perf_counter() - t0) * 1e3)
    out["config3"] = {"n_obs": N_OBS, "dim": DIM, "tuples": int(len(off3) - 1), "variables": int(len(x3)), "evaluations": len(t),
                      "ms_per_objective_and_gradient_median": float(np.median(t)), "ms_50_evaluations": float(np.sum(t)),
                      "what": "slsgp_map_objective_pref with hyper-parameters (Gram, Cholesky, inverse, alpha, BTL, all D + 2 + N gradient entries), "
                              "host wall clock per call incl. the copies of x and the gradient"}

    # ---- the same objective at the sizes an optimiser run actually sees (N = 3 ... a few hundred): latency per evaluation
    lat = []
    for Nl, Dl in ((30, 6), (64, 6), (80, 6), (128, 16), (200, 6), (400, 16), (400, 64)):
        Xl = synth.make_X(Nl, Dl, "uniform", seed=1)
        offl, idxl = synth.make_tuples(Xl)
        ctx.set_data(Xl)
        ctx.set_preferences(offl, idxl)
        xl = np.concatenate([0.05 * rng.standard_normal(Nl), [0.5, 0.005], np.full(Dl, 0.5)])
        for it in range(5):
            ctx.map_objective_pref(KERNEL_SE, xl * (1.0 + 1e-3 * it), True, 0.5, 0.5, 0.005, 0.25, 0.01)
        t0 = time.perf_counter()
        for it in range(40):
            ctx.map_objective_pref(KERNEL_SE, xl * (1.0 + 1e-4 * it), True, 0.5, 0.5, 0.005, 0.25, 0.01)
        lat.append({"n_obs": Nl, "dim": Dl, "us_per_objective_and_gradient": (time.perf_counter() - t0) / 40 * 1e6})
    out["map_objective_latency"] = {"rows": lat, "what": "slsgp_map_objective_pref with hyper-parameters, host wall clock per call (mean of 40): N <= 64 is the "
                                    "single-launch small-model kernel, above it the general path with one stream synchronisation per evaluation"}

    # ---- configs 1 and 5: the optimiser loops through the C++ host layer (default search driver), simulated user of the nd demo
    try:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import loop_support as LS
        b200 = LS.LoopLib("b200")
        demo_hyper = (0.5, 0.5, 0.001, 0.1, 0.01)
        driver = {0: "native", 1: "hybrid", 2: "reference"}[pkg.hostlib.get_search_driver()]
        LS.run_sls_loop(b200, 6, 3, 1, hyper=demo_hyper)  # warm-up
        log1 = LS.run_sls_loop(b200, 6, 15, 1, hyper=demo_hyper)
        c1 = {"dim": 6, "iterations": 15, "search_driver": driver, "total_s": sum(r["ms"] for r in log1) / 1e3,
              "ms_per_iteration": [round(r["ms"], 2) for r in log1], "final_objective": log1[-1]["objective"],
              "what": "SequentialLineSearchOptimizer::SubmitFeedbackData x 15 (Matern-5/2, MAP hyper-parameters, slider enlargement), host wall clock"}
        if LS.ref_loop_available():
            ref = LS.LoopLib("ref")
            logr = LS.run_sls_loop(ref, 6, 15, 1, hyper=demo_hyper)
            c1["reference_cpu"] = {"total_s": sum(r["ms"] for r in logr) / 1e3, "ms_per_iteration": [round(r["ms"], 2) for r in logr],
                                   "final_objective": logr[-1]["objective"], "cores": 1,
                                   "what": "the unmodified reference loop (oracle/_ref/libsls_ref_loop.so: reference sources + NLopt 2.10 + eigen-lite), same seed and user"}
        out["config1"] = c1
        log5 = LS.run_sls_loop(b200, 64, 200, 1, kt=LS.SE, hyper=demo_hyper)
        ms5 = [r["ms"] for r in log5]
        out["config5"] = {"dim": 64, "iterations": 200, "search_driver": driver, "n_gpus": 1, "total_s": sum(ms5) / 1e3,
                          "ms_per_iteration_median": float(np.median(ms5)), "ms_per_iteration_p90": float(np.percentile(ms5, 90)),
                          "ms_last_iteration": ms5[-1], "n_points_final": log5[-1]["n_points"], "final_objective": log5[-1]["objective"],
                          "what": "SequentialLineSearchOptimizer loop, D = 64, SE kernel, EI, MAP hyper-parameters, 200 x SubmitFeedbackData on one GPU "
                                  "(8-GPU candidate shard: tools/multi_gpu_library.py, profiles/)"}
    except Exception as e:
        out["loops"] = {"unavailable": repr(e)}

    # ---- N grid at D = 16: Gram + Cholesky and the tensor sweep
    grid = []
    for Ng in (256, 1024, 4096, 8192):
        Xg, thg = synth.make_X(Ng, DIM, "uniform", seed=1), synth.make_theta(DIM, "default")
        ctx.set_data(Xg)
        yg = synth.make_y(Xg, seed=3)
        for _ in range(3):
            ctx.invalidate()
            ctx.gram(KERNEL_SE, thg, NOISE, want=False)
            ctx.factor()
            ctx.inverse(want=False)
            ctx.solve_alpha(yg)
        row = {"n_obs": Ng, "dim": DIM, "gram_ms": ctx.phase_ms("gram"), "cholesky_ms": ctx.phase_ms("factor"), "inverse_ms": ctx.phase_ms("inverse")}
        row["cusolver_potrf_ms"] = cusolver_potrf_ms(torch, ctx.gram(KERNEL_SE, thg, NOISE, want=True), reps=5)
        Mg = 1 << 18
        qg = torch.rand((Mg, DIM), dtype=torch.float64, device="cuda")
        vg, gg = torch.empty(Mg, dtype=torch.float64, device="cuda"), torch.empty((Mg, DIM), dtype=torch.float64, device="cuda")
        ctx.set_sweep_mode(pkg.SWEEP_TENSOR)
        ms = device_ms(lambda: (ctx.acq_batch_device(ACQ_EI, 1.0, qg.data_ptr(), Mg, d_val=vg.data_ptr(), d_grad=gg.data_ptr()), ctx.synchronize()), 2)
        row["ei_evals_per_s_tensor"] = Mg / (ms * 1e-3)
        ctx.set_sweep_mode(pkg.SWEEP_FP64)
        grid.append(row)
        del qg, vg, gg
    out["grid_d16"] = grid
    return out


def ctx_xp(D):
    for xp in (8, 12, 20, 36, 68):
        if D + 1 <= xp:
            return xp
    return 0


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        main_reference(a)
    else:
        main_graft(a)
