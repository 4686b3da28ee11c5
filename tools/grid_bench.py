"""Throughput grid of north_star ("synthetic N in {256..8192}, D in {6..64}"): for every (N, D) the model build phases
(Gram, Cholesky, inverse), the MAP objective + gradient, and the EI value + gradient sweep in the three arithmetic modes
that hold parity (FP64; tensor 3-pass for the SE and the Matern kernel), on one GPU. Not the bench: a table for profiles/.
    python tools/grid_bench.py [--quick]"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tools")]
import synth  # noqa: E402

pkg = importlib.import_module("sequential-line-search_b200")
quick = "--quick" in sys.argv
Ns = (256, 2048) if quick else (256, 512, 1024, 2048, 4096, 8192)
Ds = (6, 16) if quick else (6, 8, 16, 32, 64)
peak = 1382.8
try:
    import json
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"]
except Exception:
    pass

ctx = pkg.Context(0)
stream = torch.cuda.Stream()
ctx.set_stream(stream.cuda_stream)


def sweep_rate(mode, M, D, reps=2):
    ctx.set_sweep_mode(mode)
    Xq = torch.rand((M, D), dtype=torch.float64, device="cuda")
    val = torch.empty(M, dtype=torch.float64, device="cuda")
    grad = torch.empty((M, D), dtype=torch.float64, device="cuda")
    ctx.acq_batch_device(0, 1.0, Xq.data_ptr(), M, d_val=val.data_ptr(), d_grad=grad.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(reps):
        ctx.acq_batch_device(0, 1.0, Xq.data_ptr(), M, d_val=val.data_ptr(), d_grad=grad.data_ptr())
    e1.record(stream)
    torch.cuda.synchronize()
    return reps * M / (e0.elapsed_time(e1) * 1e-3)


print(f"# one B200; tensor peak used for the fraction: {peak:.0f} TFLOP/s (bf16 sustained, MEASURED_PEAKS.json)")
print("#    N   D | gram ms  chol ms  inv ms | MAP obj+grad ms | EI evals/s: FP64     SE tensor (executed TFLOP/s, frac)   Matern tensor")
for N in Ns:
    for D in Ds:
        X, theta = synth.make_X(N, D, "uniform"), synth.make_theta(D, "default")
        y = synth.make_y(X)
        ph = {}
        for kt in (0, 1):
            ctx.fit(X, kt, theta, 0.005, y)
            for it in range(3):
                ctx.invalidate()
                ctx.gram(kt, theta, 0.005, want=False)
                ctx.factor()
                ctx.inverse(want=False)
                ctx.solve_alpha(y)
            if kt == 0:
                ph = {k: ctx.phase_ms(k) for k in ("gram", "factor", "inverse")}
                offsets, idx = synth.make_tuples(X)
                ctx.set_preferences(offsets, idx)
                x0 = np.concatenate([0.05 * np.random.default_rng(0).standard_normal(N), [0.5, 0.005], np.full(D, 0.5)])
                for it in range(3):
                    ctx.map_objective_pref(0, x0, True, 0.5, 0.5, 0.005, 0.25, 0.01)
                map_ms = ctx.phase_ms("map")
                ctx.fit(X, kt, theta, 0.005, y)
                M64 = max(4096, min(65536, int(2e11 / (2.0 * N * N))))
                r64 = sweep_rate(pkg.SWEEP_FP64, M64, D)
                Mt = max(37888, min(37888 * 8, int(4e13 / (6.0 * N * N)) // 37888 * 37888))
                rse = sweep_rate(pkg.SWEEP_TENSOR, Mt, D)
            else:
                rma = sweep_rate(pkg.SWEEP_TENSOR, Mt, D) if D <= 66 else float("nan")
        ldt = (N + 255) // 256 * 256
        ex = 3 * 2.0 * ldt * ldt * rse * 1e-12  # executed on the tensor pipe (padded to the 256-column blocks)
        print(f"{N:6d} {D:3d} | {ph['gram']:7.3f} {ph['factor']:8.3f} {ph['inverse']:7.3f} | {map_ms:15.3f} | {r64:12.3e} "
              f"{rse:12.3e} ({ex:6.0f}, {ex / peak:4.2f}) {rma:12.3e}", flush=True)
