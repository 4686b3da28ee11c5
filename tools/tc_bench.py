"""Per-kernel timings of the tensor-core sweep on the GPU box (not the bench): python tools/tc_bench.py [N] [D] [shards]
Prints, per sweep mode, the average CUDA-event time of tc_kstar / tc_gemm / sweep_finish per shard and the rates."""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tools")]
import synth  # noqa: E402

pkg = importlib.import_module("sequential-line-search_b200")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
D = int(sys.argv[2]) if len(sys.argv) > 2 else 16
shards = int(sys.argv[3]) if len(sys.argv) > 3 else 8
cap = int(os.environ.get("SLSGP_TC_SHARD", 148 * 128 * 2))
M = cap * shards
ctx = pkg.Context(0)
stream = torch.cuda.Stream()
ctx.set_stream(stream.cuda_stream)
X, theta = synth.make_X(N, D, "uniform"), synth.make_theta(D, "default")
ctx.fit(X, 0, theta, 0.005, synth.make_y(X))
Xq = torch.rand((M, D), dtype=torch.float64, device="cuda")
val = torch.empty(M, dtype=torch.float64, device="cuda")
grad = torch.empty((M, D), dtype=torch.float64, device="cuda")
F = 2.0 * N * N
for name, mode, passes in (("tensor", pkg.SWEEP_TENSOR, 3), ("tensor_x2", pkg.SWEEP_TENSOR_X2, 2), ("tensor_x1", pkg.SWEEP_TENSOR_X1, 1)):
    ctx.set_sweep_mode(mode)
    ctx.acq_batch_device(0, 1.0, Xq.data_ptr(), M, d_val=val.data_ptr(), d_grad=grad.data_ptr())
    torch.cuda.synchronize()
    ctx.profile_enable(True)
    for k in ("tc_gemm", "tc_kstar", "sweep_finish"):
        ctx.profile_read(k)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    reps = 3
    for r in range(reps):
        ctx.acq_batch_device(0, 1.0, Xq.data_ptr(), M, d_val=val.data_ptr(), d_grad=grad.data_ptr())
    e1.record(stream)
    torch.cuda.synchronize()
    tot = e0.elapsed_time(e1)
    prof = {k: ctx.profile_read(k) for k in ("tc_gemm", "tc_kstar", "sweep_finish")}
    ctx.profile_enable(False)
    g_ms, g_n = prof["tc_gemm"]
    per = g_ms / g_n
    print(f"{name:10s} pair={os.environ.get('SLSGP_TC_PAIR', 'default')} total {reps * M / tot * 1e3:.3e} evals/s | per shard of {cap}: "
          f"gemm {per:.4f} ms ({cap / per * 1e3:.3e} cand/s, executed {passes * F * cap / per * 1e-9:.0f} TFLOP/s) "
          f"kstar {prof['tc_kstar'][0] / prof['tc_kstar'][1]:.4f} ms  finish {prof['sweep_finish'][0] / prof['sweep_finish'][1]:.4f} ms", flush=True)
