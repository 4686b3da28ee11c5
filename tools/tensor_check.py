"""Tensor-core sweep vs the FP64 sweep on the same model and candidates (GPU box only):
    python tools/tensor_check.py N D M [kind] [reps]
Prints max / rms errors of mu, sigma, EI and grad EI relative to the largest FP64 magnitude, and device timings."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tools")]
import synth  # noqa: E402

pkg = importlib.import_module("sequential-line-search_b200")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 96
D = int(sys.argv[2]) if len(sys.argv) > 2 else 6
M = int(sys.argv[3]) if len(sys.argv) > 3 else 200
kind = sys.argv[4] if len(sys.argv) > 4 else "uniform"
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3

ctx = pkg.Context(0)
X = synth.make_X(N, D, kind)
theta = synth.make_theta(D, "default")
y = synth.make_y(X)
ctx.fit(X, 0, theta, 0.005, y)
Q = synth.make_queries(M, D)


def run(mode):
    ctx.set_sweep_mode(mode)
    out = {}
    for r in range(reps):
        mu, sg, dmu, dsg = ctx.posterior_batch(Q)
        t_post = ctx.phase_ms("sweep")
        val, grad = ctx.acq_batch(0, 1.0, Q)
        t_acq = ctx.phase_ms("sweep")
    out.update(mu=mu, sigma=sg, dmu=dmu, dsigma=dsg, val=val, grad=grad, t_post=t_post, t_acq=t_acq)
    return out


a = run(pkg.SWEEP_FP64)
print(f"N={N} D={D} M={M} kind={kind}: fp64 acq sweep {a['t_acq']:.3f} ms ({M / a['t_acq'] * 1e3:.3e} evals/s)", flush=True)
for name, mode in (("tensor (3-pass)", pkg.SWEEP_TENSOR), ("tensor_x2", pkg.SWEEP_TENSOR_X2), ("tensor_x1", pkg.SWEEP_TENSOR_X1)):
    b = run(mode)
    print(f"{name}: acq sweep {b['t_acq']:.3f} ms ({M / b['t_acq'] * 1e3:.3e} evals/s)", flush=True)
    for k in ("mu", "sigma", "dmu", "dsigma", "val", "grad"):
        ref, got = a[k], b[k]
        scale = max(np.max(np.abs(ref)), 1e-300)
        err = np.abs(got - ref)
        print(f"  {k:7s} max|ref| {scale:.4e}  max err/scale {np.nanmax(err) / scale:.3e}  rms err/scale "
              f"{np.sqrt(np.nanmean(err ** 2)) / scale:.3e}  nan(got) {int(np.isnan(got).sum())}")
    ctx.profile_enable(True)
    ctx.acq_batch(0, 1.0, Q)
    for k in ("tc_kstar", "tc_gemm", "sweep_finish"):
        ms, n = ctx.profile_read(k)
        print(f"  {k}: {ms:.3f} ms over {n} launches")
    ctx.profile_enable(False)
print("launches", ctx.launch_count())
