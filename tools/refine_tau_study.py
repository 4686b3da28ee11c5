"""Second tier of the tensor sweep: what the threshold tau costs and buys (GPU box). For dense low-dimensional data most candidates
have a small posterior variance; with tau = 0.1 nearly all of them are re-evaluated in IEEE double and the sweep runs at the FP64
rate. Prints, per (N, D, tau): the share of candidates below tau * a, the sweep rate, and the largest error of every output of the
tensor path (3 passes) against the FP64 sweep, relative to the largest reference entry. usage: python tools/refine_tau_study.py"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tools")]
import synth  # noqa: E402

pkg = importlib.import_module("sequential-line-search_b200")
ctx = pkg.Context(0)


def rel(a, b):
    return float(np.max(np.abs(a - b))) / max(float(np.max(np.abs(b))), 1e-300)


for (N, D, kind, kt) in ((2048, 6, "uniform", 0), (2048, 8, "uniform", 0), (2048, 8, "uniform", 1), (512, 8, "uniform", 0), (2048, 16, "sls", 0), (700, 5, "sls", 1), (300, 4, "uniform", 0)):
    X, theta = synth.make_X(N, D, kind), synth.make_theta(D, "perturbed" if kind == "sls" else "default")
    ctx.fit(X, kt, theta, 0.005, synth.make_y(X))
    M = 37888 * 2
    Q = synth.f64(np.concatenate([synth.make_queries(M - 64, D), X[:, :64] + 1e-3], axis=1))
    ctx.set_sweep_mode(pkg.SWEEP_FP64)
    mu0, s0, dmu0, ds0 = ctx.posterior_batch(Q)
    v0, g0 = ctx.acq_batch(0, 1.0, Q)
    u0, gu0 = ctx.acq_batch(1, 2.0, Q)
    ctx.set_sweep_mode(pkg.SWEEP_TENSOR)
    dq = torch.from_numpy(np.ascontiguousarray(Q.T)).cuda()
    dv, dg = torch.empty(M, dtype=torch.float64, device="cuda"), torch.empty((M, D), dtype=torch.float64, device="cuda")
    print(f"N = {N} D = {D} {kind} kernel {kt}: sigma^2 / a quantiles (1 %, 10 %, 50 %) = {np.quantile(s0 ** 2 / theta[0], [0.01, 0.1, 0.5])}")
    for tau in (0.1, 0.05, 0.02, 0.01, 0.003, 0.001, 0.0):
        ctx.set_refine_threshold(tau)
        mu, s, dmu, ds = ctx.posterior_batch(Q)
        v, g = ctx.acq_batch(0, 1.0, Q)
        u, gu = ctx.acq_batch(1, 2.0, Q)
        ctx.acq_batch_device(0, 1.0, dq.data_ptr(), M, d_val=dv.data_ptr(), d_grad=dg.data_ptr())
        ctx.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            ctx.acq_batch_device(0, 1.0, dq.data_ptr(), M, d_val=dv.data_ptr(), d_grad=dg.data_ptr())
        ctx.synchronize()
        e1.record()
        torch.cuda.synchronize()
        rate = 3 * M / (e0.elapsed_time(e1) * 1e-3)
        share = float(np.mean(s0 ** 2 < tau * theta[0]))
        errs = {"mu": rel(mu, mu0), "sigma": rel(s, s0), "dmu": rel(dmu, dmu0), "dsigma": rel(ds, ds0), "EI": rel(v, v0), "gEI": rel(g, g0), "UCB": rel(u, u0), "gUCB": rel(gu, gu0)}
        print(f"  tau {tau:5.3f}: {100 * share:5.1f} % re-evaluated, {rate:.3e} evals/s, worst error {max(errs.values()):.1e} (" + " ".join(f"{k} {e:.0e}" for k, e in errs.items()) + ")")
    ctx.set_refine_threshold(0.1)
ctx.close()
