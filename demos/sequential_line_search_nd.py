#!/usr/bin/env python
"""The reference's headless demo (demos/sequential_line_search_nd/main.cpp:17-118; README example: D = 6, 15
iterations) on the B200 build: a synthetic user who always picks the best point of the slider under
f(x) = exp(-|x - 0.4|^2). Prints one line per iteration and the found maximiser.

    python demos/sequential_line_search_nd.py [--dims 6] [--iters 15] [--kernel se|matern] [--seed 0]
"""
import argparse
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("sequential-line-search_b200")
pkg.build_python_module()
sys.path.insert(0, pkg.LIB_DIR)
import pySequentialLineSearch as sls  # noqa: E402


def objective(x):
    return float(np.exp(-np.sum((x - 0.4) ** 2)))


def run(dims=6, iters=15, kernel="matern", seed=0, verbose=True):
    rng = np.random.default_rng(seed)
    opt = sls.SequentialLineSearchOptimizer(
        num_dims=dims, use_slider_enlargement=True, use_map_hyperparams=True,
        kernel_type=sls.KernelType.ArdMatern52Kernel if kernel == "matern" else sls.KernelType.ArdSquaredExponentialKernel,
        initial_query_generator=lambda n: (rng.random(n), rng.random(n)))
    opt.set_hyperparams(kernel_signal_var=0.5, kernel_length_scale=0.5, noise_level=0.001, kernel_hyperparams_prior_var=0.1, btl_scale=0.01)
    history = []
    for it in range(iters):
        ts = np.linspace(0.0, 1.0, 1001)
        ys = [objective(opt.calc_point_from_slider_position(t)) for t in ts]
        t_best = float(ts[int(np.argmax(ys))])
        x = opt.calc_point_from_slider_position(t_best)
        t0 = time.time()
        opt.submit_feedback_data(t_best)
        dt = time.time() - t0
        history.append(dict(y=max(ys), residual=float(np.linalg.norm(x - 0.4)), seconds=dt, slider=opt.get_slider_ends()))
        if verbose:
            print(f"iter {it + 1:2d}: y = {max(ys):.4f}  |x - x*| = {history[-1]['residual']:.4f}  submit {dt * 1e3:.1f} ms", flush=True)
    x_star = opt.get_maximizer()
    if verbose:
        print("found maximiser:", np.round(x_star, 3), " f =", round(objective(x_star), 4))
    return history, x_star, opt


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--dims", type=int, default=6)
    ap.add_argument("--iters", type=int, default=15)
    ap.add_argument("--kernel", choices=["se", "matern"], default="matern")
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    run(a.dims, a.iters, a.kernel, a.seed)
