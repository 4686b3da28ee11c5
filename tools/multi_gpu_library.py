"""Multi-GPU INSIDE the library (slsgp_ctx_create_multi), measured in one process: config 4 (N = 2048, D = 16, 2^24 candidates,
arg-max and arg-max + ascent) on 1 .. G devices, and config 5 (the full SequentialLineSearchOptimizer loop, D = 64, SE kernel, EI,
200 iterations) through the pySequentialLineSearch module on 1 and G devices. usage: python tools/multi_gpu_library.py [--iters 200]"""
import argparse
import importlib
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "tests")]
import synth  # noqa: E402

pkg = importlib.import_module("sequential-line-search_b200")


def config4(devices, count=1 << 24, reps=3):
    ctx = pkg.Context(devices if len(devices) > 1 else devices[0])
    X, theta = synth.make_X(2048, 16, "uniform"), synth.make_theta(16, "default")
    ctx.fit(X, 0, theta, 0.005, synth.make_y(X))
    ctx.set_sweep_mode(pkg.SWEEP_TENSOR)
    ctx.acq_argmax(0, 1.0, 7, 0, count // 8)  # warm-up: replicas, operands, workspaces
    t = []
    for r in range(reps):
        t0 = time.perf_counter()
        x, v, idx, _ = ctx.acq_argmax(0, 1.0, 7, 0, count)
        t.append(time.perf_counter() - t0)
    t0 = time.perf_counter()
    xm, vm, _, vs = ctx.acq_maximize(0, 1.0, 7, 0, count, n_starts=4096, n_iters=40)
    tm = time.perf_counter() - t0
    ctx.close()
    return min(t), idx, v, tm, vm


def config5(devices, iters, D=64):
    import loop_support as LS
    pkg.hostlib.set_devices(devices)
    pkg.hostlib.set_search_driver(pkg.hostlib.HYBRID if pkg.hostlib.nlopt_available() else pkg.hostlib.NATIVE)
    sys.path.insert(0, pkg.LIB_DIR)
    pkg.build_python_module()
    import pySequentialLineSearch as sls
    opt = sls.SequentialLineSearchOptimizer(num_dims=D, use_map_hyperparams=True, kernel_type=sls.KernelType.ArdSquaredExponentialKernel)
    opt.set_hyperparams(0.5, 0.5, 0.001, 0.1, 0.01)
    ms = []
    for it in range(iters):
        e0, e1 = opt.get_slider_ends()
        t = LS.best_slider_position(np.asarray(e0), np.asarray(e1))
        t0 = time.perf_counter()
        opt.submit_feedback_data(t)
        ms.append((time.perf_counter() - t0) * 1e3)
    return ms, LS.demo_objective(np.asarray(opt.get_maximizer())), opt.get_raw_data_points().shape[1]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=200)
    args = ap.parse_args()
    G = torch.cuda.device_count()
    print(f"{G} GPUs visible; one process, slsgp_ctx_create_multi")
    print("config 4: N=2048 D=16, 2^24 counter-based candidates, tensor sweep (3-pass) + arg-max; then arg-max + 4096-start ascent")
    base = None
    for g in [n for n in (1, 2, 4, 8) if n <= G]:
        t, idx, v, tm, vm = config4(list(range(g)))
        base = base or t
        print(f"  {g} GPU(s): arg-max {t * 1e3:8.1f} ms = {(1 << 24) / t:.3e} evals/s  (speed-up {base / t:4.2f}x, efficiency {base / t / g:4.2f})  winner {idx} EI {v:.6f}"
              f" | maximise {tm * 1e3:8.1f} ms, EI {vm:.6f}")
    print(f"config 5: SequentialLineSearchOptimizer loop through pySequentialLineSearch, D=64, SE kernel, EI, MAP hyper-parameters, {args.iters} iterations")
    for g in sorted({1, G}):
        ms, f, n = config5(list(range(g)), args.iters)
        q = np.percentile(ms, [50, 90])
        print(f"  {g} GPU(s): total {sum(ms) / 1e3:7.2f} s, per iteration median {q[0]:7.1f} ms, p90 {q[1]:7.1f} ms, last {ms[-1]:7.1f} ms (N = {n}), f(x+) = {f:.4f}")


if __name__ == "__main__":
    main()
