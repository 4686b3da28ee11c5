"""slsgp_acq_maximize at the sizes the optimisers reach (GPU box): time vs number of ascent iterations, and the convergence of the
best value. usage: python tools/maximize_probe.py [N] [D] [count]"""
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tools")]
import synth  # noqa: E402

pkg = importlib.import_module("sequential-line-search_b200")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 60
D = int(sys.argv[2]) if len(sys.argv) > 2 else 64
count = int(sys.argv[3]) if len(sys.argv) > 3 else 3200 * 128
ctx = pkg.Context(0)
X, theta = synth.make_X(N, D, "sls"), synth.make_theta(D, "default")
ctx.fit(X, 0, theta, 0.001, synth.make_y(X))
ctx.set_sweep_mode(pkg.SWEEP_TENSOR)
ctx.acq_maximize(0, 1.0, 3, 0, count, 1024, 5)
print(f"N = {N}, D = {D}, {count} candidates, 1024 starts")
for iters in (0, 10, 20, 40, 80, 120, 200):
    t0 = time.perf_counter()
    x, v, g, vs = ctx.acq_maximize(0, 1.0, 3, 0, count, 1024, iters)
    dt = (time.perf_counter() - t0) * 1e3
    gp = np.where(((x <= 0) & (g < 0)) | ((x >= 1) & (g > 0)), 0.0, g)
    print(f"  n_iters {iters:4d}: {dt:7.2f} ms   EI {v:.12g} (sweep alone {vs:.6g})   |projected grad| {np.linalg.norm(gp):.2e}")
ctx.close()
