"""Host wall-clock latency of one MAP objective + gradient evaluation at small N (the regime of the optimisers' inner loops),
for the fused single-launch path (csrc/small.cuh) and, with SLSGP_SMALL_FUSED=0, the general multi-launch path."""
import importlib
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
import synth  # noqa: E402

pkg = importlib.import_module("sequential-line-search_b200")
ctx = pkg.Context(0)
D = 6
print("SLSGP_SMALL_FUSED =", os.environ.get("SLSGP_SMALL_FUSED", "(default: on)"))
for N in (4, 9, 16, 24, 32, 48, 64, 80):
    X = synth.make_X(N, D, "sls")
    y = synth.make_y(X)
    x = np.concatenate([[0.5, 0.005], np.full(D, 0.5)])
    offsets, idx = synth.make_tuples(X)
    ctx.set_data(X)
    ctx.set_preferences(offsets, idx)
    xp = np.concatenate([0.1 * y, x])
    res = []
    for fn in (lambda: ctx.map_objective_gpr(1, y, x), lambda: ctx.map_objective_pref(1, xp, True, 0.5, 0.5, 0.005, 0.25, 0.01)):
        for _ in range(20):
            fn()
        t0 = time.perf_counter()
        for _ in range(200):
            fn()
        res.append((time.perf_counter() - t0) / 200 * 1e6)
    print(f"N={N:3d} D={D}: GPR objective {res[0]:6.1f} us   preference objective (with hyper-parameters) {res[1]:6.1f} us")
ctx.close()
