"""Ad-hoc phase timings on the GPU box (not the bench): python tools/quick_timing.py [N] [D] [M]"""
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tools")]
import synth as S  # noqa: E402  (input generators only; no checker code)

pkg = importlib.import_module("sequential-line-search_b200")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
D = int(sys.argv[2]) if len(sys.argv) > 2 else 16
M = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 16
ctx = pkg.Context(0)
X, theta = S.make_X(N, D, "uniform"), S.make_theta(D, "default")
y = S.make_y(X)
ctx.set_data(X)
for it in range(3):
    ctx.invalidate()
    ctx.gram(0, theta, 0.005, want=False)
    ctx.factor()
    ctx.inverse(want=False)
    ctx.solve_alpha(y)
    print(f"iter {it}: gram {ctx.phase_ms('gram'):.3f} ms  factor {ctx.phase_ms('factor'):.3f} ms  "
          f"inverse {ctx.phase_ms('inverse'):.3f} ms  alpha {ctx.phase_ms('alpha'):.3f} ms", flush=True)
for it in range(3):
    t0 = time.time()
    x, v, idx, _ = ctx.acq_argmax(0, 1.0, 1, 0, M)
    dt = time.time() - t0
    ms = ctx.phase_ms("sweep")
    print(f"sweep M={M}: {ms:.2f} ms device ({M / ms * 1e3:.3e} evals/s value-only), wall {dt * 1e3:.1f} ms", flush=True)
Q = S.make_queries(M, D)
for it in range(2):
    t0 = time.time()
    val, grad = ctx.acq_batch(0, 1.0, Q)
    dt = time.time() - t0
    ms = ctx.phase_ms("sweep")
    print(f"acq_batch (host buffers, value+grad) M={M}: {ms:.2f} ms device ({M / ms * 1e3:.3e} evals/s), wall {dt * 1e3:.1f} ms", flush=True)
if N >= 512:
    offsets, idx = S.make_tuples(X)
    ctx.set_preferences(offsets, idx)
    x0 = np.concatenate([0.05 * np.random.default_rng(0).standard_normal(N), [0.5, 0.005], np.full(D, 0.5)])
    for it in range(3):
        t0 = time.time()
        f, g = ctx.map_objective_pref(0, x0, True, 0.5, 0.5, 0.005, 0.25, 0.01)
        print(f"map objective+grad: {ctx.phase_ms('map'):.2f} ms device, wall {(time.time() - t0) * 1e3:.1f} ms  f={f:.6f}", flush=True)
print("launches", ctx.launch_count())
