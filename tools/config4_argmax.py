"""BASELINE config 4 at full size on one GPU: 2^24 counter-based candidates at N = 2048, D = 16 through slsgp_acq_argmax
(global stage) and slsgp_acq_maximize (global stage + batched ascent), tensor sweep. Prints wall-clock and device times."""
import importlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tools")]
import synth  # noqa: E402

pkg = importlib.import_module("sequential-line-search_b200")
N, D, M = 2048, 16, 1 << 24
ctx = pkg.Context(0)
X, theta = synth.make_X(N, D, "uniform"), synth.make_theta(D, "default")
ctx.fit(X, 0, theta, 0.005, synth.make_y(X))
ctx.set_sweep_mode(pkg.SWEEP_TENSOR)
ctx.acq_argmax(0, 1.0, 1, 0, 1 << 18)  # warm-up (allocations, operand packing)
for rep in range(2):
    t0 = time.time()
    x, v, idx, _ = ctx.acq_argmax(0, 1.0, 7, 0, M)
    dt = time.time() - t0
    print(f"acq_argmax  M=2^24: wall {dt:.3f} s, device {ctx.phase_ms('sweep') * 1e-3:.3f} s -> {M / dt:.3e} evals/s (values only), best {v:.6f} at {idx}", flush=True)
t0 = time.time()
x2, v2, g2, vs = ctx.acq_maximize(0, 1.0, 7, 0, M, n_starts=4096, n_iters=40)
dt = time.time() - t0
print(f"acq_maximize M=2^24, 4096 starts x 40 iterations: wall {dt:.3f} s; sweep best {vs:.6f} -> refined {v2:.6f}, |grad| {abs(g2).max():.2e}")
