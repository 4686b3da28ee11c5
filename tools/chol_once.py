"""One Gram + Cholesky at N (argv[1]) with the current environment switches; for ncu launch lists (not part of the product path)."""
import importlib
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
import synth  # noqa: E402

pkg = importlib.import_module("sequential-line-search_b200")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ctx = pkg.Context(0)
ctx.set_data(synth.make_X(n, 16, "uniform"))
for _ in range(reps):
    ctx.invalidate()
    ctx.gram(0, synth.make_theta(16, "default"), 0.005, want=False)
    ctx.factor()
print("factor ms", ctx.lib.slsgp_last_phase_ms(ctx.h, b"factor"))
ctx.close()
