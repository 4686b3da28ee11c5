"""A few evaluations of slsgp_map_objective_pref (with hyper-parameters) and of the whitened objective at N = 2048, D = 16 and at
a small N: the launch sequence ncu captures for the K5 / K6 kernels. usage: python tools/map_objective_once.py [N] [D]"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tools")]
import synth  # noqa: E402

pkg = importlib.import_module("sequential-line-search_b200")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
D = int(sys.argv[2]) if len(sys.argv) > 2 else 16
ctx = pkg.Context(0)
X = synth.make_X(N, D, "uniform")
off, idx = synth.make_tuples(X)
ctx.set_data(X)
ctx.set_preferences(off, idx)
rng = np.random.default_rng(0)
x = np.concatenate([0.05 * rng.standard_normal(N), [0.5, 0.005], np.full(D, 0.5)])
for it in range(4):
    f, g = ctx.map_objective_pref(0, x * (1 + 0.01 * it), True, 0.5, 0.5, 0.005, 0.25, 0.01)
print("objective", f, "ms", ctx.phase_ms("map"))
ctx.gram(0, synth.make_theta(D), 0.005, want=False)
ctx.factor()
for it in range(4):
    f, gz, y = ctx.map_objective_pref_whitened(0.01 * rng.standard_normal(N), 0.01)
print("whitened objective", f, "ms", ctx.phase_ms("map"))
ctx.close()
