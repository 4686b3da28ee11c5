"""Latency of one MAP objective + gradient evaluation (slsgp_map_objective_pref with hyper-parameters) on the general path, by N
and D (GPU box), host wall clock per call. usage: python tools/map_latency_probe.py"""
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tools")]
import synth  # noqa: E402

pkg = importlib.import_module("sequential-line-search_b200")
ctx = pkg.Context(0)
rng = np.random.default_rng(0)
for D in (6, 64):
    for N in (40, 80, 128, 200, 400):
        X = synth.make_X(N, D, "uniform")
        off, idx = synth.make_tuples(X)
        ctx.set_data(X)
        ctx.set_preferences(off, idx)
        x = np.concatenate([0.05 * rng.standard_normal(N), [0.5, 0.005], np.full(D, 0.5)])
        for it in range(5):
            ctx.map_objective_pref(0, x * (1 + 0.001 * it), True, 0.5, 0.5, 0.005, 0.25, 0.01)
        t0 = time.perf_counter()
        reps = 50
        for it in range(reps):
            ctx.map_objective_pref(0, x * (1 + 0.0001 * it), True, 0.5, 0.5, 0.005, 0.25, 0.01)
        dt = (time.perf_counter() - t0) / reps
        print(f"D = {D:3d} N = {N:4d}: {dt * 1e6:7.1f} us per objective + gradient evaluation (device phase {ctx.phase_ms('map') * 1e3:6.1f} us)")
ctx.close()
