"""SURVEY.md 8(f) rank 3, measured: what extending the factor by the appended columns would save per SubmitFeedbackData.
For N = 100 .. 600 (D = 16): (a) three slsgp_append_point calls (the three points an iteration adds), (b) the from-scratch model
build the regressor constructor does today (set_data + gram + factor + inverse + alpha), (c) a whole SubmitFeedbackData of the
SequentialLineSearchOptimizer at that N with FIXED hyper-parameters (the only mode in which K survives an iteration: with
use_map_hyperparams, the default, every MAP evaluation rebuilds K). usage: python tools/incremental_refit_study.py"""
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "tests")]
import synth  # noqa: E402
import loop_support as LS  # noqa: E402

pkg = importlib.import_module("sequential-line-search_b200")


def main():
    D = 16
    ctx = pkg.Context(0)
    rows = {}
    for N in (100, 200, 300, 400, 500, 600):
        X, theta = synth.make_X(N + 3, D, "sls"), synth.make_theta(D, "default")
        y = synth.make_y(X)
        app, fit = [], []
        for rep in range(5):
            ctx.fit(X[:, :N], 0, theta, 0.005, y[:N])
            t0 = time.perf_counter()
            for n in range(N, N + 3):
                ctx.append_point(X[:, n], y[n], want=False)
            app.append(time.perf_counter() - t0)
            t0 = time.perf_counter()
            ctx.fit(X, 0, theta, 0.005, y)
            fit.append(time.perf_counter() - t0)
        rows[N] = (1e3 * np.median(app), 1e3 * np.median(fit))
    ctx.close()
    # whole iterations with fixed hyper-parameters: one long loop, iteration time sampled where N crosses the sizes above
    b200 = LS.LoopLib("b200")
    log = LS.run_sls_loop(b200, D, 300, 1, kt=LS.SE, use_map=False, hyper=(0.5, 0.5, 0.001, 0.1, 0.01))
    # the wired path (sequential_line_search::SetIncrementalRefit): the same loop again with the previous regressor's model handed on
    pkg.hostlib.set_incremental_refit(True)
    log_inc = LS.run_sls_loop(b200, D, 300, 1, kt=LS.SE, use_map=False, hyper=(0.5, 0.5, 0.001, 0.1, 0.01))
    pkg.hostlib.set_incremental_refit(False)
    print(f"search driver: {pkg.hostlib.get_search_driver()} (0 native, 1 hybrid, 2 reference); D = {D}, SE kernel, fixed hyper-parameters")
    print(f"{'N':>5s} {'3 x append_point ms':>20s} {'model rebuild ms':>18s} {'saving ms':>10s} {'SubmitFeedbackData ms':>22s} {'saving / iteration':>19s}")
    for N, (a, f) in rows.items():
        near = [r["ms"] for r in log if abs(r["n_points"] - N) <= 6]
        it = float(np.median(near)) if near else float("nan")
        print(f"{N:5d} {a:20.3f} {f:18.3f} {f - a:10.3f} {it:22.2f} {100 * (f - a) / it if near else float('nan'):18.1f}%")
    print("\nwired: SequentialLineSearchOptimizer loop of 300 iterations, rebuild per iteration vs SetIncrementalRefit(true) (same seed and user)")
    print(f"{'N':>5s} {'rebuild ms / iteration':>24s} {'incremental ms / iteration':>28s}")
    for N in rows:
        a = [r["ms"] for r in log if abs(r["n_points"] - N) <= 10]
        b = [r["ms"] for r in log_inc if abs(r["n_points"] - N) <= 10]
        if a and b:
            print(f"{N:5d} {np.median(a):24.2f} {np.median(b):28.2f}")
    print(f"whole loop: {sum(r['ms'] for r in log) / 1e3:.2f} s rebuilding, {sum(r['ms'] for r in log_inc) / 1e3:.2f} s incremental; "
          f"final objective {log[-1]['objective']:.4f} / {log_inc[-1]['objective']:.4f}")


if __name__ == "__main__":
    main()
