"""Per-iteration wall time of SubmitFeedbackData: the real reference loop on the host CPU (oracle/_ref/libsls_ref_loop.so:
the unmodified sources + NLopt, one thread, as the reference runs) beside the B200 host layer under each search driver.
Shapes: config 1 (nd demo, D = 6, 15 iterations) and the config-5 shape (D = 64, SE kernel, EI; the reference side is cut
off after --ref-budget seconds because its cost per iteration grows like N^3 x evaluations). TEST TOOLING (loads oracle/).

usage: python tools/loop_timing.py [--d64-iters 200] [--ref-budget 120] [--out gpurun_out/loop_timing.txt]
"""
import argparse
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import loop_support as LS  # noqa: E402

DEMO_HYPER = (0.5, 0.5, 0.001, 0.1, 0.01)


def run(L, D, iters, seed, kt, budget_s=None, hyper=DEMO_HYPER):
    L.srand(seed)
    opt = L.sls(D, True, True, kt, LS.EI)
    opt.set_hyperparams(*hyper)
    rows, t_all = [], time.perf_counter()
    for it in range(iters):
        e0, e1 = opt.slider_ends()
        t = LS.best_slider_position(e0, e1)
        t0 = time.perf_counter()
        opt.submit(t)
        rows.append((it, opt.num_points(), (time.perf_counter() - t0) * 1e3, LS.demo_objective(opt.maximizer())))
        if budget_s is not None and time.perf_counter() - t_all > budget_s:
            break
    opt.close()
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--d64-iters", type=int, default=200)
    ap.add_argument("--ref-budget", type=float, default=120.0)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "loop_timing.txt"))
    args = ap.parse_args()
    pkg = importlib.import_module("sequential-line-search_b200")
    ref, b200 = LS.LoopLib("ref"), LS.LoopLib("b200")
    lines = [f"host cores: {os.cpu_count()} (the reference loop is single-threaded)"]
    for name, D, iters, kt in (("config 1: nd demo, D=6, 15 iterations, Matern-5/2, MAP hyper-parameters, EI", 6, 15, LS.MATERN),
                               (f"config-5 shape: D=64, {args.d64_iters} iterations, SE kernel, MAP hyper-parameters, EI", 64, args.d64_iters, LS.SE)):
        lines.append("")
        lines.append(name)
        cols = {"reference CPU": run(ref, D, iters, 1, kt, budget_s=args.ref_budget)}
        for mode, label in ((pkg.hostlib.REFERENCE, "B200 reference-driver"), (pkg.hostlib.HYBRID, "B200 hybrid"), (pkg.hostlib.NATIVE, "B200 native")):
            pkg.hostlib.set_search_driver(mode)
            run(b200, D, 2, 1, kt)  # warm-up: context creation, first-launch module load
            cols[label] = run(b200, D, iters, 1, kt)
        names = list(cols)
        lines.append("iter " + "".join(f"{n:>34s}" for n in names))
        lines.append("     " + "".join(f"{'N':>8s}{'ms':>12s}{'f(x+)':>14s}" for _ in names))
        for it in range(iters):
            if it >= 30 and it % 10 != 9:
                continue
            row = f"{it:4d} "
            for n in names:
                r = cols[n][it] if it < len(cols[n]) else None
                row += f"{r[1]:8d}{r[2]:12.2f}{r[3]:14.4f}" if r else f"{'-':>8s}{'-':>12s}{'-':>14s}"
            lines.append(row)
        for n in names:
            ms = [r[2] for r in cols[n]]
            lines.append(f"  {n}: {len(ms)} iterations, total {sum(ms) / 1e3:.2f} s, median {np.median(ms):.2f} ms, last {ms[-1]:.2f} ms (N = {cols[n][-1][1]}), "
                         f"final f(x+) = {cols[n][-1][3]:.4f}")
        n_common = min(len(cols["reference CPU"]), iters)
        for n in names[1:]:
            ratio = sum(r[2] for r in cols["reference CPU"][:n_common]) / sum(r[2] for r in cols[n][:n_common])
            lines.append(f"  reference CPU / {n} over the first {n_common} iterations: {ratio:.1f}x")
    text = "\n".join(lines)
    print(text)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        f.write(text + "\n")


if __name__ == "__main__":
    main()
