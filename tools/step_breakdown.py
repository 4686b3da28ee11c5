"""Where SubmitFeedbackData spends its time, per iteration (GPU box): MAP fit / acquisition search / slider, and the number of MAP
objective evaluations. usage: python tools/step_breakdown.py [D] [iters] [driver: native|hybrid|reference] [kernel 0|1]"""
import ctypes as C
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "tests")]
import loop_support as LS  # noqa: E402

pkg = importlib.import_module("sequential-line-search_b200")
D = int(sys.argv[1]) if len(sys.argv) > 1 else 64
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 40
driver = sys.argv[3] if len(sys.argv) > 3 else "hybrid"
kt = int(sys.argv[4]) if len(sys.argv) > 4 else LS.SE
pkg.hostlib.set_search_driver({"native": 0, "hybrid": 1, "reference": 2}[driver])
L = LS.LoopLib("b200")
L.srand(1)
opt = L.sls(D, True, True, kt, LS.EI)
opt.set_hyperparams(0.5, 0.5, 0.001, 0.1, 0.01)
out = (C.c_double * 4)()
print(f"D = {D}, driver {driver}, kernel {kt}: iteration, N, total ms | MAP fit ms (evaluations) | search ms | slider ms")
tot = np.zeros(4)
for it in range(iters):
    e0, e1 = opt.slider_ends()
    t0 = time.perf_counter()
    opt.submit(LS.best_slider_position(e0, e1))
    ms = (time.perf_counter() - t0) * 1e3
    L.lib.b200_sls_last_step_timings(C.c_void_p(opt.h), out)
    tot += np.array([ms, out[0], out[1], out[2]])
    print(f"{it:4d} {opt.num_points():4d} {ms:8.1f} | {out[0]:8.1f} ({int(out[3]):4d}) | {out[1]:8.1f} | {out[2]:6.2f}")
print(f"sum: total {tot[0]:.0f} ms, MAP fit {tot[1]:.0f}, search {tot[2]:.0f}, slider {tot[3]:.1f}")
