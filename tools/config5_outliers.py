import os, sys, importlib
ROOT = "/root/repo" if os.path.isdir("/root/repo/tests") else os.getcwd()
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")]
import loop_support as LS
b200 = LS.LoopLib("b200")
h = (0.5, 0.5, 0.001, 0.1, 0.01)
LS.run_sls_loop(b200, 6, 3, 1, hyper=h)
log = LS.run_sls_loop(b200, 64, 200, 1, kt=LS.SE, hyper=h)
ms = [r["ms"] for r in log]
import numpy as np
print("total", sum(ms) / 1e3, "median", np.median(ms))
order = np.argsort(ms)[::-1][:12]
print("slowest iterations (iter, n_points, ms):", [(int(i), log[i]["n_points"], round(ms[i], 1)) for i in order])
print("first 10:", [round(m, 1) for m in ms[:10]])
