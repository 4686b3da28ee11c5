"""Config 5 (D = 64, 200 x SubmitFeedbackData, SE kernel, EI, MAP hyper-parameters) through the C++ host layer on one GPU, with
the slowest iterations listed: they sit where the number of data points crosses a multiple of 64 (65, 129, 195, 257, 321, 385) -
the leading dimension of every N x N device buffer grows there and each pooled context re-allocates. The 2x allocation headroom of
slsgp.cu:ensure() came out of this listing (12.0 -> 11.2 s in total on the same box). Not part of the product path.
    python tools/config5_outliers.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")]
import loop_support as LS  # noqa: E402

b200 = LS.LoopLib("b200")
hyper = (0.5, 0.5, 0.001, 0.1, 0.01)
LS.run_sls_loop(b200, 6, 3, 1, hyper=hyper)  # warm-up
log = LS.run_sls_loop(b200, 64, 200, 1, kt=LS.SE, hyper=hyper)
ms = [r["ms"] for r in log]
print("total", sum(ms) / 1e3, "s, median", np.median(ms), "ms per iteration")
order = np.argsort(ms)[::-1][:12]
print("slowest iterations (iteration, data points, ms):", [(int(i), log[i]["n_points"], round(ms[i], 1)) for i in order])
print("first 10 iterations, ms:", [round(m, 1) for m in ms[:10]])
