"""Where the time of one SubmitFeedbackData goes (GPU box): MAP fit vs acquisition search, per data size.
python tools/submit_timing.py [D] [kernel: se|matern]"""
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tools")]
import synth  # noqa: E402

pkg = importlib.import_module("sequential-line-search_b200")
host = pkg.hostlib.Host()
D = int(sys.argv[1]) if len(sys.argv) > 1 else 6
kt = 0 if (len(sys.argv) > 2 and sys.argv[2] == "se") else 1
for N in (9, 30, 60, 150, 300):
    X = synth.make_X(N, D, "sls")
    offsets, idx = synth.make_tuples(X)
    for use_map in (False, True):
        t0 = time.time()
        h = host.pref_create(kt, X, offsets, idx, use_map, 0.5, 0.5, 0.005, 0.25, 0.01)
        t1 = time.time()
        reg = host.pref_regressor(h)
        x = host.find_next_point(reg, D, n_global=50 * D, n_local=10 * D)
        t2 = time.time()
        print(f"N={N:4d} D={D} use_map={use_map!s:5}: MAP fit {1e3 * (t1 - t0):7.1f} ms ({host.pref_num_map_evaluations(h)} evaluations)   "
              f"FindNextPoint {1e3 * (t2 - t1):7.1f} ms", flush=True)
        host.pref_destroy(h)
