"""BASELINE config 3: PreferenceRegressor MAP fit with hyper-parameters at N = 2048, D = 16 (683 preference triples,
2066 variables) through the C++ host layer. Prints wall-clock time and the number of objective evaluations."""
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
N, D = 2048, 16
X = synth.make_X(N, D, "sls")
offsets, idx = synth.make_tuples(X)
h0 = host.pref_create(0, X[:, :30], offsets[:11], idx[:offsets[10]], False, 0.5, 0.5, 0.005, 0.25, 0.01)  # warm-up: context, allocations
for use_map in (False, True):
    t0 = time.time()
    h = host.pref_create(0, X, offsets, idx, use_map, 0.5, 0.5, 0.005, 0.25, 0.01)
    dt = time.time() - t0
    st = host.pref_state(h, N, D)
    print(f"N={N} D={D} P={len(offsets) - 1} use_map_hyperparams={use_map}: fit {dt:.3f} s, {host.pref_num_map_evaluations(h)} objective evaluations, "
          f"a={st['theta'][0]:.4f} b={st['b']:.5f} r[:4]={np.round(st['theta'][1:5], 4)} max|y|={np.abs(st['y']).max():.4f}", flush=True)
    host.pref_destroy(h)
