"""Time slsgp_append_point (bordered O(N^2) update) against a from-scratch refit (set_data + gram + factor + inverse +
alpha) of the grown data set: the step FindNextPoints repeats once per pending option."""
import importlib
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
import synth  # noqa: E402

pkg = importlib.import_module("sequential-line-search_b200")


def main():
    D = 16
    ctx = pkg.Context(0)
    for N in (256, 1024, 2040):
        X, theta = synth.make_X(N + 8, D, "sls"), synth.make_theta(D, "default")
        y = synth.make_y(X)
        ctx.fit(X[:, :N], 0, theta, 0.005, y[:N])
        t_app = []
        for n in range(N, N + 8):               # N + 8 stays inside the padded leading dimension for these N
            t0 = time.perf_counter()
            ctx.append_point(X[:, n], y[n], want=False)
            t_app.append(time.perf_counter() - t0)
        t_fit = []
        for _ in range(5):
            t0 = time.perf_counter()
            ctx.fit(X, 0, theta, 0.005, y)
            t_fit.append(time.perf_counter() - t0)
        print(f"N={N:5d} D={D}: append_point {1e3 * np.median(t_app):7.3f} ms   full refit {1e3 * np.median(t_fit):7.3f} ms "
              f"(host wall clock, medians)")
    ctx.close()


if __name__ == "__main__":
    main()
