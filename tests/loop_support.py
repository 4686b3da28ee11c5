"""ctypes binding for the optimiser-level C facade (sequential-line-search_b200/host/src/loop_capi.inl), which is compiled
twice: with the ``ref_`` prefix into oracle/_ref/libsls_ref_loop.so (the reference's unmodified sources + the real NLopt;
TEST INFRASTRUCTURE) and with the ``b200_`` prefix into libsls_b200_host.so (the product's host layer on the GPU). The
step-level parity tests and tools/loop_timing.py drive both sides through the one class below.
"""
from __future__ import annotations

import ctypes as C
import os
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_LOOP_PATH = os.path.join(ROOT, "oracle", "_ref", "libsls_ref_loop.so")
REF_LOOP_FMA_PATH = os.path.join(ROOT, "oracle", "_ref", "libsls_ref_loop_fma.so")  # same sources, FMA contraction allowed
HOST_PATH = os.path.join(ROOT, "sequential-line-search_b200", "lib", "libsls_b200_host.so")

c_dp = C.POINTER(C.c_double)
c_up = C.POINTER(C.c_uint)
SE, MATERN = 0, 1
EI, UCB = 0, 1


def _f64(a):
    return np.require(np.asarray(a, dtype=np.float64), requirements=["F", "A"])


def _p(a):
    return None if a is None else a.ctypes.data_as(c_dp)


def ref_loop_available() -> bool:
    return os.path.exists(REF_LOOP_PATH)


class LoopLib:
    """One side of the comparison. ``side`` is "ref" (CPU reference), "ref_fma" (the same reference sources compiled with FMA
    contraction: the yardstick for the reference's own reproducibility) or "b200" (GPU host layer)."""

    _POINTER = ("sls_create", "pbo_create", "pref_fit", "pref_fit_regressor", "gpr_fit", "gpr_given", "gpr_fit_regressor")
    _DOUBLE = ("sls_query", "loop_acq_value", "loop_predict")

    def __init__(self, side: str):
        assert side in ("ref", "ref_fma", "b200")
        self.lib = C.CDLL({"ref": REF_LOOP_PATH, "ref_fma": REF_LOOP_FMA_PATH, "b200": HOST_PATH}[side])
        self.side = side = "ref" if side == "ref_fma" else side
        for n in self._POINTER:
            self.fn(n).restype = C.c_void_p
        for n in self._DOUBLE:
            self.fn(n).restype = C.c_double
        err = self.lib.ref_loop_last_error if side == "ref" else self.lib.b200_last_error
        err.restype = C.c_char_p
        self._err = err

    def fn(self, name):
        return getattr(self.lib, f"{self.side}_{name}")

    def check(self, ok, what):
        if not ok:
            raise RuntimeError(f"{self.side}_{what}: {self._err().decode()}")

    def srand(self, seed: int):
        self.fn("srand")(C.c_uint(seed))

    # ---- regressors -------------------------------------------------------------------------------------------------
    def pref_fit(self, kt, X, tuples, use_map, a=0.5, r=0.5, b=0.005, var=0.25, btl=0.01, num_iters=100):
        X = _f64(X)
        D, N = X.shape
        offsets = np.zeros(len(tuples) + 1, dtype=np.uint32)
        offsets[1:] = np.cumsum([len(t) for t in tuples])
        idx = np.array([i for t in tuples for i in t], dtype=np.uint32)
        h = self.fn("pref_fit")(kt, D, N, _p(X), len(tuples), offsets.ctypes.data_as(c_up), idx.ctypes.data_as(c_up), int(use_map),
                                C.c_double(a), C.c_double(r), C.c_double(b), C.c_double(var), C.c_double(btl), C.c_uint(num_iters))
        self.check(h, "pref_fit")
        return PrefFit(self, h, D, N)

    def gpr_fit(self, kt, X, y):
        X, y = _f64(X), _f64(y)
        D, N = X.shape
        theta, b = np.zeros(D + 1), C.c_double(0.0)
        h = self.fn("gpr_fit")(kt, D, N, _p(X), _p(y), _p(theta), C.byref(b))
        self.check(h, "gpr_fit")
        return GprFit(self, h, D, theta, b.value)

    def gpr_given(self, kt, X, y, theta, b):
        X, y, theta = _f64(X), _f64(y), _f64(theta)
        D, N = X.shape
        h = self.fn("gpr_given")(kt, D, N, _p(X), _p(y), _p(theta), C.c_double(b))
        self.check(h, "gpr_given")
        return GprFit(self, h, D, theta, b)

    # ---- acquisition through the Regressor interface ----------------------------------------------------------------
    def find_next_point(self, reg, D, n_global, n_local, acq=EI, beta=1.0):
        x = np.zeros(D)
        self.check(self.fn("loop_find_next_point")(C.c_void_p(reg), D, C.c_uint(n_global), C.c_uint(n_local), acq, C.c_double(beta), _p(x)) == 0,
                   "loop_find_next_point")
        return x

    def find_next_points(self, reg, D, n_points, n_global, n_local, acq=EI, beta=1.0):
        X = np.zeros((n_points, D))
        self.check(self.fn("loop_find_next_points")(C.c_void_p(reg), D, C.c_uint(n_points), C.c_uint(n_global), C.c_uint(n_local), acq,
                                                    C.c_double(beta), _p(X)) == 0, "loop_find_next_points")
        return X

    def acq_value(self, reg, x, acq=EI, beta=1.0):
        x = _f64(x)
        return self.fn("loop_acq_value")(C.c_void_p(reg), len(x), acq, C.c_double(beta), _p(x))

    def acq_derivative(self, reg, x, acq=EI, beta=1.0):
        x = _f64(x)
        g = np.zeros(len(x))
        self.check(self.fn("loop_acq_derivative")(C.c_void_p(reg), len(x), acq, C.c_double(beta), _p(x), _p(g)) == 0, "loop_acq_derivative")
        return g

    def predict(self, reg, x, what=0):
        x = _f64(x)
        return self.fn("loop_predict")(C.c_void_p(reg), len(x), what, _p(x))

    def slider(self, e0, e1, enlarge=True):
        e0, e1 = _f64(e0), _f64(e1)
        o0, o1 = np.zeros(len(e0)), np.zeros(len(e0))
        self.fn("loop_slider")(len(e0), _p(e0), _p(e1), int(enlarge), _p(o0), _p(o1))
        return o0, o1

    # ---- front-ends -----------------------------------------------------------------------------------------------------
    def sls(self, D, enlarge=True, use_map=True, kt=MATERN, acq=EI, strategy=0, init_ends=None):
        ends = None if init_ends is None else _f64(np.concatenate([np.asarray(init_ends[0], float), np.asarray(init_ends[1], float)]))
        h = self.fn("sls_create")(D, int(enlarge), int(use_map), kt, acq, strategy, _p(ends))
        self.check(h, "sls_create")
        return Sls(self, h, D)

    def pbo(self, D, use_map=True, kt=MATERN, acq=EI, strategy=0, num_options=2, init_options=None):
        opts = None if init_options is None else _f64(np.asarray(init_options, float).reshape(num_options, D).T).T.copy()
        h = self.fn("pbo_create")(D, int(use_map), kt, acq, strategy, num_options, None if opts is None else opts.ctypes.data_as(c_dp))
        self.check(h, "pbo_create")
        return Pbo(self, h, D, num_options)


class PrefFit:
    def __init__(self, L, h, D, N):
        self.L, self.h, self.D, self.N = L, h, D, N
        self.reg = L.fn("pref_fit_regressor")(C.c_void_p(h))

    def state(self):
        y, theta, b = np.zeros(self.N), np.zeros(self.D + 1), C.c_double(0.0)
        self.L.fn("pref_fit_get_state")(C.c_void_p(self.h), _p(y), _p(theta), C.byref(b))
        return y, theta, b.value

    def find_arg_max(self):
        x = np.zeros(self.D)
        self.L.fn("pref_fit_find_arg_max")(C.c_void_p(self.h), _p(x))
        return x

    def close(self):
        if self.h:
            self.L.fn("pref_fit_destroy")(C.c_void_p(self.h))
            self.h = None


class GprFit:
    def __init__(self, L, h, D, theta, b):
        self.L, self.h, self.D, self.theta, self.b = L, h, D, np.array(theta), b
        self.reg = L.fn("gpr_fit_regressor")(C.c_void_p(h))

    def close(self):
        if self.h:
            self.L.fn("gpr_fit_destroy")(C.c_void_p(self.h))
            self.h = None


class Sls:
    def __init__(self, L, h, D):
        self.L, self.h, self.D = L, h, D

    def set_hyperparams(self, a=0.5, r=0.5, b=0.005, var=0.25, btl=0.01):
        self.L.fn("sls_set_hyperparams")(C.c_void_p(self.h), C.c_double(a), C.c_double(r), C.c_double(b), C.c_double(var), C.c_double(btl))

    def submit(self, position, n_map=-1, n_global=0, n_local=0):
        self.L.check(self.L.fn("sls_submit")(C.c_void_p(self.h), C.c_double(position), n_map, n_global, n_local) == 0, "sls_submit")

    def slider_ends(self):
        e0, e1 = np.zeros(self.D), np.zeros(self.D)
        self.L.fn("sls_get_slider_ends")(C.c_void_p(self.h), _p(e0), _p(e1))
        return e0, e1

    def maximizer(self):
        x = np.zeros(self.D)
        self.L.fn("sls_get_maximizer")(C.c_void_p(self.h), _p(x))
        return x

    def calc_point(self, position):
        x = np.zeros(self.D)
        self.L.fn("sls_calc_point")(C.c_void_p(self.h), C.c_double(position), _p(x))
        return x

    def num_points(self):
        return self.L.fn("sls_num_points")(C.c_void_p(self.h))

    def raw_data_points(self):
        X = np.zeros((self.D, self.num_points()), order="F")
        self.L.fn("sls_get_raw_data_points")(C.c_void_p(self.h), _p(X))
        return X

    def query(self, what, x):
        return self.L.fn("sls_query")(C.c_void_p(self.h), what, _p(_f64(x)))

    def close(self):
        if self.h:
            self.L.fn("sls_destroy")(C.c_void_p(self.h))
            self.h = None


class Pbo:
    def __init__(self, L, h, D, num_options):
        self.L, self.h, self.D, self.n = L, h, D, num_options

    def set_hyperparams(self, a=0.5, r=0.5, b=0.005, var=0.25, btl=0.01):
        self.L.fn("pbo_set_hyperparams")(C.c_void_p(self.h), C.c_double(a), C.c_double(r), C.c_double(b), C.c_double(var), C.c_double(btl))

    def submit(self, option_index, n_map=0):
        self.L.check(self.L.fn("pbo_submit")(C.c_void_p(self.h), option_index, n_map) == 0, "pbo_submit")

    def determine_next_query(self, n_global=0, n_local=0):
        self.L.check(self.L.fn("pbo_determine_next_query")(C.c_void_p(self.h), n_global, n_local) == 0, "pbo_determine_next_query")

    def current_options(self):
        out = np.zeros((self.n, self.D))
        self.L.fn("pbo_get_current_options")(C.c_void_p(self.h), _p(out))
        return out

    def close(self):
        if self.h:
            self.L.fn("pbo_destroy")(C.c_void_p(self.h))
            self.h = None


# ---- the simulated user of the reference's nd demo (demos/sequential_line_search_nd/main.cpp:26-34, 63-76) ----------
def demo_objective(x, centre=0.4):
    x = np.asarray(x, float)
    return float(np.exp(-np.sum((x - centre) ** 2)))


def best_slider_position(e0, e1, centre=0.4):
    """arg max over t in [0, 1] of exp(-|e0 + t (e1 - e0) - c|^2): the closed form of the demo's 1e-5 line scan."""
    d = e1 - e0
    dd = float(d @ d)
    if dd == 0.0:
        return 0.0
    return float(np.clip(((centre - e0) @ d) / dd, 0.0, 1.0))


def run_sls_loop(L: LoopLib, D, iters, seed, kt=MATERN, acq=EI, use_map=True, enlarge=True, hyper=None, budgets=None, positions=None):
    """The nd demo's loop. Returns per-iteration dicts: slider ends after the submit, chosen position, wall ms, objective.
    `positions` (a list) replays a fixed sequence of slider positions instead of the simulated user."""
    L.srand(seed)
    opt = L.sls(D, enlarge, use_map, kt, acq)
    if hyper:
        opt.set_hyperparams(*hyper)
    log = []
    for it in range(iters):
        e0, e1 = opt.slider_ends()
        t_star = positions[it] if positions is not None else best_slider_position(e0, e1)
        t0 = time.perf_counter()
        if budgets:
            opt.submit(t_star, *budgets)
        else:
            opt.submit(t_star)
        ms = (time.perf_counter() - t0) * 1e3
        n0, n1 = opt.slider_ends()
        log.append({"iter": it, "position": t_star, "end_0": n0, "end_1": n1, "ms": ms, "n_points": opt.num_points(),
                    "objective": demo_objective(opt.maximizer())})
    opt.close()
    return log
