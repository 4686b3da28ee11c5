"""Shared test plumbing: ctypes bindings for the CPU checkers and the synthetic-input generators.

* ``Oracle``  – oracle/libslsgp_oracle.so, the plain-C restatement (oracle/slsgp_oracle.c).
* ``Ref``     – oracle/_ref/libsls_ref_probe.so, the reference's unmodified sources compiled here
                (oracle/Makefile). Present in this container and, as a prebuilt file, on the GPU box.
Both are TEST INFRASTRUCTURE: nothing under sequential-line-search_b200/ imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

c_dp = C.POINTER(C.c_double)
c_up = C.POINTER(C.c_uint)

SE, MATERN = 0, 1
EI, UCB = 0, 1


def _p(a):
    return None if a is None else a.ctypes.data_as(c_dp)


def _u(a):
    return a.ctypes.data_as(c_up)


def build_oracle():
    """(Re)build the checkers with oracle/Makefile. `ref` is a no-op when /root/reference is absent."""
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, "oracle", "ref"], check=True)


sys.path.insert(0, os.path.join(ROOT, "tools"))
from synth import make_X, make_theta, nd_demo_objective, make_y, make_tuples, make_queries, f64  # noqa: E402,F401

# ----------------------------------------------------------------------------------------------------------------
# plain-C oracle
# ----------------------------------------------------------------------------------------------------------------
class _Model(C.Structure):
    _fields_ = [("kernel_type", C.c_int), ("D", C.c_int), ("N", C.c_int), ("X", c_dp), ("theta", c_dp),
                ("b", C.c_double), ("y", c_dp), ("L", c_dp)]


class Oracle:
    def __init__(self):
        path = os.path.join(ORACLE_DIR, "libslsgp_oracle.so")
        if not os.path.exists(path):
            build_oracle()
        self.lib = lib = C.CDLL(path)
        for name in ("kernel", "logdet", "predict_mu", "predict_sigma", "acq_value", "btl", "log_lognormal",
                     "log_lognormal_derivative", "map_objective_pref", "map_objective_gpr"):
            getattr(lib, "slsgp_oracle_" + name).restype = C.c_double

    # kernels -------------------------------------------------------------------------------------------------
    def kernel(self, kt, xa, xb, theta):
        xa, xb, theta = f64(xa), f64(xb), f64(theta)
        return self.lib.slsgp_oracle_kernel(kt, len(xa), _p(xa), _p(xb), _p(theta))

    def kernel_theta_derivative(self, kt, xa, xb, theta):
        xa, xb, theta = f64(xa), f64(xb), f64(theta)
        out = np.empty(len(theta))
        self.lib.slsgp_oracle_kernel_theta_derivative(kt, len(xa), _p(xa), _p(xb), _p(theta), _p(out))
        return out

    def kernel_first_arg_derivative(self, kt, xa, xb, theta):
        xa, xb, theta = f64(xa), f64(xb), f64(theta)
        out = np.empty(len(xa))
        self.lib.slsgp_oracle_kernel_first_arg_derivative(kt, len(xa), _p(xa), _p(xb), _p(theta), _p(out))
        return out

    def large_ky(self, kt, X, theta, b):
        X, theta = f64(X), f64(theta)
        D, N = X.shape
        K = np.empty((N, N), order="F")
        self.lib.slsgp_oracle_large_ky(kt, D, N, _p(X), _p(theta), C.c_double(b), _p(K))
        return K

    def small_k(self, kt, X, theta, x):
        X, theta, x = f64(X), f64(theta), f64(x)
        D, N = X.shape
        k = np.empty(N)
        self.lib.slsgp_oracle_small_k(kt, D, N, _p(X), _p(theta), _p(x), _p(k))
        return k

    def small_k_x_derivative(self, kt, X, theta, x):
        X, theta, x = f64(X), f64(theta), f64(x)
        D, N = X.shape
        J = np.empty((D, N), order="F")
        self.lib.slsgp_oracle_small_k_x_derivative(kt, D, N, _p(X), _p(theta), _p(x), _p(J))
        return J

    def large_ky_theta_derivative(self, kt, X, theta):
        X, theta = f64(X), f64(theta)
        D, N = X.shape
        out = np.empty((D + 1, N * N))
        self.lib.slsgp_oracle_large_ky_theta_derivative(kt, D, N, _p(X), _p(theta), _p(out))
        return out.reshape(D + 1, N, N).transpose(0, 2, 1)  # [t] -> (i, j) of column-major N x N

    # linear algebra ------------------------------------------------------------------------------------------
    def cholesky(self, K):
        K = f64(K)
        N = K.shape[0]
        L = np.empty((N, N), order="F")
        status = self.lib.slsgp_oracle_cholesky(N, _p(K), _p(L))
        return L, status

    def llt_solve(self, L, B):
        L = f64(L)
        B = f64(np.array(B, dtype=np.float64, copy=True))
        nrhs = 1 if B.ndim == 1 else B.shape[1]
        self.lib.slsgp_oracle_llt_solve(L.shape[0], _p(L), nrhs, _p(B))
        return B

    def logdet(self, L):
        L = f64(L)
        return self.lib.slsgp_oracle_logdet(L.shape[0], _p(L))

    def inverse(self, K):
        K = f64(K)
        N = K.shape[0]
        Kinv = np.empty((N, N), order="F")
        self.lib.slsgp_oracle_inverse(N, _p(K), _p(Kinv))
        return Kinv

    # model ---------------------------------------------------------------------------------------------------
    def model(self, kt, X, theta, b, y):
        X, theta, y = f64(X), f64(theta), f64(y)
        K = self.large_ky(kt, X, theta, b)
        L, status = self.cholesky(K)
        assert status == 0
        m = _Model(kt, X.shape[0], X.shape[1], _p(X), _p(theta), b, _p(y), _p(L))
        m._keep = (X, theta, y, L, K)
        return m

    def predict(self, m, x):
        x = f64(x)
        D = m.D
        dmu, dsg = np.empty(D), np.empty(D)
        mu = self.lib.slsgp_oracle_predict_mu(C.byref(m), _p(x))
        sg = self.lib.slsgp_oracle_predict_sigma(C.byref(m), _p(x))
        self.lib.slsgp_oracle_predict_mu_derivative(C.byref(m), _p(x), _p(dmu))
        self.lib.slsgp_oracle_predict_sigma_derivative(C.byref(m), _p(x), _p(dsg))
        return mu, sg, dmu, dsg

    def f_best(self, m):
        fb = C.c_double()
        i = self.lib.slsgp_oracle_predict_maximum_point_from_data(C.byref(m), C.byref(fb))
        return i, fb.value

    def acq(self, m, acq_type, beta, f_best, x):
        x = f64(x)
        g = np.empty(m.D)
        v = self.lib.slsgp_oracle_acq_value(C.byref(m), acq_type, C.c_double(beta), C.c_double(f_best), _p(x))
        self.lib.slsgp_oracle_acq_derivative(C.byref(m), acq_type, C.c_double(beta), C.c_double(f_best), _p(x), _p(g))
        return v, g

    def acq_batch(self, m, acq_type, beta, f_best, Xq):
        Xq = f64(Xq)
        D, M = Xq.shape
        out = dict(mu=np.empty(M), sigma=np.empty(M), dmu=np.empty((D, M), order="F"),
                   dsigma=np.empty((D, M), order="F"), val=np.empty(M), grad=np.empty((D, M), order="F"))
        self.lib.slsgp_oracle_acq_batch(C.byref(m), acq_type, C.c_double(beta), C.c_double(f_best), C.c_longlong(M),
                                        _p(Xq), _p(out["mu"]), _p(out["sigma"]), _p(out["dmu"]), _p(out["dsigma"]),
                                        _p(out["val"]), _p(out["grad"]))
        return out

    # likelihood / priors / MAP objectives --------------------------------------------------------------------
    def btl(self, f, scale):
        f = f64(f)
        d = np.empty(len(f))
        v = self.lib.slsgp_oracle_btl(len(f), _p(f), C.c_double(scale))
        self.lib.slsgp_oracle_btl_derivative(len(f), _p(f), C.c_double(scale), _p(d))
        return v, d

    def map_objective_pref(self, kt, X, offsets, idx, use_map, a, r, b, prior_var, btl_scale, x, want_grad=True):
        X, x = f64(X), f64(x)
        D, N = X.shape
        g = np.empty(len(x)) if want_grad else None
        f = self.lib.slsgp_oracle_map_objective_pref(kt, D, N, _p(X), len(offsets) - 1, _u(offsets), _u(idx),
                                                     int(use_map), C.c_double(a), C.c_double(r), C.c_double(b),
                                                     C.c_double(prior_var), C.c_double(btl_scale), _p(x), _p(g))
        return f, g

    def map_objective_pref_noiseless(self, kt, X, offsets, idx, use_map, a, r, b, prior_var, btl_scale, x, want_grad=True):
        """The objective of the reference's SEQUENTIAL_LINE_SEARCH_USE_NOISELESS_FORMULATION build."""
        X, x = f64(X), f64(x)
        D, N = X.shape
        g = np.empty(len(x)) if want_grad else None
        self.lib.slsgp_oracle_map_objective_pref_noiseless.restype = C.c_double
        f = self.lib.slsgp_oracle_map_objective_pref_noiseless(kt, D, N, _p(X), len(offsets) - 1, _u(offsets), _u(idx), int(use_map), C.c_double(a),
                                                               C.c_double(r), C.c_double(b), C.c_double(prior_var), C.c_double(btl_scale), _p(x), _p(g))
        return f, g

    def map_objective_gpr(self, kt, X, y, x, want_grad=True):
        X, y, x = f64(X), f64(y), f64(x)
        D, N = X.shape
        g = np.empty(len(x)) if want_grad else None
        f = self.lib.slsgp_oracle_map_objective_gpr(kt, D, N, _p(X), _p(y), _p(x), _p(g))
        return f, g


# ----------------------------------------------------------------------------------------------------------------
# the reference itself (compiled from /root/reference by oracle/Makefile)
# ----------------------------------------------------------------------------------------------------------------
REF_PATH = os.path.join(ORACLE_DIR, "_ref", "libsls_ref_probe.so")


def ref_available():
    return os.path.exists(REF_PATH)


REF_NOISELESS_PATH = os.path.join(ORACLE_DIR, "_ref", "libsls_ref_probe_noiseless.so")  # built with SEQUENTIAL_LINE_SEARCH_USE_NOISELESS_FORMULATION


class Ref:
    def __init__(self, path=None):
        self.lib = lib = C.CDLL(path or REF_PATH)
        for name in ("ref_pref_create", "ref_pref_regressor", "ref_gpr_create", "ref_gpr_regressor"):
            getattr(lib, name).restype = C.c_void_p
        for name in ("ref_pref_objective", "ref_predict_mu", "ref_predict_sigma", "ref_acq_value", "ref_btl"):
            getattr(lib, name).restype = C.c_double
        lib.ref_pref_opt_dim.restype = C.c_uint

    def kernel(self, kt, xa, xb, theta):
        xa, xb, theta = f64(xa), f64(xb), f64(theta)
        D = len(xa)
        k = C.c_double()
        dth, dx = np.empty(D + 1), np.empty(D)
        self.lib.ref_kernel(kt, D, _p(xa), _p(xb), _p(theta), C.byref(k), _p(dth), _p(dx))
        return k.value, dth, dx

    def large_ky(self, kt, X, theta, b):
        X, theta = f64(X), f64(theta)
        D, N = X.shape
        K = np.empty((N, N), order="F")
        self.lib.ref_calc_large_ky(kt, D, N, _p(X), _p(theta), C.c_double(b), _p(K))
        return K

    def small_k(self, kt, X, theta, x):
        X, theta, x = f64(X), f64(theta), f64(x)
        D, N = X.shape
        k, J = np.empty(N), np.empty((D, N), order="F")
        self.lib.ref_calc_small_k(kt, D, N, _p(X), _p(theta), _p(x), _p(k), _p(J))
        return k, J

    def large_ky_theta_derivative(self, kt, X, theta):
        X, theta = f64(X), f64(theta)
        D, N = X.shape
        out = np.empty((D + 1, N * N))
        self.lib.ref_calc_large_ky_theta_derivative(kt, D, N, _p(X), _p(theta), _p(out))
        return out.reshape(D + 1, N, N).transpose(0, 2, 1)

    # PreferenceRegressor -------------------------------------------------------------------------------------
    def pref_create(self, kt, X, offsets, idx, use_map, a, r, b, prior_var, btl_scale, solution):
        X, solution = f64(X), f64(solution)
        D, N = X.shape
        h = self.lib.ref_pref_create(kt, D, N, _p(X), len(offsets) - 1, _u(offsets), _u(idx), int(use_map),
                                     C.c_double(a), C.c_double(r), C.c_double(b), C.c_double(prior_var),
                                     C.c_double(btl_scale), _p(solution))
        return C.c_void_p(h)

    def pref_destroy(self, h):
        self.lib.ref_pref_destroy(h)

    def pref_regressor(self, h):
        return C.c_void_p(self.lib.ref_pref_regressor(h))

    def pref_objective(self, h, x, want_grad=True):
        x = f64(x)
        assert len(x) == self.lib.ref_pref_opt_dim(h)
        g = np.empty(len(x)) if want_grad else None
        return self.lib.ref_pref_objective(h, _p(x), _p(g)), g

    def pref_state(self, h, N, D):
        y, theta, b = np.empty(N), np.empty(D + 1), C.c_double()
        K, L = np.empty((N, N), order="F"), np.empty((N, N), order="F")
        self.lib.ref_pref_get_state(h, _p(y), _p(theta), C.byref(b), _p(K), _p(L))
        return dict(y=y, theta=theta, b=b.value, K=K, L=L)

    # GaussianProcessRegressor --------------------------------------------------------------------------------
    def gpr_create(self, kt, X, y, theta, b):
        X, y, theta = f64(X), f64(y), f64(theta)
        D, N = X.shape
        return C.c_void_p(self.lib.ref_gpr_create(kt, D, N, _p(X), _p(y), _p(theta), C.c_double(b)))

    def gpr_destroy(self, h):
        self.lib.ref_gpr_destroy(h)

    def gpr_regressor(self, h):
        return C.c_void_p(self.lib.ref_gpr_regressor(h))

    def gpr_state(self, h, N):
        K, Kinv = np.empty((N, N), order="F"), np.empty((N, N), order="F")
        self.lib.ref_gpr_get_state(h, _p(K), _p(Kinv))
        return K, Kinv

    def gpr_objective(self, kt, X, y, points, want_grad=True):
        X, y = f64(X), f64(y)
        points = np.ascontiguousarray(points, dtype=np.float64)
        D, N = X.shape
        n = points.shape[0]
        f = np.empty(n)
        g = np.empty((n, D + 2)) if want_grad else None
        self.lib.ref_gpr_objective(kt, D, N, _p(X), _p(y), n, points.ctypes.data_as(c_dp), _p(f),
                                   None if g is None else g.ctypes.data_as(c_dp))
        return f, g

    # Regressor virtuals + acquisition ------------------------------------------------------------------------
    def predict(self, reg, x):
        x = f64(x)
        D = len(x)
        dmu, dsg = np.empty(D), np.empty(D)
        mu = self.lib.ref_predict_mu(reg, D, _p(x))
        sg = self.lib.ref_predict_sigma(reg, D, _p(x))
        self.lib.ref_predict_mu_derivative(reg, D, _p(x), _p(dmu))
        self.lib.ref_predict_sigma_derivative(reg, D, _p(x), _p(dsg))
        return mu, sg, dmu, dsg

    def x_best(self, reg, D):
        out = np.empty(D)
        self.lib.ref_predict_maximum_point_from_data(reg, D, _p(out))
        return out

    def acq(self, reg, acq_type, beta, x, want_grad=True):
        x = f64(x)
        D = len(x)
        v = self.lib.ref_acq_value(reg, D, acq_type, C.c_double(beta), _p(x))
        g = None
        if want_grad:
            g = np.empty(D)
            self.lib.ref_acq_derivative(reg, D, acq_type, C.c_double(beta), _p(x), _p(g))
        return v, g

    def btl(self, f, scale):
        f = f64(f)
        d = np.empty(len(f))
        return self.lib.ref_btl(len(f), _p(f), C.c_double(scale), _p(d)), d

    def data_manager_run(self, batches, eps=1e-4):
        """PreferenceDataManager::AddNewPoints over a list of (D x k) batches (first column preferred)."""
        D = batches[0].shape[0]
        sizes = np.asarray([b.shape[1] for b in batches], dtype=np.int32)
        pts = f64(np.concatenate([np.asarray(b, dtype=np.float64).T.reshape(-1) for b in batches]))
        total = int(sizes.sum())
        X = np.empty(D * total)
        offsets, idx = np.zeros(len(batches) + 1, dtype=np.uint32), np.zeros(total, dtype=np.uint32)
        n = self.lib.ref_data_manager_run(D, len(batches), sizes.ctypes.data_as(C.POINTER(C.c_int)), _p(pts), C.c_double(eps),
                                          _p(X), _u(offsets), _u(idx))
        return X[:D * n].reshape((D, n), order="F"), offsets, idx


def rel_err(a, b, floor=1e-300):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), floor))
