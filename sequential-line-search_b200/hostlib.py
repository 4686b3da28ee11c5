"""ctypes view of libsls_b200_host.so, the C++ host layer (host/): the reference's Regressor / acquisition_func
interface served by libslsgp. The extern "C" facade (host/src/capi.cpp) offers the same handles the test suite binds
onto the reference classes, so tests drive both with identical arguments. Used by tests and smoke only; C++ users link
the library and include host/include/sequential-line-search/*.hpp directly."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .build import HOST_LIB_PATH, build_host

c_dp = C.POINTER(C.c_double)
c_up = C.POINTER(C.c_uint)
_lib = None

_DOUBLE = ("b200_sls_model_vs_rebuilt_copy", "b200_time_acq_values", "b200_pref_objective", "b200_predict_mu", "b200_predict_sigma", "b200_acq_value", "b200_utils_btl")
_POINTER = ("b200_gpr_copy", "b200_gpr_create", "b200_gpr_create_map", "b200_gpr_regressor", "b200_pref_create", "b200_pref_create_warm", "b200_pref_regressor")
HOST_SYMBOLS = [
    "b200_last_error", "b200_kernel", "b200_calc_large_ky", "b200_gpr_create", "b200_gpr_create_map", "b200_gpr_destroy", "b200_gpr_append_point",
    "b200_gpr_regressor", "b200_gpr_get_state", "b200_pref_create", "b200_pref_create_warm", "b200_pref_destroy", "b200_pref_regressor",
    "b200_pref_objective", "b200_pref_num_map_evaluations", "b200_pref_get_state", "b200_pref_find_arg_max",
    "b200_pref_damp_data", "b200_predict_mu", "b200_predict_sigma", "b200_predict_mu_derivative",
    "b200_predict_sigma_derivative", "b200_predict_maximum_point_from_data", "b200_predict_batch", "b200_acq_value",
    "b200_acq_derivative", "b200_acq_values", "b200_find_next_point", "b200_find_next_points", "b200_data_manager_run", "b200_slider", "b200_test_minimize",
    "b200_utils_btl", "b200_utils_random_vector", "b200_utils_export_csv",
    "b200_nlopt_available", "b200_get_search_driver", "b200_set_search_driver", "b200_calc_small_k", "b200_calc_large_ky_theta_derivative",
    "b200_release_device_resources", "b200_gpr_copy", "b200_gpr_num_points", "b200_regressor_set_sweep_mode", "b200_set_devices", "b200_get_device_count", "b200_time_acq_values",
    "b200_get_incremental_refit", "b200_set_incremental_refit", "b200_pref_num_points_kept", "b200_sls_num_points_kept", "b200_sls_last_step_timings", "b200_sls_model_vs_rebuilt_copy",
] + ["b200_" + n for n in (
    # host/src/loop_capi.inl: optimiser front-ends and driver-dependent entry points (bound by tests/loop_support.py)
    "srand sls_create sls_destroy sls_set_hyperparams sls_set_ucb_hyperparam sls_submit sls_get_slider_ends sls_get_maximizer sls_calc_point "
    "sls_num_points sls_get_raw_data_points sls_query pbo_create pbo_destroy pbo_set_hyperparams pbo_submit pbo_determine_next_query "
    "pbo_get_current_options pbo_get_maximizer pbo_num_points pref_fit pref_fit_destroy pref_fit_regressor pref_fit_get_state "
    "pref_fit_find_arg_max gpr_fit gpr_given gpr_fit_destroy gpr_fit_regressor loop_find_next_point loop_find_next_points loop_acq_value "
    "loop_acq_derivative loop_predict loop_slider").split()]

NATIVE, HYBRID, REFERENCE = 0, 1, 2  # sequential_line_search::SearchDriver


def nlopt_available() -> bool:
    return bool(load_host_library().b200_nlopt_available())


def set_search_driver(mode: int) -> None:
    if load_host_library().b200_set_search_driver(mode) != 0:
        raise RuntimeError(load_host_library().b200_last_error().decode())


def set_devices(ids) -> None:
    """The GPUs new regressors are built on (sequential_line_search::SetDevices); more than one = a multi-GPU group."""
    arr = (C.c_int * len(ids))(*ids)
    if load_host_library().b200_set_devices(arr, len(ids)) != 0:
        raise RuntimeError(load_host_library().b200_last_error().decode())


def set_incremental_refit(on: bool) -> None:
    """sequential_line_search::SetIncrementalRefit (driver.hpp): extend the previous iteration's factored model instead of rebuilding."""
    load_host_library().b200_set_incremental_refit(1 if on else 0)


def get_incremental_refit() -> bool:
    return bool(load_host_library().b200_get_incremental_refit())


def get_search_driver() -> int:
    return load_host_library().b200_get_search_driver()


def load_host_library(build_if_missing: bool = True) -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(HOST_LIB_PATH):
        if not build_if_missing:
            raise RuntimeError(f"{HOST_LIB_PATH} is missing; run `python __graft_entry__.py build`")
        build_host()
    lib = C.CDLL(HOST_LIB_PATH)
    for name in HOST_SYMBOLS:
        getattr(lib, name)  # AttributeError == facade / library mismatch
    for name in _DOUBLE:
        getattr(lib, name).restype = C.c_double
    for name in _POINTER:
        getattr(lib, name).restype = C.c_void_p
    lib.b200_last_error.restype = C.c_char_p
    lib.b200_pref_num_map_evaluations.restype = C.c_uint
    _lib = lib
    return lib


def _f64(a):
    return np.require(np.asarray(a, dtype=np.float64), requirements=["F", "A"])


def _p(a):
    return None if a is None else a.ctypes.data_as(c_dp)


class HostError(RuntimeError):
    pass


class Host:
    """Thin wrapper: numpy in / numpy out, raises HostError with the C++ exception text."""

    def __init__(self):
        self.lib = load_host_library()

    def _ok(self, cond):
        if not cond:
            raise HostError(self.lib.b200_last_error().decode())

    def kernel(self, kt, xa, xb, theta):
        xa, xb, theta = _f64(xa), _f64(xb), _f64(theta)
        D = len(xa)
        k, dth, dx = C.c_double(), np.empty(D + 1), np.empty(D)
        self.lib.b200_kernel(kt, D, _p(xa), _p(xb), _p(theta), C.byref(k), _p(dth), _p(dx))
        return k.value, dth, dx

    def large_ky(self, kt, X, theta, b):
        X, theta = _f64(X), _f64(theta)
        D, N = X.shape
        K = np.empty((N, N), order="F")
        self._ok(self.lib.b200_calc_large_ky(kt, D, N, _p(X), _p(theta), C.c_double(b), _p(K)) == 0)
        return K

    def small_k(self, kt, X, theta, x):
        """CalcSmallK and CalcSmallKSmallXDerivative (the reference's L1 free functions)."""
        X, theta, x = _f64(X), _f64(theta), _f64(x)
        D, N = X.shape
        k, J = np.empty(N), np.empty((D, N), order="F")
        self._ok(self.lib.b200_calc_small_k(kt, D, N, _p(X), _p(theta), _p(x), _p(k), _p(J)) == 0)
        return k, J

    def large_ky_theta_derivative(self, kt, X, theta):
        """CalcLargeKYThetaDerivative: [D + 1, N, N] (also checks CalcLargeKYNoiseLevelDerivative == I)."""
        X, theta = _f64(X), _f64(theta)
        D, N = X.shape
        out = np.empty((D + 1, N * N))
        self._ok(self.lib.b200_calc_large_ky_theta_derivative(kt, D, N, _p(X), _p(theta), _p(out)) == 0)
        return out.reshape(D + 1, N, N).transpose(0, 2, 1)

    def time_acq_values(self, reg, D, M, acq_type=0, beta=1.0, reps=3):
        """Seconds per call of acquisition_func::CalcAcquisitionValues on M candidates, timed inside the C++ layer."""
        t = self.lib.b200_time_acq_values(reg, D, M, acq_type, C.c_double(beta), reps)
        self._ok(t == t)
        return t

    def regressor_set_sweep_mode(self, reg, mode):
        self._ok(self.lib.b200_regressor_set_sweep_mode(reg, int(mode)) == 0)

    def gpr_copy(self, h):
        c = self.lib.b200_gpr_copy(h)
        self._ok(c)
        return C.c_void_p(c)

    def gpr_num_points(self, h):
        return self.lib.b200_gpr_num_points(h)

    # GaussianProcessRegressor
    def gpr_create(self, kt, X, y, theta=None, b=None):
        X, y = _f64(X), _f64(y)
        D, N = X.shape
        if theta is None:
            h = self.lib.b200_gpr_create_map(kt, D, N, _p(X), _p(y))
        else:
            h = self.lib.b200_gpr_create(kt, D, N, _p(X), _p(y), _p(_f64(theta)), C.c_double(b))
        self._ok(h)
        return C.c_void_p(h)

    def gpr_append_point(self, h, x, y):
        x = _f64(x)
        self._ok(self.lib.b200_gpr_append_point(h, x.size, _p(x), C.c_double(y)) == 0)

    def gpr_destroy(self, h):
        self.lib.b200_gpr_destroy(h)

    def gpr_regressor(self, h):
        return C.c_void_p(self.lib.b200_gpr_regressor(h))

    def gpr_state(self, h, N, D):
        K, Kinv = np.empty((N, N), order="F"), np.empty((N, N), order="F")
        theta, b = np.empty(D + 1), C.c_double()
        self.lib.b200_gpr_get_state(h, _p(K), _p(Kinv), _p(theta), C.byref(b))
        return dict(K=K, Kinv=Kinv, theta=theta, b=b.value)

    # PreferenceRegressor
    def pref_create(self, kt, X, offsets, idx, use_map, a, r, b, prior_var, btl_scale, num_iters=100, warm_from=None):
        X = _f64(X)
        D, N = X.shape
        offsets, idx = np.ascontiguousarray(offsets, dtype=np.uint32), np.ascontiguousarray(idx, dtype=np.uint32)
        if warm_from is not None:
            h = self.lib.b200_pref_create_warm(warm_from, kt, D, N, _p(X), len(offsets) - 1, offsets.ctypes.data_as(c_up),
                                               idx.ctypes.data_as(c_up), int(use_map), C.c_double(a), C.c_double(r), C.c_double(b),
                                               C.c_double(prior_var), C.c_double(btl_scale), C.c_uint(num_iters))
            self._ok(h)
            return C.c_void_p(h)
        h = self.lib.b200_pref_create(kt, D, N, _p(X), len(offsets) - 1, offsets.ctypes.data_as(c_up), idx.ctypes.data_as(c_up),
                                      int(use_map), C.c_double(a), C.c_double(r), C.c_double(b), C.c_double(prior_var),
                                      C.c_double(btl_scale), C.c_uint(num_iters))
        self._ok(h)
        return C.c_void_p(h)

    def pref_destroy(self, h):
        self.lib.b200_pref_destroy(h)

    def pref_regressor(self, h):
        return C.c_void_p(self.lib.b200_pref_regressor(h))

    def pref_objective(self, h, x, want_grad=True):
        x = _f64(x)
        g = np.empty(len(x)) if want_grad else None
        f = self.lib.b200_pref_objective(h, _p(x), len(x), _p(g))
        self._ok(f == f)
        return f, g

    def pref_num_map_evaluations(self, h):
        return int(self.lib.b200_pref_num_map_evaluations(h))

    def pref_state(self, h, N, D):
        y, theta, b = np.empty(N), np.empty(D + 1), C.c_double()
        K, L = np.empty((N, N), order="F"), np.empty((N, N), order="F")
        self.lib.b200_pref_get_state(h, _p(y), _p(theta), C.byref(b), _p(K), _p(L))
        return dict(y=y, theta=theta, b=b.value, K=K, L=L)

    def pref_find_arg_max(self, h, D):
        out = np.empty(D)
        self.lib.b200_pref_find_arg_max(h, _p(out))
        return out

    def pref_damp_data(self, h, directory, prefix=""):
        self._ok(self.lib.b200_pref_damp_data(h, directory.encode(), prefix.encode()) == 0)

    # host bookkeeping
    def data_manager_run(self, batches, eps=1e-4):
        """batches: list of (D x k) arrays, first column preferred. Returns (X, offsets, idx) after AddNewPoints of each."""
        D = batches[0].shape[0]
        sizes = np.asarray([b.shape[1] for b in batches], dtype=np.int32)
        pts = _f64(np.concatenate([np.asarray(b, dtype=np.float64).T.reshape(-1) for b in batches]))
        total = int(sizes.sum())
        X = np.empty(D * total)
        offsets, idx = np.zeros(len(batches) + 1, dtype=np.uint32), np.zeros(total, dtype=np.uint32)
        n = self.lib.b200_data_manager_run(D, len(batches), sizes.ctypes.data_as(C.POINTER(C.c_int)), _p(pts), C.c_double(eps), _p(X),
                                           offsets.ctypes.data_as(c_up), idx.ctypes.data_as(c_up))
        return X[:D * n].reshape((D, n), order="F"), offsets, idx

    def slider(self, end_0, end_1, enlarge=True, scale=1.25, minimum_length=0.25):
        end_0, end_1 = _f64(end_0), _f64(end_1)
        a, b = np.empty(len(end_0)), np.empty(len(end_0))
        self.lib.b200_slider(len(end_0), _p(end_0), _p(end_1), int(enlarge), C.c_double(scale), C.c_double(minimum_length), _p(a), _p(b))
        return a, b

    # utils
    def btl(self, f, scale=1.0):
        f = _f64(f)
        d = np.empty(len(f))
        return self.lib.b200_utils_btl(len(f), _p(f), C.c_double(scale), _p(d)), d

    def random_vector(self, n):
        out = np.empty(n)
        self.lib.b200_utils_random_vector(C.c_uint(n), _p(out))
        return out

    def export_csv(self, path, X):
        X = _f64(X)
        self._ok(self.lib.b200_utils_export_csv(str(path).encode(), X.shape[0], X.shape[1], _p(X)) == 0)

    def test_minimize(self, problem, x0, max_evals=2000):
        x0 = _f64(x0)
        x, f = np.empty(len(x0)), C.c_double()
        evals = self.lib.b200_test_minimize(problem, len(x0), _p(x0), C.c_uint(max_evals), _p(x), C.byref(f))
        return x, f.value, int(evals)

    # Regressor virtuals + acquisition
    def predict(self, reg, x):
        x = _f64(x)
        D = len(x)
        dmu, dsg = np.empty(D), np.empty(D)
        mu = self.lib.b200_predict_mu(reg, D, _p(x))
        sg = self.lib.b200_predict_sigma(reg, D, _p(x))
        self._ok(mu == mu)
        self._ok(self.lib.b200_predict_mu_derivative(reg, D, _p(x), _p(dmu)) == 0)
        self._ok(self.lib.b200_predict_sigma_derivative(reg, D, _p(x), _p(dsg)) == 0)
        return mu, sg, dmu, dsg

    def predict_batch(self, reg, Xq):
        Xq = _f64(Xq)
        D, M = Xq.shape
        mu, sg = np.empty(M), np.empty(M)
        dmu, dsg = np.empty((D, M), order="F"), np.empty((D, M), order="F")
        self._ok(self.lib.b200_predict_batch(reg, D, M, _p(Xq), _p(mu), _p(sg), _p(dmu), _p(dsg)) == 0)
        return mu, sg, dmu, dsg

    def x_best(self, reg, D):
        out = np.empty(D)
        self._ok(self.lib.b200_predict_maximum_point_from_data(reg, D, _p(out)) == 0)
        return out

    def acq(self, reg, acq_type, beta, x, want_grad=True):
        x = _f64(x)
        D = len(x)
        v = self.lib.b200_acq_value(reg, D, acq_type, C.c_double(beta), _p(x))
        self._ok(v == v)
        g = None
        if want_grad:
            g = np.empty(D)
            self._ok(self.lib.b200_acq_derivative(reg, D, acq_type, C.c_double(beta), _p(x), _p(g)) == 0)
        return v, g

    def acq_values(self, reg, acq_type, beta, Xq, want_grad=True):
        Xq = _f64(Xq)
        D, M = Xq.shape
        val = np.empty(M)
        grad = np.empty((D, M), order="F") if want_grad else None
        self._ok(self.lib.b200_acq_values(reg, D, M, acq_type, C.c_double(beta), _p(Xq), _p(val), _p(grad)) == 0)
        return val, grad

    def find_next_point(self, reg, D, n_global=100, n_local=50, acq_type=0, beta=1.0):
        out = np.empty(D)
        self._ok(self.lib.b200_find_next_point(reg, D, n_global, n_local, acq_type, C.c_double(beta), _p(out)) == 0)
        return out

    def find_next_points(self, reg, D, n_points, n_global=100, n_local=50, acq_type=0, beta=1.0):
        out = np.empty((D, n_points), order="F")
        self._ok(self.lib.b200_find_next_points(reg, D, n_points, n_global, n_local, acq_type, C.c_double(beta), _p(out)) == 0)
        return out
