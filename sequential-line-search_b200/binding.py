"""ctypes binding of include/slsgp.h — the reference-side stub a Python host would use.

Thin by design: numpy arrays in, numpy arrays out, every call goes straight to the C ABI. There is no CPU path:
if libslsgp.so is missing it is built with nvcc; if no CUDA device is usable, creating a Context raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .build import LIB_PATH as _LIB_PATH, build as _build_library

c_dp = C.POINTER(C.c_double)
c_u32p = C.POINTER(C.c_uint32)

OK, ERR_INVALID, ERR_STATE, ERR_NOT_SPD, ERR_NAN, ERR_CUDA, ERR_NOMEM = range(7)
KERNEL_ARD_SQUARED_EXP, KERNEL_ARD_MATERN52 = 0, 1
ACQ_EXPECTED_IMPROVEMENT, ACQ_GP_UCB = 0, 1
SWEEP_FP64, SWEEP_TENSOR, SWEEP_TENSOR_X2, SWEEP_TENSOR_X1 = 0, 1, 2, 3
COMPAT_SE_XGRAD_2X, COMPAT_NOISELESS = 1, 2

# every symbol include/slsgp.h declares: (name, restype, argtypes)
_API = [
    ("slsgp_ctx_create", C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    ("slsgp_ctx_create_multi", C.c_int, [C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_void_p)]),
    ("slsgp_ctx_device_count", C.c_int, [C.c_void_p]),
    ("slsgp_ctx_destroy", C.c_int, [C.c_void_p]),
    ("slsgp_last_error", C.c_char_p, [C.c_void_p]),
    ("slsgp_status_string", C.c_char_p, [C.c_int]),
    ("slsgp_set_compat_flags", C.c_int, [C.c_void_p, C.c_uint]),
    ("slsgp_set_sweep_mode", C.c_int, [C.c_void_p, C.c_int]),
    ("slsgp_get_sweep_mode", C.c_int, [C.c_void_p, C.POINTER(C.c_int)]),
    ("slsgp_set_refine_threshold", C.c_int, [C.c_void_p, C.c_double]),
    ("slsgp_set_stream", C.c_int, [C.c_void_p, C.c_void_p]),
    ("slsgp_synchronize", C.c_int, [C.c_void_p]),
    ("slsgp_set_data", C.c_int, [C.c_void_p, c_dp, C.c_int, C.c_int]),
    ("slsgp_set_data_extend", C.c_int, [C.c_void_p, c_dp, C.c_int, C.c_int, C.POINTER(C.c_int)]),
    ("slsgp_invalidate", C.c_int, [C.c_void_p]),
    ("slsgp_gram", C.c_int, [C.c_void_p, C.c_int, c_dp, C.c_double, c_dp]),
    ("slsgp_factor", C.c_int, [C.c_void_p, c_dp, c_dp]),
    ("slsgp_inverse", C.c_int, [C.c_void_p, c_dp]),
    ("slsgp_solve_alpha", C.c_int, [C.c_void_p, c_dp, c_dp]),
    ("slsgp_append_point", C.c_int, [C.c_void_p, c_dp, C.c_double, c_dp, c_dp]),
    ("slsgp_get_f_best", C.c_int, [C.c_void_p, c_dp, C.POINTER(C.c_int)]),
    ("slsgp_posterior_batch", C.c_int, [C.c_void_p, c_dp, C.c_int64, c_dp, c_dp, c_dp, c_dp]),
    ("slsgp_acq_batch", C.c_int, [C.c_void_p, C.c_int, C.c_double, c_dp, C.c_int64, c_dp, c_dp]),
    ("slsgp_acq_from_posterior", C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int64, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp]),
    ("slsgp_acq_batch_device", C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_int64] + [C.c_void_p] * 6),
    ("slsgp_acq_argmax", C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_uint64, C.c_int64, C.c_int64, c_dp, c_dp,
                                   C.POINTER(C.c_int64), c_dp]),
    ("slsgp_acq_maximize", C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_uint64, C.c_int64, C.c_int64, C.c_int, C.c_int, c_dp, c_dp,
                                     c_dp, c_dp]),
    ("slsgp_pair_acq_argmax", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_uint64, C.c_int64, C.c_int64, c_dp, c_dp, C.POINTER(C.c_int64)]),
    ("slsgp_argmax_device", C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, c_dp, C.POINTER(C.c_int64)]),
    ("slsgp_candidates", C.c_int, [C.c_void_p, C.c_uint64, C.c_int64, C.c_int64, c_dp]),
    ("slsgp_set_preferences", C.c_int, [C.c_void_p, c_u32p, c_u32p, C.c_int]),
    ("slsgp_map_objective_pref", C.c_int, [C.c_void_p, C.c_int, c_dp, C.c_int, C.c_int, C.c_double, C.c_double,
                                           C.c_double, C.c_double, C.c_double, c_dp, c_dp]),
    ("slsgp_map_objective_pref_whitened", C.c_int, [C.c_void_p, c_dp, C.c_double, c_dp, c_dp, c_dp]),
    ("slsgp_whiten", C.c_int, [C.c_void_p, c_dp, c_dp]),
    ("slsgp_map_objective_gpr", C.c_int, [C.c_void_p, C.c_int, c_dp, c_dp, c_dp, c_dp]),
    ("slsgp_trim", C.c_int, [C.c_void_p, C.c_size_t]),
    ("slsgp_small_k", C.c_int, [C.c_void_p, C.c_int, c_dp, c_dp, c_dp, c_dp]),
    ("slsgp_gram_theta_derivative", C.c_int, [C.c_void_p, C.c_int, c_dp, c_dp]),
    ("slsgp_launch_count", C.c_uint64, [C.c_void_p]),
    ("slsgp_last_phase_ms", C.c_double, [C.c_void_p, C.c_char_p]),
    ("slsgp_profile_enable", C.c_int, [C.c_void_p, C.c_int]),
    ("slsgp_profile_read", C.c_int, [C.c_void_p, C.c_char_p, c_dp, C.POINTER(C.c_uint64)]),
]
API_SYMBOLS = [name for name, _, _ in _API]

_lib = None


def load_library(build_if_missing: bool = True) -> C.CDLL:
    """dlopen libslsgp.so (building it first if needed) and declare every prototype."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("SLSGP_LIB", _LIB_PATH)  # override: A/B runs of two builds on the GPU box
    if not os.path.exists(path):
        if not build_if_missing:
            raise RuntimeError(f"{path} is missing; run `python __graft_entry__.py build`")
        _build_library()
    lib = C.CDLL(path)
    for name, restype, argtypes in _API:
        fn = getattr(lib, name)  # AttributeError here == header / library mismatch
        fn.restype, fn.argtypes = restype, argtypes
    _lib = lib
    return lib


class SlsgpError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"libslsgp status {status}: {message}")
        self.status = status


def _f64(a, order="F"):
    return np.require(np.asarray(a, dtype=np.float64), requirements=[order, "A"])


def _p(a):
    return None if a is None else a.ctypes.data_as(c_dp)


class Context:
    """One libslsgp device context. Mirrors the C ABI one to one; see include/slsgp.h for semantics."""

    def __init__(self, device=0):
        """device: one device index, or a list of indices for a multi-GPU group (slsgp_ctx_create_multi; the first is the primary)."""
        self.lib = load_library()
        h = C.c_void_p()
        if isinstance(device, (list, tuple)):
            ids = (C.c_int * len(device))(*device)
            st = self.lib.slsgp_ctx_create_multi(ids, len(device), C.byref(h))
        else:
            st = self.lib.slsgp_ctx_create(device, C.byref(h))
        if st != OK:
            raise SlsgpError(st, "slsgp_ctx_create failed: no usable CUDA device (libslsgp has no CPU fallback)")
        self.h = h
        self.N = self.D = 0

    def device_count(self):
        return int(self.lib.slsgp_ctx_device_count(self.h))

    def close(self):
        if getattr(self, "h", None):
            self.lib.slsgp_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, st):
        if st != OK:
            raise SlsgpError(st, self.lib.slsgp_last_error(self.h).decode())

    # ---- configuration
    def set_compat_flags(self, flags):
        self._check(self.lib.slsgp_set_compat_flags(self.h, flags))

    def set_sweep_mode(self, mode):
        self._check(self.lib.slsgp_set_sweep_mode(self.h, mode))

    def set_stream(self, cuda_stream):
        self._check(self.lib.slsgp_set_stream(self.h, C.c_void_p(cuda_stream)))

    def synchronize(self):
        self._check(self.lib.slsgp_synchronize(self.h))

    # ---- model
    def set_data(self, X):
        X = _f64(X)
        self.D, self.N = X.shape
        self._check(self.lib.slsgp_set_data(self.h, _p(X), self.N, self.D))

    def set_data_extend(self, X):
        """slsgp_set_data_extend: returns the number of leading data points whose model was kept (0 = replaced)."""
        X = _f64(X)
        self.D, self.N = X.shape
        kept = C.c_int(0)
        self._check(self.lib.slsgp_set_data_extend(self.h, _p(X), self.N, self.D, C.byref(kept)))
        return kept.value

    def invalidate(self):
        """slsgp_invalidate: the next gram / factor / inverse recompute even for unchanged hyper-parameters (timing loops)."""
        self._check(self.lib.slsgp_invalidate(self.h))

    def gram(self, kernel_type, theta, noise, want=True):
        theta = _f64(theta)
        K = np.empty((self.N, self.N), order="F") if want else None
        self._check(self.lib.slsgp_gram(self.h, kernel_type, _p(theta), noise, _p(K)))
        return K

    def factor(self, want_L=False):
        logdet = C.c_double()
        L = np.empty((self.N, self.N), order="F") if want_L else None
        self._check(self.lib.slsgp_factor(self.h, C.byref(logdet), _p(L)))
        return (logdet.value, L) if want_L else logdet.value

    def inverse(self, want=True):
        Kinv = np.empty((self.N, self.N), order="F") if want else None
        self._check(self.lib.slsgp_inverse(self.h, _p(Kinv)))
        return Kinv

    def solve_alpha(self, y):
        y = _f64(y)
        alpha = np.empty(self.N)
        self._check(self.lib.slsgp_solve_alpha(self.h, _p(y), _p(alpha)))
        return alpha

    def append_point(self, x, y_new, want=True):
        """slsgp_append_point: bordered O(N^2) update by one data point; returns (new K column, new Kinv) when asked."""
        x = _f64(x)
        if x.size != self.D:
            raise ValueError("x has the wrong length")
        kcol = np.empty(self.N + 1) if want else None
        Kinv = np.empty((self.N + 1, self.N + 1), order="F") if want else None
        self._check(self.lib.slsgp_append_point(self.h, _p(x), float(y_new), _p(kcol) if want else None, _p(Kinv) if want else None))
        self.N += 1
        return (kcol, Kinv) if want else None

    def f_best(self):
        f, i = C.c_double(), C.c_int()
        self._check(self.lib.slsgp_get_f_best(self.h, C.byref(f), C.byref(i)))
        return f.value, i.value

    def fit(self, X, kernel_type, theta, noise, y):
        """set_data + gram + factor + solve_alpha: what a regressor constructor does after MAP."""
        self.set_data(X)
        self.gram(kernel_type, theta, noise, want=False)
        self.factor()
        return self.solve_alpha(y)

    # ---- sweep
    def posterior_batch(self, Xq, grads=True):
        Xq = _f64(Xq)
        D, M = Xq.shape
        mu, sigma = np.empty(M), np.empty(M)
        dmu = np.empty((D, M), order="F") if grads else None
        dsg = np.empty((D, M), order="F") if grads else None
        self._check(self.lib.slsgp_posterior_batch(self.h, _p(Xq), M, _p(mu), _p(sigma), _p(dmu), _p(dsg)))
        return mu, sigma, dmu, dsg

    def acq_batch(self, acq_type, ucb_beta, Xq, grads=True):
        Xq = _f64(Xq)
        D, M = Xq.shape
        val = np.empty(M)
        grad = np.empty((D, M), order="F") if grads else None
        self._check(self.lib.slsgp_acq_batch(self.h, acq_type, ucb_beta, _p(Xq), M, _p(val), _p(grad)))
        return val, grad

    def acq_from_posterior(self, acq_type, ucb_beta, f_best, mu, sigma, dmu=None, dsigma=None):
        """Acquisition value (and gradient when dmu / dsigma are given) from an externally supplied posterior."""
        mu, sigma = _f64(mu), _f64(sigma)
        M = mu.shape[0]
        grads = dmu is not None
        D = _f64(dmu).shape[0] if grads else 1
        val = np.empty(M)
        grad = np.empty((D, M), order="F") if grads else None
        self._check(self.lib.slsgp_acq_from_posterior(self.h, acq_type, ucb_beta, f_best, D, M, _p(mu), _p(sigma),
                                                      _p(_f64(dmu)) if grads else None, _p(_f64(dsigma)) if grads else None,
                                                      _p(val), _p(grad) if grads else None))
        return val, grad

    def acq_batch_device(self, acq_type, ucb_beta, d_Xq, M, d_mu=0, d_sigma=0, d_dmu=0, d_dsigma=0, d_val=0, d_grad=0):
        """All arguments are raw device addresses (ints), e.g. torch.Tensor.data_ptr(). Asynchronous."""
        v = lambda a: C.c_void_p(a) if a else None
        self._check(self.lib.slsgp_acq_batch_device(self.h, acq_type, ucb_beta, v(d_Xq), M, v(d_mu), v(d_sigma),
                                                    v(d_dmu), v(d_dsigma), v(d_val), v(d_grad)))

    def acq_argmax(self, acq_type, ucb_beta, seed, first, count, want_grad=False):
        x = np.empty(self.D)
        val, idx = C.c_double(), C.c_int64()
        grad = np.empty(self.D) if want_grad else None
        self._check(self.lib.slsgp_acq_argmax(self.h, acq_type, ucb_beta, seed, first, count, _p(x), C.byref(val),
                                              C.byref(idx), _p(grad)))
        return x, val.value, idx.value, grad

    def acq_maximize(self, acq_type, ucb_beta, seed, first, count, n_starts=1024, n_iters=40):
        """Sweep + batched multi-start ascent on the device. Returns (x_best, value, gradient at x_best, best sweep value)."""
        x, g = np.empty(self.D), np.empty(self.D)
        v, vs = C.c_double(), C.c_double()
        self._check(self.lib.slsgp_acq_maximize(self.h, acq_type, ucb_beta, seed, first, count, n_starts, n_iters, _p(x), C.byref(v),
                                                _p(g), C.byref(vs)))
        return x, v.value, g, vs.value

    def small_k(self, kernel_type, theta, x, want_derivative=True):
        """CalcSmallK / CalcSmallKSmallXDerivative on the context's X: (k [N], dk/dx [D x N] or None)."""
        theta, x = _f64(theta), _f64(x)
        k = np.empty(self.N)
        dk = np.empty((self.D, self.N), order="F") if want_derivative else None
        self._check(self.lib.slsgp_small_k(self.h, kernel_type, _p(theta), _p(x), _p(k), _p(dk)))
        return k, dk

    def gram_theta_derivative(self, kernel_type, theta):
        """CalcLargeKYThetaDerivative: array [D + 1, N, N]."""
        theta = _f64(theta)
        out = np.empty((self.D + 1, self.N * self.N))
        self._check(self.lib.slsgp_gram_theta_derivative(self.h, kernel_type, _p(theta), _p(out)))
        return out.reshape(self.D + 1, self.N, self.N).transpose(0, 2, 1)

    def set_refine_threshold(self, tau):
        self._check(self.lib.slsgp_set_refine_threshold(self.h, float(tau)))

    def trim(self, keep_bytes=0):
        self._check(self.lib.slsgp_trim(self.h, keep_bytes))

    def pair_acq_argmax(self, sigma_ctx, acq_type, ucb_beta, seed, first, count):
        """mu from this context, sigma from `sigma_ctx` (same device): (x, value, index) of the best candidate."""
        x = np.empty(self.D)
        val, idx = C.c_double(), C.c_int64()
        self._check(self.lib.slsgp_pair_acq_argmax(self.h, sigma_ctx.h, acq_type, ucb_beta, seed, first, count, _p(x), C.byref(val), C.byref(idx)))
        return x, val.value, idx.value

    def argmax_device(self, d_val, count, index0=0):
        val, idx = C.c_double(), C.c_int64()
        self._check(self.lib.slsgp_argmax_device(self.h, C.c_void_p(d_val), count, index0, C.byref(val), C.byref(idx)))
        return val.value, idx.value

    def candidates(self, seed, first, count):
        Xq = np.empty((self.D, count), order="F")
        self._check(self.lib.slsgp_candidates(self.h, seed, first, count, _p(Xq)))
        return Xq

    # ---- MAP objectives
    def set_preferences(self, offsets, idx):
        offsets = np.ascontiguousarray(offsets, dtype=np.uint32)
        idx = np.ascontiguousarray(idx, dtype=np.uint32)
        self._check(self.lib.slsgp_set_preferences(self.h, offsets.ctypes.data_as(c_u32p), idx.ctypes.data_as(c_u32p),
                                                   len(offsets) - 1))

    def map_objective_pref(self, kernel_type, x, use_map, a, r, b, prior_var, btl_scale, want_grad=True):
        x = _f64(x)
        f = C.c_double()
        g = np.empty(len(x)) if want_grad else None
        self._check(self.lib.slsgp_map_objective_pref(self.h, kernel_type, _p(x), len(x), int(use_map), a, r, b,
                                                      prior_var, btl_scale, C.byref(f), _p(g)))
        return f.value, g

    def map_objective_pref_whitened(self, z, btl_scale, want_grad=True):
        """F(z), grad_z F and y = L z for the current factor (fixed hyper-parameters)."""
        z = _f64(z)
        f = C.c_double()
        g = np.empty(len(z)) if want_grad else None
        y = np.empty(len(z))
        self._check(self.lib.slsgp_map_objective_pref_whitened(self.h, _p(z), btl_scale, C.byref(f), _p(g), _p(y)))
        return f.value, g, y

    def whiten(self, y):
        y = _f64(y)
        z = np.empty(len(y))
        self._check(self.lib.slsgp_whiten(self.h, _p(y), _p(z)))
        return z

    def map_objective_gpr(self, kernel_type, y, x, want_grad=True):
        y, x = _f64(y), _f64(x)
        f = C.c_double()
        g = np.empty(len(x)) if want_grad else None
        self._check(self.lib.slsgp_map_objective_gpr(self.h, kernel_type, _p(y), _p(x), C.byref(f), _p(g)))
        return f.value, g

    # ---- introspection
    def launch_count(self):
        return int(self.lib.slsgp_launch_count(self.h))

    def profile_enable(self, on=True):
        self._check(self.lib.slsgp_profile_enable(self.h, int(on)))

    def profile_read(self, kernel):
        ms, n = C.c_double(), C.c_uint64()
        self._check(self.lib.slsgp_profile_read(self.h, kernel.encode(), C.byref(ms), C.byref(n)))
        return ms.value, int(n.value)

    def phase_ms(self, name):
        return float(self.lib.slsgp_last_phase_ms(self.h, name.encode()))
