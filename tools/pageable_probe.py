"""Where the pageable end-to-end path spends its time (GPU box): memcpy bandwidth of the host, slsgp_acq_batch with pageable and
with pinned numpy buffers, and the C++ facade call. usage: python tools/pageable_probe.py"""
import ctypes as C
import importlib
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tools")]
import synth  # noqa: E402

pkg = importlib.import_module("sequential-line-search_b200")
N, D, M = 2048, 16, 1 << 20
X, theta = synth.make_X(N, D, "uniform"), synth.make_theta(D, "default")
y = synth.make_y(X)

a = np.random.rand(M * D)
b = np.empty_like(a)
b[:] = 0
t0 = time.perf_counter()
for _ in range(5):
    b[:] = a
dt = (time.perf_counter() - t0) / 5
print(f"numpy copy of {a.nbytes / 1e6:.0f} MB (touched): {a.nbytes / dt / 1e9:.1f} GB/s")
t0 = time.perf_counter()
c = np.empty_like(a)
c[:] = a
dt = time.perf_counter() - t0
print(f"numpy copy into fresh pages: {a.nbytes / dt / 1e9:.1f} GB/s")

ctx = pkg.Context(0)
ctx.fit(X, 0, theta, 0.005, y)
ctx.set_sweep_mode(pkg.SWEEP_TENSOR)
Q = synth.f64(np.random.rand(D, M))
val, grad = ctx.acq_batch(0, 1.0, Q)  # warm-up (allocations, pinned ring)
for label in ("pageable numpy in / fresh numpy out", "again"):
    t0 = time.perf_counter()
    val, grad = ctx.acq_batch(0, 1.0, Q)
    dt = time.perf_counter() - t0
    print(f"slsgp_acq_batch, {label}: {M / dt:.3e} evals/s ({dt * 1e3:.1f} ms)")
# pre-touched outputs, raw call
lib = ctx.lib
vo, go = np.zeros(M), np.zeros((D, M), order="F")
dp = C.POINTER(C.c_double)
for _ in range(3):
    t0 = time.perf_counter()
    lib.slsgp_acq_batch(ctx.h, 0, C.c_double(1.0), Q.ctypes.data_as(dp), C.c_int64(M), vo.ctypes.data_as(dp), go.ctypes.data_as(dp))
    dt = time.perf_counter() - t0
    print(f"slsgp_acq_batch, pageable in, pre-touched pageable out: {M / dt:.3e} evals/s ({dt * 1e3:.1f} ms)")
tq = torch.from_numpy(np.ascontiguousarray(Q.T)).pin_memory()
tv, tg = torch.empty(M, dtype=torch.float64).pin_memory(), torch.empty((M, D), dtype=torch.float64).pin_memory()
for _ in range(3):
    t0 = time.perf_counter()
    lib.slsgp_acq_batch(ctx.h, 0, C.c_double(1.0), C.cast(tq.data_ptr(), dp), C.c_int64(M), C.cast(tv.data_ptr(), dp), C.cast(tg.data_ptr(), dp))
    dt = time.perf_counter() - t0
    print(f"slsgp_acq_batch, pinned in / out: {M / dt:.3e} evals/s ({dt * 1e3:.1f} ms)")
ctx.close()
host = pkg.hostlib.Host()
hreg = host.gpr_create(0, X, y, theta, 0.005)
reg = host.gpr_regressor(hreg)
host.regressor_set_sweep_mode(reg, pkg.SWEEP_TENSOR)
sec = host.time_acq_values(reg, D, M, 0, 1.0, 4)
print(f"C++ CalcAcquisitionValues (Eigen in / out): {M / sec:.3e} evals/s ({sec * 1e3:.1f} ms)")
host.gpr_destroy(hreg)
