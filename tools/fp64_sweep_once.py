"""One model build (N=2048, D=16) and ONE FP64 sweep shard of 16384 candidates: the smallest process that launches the sweep's
DMMA GEMM at its production shape (for `ncu -k regex:gemm64_dmma -s 11 -c 1`: the 11 earlier launches are trtri + K^-1)."""
import importlib
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
import synth  # noqa: E402

pkg = importlib.import_module("sequential-line-search_b200")
N, D, M = 2048, 16, 16384
X, theta = synth.make_X(N, D, "uniform"), synth.make_theta(D, "default")
ctx = pkg.Context(0)
ctx.fit(X, 0, theta, 0.005, synth.make_y(X))
val, grad = ctx.acq_batch(0, 1.0, synth.make_queries(M, D))
print("ok", float(val.max()))
ctx.close()
