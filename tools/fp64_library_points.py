"""Library reference points for the FP64 kernels on this box (not part of the product path): cuBLAS DGEMM throughput (the
measured FP64 tensor-pipe peak SURVEY.md 8(d) asks for; MEASURED_PEAKS.json has no FP64 figure) and cuSOLVER potrf /
potri-equivalent times through torch, next to libslsgp's own Gram + Cholesky + inverse at the same sizes."""
import importlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
import synth  # noqa: E402

pkg = importlib.import_module("sequential-line-search_b200")


def timed(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    dev = torch.device("cuda:0")
    for n in (2048, 4096, 8192):
        A = torch.randn(n, n, dtype=torch.float64, device=dev)
        B = torch.randn(n, n, dtype=torch.float64, device=dev)
        C = torch.empty_like(A)
        ms = timed(lambda: torch.matmul(A, B, out=C), reps=5 if n == 8192 else 20)
        print(f"cuBLAS DGEMM n={n}: {ms:8.3f} ms  {2.0 * n ** 3 / ms / 1e9:7.2f} TFLOP/s")
    ctx = pkg.Context(0)
    for n in (512, 2048, 4096, 8192):
        X, theta = synth.make_X(n, 16, "uniform"), synth.make_theta(16, "default")
        ctx.set_data(X)
        K = torch.from_numpy(np.ascontiguousarray(ctx.gram(0, theta, 0.005))).to(dev)
        ms_potrf = timed(lambda: torch.linalg.cholesky(K), reps=10)
        L = torch.linalg.cholesky(K)
        ms_potri = timed(lambda: torch.cholesky_inverse(L), reps=5)
        ours = {}
        for _ in range(3):
            ctx.invalidate()
            ctx.gram(0, theta, 0.005, want=False)
            ctx.factor()
            ctx.lib.slsgp_inverse(ctx.h, None)
            for ph in ("gram", "factor", "inverse"):
                ours[ph] = ctx.lib.slsgp_last_phase_ms(ctx.h, ph.encode())
        print(f"N={n}: cuSOLVER potrf (torch.linalg.cholesky) {ms_potrf:7.3f} ms, cholesky_inverse {ms_potri:7.3f} ms | "
              f"libslsgp gram {ours['gram']:.3f} ms, factor {ours['factor']:.3f} ms, inverse {ours['inverse']:.3f} ms")
    ctx.close()


if __name__ == "__main__":
    main()
