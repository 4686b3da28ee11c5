"""Gram kernel (K1) timing: median CUDA-event time of slsgp_gram over N x D, as GB/s of the algorithmic bytes
(N D + N^2 doubles) against the measured HBM peak. usage: python tools/gram_bench.py [--sizes 2048,4096,8192] [--dims 16]"""
import argparse
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tools")]
import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="2048,4096,8192")
    ap.add_argument("--dims", default="16")
    ap.add_argument("--kernel", type=int, default=0)
    args = ap.parse_args()
    pkg = importlib.import_module("sequential-line-search_b200")
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    ctx = pkg.Context(0)
    print(f"layout: {'first (SLSGP_GRAM_V1=1)' if os.environ.get('SLSGP_GRAM_V1') == '1' else 'second'}; kernel type {args.kernel}; HBM peak {peak} GB/s")
    for D in map(int, args.dims.split(",")):
        for N in map(int, args.sizes.split(",")):
            X, theta = synth.make_X(N, D, "uniform"), synth.make_theta(D, "default")
            ctx.set_data(X)
            t = []
            for it in range(25):
                ctx.invalidate()
                ctx.gram(args.kernel, theta, 0.005, want=False)
                if it >= 5:
                    t.append(ctx.phase_ms("gram"))
            ms = float(np.median(t))
            gb = 8.0 * (N * D + N * N) / 1e9
            print(f"N={N:5d} D={D:3d}: {ms * 1e3:8.1f} us  {gb / (ms * 1e-3):8.1f} GB/s  {gb / (ms * 1e-3) / peak * 100:5.1f} % of HBM peak  (min {min(t) * 1e3:.1f} us)")
    ctx.close()


if __name__ == "__main__":
    main()
