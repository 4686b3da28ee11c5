"""Cholesky factor time of libslsgp under its A/B switches next to cuSOLVER potrf (torch.linalg.cholesky) on the same Gram matrix.

Settings (environment, read when a context is created): SLSGP_CHOL_PIVOTS (1 or 2 pivots per barrier in the diagonal tile),
SLSGP_CHOL_TWO_LEVEL_FROM (block columns from which the two-level form is used), SLSGP_CHOL_PANEL (block columns per panel),
SLSGP_CHOL_LOOKAHEAD (trailing update beyond the next panel on a second stream). Times: `factor` phase of the library (copy K -> L,
all steps, zeroing of the upper triangle, log-determinant), median of `reps` factorisations; not part of the product path."""
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


def factor_ms(n, env, reps=7):
    for k in ("SLSGP_CHOL_PIVOTS", "SLSGP_CHOL_TWO_LEVEL_FROM", "SLSGP_CHOL_PANEL", "SLSGP_CHOL_LOOKAHEAD", "SLSGP_CHOL_SWITCH_REM"):
        os.environ.pop(k, None)
    os.environ.update(env)
    ctx = pkg.Context(0)
    try:
        X, theta = synth.make_X(n, 16, "uniform"), synth.make_theta(16, "default")
        ctx.set_data(X)
        t, inv = [], []
        for _ in range(reps + 2):
            ctx.invalidate()
            ctx.gram(0, theta, 0.005, want=False)
            ctx.factor()
            t.append(ctx.lib.slsgp_last_phase_ms(ctx.h, b"factor"))
            ctx.lib.slsgp_inverse(ctx.h, None)
            inv.append(ctx.lib.slsgp_last_phase_ms(ctx.h, b"inverse"))
        return float(np.median(t[2:])), float(np.median(inv[2:]))
    finally:
        ctx.close()


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [512, 1024, 2048, 4096, 8192]
    dev = torch.device("cuda:0")
    for n in sizes:
        ctx = pkg.Context(0)
        X, theta = synth.make_X(n, 16, "uniform"), synth.make_theta(16, "default")
        ctx.set_data(X)
        K = torch.from_numpy(np.ascontiguousarray(ctx.gram(0, theta, 0.005))).to(dev)
        ctx.close()
        for _ in range(3):
            torch.linalg.cholesky(K)
        torch.cuda.synchronize()
        ts = []
        for _ in range(7):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            torch.linalg.cholesky(K)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        del K
        print(f"N={n}: cuSOLVER potrf {np.median(ts):7.3f} ms", flush=True)
        off = "100000"
        settings = [("single level, 1 pivot / barrier", {"SLSGP_CHOL_PIVOTS": "1", "SLSGP_CHOL_TWO_LEVEL_FROM": off}),
                    ("single level, 2 pivots / barrier", {"SLSGP_CHOL_PIVOTS": "2", "SLSGP_CHOL_TWO_LEVEL_FROM": off})]
        if n >= 1024:
            for pb in ((8, 16) if n >= 4096 else (4, 8)):
                settings.append((f"two-level to the end, panel {pb}", {"SLSGP_CHOL_TWO_LEVEL_FROM": "3", "SLSGP_CHOL_PANEL": str(pb), "SLSGP_CHOL_SWITCH_REM": "0"}))
            settings.append(("two-level to the end, panel 4, one stream", {"SLSGP_CHOL_TWO_LEVEL_FROM": "3", "SLSGP_CHOL_PANEL": "4", "SLSGP_CHOL_LOOKAHEAD": "0", "SLSGP_CHOL_SWITCH_REM": "0"}))
        if n >= 4096:
            for pb in (8,):
                for sw in (32, 48):
                    settings.append((f"two-level, panel {pb}, single level for the last {sw}", {"SLSGP_CHOL_TWO_LEVEL_FROM": "3", "SLSGP_CHOL_PANEL": str(pb), "SLSGP_CHOL_SWITCH_REM": str(sw)}))
        settings.append(("library default", {}))
        for name, env in settings:
            f, i = factor_ms(n, env)
            print(f"    {name:52s} factor {f:7.3f} ms   inverse {i:7.3f} ms", flush=True)


if __name__ == "__main__":
    main()
