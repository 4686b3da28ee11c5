"""Oracle parity at the sizes the benchmark runs ("Weak" #2 / "Next" #4): the FP64 and the tensor sweep against the
plain-C oracle at config 4 (N = 2048, D = 16) and config 2 (N = 512, D = 8), a few candidates against the REFERENCE itself
(oracle/_ref: the unmodified CalcAcquisitionValue{,Derivative}, ~10 s of CPU per candidate at N = 2048), the MAP objective of
the PreferenceRegressor at N = 2048 against the oracle, and K / L / K^-1 identities at N = 4096 and 8192.

Tolerances: FP64 path 1e-5 relative, tensor path 1e-3 relative (north_star), as max |err| / max |ref| per output array."""
import importlib

import numpy as np
import pytest

import support as S

pkg = importlib.import_module("sequential-line-search_b200")
pytestmark = pytest.mark.gpu
RT64, RT32 = 1e-5, 1e-3


def _err(got, want):
    got, want = np.asarray(got), np.asarray(want)
    assert not np.isnan(got).any()
    return float(np.max(np.abs(got - want))) / max(float(np.max(np.abs(want))), 1e-300)


def _candidates(X, M, D):
    """Uniform candidates plus candidates next to data points (where sigma^2 = a - k K^-1 k cancels) and outside the box."""
    rng = np.random.default_rng(17)
    near = X[:, :8] + 1e-3 * rng.standard_normal((D, 8))
    return S.f64(np.concatenate([S.make_queries(M - 10, D), near, X[:, 5:6], np.full((D, 1), 1.3)], axis=1))


@pytest.mark.parametrize("kt,D,N,xkind", [(S.SE, 16, 2048, "uniform"), (S.SE, 16, 2048, "sls"), (S.MATERN, 16, 2048, "sls"), (S.SE, 8, 512, "uniform"),
                                          (S.MATERN, 8, 512, "sls")])
def test_sweeps_match_the_oracle_at_benchmark_sizes(oracle, kt, D, N, xkind):
    X, theta, noise = S.make_X(N, D, xkind), S.make_theta(D, "default"), 0.005
    y = S.make_y(X)
    ctx = pkg.Context(0)
    try:
        ctx.fit(X, kt, theta, noise, y)
        m = oracle.model(kt, X, theta, noise, y)
        idx_o, f_best = oracle.f_best(m)
        Q = _candidates(X, 64, D)
        for acq, beta in ((S.EI, 1.0), (S.UCB, 2.0)):
            want = oracle.acq_batch(m, acq, beta, f_best, Q)
            for mode, tol, name in ((pkg.SWEEP_FP64, RT64, "fp64"), (pkg.SWEEP_TENSOR, RT32, "tensor")):
                ctx.set_sweep_mode(mode)
                mu, sigma, dmu, dsigma = ctx.posterior_batch(Q)
                val, grad = ctx.acq_batch(acq, beta, Q)
                errs = {"mu": _err(mu, want["mu"]), "sigma": _err(sigma, want["sigma"]), "dmu": _err(dmu, want["dmu"]),
                        "dsigma": _err(dsigma, want["dsigma"]), "val": _err(val, want["val"]), "grad": _err(grad, want["grad"])}
                print(f"\nN={N} D={D} kernel={kt} {xkind} acq={acq} {name}: " + " ".join(f"{k}={v:.1e}" for k, v in errs.items()))
                for k, v in errs.items():
                    assert v <= tol, (name, k, v)
        ctx.set_sweep_mode(pkg.SWEEP_FP64)
        f_dev, idx_dev = ctx.f_best()
        assert idx_dev == idx_o and abs(f_dev - f_best) <= RT64 * abs(f_best)
    finally:
        ctx.close()


def test_config4_candidates_against_the_reference_itself(ref):
    """N = 2048, D = 16: EI and its gradient for three candidates through the reference's own CalcAcquisitionValue /
    CalcAcquisitionValueDerivative (each an O(N^3) call there), against the FP64 and tensor sweeps."""
    kt, D, N = S.SE, 16, 2048
    X, theta, noise = S.make_X(N, D, "uniform"), S.make_theta(D, "default"), 0.005
    y = S.make_y(X)
    Q = _candidates(X, 12, D)[:, [0, 3, 5]]  # two uniform candidates and one next to a data point
    h = ref.gpr_create(kt, X, y, theta, noise)
    try:
        reg = ref.gpr_regressor(h)
        want = [ref.acq(reg, S.EI, 1.0, Q[:, m]) for m in range(Q.shape[1])]
    finally:
        ref.gpr_destroy(h)
    want_v, want_g = np.array([w[0] for w in want]), np.array([w[1] for w in want]).T
    ctx = pkg.Context(0)
    try:
        ctx.fit(X, kt, theta, noise, y)
        for mode, tol in ((pkg.SWEEP_FP64, RT64), (pkg.SWEEP_TENSOR, RT32)):
            ctx.set_sweep_mode(mode)
            val, grad = ctx.acq_batch(S.EI, 1.0, Q)
            print(f"\nvs the reference, mode {mode}: EI err {_err(val, want_v):.1e}, grad err {_err(grad, want_g):.1e}")
            assert _err(val, want_v) <= tol and _err(grad, want_g) <= tol
    finally:
        ctx.close()


def test_preference_map_objective_at_config3_size(oracle):
    """config 3: N = 2048, D = 16, 683 triplets, hyper-parameters in the variable vector (2066 variables): F and its gradient at
    one point against the oracle (an O(N^3) evaluation on the CPU)."""
    kt, D, N = S.SE, 16, 2048
    X = S.make_X(N, D, "uniform")
    offsets, idx = S.make_tuples(X)
    a, r, b, var, btl = 0.5, 0.5, 0.005, 0.25, 0.01
    rng = np.random.default_rng(21)
    x = np.concatenate([0.05 * rng.standard_normal(N), [0.45, 0.006], rng.uniform(0.4, 0.7, D)])
    want_f, want_g = oracle.map_objective_pref(kt, X, offsets, idx, True, a, r, b, var, btl, x)
    ctx = pkg.Context(0)
    try:
        ctx.set_data(X)
        ctx.set_preferences(offsets, idx)
        f, g = ctx.map_objective_pref(kt, x, True, a, r, b, var, btl)
    finally:
        ctx.close()
    e_y, e_h = _err(g[:N], want_g[:N]), float(np.max(np.abs(g[N:] - want_g[N:]) / np.maximum(np.abs(want_g[N:]), 1e-3 * np.max(np.abs(want_g[N:])))))
    print(f"\nMAP objective N={N}: |f - f_oracle| / |f| = {abs(f - want_f) / abs(want_f):.1e}, grad_y err {e_y:.1e}, hyper-gradient err {e_h:.1e}")
    assert abs(f - want_f) <= RT64 * abs(want_f) and e_y <= RT64 and e_h <= RT64


@pytest.mark.parametrize("N", [4096, 8192])
def test_factor_and_inverse_identities_at_large_n(N):
    """No CPU checker finishes an N = 8192 Cholesky in test time; the factor and the inverse are checked by what defines them,
    with products formed in float64 numpy on row / column samples: (L L^T)[S, :] = K[S, :], (K K^-1)[S, :] = I[S, :], logdet."""
    kt, D = S.SE, 16
    X, theta, noise = S.make_X(N, D, "uniform"), S.make_theta(D, "default"), 0.005
    ctx = pkg.Context(0)
    try:
        ctx.set_data(X)
        K = ctx.gram(kt, theta, noise, want=True)
        logdet, L = ctx.factor(want_L=True)
        Kinv = ctx.inverse(want=True)
    finally:
        ctx.close()
    rows = np.random.default_rng(3).choice(N, 96, replace=False)
    for r in rows:
        assert not L[r, r + 1:].any()  # strictly upper part of the returned factor is zero
    e_llt = np.max(np.abs(L[rows] @ L.T - K[rows])) / np.max(np.abs(K))
    e_inv = np.max(np.abs(K[rows] @ Kinv - np.eye(N)[rows]))
    e_sym = np.max(np.abs(Kinv[rows] - Kinv[:, rows].T)) / np.max(np.abs(Kinv))
    e_ld = abs(logdet - 2.0 * np.sum(np.log(np.diag(L)))) / abs(logdet)
    print(f"\nN={N}: |LL^T - K| / |K| = {e_llt:.1e}, |K K^-1 - I| = {e_inv:.1e}, asymmetry of K^-1 {e_sym:.1e}, logdet {e_ld:.1e}")
    assert e_llt < 1e-13 and e_inv < 1e-9 and e_sym < 1e-12 and e_ld < 1e-12


@pytest.mark.parametrize("panel,look_ahead,switch_rem", [(2, 1, 0), (4, 1, 0), (4, 0, 0), (3, 1, 0), (2, 1, 5), (3, 1, 7), (4, 0, 6)])
@pytest.mark.parametrize("N", [700, 1100])
def test_two_level_cholesky_equals_the_single_level_one(monkeypatch, oracle, N, panel, look_ahead, switch_rem):
    """The two-level factorisation (panels + one rank-64p update per panel on two streams, slsgp.cu:do_factor) is the default from
    N = 4096 on; forced here at sizes the plain-C oracle factors in a second. Same matrix, same 64-wide block columns: L differs
    from the single-level sweep only by the summation order inside the trailing update (checked at 1e-13 against it, and
    against the oracle's Cholesky), W = L^-1 and K^-1 follow."""
    kt, D = S.MATERN, 6
    X, theta, noise = S.make_X(N, D, "sls"), S.make_theta(D, "perturbed"), 0.005
    got = {}
    for two_level in (0, 1):
        monkeypatch.setenv("SLSGP_CHOL_TWO_LEVEL_FROM", "3" if two_level else "100000")
        monkeypatch.setenv("SLSGP_CHOL_PANEL", str(panel))
        monkeypatch.setenv("SLSGP_CHOL_LOOKAHEAD", str(look_ahead))
        monkeypatch.setenv("SLSGP_CHOL_SWITCH_REM", str(switch_rem))  # block columns left to the single-level sweep at the end
        ctx = pkg.Context(0)
        try:
            ctx.set_data(X)
            K = ctx.gram(kt, theta, noise, want=True)
            for _ in range(2):  # the second factorisation reuses the streams and events of the first
                ctx.invalidate()
                ctx.gram(kt, theta, noise, want=False)
                logdet, L = ctx.factor(want_L=True)
            got[two_level] = (logdet, L, ctx.inverse(want=True))
        finally:
            ctx.close()
    L_o, status = oracle.cholesky(oracle.large_ky(kt, X, theta, noise))
    assert status == 0
    e_o = _err(got[1][1], L_o)
    e_l = _err(got[1][1], got[0][1])
    e_i = _err(got[1][2], got[0][2])
    print(f"\nN={N} panel={panel} look-ahead={look_ahead} switch={switch_rem}: L vs oracle {e_o:.1e}, vs single level {e_l:.1e}, K^-1 vs single level {e_i:.1e}")
    assert e_o < 1e-9 and e_l < 1e-13 and e_i < 1e-9
    assert abs(got[1][0] - got[0][0]) <= 1e-12 * abs(got[0][0])
    assert not np.triu(got[1][1], 1).any()


def test_pair_pivot_tile_is_bit_identical_to_the_single_pivot_tile(monkeypatch):
    """potf2_inverse_regs_pair (two pivots per barrier, the default) performs the same operations in the same order on every
    entry as potf2_inverse_regs; the factors must agree bit for bit, including the first non-positive pivot it reports."""
    kt, D, N = S.SE, 8, 448
    X, theta, noise = S.make_X(N, D, "uniform"), S.make_theta(D, "default"), 0.005
    got = {}
    for piv in (1, 2):
        monkeypatch.setenv("SLSGP_CHOL_PIVOTS", str(piv))
        ctx = pkg.Context(0)
        try:
            ctx.set_data(X)
            ctx.gram(kt, theta, noise, want=False)
            logdet, L = ctx.factor(want_L=True)
            got[piv] = (logdet, L, ctx.inverse(want=True))
            Xd = X.copy()
            Xd[:, 100] = Xd[:, 37]  # duplicated point, slightly negative noise: the Schur complement at 100 is negative
            ctx.set_data(Xd)
            ctx.gram(kt, theta, -1e-3, want=False)
            with pytest.raises(pkg.SlsgpError) as ei:
                ctx.factor()
            got[piv] += (str(ei.value),)
        finally:
            ctx.close()
    assert got[1][0] == got[2][0] and np.array_equal(got[1][1], got[2][1]) and np.array_equal(got[1][2], got[2][2])
    assert got[1][3] == got[2][3] and "pivot" in got[2][3]
