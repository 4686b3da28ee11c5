"""The parts of the reference's public surface closed in round 2 (VERDICT r1 "Missing" #3), on the GPU:
the L1 free functions CalcSmallK / CalcSmallKSmallXDerivative / CalcLargeKYThetaDerivative / CalcLargeKYNoiseLevelDerivative
(include/sequential-line-search/regressor.hpp:42-70), the noiseless formulation, deep copies of device-backed regressors,
slsgp_trim and the sweep-mode guard."""
import importlib

import numpy as np
import pytest

import support as S

pkg = importlib.import_module("sequential-line-search_b200")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def host():
    return pkg.hostlib.Host()


@pytest.mark.parametrize("kt,D,N", [(S.SE, 6, 40), (S.MATERN, 6, 40), (S.SE, 64, 64), (S.MATERN, 3, 1), (S.SE, 16, 300)])
def test_l1_free_functions_match_the_reference(host, ref, oracle, kt, D, N):
    X, theta = S.make_X(N, D, "sls"), S.make_theta(D, "perturbed")
    for x in list(S.make_queries(3, D).T) + [X[:, 0].copy()]:  # the last one coincides with a data point (Matern guard, r = 0)
        k, J = host.small_k(kt, X, theta, x)
        k_r, J_r = ref.small_k(kt, X, theta, x)
        assert S.rel_err(k, k_r) < 1e-13
        assert np.max(np.abs(J - J_r)) <= 1e-12 * max(np.max(np.abs(J_r)), 1e-300)
    T = host.large_ky_theta_derivative(kt, X, theta)
    T_r = ref.large_ky_theta_derivative(kt, X, theta)
    assert T.shape == T_r.shape == (D + 1, N, N)
    for t in range(D + 1):
        assert np.max(np.abs(T[t] - T_r[t])) <= 1e-12 * max(np.max(np.abs(T_r[t])), 1e-300), t
    # the C ABI directly, on a context that holds a fitted model: the model must be left untouched
    ctx = pkg.Context(0)
    y = S.make_y(X)
    ctx.fit(X, kt, S.make_theta(D, "default"), 0.005, y)
    Q = S.make_queries(4, D)
    before = ctx.acq_batch(0, 1.0, Q)
    k, J = ctx.small_k(kt, theta, Q[:, 0])
    assert S.rel_err(k, oracle.small_k(kt, X, theta, Q[:, 0])) < 1e-13
    assert S.rel_err(J, oracle.small_k_x_derivative(kt, X, theta, Q[:, 0])) < 1e-12
    assert S.rel_err(ctx.gram_theta_derivative(kt, theta), oracle.large_ky_theta_derivative(kt, X, theta)) < 1e-12
    after = ctx.acq_batch(0, 1.0, Q)
    np.testing.assert_array_equal(before[0], after[0])
    np.testing.assert_array_equal(before[1], after[1])
    ctx.close()


@pytest.mark.parametrize("kt,D,N", [(S.SE, 4, 17), (S.MATERN, 6, 30), (S.MATERN, 5, 90)])
def test_noiseless_formulation_on_the_device(oracle, kt, D, N):
    """SLSGP_COMPAT_NOISELESS against the oracle's restatement of the reference's noiseless build (pinned on the CPU side by
    tests/test_oracle_vs_ref.py::test_noiseless_formulation_objective). 1e-5 relative, north_star FP64."""
    X = S.make_X(N, D, "uniform")
    offsets, idx = S.make_tuples(X)
    a, r, b, var, btl = 0.5, 0.5, 0.005, 0.25, 0.01
    rng = np.random.default_rng(8)
    theta = S.make_theta(D, "perturbed")
    x = np.concatenate([0.05 * rng.standard_normal(N), [theta[0], 0.0123], 0.3 * theta[1:]])
    ctx = pkg.Context(0)
    ctx.set_data(X)
    ctx.set_preferences(offsets, idx)
    ctx.set_compat_flags(pkg.COMPAT_SE_XGRAD_2X | pkg.COMPAT_NOISELESS)
    f, g = ctx.map_objective_pref(kt, x, True, a, r, b, var, btl)
    of, og = oracle.map_objective_pref_noiseless(kt, X, offsets, idx, True, a, r, b, var, btl, x)
    assert abs(f - of) <= 1e-5 * abs(of)
    assert np.max(np.abs(g - og)) <= 1e-5 * np.max(np.abs(og)) and g[N + 1] == 0.0
    ctx.set_compat_flags(pkg.COMPAT_SE_XGRAD_2X)
    f_std, _ = ctx.map_objective_pref(kt, x, True, a, r, b, var, btl)
    sf, _ = oracle.map_objective_pref(kt, X, offsets, idx, True, a, r, b, var, btl, x)
    assert abs(f_std - sf) <= 1e-5 * abs(sf) and abs(f_std - f) > 1e-6 * abs(f)
    ctx.close()


def test_copies_of_a_regressor_are_independent(host, ref):
    """copies used to share one device context. Now a copy refits its own model; AppendPoint on either side leaves
    the other untouched, and both keep answering like the reference regressor of their own data."""
    kt, D, N = S.MATERN, 5, 33
    X, theta, b = S.make_X(N + 2, D, "sls"), S.make_theta(D, "perturbed"), 0.005
    y = S.make_y(X)
    h = host.gpr_create(kt, X[:, :N], y[:N], theta, b)
    c = host.gpr_copy(h)
    r_small, r_big = ref.gpr_create(kt, X[:, :N], y[:N], theta, b), ref.gpr_create(kt, X[:, :N + 1], y[:N + 1], theta, b)
    try:
        host.gpr_append_point(c, X[:, N], y[N])          # grows the COPY only
        assert host.gpr_num_points(h) == N and host.gpr_num_points(c) == N + 1
        Q = S.make_queries(6, D)
        for m in range(Q.shape[1]):
            for got, want in ((host.predict(host.gpr_regressor(h), Q[:, m]), ref.predict(ref.gpr_regressor(r_small), Q[:, m])),
                              (host.predict(host.gpr_regressor(c), Q[:, m]), ref.predict(ref.gpr_regressor(r_big), Q[:, m]))):
                assert abs(got[0] - want[0]) <= 1e-5 * max(abs(want[0]), 1e-3) and abs(got[1] - want[1]) <= 1e-5 * max(abs(want[1]), 1e-3)
        host.gpr_append_point(h, X[:, N + 1], y[N + 1])  # and the other way round
        assert host.gpr_num_points(h) == N + 1 and host.gpr_num_points(c) == N + 1
        c2 = host.gpr_copy(c)                             # a copy of a copy, used without ever touching its source again
        host.gpr_destroy(c)
        got, want = host.predict(host.gpr_regressor(c2), Q[:, 0]), ref.predict(ref.gpr_regressor(r_big), Q[:, 0])
        assert abs(got[0] - want[0]) <= 1e-5 * max(abs(want[0]), 1e-3)
        host.gpr_destroy(c2)
    finally:
        host.gpr_destroy(h)
        ref.gpr_destroy(r_small)
        ref.gpr_destroy(r_big)


def test_trim_releases_the_sweep_workspace_and_keeps_the_model(oracle):
    kt, D, N = S.SE, 6, 50
    X, theta = S.make_X(N, D, "sls"), S.make_theta(D, "default")
    y = S.make_y(X)
    ctx = pkg.Context(0)
    ctx.fit(X, kt, theta, 0.005, y)
    Q = S.make_queries(3000, D)
    v0, g0 = ctx.acq_batch(0, 1.0, Q)
    ctx.trim(0)
    v1, g1 = ctx.acq_batch(0, 1.0, Q)
    np.testing.assert_array_equal(v0, v1)
    np.testing.assert_array_equal(g0, g1)
    ctx.close()
    pkg.hostlib.load_host_library().b200_release_device_resources()


@pytest.mark.parametrize("acq,beta,mode", [(S.EI, 1.0, "fp64"), (S.UCB, 2.0, "fp64"), (S.EI, 1.0, "tensor")])
def test_pair_acq_argmax_equals_the_explicit_pipeline(acq, beta, mode):
    """slsgp_pair_acq_argmax (device-resident global stage of FindNextPoints: mu from one model, sigma from another) against the
    same steps done one by one through host buffers: slsgp_candidates, two slsgp_posterior_batch calls, slsgp_acq_from_posterior."""
    kt, D, N = S.MATERN, 5, 40
    X, theta = S.make_X(N + 3, D, "sls"), S.make_theta(D, "perturbed")
    y = S.make_y(X)
    a, b = pkg.Context(0), pkg.Context(0)
    try:
        a.fit(X[:, :N], kt, theta, 0.005, y[:N])
        b.fit(X, kt, theta, 0.005, y)          # the model that already holds three pending points
        sweep = pkg.SWEEP_TENSOR if mode == "tensor" else pkg.SWEEP_FP64
        a.set_sweep_mode(sweep)
        b.set_sweep_mode(sweep)
        seed, first, count = 42, 17, 150000     # two chunks of 2^17
        x, v, idx = a.pair_acq_argmax(b, acq, beta, seed, first, count)
        Q = a.candidates(seed, first, count)
        mu, _, _, _ = a.posterior_batch(Q, grads=False)
        _, sigma, _, _ = b.posterior_batch(Q, grads=False)
        f_best, _ = a.f_best()
        val, _ = a.acq_from_posterior(acq, beta, f_best, mu, sigma)
        assert idx == first + int(np.argmax(val)) and v == val.max()
        np.testing.assert_array_equal(x, Q[:, idx - first])
    finally:
        a.close()
        b.close()
