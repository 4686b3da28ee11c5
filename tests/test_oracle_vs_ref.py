"""Pin the plain-C oracle (oracle/slsgp_oracle.c) against the reference ITSELF.

oracle/_ref/libsls_ref_probe.so is the reference's unmodified C++ (src/*.cpp + mathtoolbox) compiled in this
container against include/eigen-lite; these tests compare the restatement with it function by function on seeded
inputs. The reference publishes no golden vectors for this path, so this comparison — plus the
fixtures in tests/golden generated from the same build — is what pins parity.
Tolerance: 1e-12 relative unless stated (both sides are IEEE double; only summation order differs).
"""
import numpy as np
import pytest

import support as S

TOL = 1e-12
CASES = [(S.SE, 5, 12, "uniform", "default"), (S.MATERN, 5, 12, "uniform", "perturbed"),
         (S.SE, 16, 40, "sls", "perturbed"), (S.MATERN, 8, 33, "sls", "default")]


@pytest.mark.parametrize("kt", [S.SE, S.MATERN])
def test_kernel_pointwise(oracle, ref, kt):
    rng = np.random.default_rng(10 + kt)
    for D in (1, 3, 6, 16, 64):
        theta = S.make_theta(D, "perturbed", seed=D)
        for _ in range(20):
            xa, xb = rng.random(D), rng.random(D)
            k, dth, dx = ref.kernel(kt, xa, xb, theta)
            assert abs(oracle.kernel(kt, xa, xb, theta) - k) <= TOL * abs(k)
            assert S.rel_err(oracle.kernel_theta_derivative(kt, xa, xb, theta), dth) < TOL
            assert S.rel_err(oracle.kernel_first_arg_derivative(kt, xa, xb, theta), dx) < TOL
        # coincident points: Matern first-arg derivative is defined as 0 (kernel-functions.cpp:198)
        xa = rng.random(D)
        k, dth, dx = ref.kernel(kt, xa, xa, theta)
        assert oracle.kernel(kt, xa, xa, theta) == k == theta[0]
        np.testing.assert_array_equal(oracle.kernel_first_arg_derivative(kt, xa, xa, theta), dx)
        np.testing.assert_array_equal(dx, np.zeros(D))


def test_se_first_arg_derivative_is_twice_analytic(oracle, ref):
    """kernel-functions.cpp:92 has -2.0 where the analytic derivative has -1.0."""
    rng = np.random.default_rng(0)
    D = 4
    theta = S.make_theta(D, "perturbed")
    xa, xb = rng.random(D), rng.random(D)
    eps = 1e-6
    fd = np.array([(ref.kernel(S.SE, xa + eps * e, xb, theta)[0] - ref.kernel(S.SE, xa - eps * e, xb, theta)[0])
                   / (2 * eps) for e in np.eye(D)])
    np.testing.assert_allclose(ref.kernel(S.SE, xa, xb, theta)[2], 2.0 * fd, rtol=1e-6)
    np.testing.assert_allclose(oracle.kernel_first_arg_derivative(S.SE, xa, xb, theta), 2.0 * fd, rtol=1e-6)
    fd_m = np.array([(ref.kernel(S.MATERN, xa + eps * e, xb, theta)[0]
                      - ref.kernel(S.MATERN, xa - eps * e, xb, theta)[0]) / (2 * eps) for e in np.eye(D)])
    np.testing.assert_allclose(ref.kernel(S.MATERN, xa, xb, theta)[2], fd_m, rtol=1e-6)


@pytest.mark.parametrize("kt,D,N,xkind,tkind", CASES)
def test_kernel_matrices(oracle, ref, kt, D, N, xkind, tkind):
    X, theta = S.make_X(N, D, xkind), S.make_theta(D, tkind)
    np.testing.assert_allclose(oracle.large_ky(kt, X, theta, 0.005), ref.large_ky(kt, X, theta, 0.005),
                               rtol=TOL, atol=0)
    x = S.make_queries(1, D)[:, 0]
    k, J = ref.small_k(kt, X, theta, x)
    np.testing.assert_allclose(oracle.small_k(kt, X, theta, x), k, rtol=TOL)
    np.testing.assert_allclose(oracle.small_k_x_derivative(kt, X, theta, x), J, rtol=TOL, atol=1e-300)
    np.testing.assert_allclose(oracle.large_ky_theta_derivative(kt, X, theta),
                               ref.large_ky_theta_derivative(kt, X, theta), rtol=TOL, atol=1e-300)


@pytest.mark.parametrize("kt,D,N,xkind,tkind", CASES)
def test_gpr_predictions_and_acquisition(oracle, ref, kt, D, N, xkind, tkind):
    X, theta = S.make_X(N, D, xkind), S.make_theta(D, tkind)
    y = S.make_y(X)
    b = 0.005
    h = ref.gpr_create(kt, X, y, theta, b)
    reg = ref.gpr_regressor(h)
    m = oracle.model(kt, X, theta, b, y)
    K_ref, Kinv_ref = ref.gpr_state(h, N)
    np.testing.assert_allclose(m._keep[4], K_ref, rtol=TOL)
    assert S.rel_err(oracle.inverse(K_ref), Kinv_ref) < 1e-9  # LLT-based vs the reference's LU inverse
    i_best, f_best = oracle.f_best(m)
    np.testing.assert_array_equal(ref.x_best(reg, D), X[:, i_best])
    # queries: random points, a data point (sigma ~ sqrt(b)), and a point far outside the data
    Q = np.concatenate([S.make_queries(6, D), X[:, :1], np.full((D, 1), 3.0)], axis=1)
    for x in Q.T:
        mu, sg, dmu, dsg = ref.predict(reg, x)
        omu, osg, odmu, odsg = oracle.predict(m, x)
        assert abs(omu - mu) <= 1e-9 * max(1.0, abs(mu))
        assert abs(osg - sg) <= 1e-9
        np.testing.assert_allclose(odmu, dmu, rtol=1e-8, atol=1e-12)
        np.testing.assert_allclose(odsg, dsg, rtol=1e-7, atol=1e-12)
        for acq, beta in ((S.EI, 1.0), (S.UCB, 2.5)):
            v, g = ref.acq(reg, acq, beta, x)
            ov, og = oracle.acq(m, acq, beta, f_best, x)
            assert abs(ov - v) <= 1e-9 * max(1e-3, abs(v))
            np.testing.assert_allclose(og, g, rtol=1e-7, atol=1e-12)
    out = oracle.acq_batch(m, S.EI, 1.0, f_best, Q)
    for q, x in enumerate(Q.T):
        v, g = ref.acq(reg, S.EI, 1.0, x)
        assert abs(out["val"][q] - v) <= 1e-9 * max(1e-3, abs(v))
        np.testing.assert_allclose(out["grad"][:, q], g, rtol=1e-7, atol=1e-12)
    ref.gpr_destroy(h)


@pytest.mark.parametrize("kt,D,N,xkind,tkind", CASES)
@pytest.mark.parametrize("use_map", [False, True])
def test_preference_regressor(oracle, ref, kt, D, N, xkind, tkind, use_map):
    X = S.make_X(N, D, xkind)
    offsets, idx = S.make_tuples(X)
    a, r, b, var, btl = 0.5, 0.5, 0.005, 0.25, 0.01
    rng = np.random.default_rng(5)
    y = 0.05 * rng.standard_normal(N)
    theta = S.make_theta(D, tkind)
    sol = np.concatenate([y, [theta[0], 0.007], theta[1:]]) if use_map else y
    h = ref.pref_create(kt, X, offsets, idx, use_map, a, r, b, var, btl, sol)
    st = ref.pref_state(h, N, D)
    np.testing.assert_array_equal(st["y"], y)
    th_eff = theta if use_map else S.make_theta(D, "default")
    b_eff = 0.007 if use_map else b
    np.testing.assert_array_equal(st["theta"], th_eff)
    assert st["b"] == b_eff
    # state: K and its Cholesky factor
    m = oracle.model(kt, X, th_eff, b_eff, y)
    np.testing.assert_allclose(m._keep[4], st["K"], rtol=TOL)
    np.testing.assert_allclose(m._keep[3], st["L"], rtol=1e-9, atol=1e-13)
    # predictions through LLT::solve
    reg = ref.pref_regressor(h)
    _, f_best = oracle.f_best(m)
    for x in S.make_queries(5, D).T:
        mu, sg, dmu, dsg = ref.predict(reg, x)
        omu, osg, odmu, odsg = oracle.predict(m, x)
        assert abs(omu - mu) <= 1e-10 * max(1.0, abs(mu)) and abs(osg - sg) <= 1e-10
        np.testing.assert_allclose(odmu, dmu, rtol=1e-9, atol=1e-13)
        np.testing.assert_allclose(odsg, dsg, rtol=1e-9, atol=1e-13)
        v, g = ref.acq(reg, S.EI, 1.0, x)
        ov, og = oracle.acq(m, S.EI, 1.0, f_best, x)
        assert abs(ov - v) <= 1e-9 * max(1e-3, abs(v))
        np.testing.assert_allclose(og, g, rtol=1e-8, atol=1e-13)
    # MAP objective + gradient at a few points (incl. the default initial point of PerformMapEstimation)
    pts = [sol, sol * 1.1 + 0.01]
    pts.append(np.concatenate([np.zeros(N), [a, b], np.full(D, r)]) if use_map else np.zeros(N))
    for x in pts:
        f, g = ref.pref_objective(h, x)
        of, og = oracle.map_objective_pref(kt, X, offsets, idx, use_map, a, r, b, var, btl, x)
        assert abs(of - f) <= 1e-10 * abs(f)
        np.testing.assert_allclose(og, g, rtol=1e-8, atol=1e-9)
        f2, _ = ref.pref_objective(h, x, want_grad=False)
        assert f2 == f
    ref.pref_destroy(h)


@pytest.mark.parametrize("kt,D,N,xkind,tkind", CASES[:3])
def test_gpr_map_objective(oracle, ref, kt, D, N, xkind, tkind):
    X = S.make_X(N, D, xkind)
    y = S.make_y(X)
    rng = np.random.default_rng(6)
    pts = np.stack([np.concatenate([[0.5, 1e-4], np.full(D, 0.5)]),          # x_ini of PerformMapEstimation
                    np.concatenate([[0.8, 0.01], rng.uniform(0.2, 1.0, D)]),
                    np.concatenate([[0.3, 0.05], rng.uniform(0.3, 2.0, D)])])
    f, g = ref.gpr_objective(kt, X, y, pts)
    for p in range(len(pts)):
        of, og = oracle.map_objective_gpr(kt, X, y, pts[p])
        assert abs(of - f[p]) <= 1e-9 * abs(f[p])
        np.testing.assert_allclose(og, g[p], rtol=1e-7, atol=1e-7)


def test_btl(oracle, ref):
    rng = np.random.default_rng(7)
    for n in (2, 3, 5):
        f = 0.05 * rng.standard_normal(n)
        for scale in (0.01, 1.0):
            v, d = ref.btl(f, scale)
            ov, od = oracle.btl(f, scale)
            assert abs(ov - v) <= 1e-14 * abs(v)
            np.testing.assert_allclose(od, d, rtol=1e-13)


def test_closed_form_known_answers(oracle):
    """N = 1 closed forms: mu = k y / (a + b), sigma^2 = a - k^2 / (a + b);
    EI at Z = 0 equals sigma / sqrt(2 pi)."""
    D = 3
    X = S.f64(np.full((D, 1), 0.3))
    theta = np.array([0.5, 0.4, 0.5, 0.6])
    a, b, y = 0.5, 0.005, np.array([0.7])
    m = oracle.model(S.SE, X, theta, b, y)
    x = np.array([0.35, 0.2, 0.5])
    k = a * np.exp(-0.5 * (((x - 0.3) / theta[1:]) ** 2).sum())
    mu, sg, _, _ = oracle.predict(m, x)
    assert abs(mu - k * 0.7 / (a + b)) < 1e-15
    assert abs(sg - np.sqrt(a - k * k / (a + b))) < 1e-15
    v, _ = oracle.acq(m, S.EI, 1.0, mu, x)  # f_best := mu(x)  ->  Z = 0
    assert abs(v - sg / np.sqrt(2 * np.pi)) < 1e-15


@pytest.mark.parametrize("kt,D,N", [(S.SE, 4, 17), (S.MATERN, 6, 30)])
def test_noiseless_formulation_objective(oracle, kt, D, N):
    """SEQUENTIAL_LINE_SEARCH_USE_NOISELESS_FORMULATION (CMakeLists.txt:28-33; src/preference-regressor.cpp:48-52,141,180-188,237):
    the oracle's restatement against the reference compiled with that option (oracle/_ref/libsls_ref_probe_noiseless.so)."""
    import os
    if not os.path.exists(S.REF_NOISELESS_PATH):
        pytest.skip("oracle/_ref/libsls_ref_probe_noiseless.so not built")
    ref_nl = S.Ref(S.REF_NOISELESS_PATH)
    X = S.make_X(N, D, "uniform")  # K_f alone must be positive definite: well-separated points
    offsets, idx = S.make_tuples(X)
    a, r, b, var, btl = 0.5, 0.5, 0.005, 0.25, 0.01
    rng = np.random.default_rng(8)
    y = 0.05 * rng.standard_normal(N)
    theta = S.make_theta(D, "perturbed")
    sol = np.concatenate([y, [theta[0], 0.0123], 0.3 * theta[1:]])
    h = ref_nl.pref_create(kt, X, offsets, idx, True, a, r, b, var, btl, sol)
    try:
        assert ref_nl.pref_state(h, N, D)["b"] == 0.0  # m_noise_hyperparam = b_fixed (:389)
        for x in (sol, sol * 1.05 + 0.002):
            f, g = ref_nl.pref_objective(h, x)
            of, og = oracle.map_objective_pref_noiseless(kt, X, offsets, idx, True, a, r, b, var, btl, x)
            assert np.isfinite(f) and abs(of - f) <= 1e-9 * abs(f)
            np.testing.assert_allclose(og, g, rtol=1e-6, atol=1e-7)
            assert og[N + 1] == 0.0 and g[N + 1] == 0.0
            # and it is NOT the standard objective
            sf, _ = oracle.map_objective_pref(kt, X, offsets, idx, True, a, r, b, var, btl, x)
            assert abs(sf - f) > 1e-6 * abs(f)
    finally:
        ref_nl.pref_destroy(h)
