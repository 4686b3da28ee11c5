"""GPU parity proper: libslsgp (through the C ABI) vs. the plain-C oracle on the same seeded inputs, at sizes the
oracle finishes in seconds, plus size-independent properties at BASELINE.json's full sizes.
FP64 tolerance: north_star's 1e-5 relative, written as RT below (observed errors are orders of magnitude smaller)."""
import importlib

import numpy as np
import pytest

import support as S

pytestmark = pytest.mark.gpu
RT = 1e-5


def check(name, got, want, rtol=RT, atol=0.0):
    got, want = np.asarray(got), np.asarray(want)
    scale = max(float(np.max(np.abs(want))), 1e-300)
    err = float(np.max(np.abs(got - want)))
    assert err <= rtol * scale + atol, f"{name}: max abs err {err:.3e} vs scale {scale:.3e}"


@pytest.fixture(scope="module")
def slsb():
    return importlib.import_module("sequential-line-search_b200")


@pytest.fixture(scope="module")
def ctx(slsb):
    c = slsb.Context(0)
    yield c
    c.close()


# ragged and tile-aligned sizes: 1 point, one tile, one tile + 1, multi-level doubling with a partial last block
SIZES = [(S.SE, 4, 1), (S.SE, 6, 63), (S.MATERN, 6, 64), (S.SE, 8, 65), (S.MATERN, 16, 200), (S.SE, 16, 448),
         (S.SE, 64, 130), (S.MATERN, 3, 321)]


@pytest.mark.parametrize("kt,D,N", SIZES)
@pytest.mark.parametrize("xkind", ["uniform", "sls"])
def test_gram_factor_inverse_alpha(ctx, oracle, kt, D, N, xkind):
    X, theta = S.make_X(N, D, xkind), S.make_theta(D, "perturbed")
    y = S.make_y(X)
    noise = 0.005
    K_o = oracle.large_ky(kt, X, theta, noise)
    L_o, status = oracle.cholesky(K_o)
    assert status == 0
    ctx.set_data(X)
    K = ctx.gram(kt, theta, noise)
    check("K", K, K_o, 1e-13)
    assert np.array_equal(K, K.T)
    logdet, L = ctx.factor(want_L=True)
    check("L", L, L_o, 1e-9)
    check("logdet", logdet, oracle.logdet(L_o), 1e-11, atol=1e-9)
    Kinv = ctx.inverse()
    check("Kinv", Kinv, oracle.inverse(K_o), 1e-7)
    assert np.array_equal(Kinv, Kinv.T)
    alpha = ctx.solve_alpha(y)
    check("alpha", alpha, oracle.llt_solve(L_o, y), 1e-7)
    m = oracle.model(kt, X, theta, noise, y)
    i_best, f_best = oracle.f_best(m)
    f, i = ctx.f_best()
    assert i == i_best
    check("f_best", f, f_best, 1e-9)


@pytest.mark.parametrize("kt,D,N", SIZES)
def test_posterior_and_acquisition_sweep(ctx, oracle, kt, D, N):
    X, theta = S.make_X(N, D, "sls"), S.make_theta(D, "perturbed")
    y = S.make_y(X)
    noise = 0.005
    ctx.fit(X, kt, theta, noise, y)
    m = oracle.model(kt, X, theta, noise, y)
    _, f_best = oracle.f_best(m)
    M = 150  # not a multiple of the 64-wide candidate tile
    Q = np.concatenate([S.make_queries(M - 2, D), X[:, :1], np.full((D, 1), 1.7)], axis=1)
    for acq, beta in ((0, 1.0), (1, 2.5)):
        want = oracle.acq_batch(m, acq, beta, f_best, Q)
        if acq == 0:
            mu, sigma, dmu, dsigma = ctx.posterior_batch(Q)
            check("mu", mu, want["mu"])
            check("sigma", sigma, want["sigma"])
            check("dmu", dmu, want["dmu"])
            check("dsigma", dsigma, want["dsigma"])
        val, grad = ctx.acq_batch(acq, beta, Q)
        check("val", val, want["val"], atol=1e-300)
        check("grad", grad, want["grad"], atol=1e-300)
        val2, _ = ctx.acq_batch(acq, beta, Q, grads=False)
        np.testing.assert_array_equal(val2, val)


def test_acq_from_posterior_reproduces_the_sweep(ctx):
    """slsgp_acq_from_posterior applies the same formulas as the tail of the sweep (used with mu and sigma from two
    different models by FindNextPoints, src/acquisition-function.cpp:63-110)."""
    kt, D, N = S.MATERN, 5, 60
    X, theta = S.make_X(N, D, "sls"), S.make_theta(D, "perturbed")
    ctx.fit(X, kt, theta, 0.005, S.make_y(X))
    f_best, _ = ctx.f_best()
    Q = np.concatenate([S.make_queries(300, D), X[:, :2]], axis=1)   # includes sigma ~ 0 points
    mu, sigma, dmu, dsigma = ctx.posterior_batch(Q)
    for acq, beta in ((0, 1.0), (1, 2.5)):
        val, grad = ctx.acq_batch(acq, beta, Q)
        v2, g2 = ctx.acq_from_posterior(acq, beta, f_best, mu, sigma, dmu, dsigma)
        np.testing.assert_allclose(v2, val, rtol=1e-12, atol=1e-300)
        np.testing.assert_allclose(g2, grad, rtol=1e-10, atol=1e-14 * np.max(np.abs(grad)))
        v3, g3 = ctx.acq_from_posterior(acq, beta, f_best, mu, sigma)
        np.testing.assert_array_equal(v3, v2)
        assert g3 is None


def test_sweep_shards_are_consistent(ctx, oracle):
    """M larger than one internal shard (16384): results must not depend on the sharding."""
    kt, D, N = S.SE, 6, 100
    X, theta = S.make_X(N, D, "uniform"), S.make_theta(D, "default")
    ctx.fit(X, kt, theta, 0.005, S.make_y(X))
    Q = S.make_queries(40000, D)
    val, grad = ctx.acq_batch(0, 1.0, Q)
    pick = np.array([0, 16383, 16384, 16385, 32768, 39999])
    v2, g2 = ctx.acq_batch(0, 1.0, Q[:, pick])
    np.testing.assert_array_equal(v2, val[pick])
    np.testing.assert_array_equal(g2, grad[:, pick])


def test_argmax_matches_explicit_sweep(ctx):
    kt, D, N = S.SE, 6, 100
    X, theta = S.make_X(N, D, "uniform"), S.make_theta(D, "default")
    ctx.fit(X, kt, theta, 0.005, S.make_y(X))
    seed, first, count = 1234, 1000, 50000
    Q = ctx.candidates(seed, first, count)
    assert Q.min() >= 0.0 and Q.max() < 1.0
    for acq, beta in ((0, 1.0), (1, 2.0)):
        val, _ = ctx.acq_batch(acq, beta, Q, grads=False)
        x, v, idx, g = ctx.acq_argmax(acq, beta, seed, first, count, want_grad=True)
        assert idx == first + int(np.argmax(val)) and v == val.max()
        np.testing.assert_array_equal(x, Q[:, idx - first])
        _, g_ref = ctx.acq_batch(acq, beta, x[:, None])
        np.testing.assert_array_equal(g, g_ref[:, 0])
        # splitting the range (what two GPUs would do) finds the same winner
        a = ctx.acq_argmax(acq, beta, seed, first, count // 2)
        b = ctx.acq_argmax(acq, beta, seed, first + count // 2, count - count // 2)
        win = a if (a[1] > b[1] or (a[1] == b[1] and a[2] < b[2])) else b
        assert win[2] == idx and win[1] == v


# N <= 64 with D <= 32 takes the single-launch small-model kernel (csrc/small.cuh); (SE, 40, 50) has too many dimensions for it
@pytest.mark.parametrize("kt,D,N", [(S.SE, 6, 30), (S.MATERN, 16, 200), (S.SE, 8, 65), (S.MATERN, 7, 33), (S.SE, 32, 64), (S.MATERN, 3, 3),
                                    (S.SE, 40, 50)])
@pytest.mark.parametrize("use_map", [False, True])
def test_map_objectives(ctx, oracle, kt, D, N, use_map):
    X = S.make_X(N, D, "sls")
    offsets, idx = S.make_tuples(X)
    a, r, b, var, btl = 0.5, 0.5, 0.005, 0.25, 0.01
    rng = np.random.default_rng(3)
    y = 0.05 * rng.standard_normal(N)
    theta = S.make_theta(D, "perturbed")
    ctx.set_data(X)
    ctx.set_preferences(offsets, idx)
    if not use_map:
        ctx.gram(kt, np.concatenate([[a], np.full(D, r)]), b, want=False)
        ctx.factor()
    pts = [np.concatenate([y, [theta[0], 0.007], theta[1:]]) if use_map else y,
           np.concatenate([np.zeros(N), [a, b], np.full(D, r)]) if use_map else np.zeros(N)]
    for x in pts:
        f_o, g_o = oracle.map_objective_pref(kt, X, offsets, idx, use_map, a, r, b, var, btl, x)
        f, g = ctx.map_objective_pref(kt, x, use_map, a, r, b, var, btl)
        check("f", f, f_o, 1e-10)
        check("grad", g, g_o)
    yv = S.make_y(X)
    for x in (np.concatenate([[0.5, 1e-4], np.full(D, 0.5)]), np.concatenate([[0.8, 0.01], theta[1:]])):
        f_o, g_o = oracle.map_objective_gpr(kt, X, yv, x)
        f, g = ctx.map_objective_gpr(kt, yv, x)
        check("gpr f", f, f_o, 1e-10)
        check("gpr grad", g, g_o)


def test_compat_flag_switches_the_se_gradient_quirk(ctx, slsb):
    kt, D, N = S.SE, 5, 40
    X, theta = S.make_X(N, D, "uniform"), S.make_theta(D, "default")
    ctx.fit(X, kt, theta, 0.005, S.make_y(X))
    Q = S.make_queries(10, D)
    _, _, dmu2, _ = ctx.posterior_batch(Q)
    ctx.set_compat_flags(0)
    _, _, dmu1, _ = ctx.posterior_batch(Q)
    ctx.set_compat_flags(slsb.COMPAT_SE_XGRAD_2X)
    np.testing.assert_allclose(dmu2, 2.0 * dmu1, rtol=1e-14)
    eps = 1e-6  # the flag-off gradient is the analytic one
    mu_p, *_ = ctx.posterior_batch(Q + eps * np.eye(D)[:, :1], grads=False)
    mu_m, *_ = ctx.posterior_batch(Q - eps * np.eye(D)[:, :1], grads=False)
    np.testing.assert_allclose(dmu1[0], (mu_p - mu_m) / (2 * eps), rtol=1e-5, atol=1e-9)


def test_error_reporting(ctx, slsb):
    c = slsb.Context(0)
    X = S.make_X(10, 3)
    with pytest.raises(slsb.SlsgpError) as e:
        c.gram(0, np.ones(4), 0.1)
    assert e.value.status == slsb.ERR_STATE
    c.set_data(X)
    with pytest.raises(slsb.SlsgpError) as e:
        c.factor()
    assert e.value.status == slsb.ERR_STATE
    bad = X.copy()
    bad[0, 0] = np.nan
    with pytest.raises(slsb.SlsgpError) as e:
        c.set_data(bad)
    assert e.value.status == slsb.ERR_NAN
    # duplicated points with zero noise: K is singular -> the factorisation must say so, not return garbage
    Xd = np.asfortranarray(np.repeat(X[:, :1], 10, axis=1))
    c.set_data(Xd)
    c.gram(0, np.concatenate([[0.5], np.full(3, 0.5)]), -1e-3, want=False)
    with pytest.raises(slsb.SlsgpError) as e:
        c.factor()
    assert e.value.status == slsb.ERR_NOT_SPD
    with pytest.raises(slsb.SlsgpError) as e:
        c.acq_batch(0, 1.0, S.make_queries(4, 3))
    assert e.value.status == slsb.ERR_STATE
    c.close()


def test_full_size_properties(ctx):
    """BASELINE config sizes (N=2048, D=16): properties that do not need the O(N^3)-per-point oracle."""
    kt, D, N = S.SE, 16, 2048
    X, theta = S.make_X(N, D, "uniform"), S.make_theta(D, "default")
    y = S.make_y(X)
    noise = 0.005
    ctx.set_data(X)
    K = ctx.gram(kt, theta, noise)
    d = (X[:, :300, None] - X[:, None, :300]) / theta[1:, None, None]
    check("K block", K[:300, :300], 0.5 * np.exp(-0.5 * (d ** 2).sum(0)) + noise * np.eye(300), 1e-13)
    logdet, L = ctx.factor(want_L=True)
    check("L L^T = K", L @ L.T, K, 1e-12)
    check("logdet", logdet, np.linalg.slogdet(K)[1], 1e-10)
    Kinv = ctx.inverse()
    check("K Kinv = I", K @ Kinv, np.eye(N), 1e-9, atol=1e-9)
    alpha = ctx.solve_alpha(y)
    check("K alpha = y", K @ alpha, y, 1e-9)
    # posterior at the data points: mu(X) = K_f alpha, sigma^2(X_i) = a - k_i^T Kinv k_i
    mu, sigma, _, _ = ctx.posterior_batch(X[:, :500])
    Kf = K - noise * np.eye(N)
    check("mu at data", mu, (Kf @ alpha)[:500], 1e-9)
    s2 = theta[0] - np.einsum("ij,ij->j", Kf[:, :500], Kinv @ Kf[:, :500])
    check("sigma at data", sigma, np.sqrt(np.maximum(s2, 0)), 1e-7)
    # EI gradient vs central differences of EI (the SE quirk doubles dmu and dsigma, so switch it off here)
    ctx.set_compat_flags(0)
    Q = S.make_queries(8, D)
    val, grad = ctx.acq_batch(1, 2.0, Q)
    eps = 1e-6
    for d_ in range(3):
        e = np.zeros((D, 1))
        e[d_] = eps
        vp, _ = ctx.acq_batch(1, 2.0, Q + e, grads=False)
        vm, _ = ctx.acq_batch(1, 2.0, Q - e, grads=False)
        np.testing.assert_allclose(grad[d_], (vp - vm) / (2 * eps), rtol=2e-5, atol=1e-8)
    ctx.set_compat_flags(1)


def test_numpy_candidate_generator_matches_the_library(ctx, slsb):
    """sharding.candidate_coords (used by the CPU multi-rank tests) is the same sequence as slsgp_candidates."""
    X = S.make_X(10, 7, "uniform")
    ctx.set_data(X)
    for seed, first, count in ((0, 0, 100), (1234, 10**9, 257), (2**63 + 5, 3, 64)):
        np.testing.assert_array_equal(ctx.candidates(seed, first, count), slsb.sharding.candidate_coords(seed, first, count, 7))


@pytest.mark.parametrize("kt,acq", [(S.SE, 0), (S.MATERN, 0), (S.SE, 1)])
def test_device_maximiser_improves_on_the_sweep_and_reaches_a_stationary_point(ctx, oracle, kt, acq):
    """slsgp_acq_maximize: sweep, best candidate per slice, batched projected ascent (FindGlobalSolution,
    src/acquisition-function.cpp:112-167). Checked against the oracle's value / gradient at the returned point."""
    D, N = 5, 60
    X, theta = S.make_X(N, D), S.make_theta(D, "perturbed")
    y = S.make_y(X)
    ctx.fit(X, kt, theta, 0.005, y)
    seed, first, count = 99, 0, 40000
    _, v_sweep, _, _ = ctx.acq_argmax(acq, 1.0, seed, first, count)
    x, v, g, v_slices = ctx.acq_maximize(acq, 1.0, seed, first, count, n_starts=256, n_iters=60)
    assert v_slices == v_sweep                       # the slice winners contain the overall sweep winner
    assert v >= v_sweep and np.all(x >= 0) and np.all(x <= 1)
    m = oracle.model(kt, X, theta, 0.005, y)
    _, f_best = oracle.f_best(m)
    v_o, g_o = oracle.acq(m, acq, 1.0, f_best, x)
    assert abs(v - v_o) <= 1e-9 * max(abs(v_o), 1e-30) + 1e-15
    g0 = np.max(np.abs(ctx.acq_batch(acq, 1.0, ctx.candidates(seed, first, 256))[1]))   # gradient scale away from optima
    np.testing.assert_allclose(g, g_o, rtol=1e-6, atol=1e-9 * g0)   # at a stationary point the gradient is all cancellation
    # first-order optimality on the box: interior coordinates have (nearly) vanished gradients
    pg = np.where(((x <= 0) & (g_o < 0)) | ((x >= 1) & (g_o > 0)), 0.0, g_o)
    assert np.max(np.abs(pg)) < 0.05 * g0, (pg, g0)
    # no iterations == the sweep winner; a split range (two GPUs) gives the same slice winners' maximum
    x0, v0, _, _ = ctx.acq_maximize(acq, 1.0, seed, first, count, n_starts=256, n_iters=0)
    assert v0 == v_sweep


@pytest.mark.parametrize("kt,D,N", [(S.SE, 4, 30), (S.MATERN, 8, 130), (S.SE, 6, 1024), (S.SE, 6, 1100)])   # fused kernel up to N = 1024
def test_whitened_map_objective_is_the_same_function(ctx, kt, D, N):
    """slsgp_map_objective_pref_whitened(z) == slsgp_map_objective_pref(y = L z, fixed hyper-parameters); its gradient is
    L^T grad_y; slsgp_whiten inverts y = L z."""
    X = S.make_X(N, D, "sls")
    offsets, idx = S.make_tuples(X)
    theta, b, btl = S.make_theta(D, "perturbed"), 0.005, 0.01
    ctx.set_data(X)
    ctx.gram(kt, theta, b, want=False)
    _, L = ctx.factor(want_L=True)
    ctx.set_preferences(offsets, idx)
    rng = np.random.default_rng(6)
    z = 0.05 * rng.standard_normal(N)
    f_w, g_w, y = ctx.map_objective_pref_whitened(z, btl)
    np.testing.assert_allclose(y, L @ z, rtol=1e-12, atol=1e-15)
    f, g_y = ctx.map_objective_pref(kt, y, False, 0.5, 0.5, b, 0.25, btl)
    assert abs(f_w - f) <= 1e-10 * abs(f)
    np.testing.assert_allclose(g_w, L.T @ g_y, rtol=1e-7, atol=1e-9 * np.max(np.abs(g_w)))
    np.testing.assert_allclose(ctx.whiten(y), z, rtol=1e-8, atol=1e-12)
    f_only, g_none, _ = ctx.map_objective_pref_whitened(z, btl, want_grad=False)
    assert f_only == f_w and g_none is None


def test_new_entry_points_validate_their_arguments(ctx, slsb):
    X = S.make_X(20, 3)
    c = slsb.Context(0)
    try:
        c.set_data(X)
        with pytest.raises(slsb.SlsgpError) as e:          # no factor yet
            c.map_objective_pref_whitened(np.zeros(20), 0.01)
        assert e.value.status == slsb.ERR_STATE
        with pytest.raises(slsb.SlsgpError) as e:
            c.whiten(np.zeros(20))
        assert e.value.status == slsb.ERR_STATE
        with pytest.raises(slsb.SlsgpError) as e:          # no model yet
            c.acq_maximize(0, 1.0, 1, 0, 100)
        assert e.value.status == slsb.ERR_STATE
        c.fit(X, S.SE, S.make_theta(3), 0.005, S.make_y(X))
        with pytest.raises(slsb.SlsgpError) as e:
            c.acq_maximize(0, 1.0, 1, 0, 0)                # empty range
        assert e.value.status == slsb.ERR_INVALID
        with pytest.raises(slsb.SlsgpError) as e:
            c.map_objective_pref_whitened(np.full(20, np.nan), 0.01)
        assert e.value.status == slsb.ERR_NAN
        # fewer candidates than requested starts: one start per candidate, still the sweep winner or better
        _, v_sweep, _, _ = c.acq_argmax(0, 1.0, 5, 0, 7)
        x, v, g, v0 = c.acq_maximize(0, 1.0, 5, 0, 7, n_starts=1024, n_iters=10)
        assert v0 == v_sweep and v >= v_sweep and x.shape == (3,)
    finally:
        c.close()


@pytest.mark.parametrize("kt,D,N0,n_add", [(S.SE, 5, 40, 6), (S.MATERN, 8, 61, 8), (S.SE, 16, 190, 5), (S.MATERN, 3, 1, 3)])
def test_append_point_equals_a_fresh_fit(ctx, oracle, kt, D, N0, n_add):
    """slsgp_append_point (bordered O(N^2) update; crosses the 64-row padding at 61 + 8 and 190 + 5) against the
    oracle's from-scratch model of the grown data set after every appended point."""
    N = N0 + n_add
    X, theta = S.make_X(N, D, "sls"), S.make_theta(D, "perturbed")
    y, noise = S.make_y(X), 0.005
    ctx.fit(X[:, :N0], kt, theta, noise, y[:N0])
    Q = S.make_queries(40, D)
    for n in range(N0, N):
        kcol, Kinv = ctx.append_point(X[:, n], y[n])
        K_o = oracle.large_ky(kt, X[:, :n + 1], theta, noise)
        check("K column", kcol, K_o[:, n], 1e-13)
        check("Kinv", Kinv, oracle.inverse(K_o), 1e-7)
        assert np.array_equal(Kinv, Kinv.T)
        m = oracle.model(kt, X[:, :n + 1], theta, noise, y[:n + 1])
        i_best, f_best = oracle.f_best(m)
        f, i = ctx.f_best()
        assert i == i_best
        check("f_best", f, f_best, 1e-9)
        mu, sigma, dmu, dsg = ctx.posterior_batch(Q)
        want = [oracle.predict(m, Q[:, j]) for j in range(Q.shape[1])]
        check("mu", mu, [w[0] for w in want])
        check("sigma", sigma, [w[1] for w in want], RT, atol=1e-9)
    # the grown state behaves like a fresh fit of all N points
    val_inc, grad_inc = ctx.acq_batch(0, 0.0, Q)
    ctx.fit(X, kt, theta, noise, y)
    val_new, grad_new = ctx.acq_batch(0, 0.0, Q)
    check("EI after appends", val_inc, val_new, 1e-7, atol=1e-12)
    check("grad EI after appends", grad_inc, grad_new, 1e-6, atol=1e-10)


def test_append_point_failure_keeps_the_model(ctx):
    """An exact duplicate with zero noise makes the bordered matrix singular: the Schur complement is zero up to rounding.
    Whether rounding leaves it positive or not, a refused append must leave the model as it was."""
    X, theta = S.make_X(20, 4, "uniform"), S.make_theta(4, "default")
    y = S.make_y(X)
    ctx.fit(X, S.SE, theta, 0.0, y)
    Q = S.make_queries(10, 4)
    before = ctx.posterior_batch(Q)
    try:
        ctx.append_point(X[:, 3], 0.1)
    except Exception as e:
        assert "Schur" in str(e) and ctx.N == 20
        for a, b in zip(before, ctx.posterior_batch(Q)):
            assert np.array_equal(a, b)
    else:
        assert ctx.N == 21
    with pytest.raises(Exception):
        ctx.append_point(np.full(4, np.nan), 0.1)
    with pytest.raises(ValueError):
        ctx.append_point(np.zeros(3), 0.1)


@pytest.mark.parametrize("kt,D,N", [(S.SE, 5, 1), (S.MATERN, 7, 33), (S.SE, 32, 64)])
def test_small_model_kernel_leaves_the_state_of_the_general_path(ctx, oracle, kt, D, N):
    """After slsgp_map_objective_gpr on a small model (one fused launch) the context must hold what slsgp_gram + slsgp_factor +
    slsgp_inverse + slsgp_solve_alpha produce: same matrices, f_best and predictions."""
    X, theta = S.make_X(N, D, "sls"), S.make_theta(D, "perturbed")
    y, noise = S.make_y(X), 0.007
    Q = S.make_queries(30, D)
    ctx.set_data(X)
    ctx.map_objective_gpr(kt, y, np.concatenate([[theta[0], noise], theta[1:]]))
    Kinv_f = ctx.inverse()                                  # stored by the fused kernel (has_inverse is set)
    fb_f = ctx.f_best()
    post_f = ctx.posterior_batch(Q)
    ctx.set_data(X)
    K = ctx.gram(kt, theta, noise)
    logdet, L = ctx.factor(want_L=True)
    Kinv = ctx.inverse()
    ctx.solve_alpha(y)
    check("Kinv", Kinv_f, Kinv, 1e-9)
    assert np.array_equal(Kinv_f, Kinv_f.T)
    assert fb_f[1] == ctx.f_best()[1]
    check("f_best", fb_f[0], ctx.f_best()[0], 1e-10)
    for name, a, b in zip(("mu", "sigma", "dmu", "dsigma"), post_f, ctx.posterior_batch(Q)):
        check(name, a, b, 1e-8, atol=1e-11)
    m = oracle.model(kt, X, theta, noise, y)
    check("Kinv vs oracle", Kinv_f, oracle.inverse(oracle.large_ky(kt, X, theta, noise)), 1e-7)


def test_small_model_kernel_reports_a_matrix_that_is_not_spd(ctx, slsb):
    X = np.tile(np.linspace(0.1, 0.9, 4)[:, None], (1, 6))   # six identical points, no noise: singular
    ctx.set_data(X)
    with pytest.raises(slsb.SlsgpError) as e:
        ctx.map_objective_gpr(S.SE, np.zeros(6), np.array([0.5, 0.0, 0.5, 0.5, 0.5, 0.5]))
    assert "SPD" in str(e.value)


def test_general_map_path_reports_a_matrix_that_is_not_spd_and_recovers(ctx, slsb, oracle):
    """The general-path MAP objectives (N > 64) synchronise once per evaluation: the pivot check of the factorisation is collected
    with the results (do_factor(defer_check)). A singular K_y must still be reported, and the next evaluation on the same context
    must be a clean one."""
    N, D = 80, 4
    X = np.tile(np.linspace(0.1, 0.9, D)[:, None], (1, N))   # identical points, (almost) no noise: singular
    ctx.set_data(X)
    hyper = np.concatenate([[0.5, 0.0], np.full(D, 0.5)])
    with pytest.raises(slsb.SlsgpError) as e:
        ctx.map_objective_gpr(S.SE, np.zeros(N), hyper)
    assert "SPD" in str(e.value) and e.value.status == slsb.ERR_NOT_SPD
    off, idx = S.make_tuples(X)
    ctx.set_preferences(off, idx)
    with pytest.raises(slsb.SlsgpError) as e:
        ctx.map_objective_pref(S.SE, np.concatenate([np.zeros(N), hyper]), True, 0.5, 0.5, 0.005, 0.25, 0.01)
    assert e.value.status == slsb.ERR_NOT_SPD
    with pytest.raises(slsb.SlsgpError):
        ctx.acq_batch(0, 1.0, S.make_queries(4, D))  # no model was left behind
    X = S.make_X(N, D, "uniform")
    y = S.make_y(X)
    hyper[1] = 0.005
    ctx.set_data(X)
    f, g = ctx.map_objective_gpr(S.SE, y, hyper)
    want_f, want_g = oracle.map_objective_gpr(S.SE, X, y, hyper)
    assert abs(f - want_f) <= 1e-9 * abs(want_f)
    check("gradient", g, want_g, 1e-7)
