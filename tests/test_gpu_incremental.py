"""SURVEY.md 8(f) rank 3, incremental refit across iterations: slsgp_set_data_extend grows a factored model by the data
points AddNewPoints appended (bordered update of K_y, L, L^-1, K_y^-1) instead of rebuilding it, a context never recomputes a
matrix it already holds, and with SetIncrementalRefit(true) the optimisers hand the previous regressor's model to the next one.
Checked against from-scratch rebuilds of the same library (which the other parity tests tie to the oracle and the reference)."""
import importlib

import numpy as np
import pytest

import loop_support as LS
import support as S

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def slsb():
    return importlib.import_module("sequential-line-search_b200")


@pytest.fixture()
def ctx(slsb):
    c = slsb.Context(0)
    yield c
    c.close()


def rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b)))) / max(float(np.max(np.abs(b))), 1e-300)


def full_model(c, X, kt, theta, noise, y, Q):
    """K, L, K^-1 and the posterior at Q of whatever model the context holds after gram / factor / inverse / solve_alpha."""
    K = c.gram(kt, theta, noise)
    _, L = c.factor(want_L=True)
    Kinv = c.inverse()
    c.solve_alpha(y)
    return K, L, Kinv, c.posterior_batch(Q)


@pytest.mark.parametrize("kt", [S.SE, S.MATERN])
@pytest.mark.parametrize("N0,extra", [(5, 3), (40, 2), (100, 3), (300, 3), (64, 0)])
def test_extended_model_equals_rebuilt_model(slsb, ctx, kt, N0, extra):
    D, noise = 6, 0.005
    X = S.make_X(N0 + extra, D, "sls")
    theta = S.make_theta(D, "perturbed")
    y = S.make_y(X)
    Q = S.make_queries(50, D)
    ctx.fit(X[:, :N0], kt, theta, noise, y[:N0])
    kept = ctx.set_data_extend(X)
    # the padded size of the model (multiples of 64) bounds what can be grown in place; beyond it the data are replaced
    assert kept == (N0 if (N0 + extra) <= -(-N0 // 64) * 64 else 0)
    got = full_model(ctx, X, kt, theta, noise, y, Q)
    fresh = slsb.Context(0)
    try:
        fresh.set_data(X)
        want = full_model(fresh, X, kt, theta, noise, y, Q)
    finally:
        fresh.close()
    assert rel(got[0], want[0]) < 1e-14, "K_y"  # the appended column comes from another kernel (libm exp) than the Gram tiles
    assert rel(got[1], want[1]) < 1e-11, "L"
    assert rel(got[2], want[2]) < 1e-9, "K^-1"  # cond(K) ~ 1e3-1e4 on the clustered data
    for g, w, name in zip(got[3], want[3], ("mu", "sigma", "dmu", "dsigma")):
        assert rel(g, w) < 1e-8, name


@pytest.mark.parametrize("N0,gone", [(30, 26), (30, 29), (130, 125), (12, 11)])
def test_extend_after_a_merge_of_coincident_points(slsb, ctx, N0, gone):
    """What AddNewPoints does when the new slider end coincides with a data point (src/preference-data-manager.cpp:14-141): the
    old point `gone` and its duplicate leave, their midpoint and the other new points are appended. The model of the first `gone`
    points is kept, everything behind it is re-appended."""
    D, noise, kt = 5, 0.005, S.MATERN
    Xall = S.make_X(N0 + 3, D, "sls")
    theta, Q = S.make_theta(D, "default"), S.make_queries(40, D)
    X_old = Xall[:, :N0]
    mid = 0.5 * (X_old[:, gone] + (X_old[:, gone] + 1e-6))
    X_new = S.f64(np.concatenate([np.delete(X_old, gone, axis=1), Xall[:, N0:N0 + 2], mid[:, None]], axis=1))
    y = S.make_y(X_new)
    ctx.fit(X_old, kt, theta, noise, S.make_y(X_old))
    assert ctx.set_data_extend(X_new) == gone
    got = full_model(ctx, X_new, kt, theta, noise, y, Q)
    fresh = slsb.Context(0)
    try:
        fresh.set_data(X_new)
        want = full_model(fresh, X_new, kt, theta, noise, y, Q)
    finally:
        fresh.close()
    assert rel(got[0], want[0]) < 1e-14, "K_y"
    assert rel(got[1], want[1]) < 1e-11, "L"
    assert rel(got[2], want[2]) < 1e-9, "K^-1"
    for g, w, name in zip(got[3], want[3], ("mu", "sigma", "dmu", "dsigma")):
        assert rel(g, w) < 1e-8, name
    # more changed columns than bordered updates are worth: replaced
    assert ctx.set_data_extend(S.f64(np.concatenate([X_new[:, :5], Xall[:, ::-1][:, :20]], axis=1))) == 0


def test_extend_falls_back_when_the_prefix_or_the_hyperparameters_differ(slsb, ctx):
    D, N0, noise = 5, 30, 0.005
    X = S.make_X(N0 + 3, D, "uniform")
    theta, y, Q = S.make_theta(D, "default"), S.make_y(X), S.make_queries(20, D)
    ctx.fit(X[:, :N0], S.SE, theta, noise, y[:N0])
    X2 = X.copy()
    X2[2, 1] += 1e-9  # an early data point moved: too much to re-append, the model is rebuilt
    assert ctx.set_data_extend(X2) == 0
    fresh = slsb.Context(0)
    try:
        fresh.set_data(X2)
        want = full_model(fresh, X2, S.SE, theta, noise, y, Q)
        got = full_model(ctx, X2, S.SE, theta, noise, y, Q)
        for g, w in zip(got[:3], want[:3]):
            np.testing.assert_array_equal(g, w)
        # kept data, other hyper-parameters at the next slsgp_gram: K_y is recomputed, not taken from the extended model
        ctx.fit(X[:, :N0], S.SE, theta, noise, y[:N0])
        assert ctx.set_data_extend(X) == N0
        theta2 = theta * 1.1
        fresh.set_data(X)
        want = full_model(fresh, X, S.SE, theta2, noise, y, Q)
        got = full_model(ctx, X, S.SE, theta2, noise, y, Q)
        for g, w in zip(got[:3], want[:3]):
            np.testing.assert_array_equal(g, w)
    finally:
        fresh.close()


def test_a_context_does_not_recompute_what_it_holds(ctx):
    D, N = 8, 200
    X, theta = S.make_X(N, D, "uniform"), S.make_theta(D, "default")
    ctx.set_data(X)
    K = ctx.gram(S.SE, theta, 0.005)
    _, L = ctx.factor(want_L=True)
    n0 = ctx.launch_count()
    K2 = ctx.gram(S.SE, theta, 0.005)
    _, L2 = ctx.factor(want_L=True)
    assert ctx.launch_count() == n0, "same hyper-parameters: K_y and its factor are answered from the context"
    np.testing.assert_array_equal(K2, K)
    np.testing.assert_array_equal(L2, L)
    ctx.gram(S.SE, theta, 0.006)
    assert ctx.launch_count() > n0
    ctx.factor()
    ctx.set_data(X)  # new data (even identical): everything is rebuilt
    n1 = ctx.launch_count()
    ctx.gram(S.SE, theta, 0.006)
    assert ctx.launch_count() > n1


def test_sequential_line_search_with_incremental_refit(slsb):
    """SequentialLineSearchOptimizer with fixed hyper-parameters, the simulated user of the nd demo, SetIncrementalRefit(true): the
    regressor of iteration i keeps the factored model of iteration i - 1 up to the data point AddNewPoints merged away, and its
    posterior equals that of a deep copy of itself, which rebuilds the model from scratch out of the same data and goodness
    values. (The slider ends of an incremental and of a rebuilding RUN cannot be compared beyond the first iterations: the search
    picks among near-equal local maxima of EI and last-bit differences of the factor flip the choice.)"""
    hl = importlib.import_module("sequential-line-search_b200.hostlib")
    L = LS.LoopLib("b200")
    L.lib.b200_sls_model_vs_rebuilt_copy.restype = LS.C.c_double
    D, iters = 5, 12
    Q = S.f64(S.make_queries(24, D))
    before = hl.get_incremental_refit()
    kept, npts, worst = [], [], []
    try:
        hl.set_incremental_refit(True)
        L.srand(11)
        opt = L.sls(D, True, False, LS.MATERN, LS.EI)
        for _ in range(iters):
            e0, e1 = opt.slider_ends()
            opt.submit(LS.best_slider_position(e0, e1))
            kept.append(L.lib.b200_sls_num_points_kept(LS.C.c_void_p(opt.h)))
            npts.append(opt.num_points())
            worst.append(L.lib.b200_sls_model_vs_rebuilt_copy(LS.C.c_void_p(opt.h), D, Q.shape[1], Q.ctypes.data_as(LS.c_dp)))
        f_inc = LS.demo_objective(opt.maximizer())
        opt.close()
        hl.set_incremental_refit(False)
        L.srand(11)
        opt = L.sls(D, True, False, LS.MATERN, LS.EI)
        for _ in range(iters):
            e0, e1 = opt.slider_ends()
            opt.submit(LS.best_slider_position(e0, e1))
            assert L.lib.b200_sls_num_points_kept(LS.C.c_void_p(opt.h)) == 0
        f_off = LS.demo_objective(opt.maximizer())
        opt.close()
    finally:
        hl.set_incremental_refit(before)
    print("\ndata points", npts, "of which kept from the previous model", kept)
    print("extended model vs rebuilt copy, worst |difference| of the posterior per iteration:", " ".join(f"{w:.1e}" for w in worst))
    print(f"objective reached after {iters} iterations: incremental {f_inc:.4f}, rebuilding {f_off:.4f}")
    assert kept[0] == 0
    for i in range(1, iters):
        assert 0 <= kept[i] <= npts[i - 1], (i, kept, npts)
    assert sum(k > 0 for k in kept) >= iters - 3
    assert max(worst) < 1e-8
