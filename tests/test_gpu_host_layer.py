"""The C++ host layer (sequential_line_search::GaussianProcessRegressor / PreferenceRegressor / acquisition_func on
libslsgp) against the reference's own classes (oracle/_ref, compiled from the unmodified sources), call for call.

Tolerance: 1e-5 relative (north_star FP64), measured as max |err| / max |ref| per quantity.
The MAP fits and searches of the Native driver are checked as what they claim to be: maximisers of the REFERENCE's objective
(given an evaluation budget that lets them converge). Parity with what the reference's own NLopt runs return is the subject of
tests/test_gpu_step_parity.py."""
import importlib
import os

import numpy as np
import pytest

import support as S

pkg = importlib.import_module("sequential-line-search_b200")
pytestmark = pytest.mark.gpu
RT = 1e-5


@pytest.fixture(scope="module")
def host():
    return pkg.hostlib.Host()


@pytest.fixture(autouse=True)
def native_driver():
    """This file exercises the host layer's OWN search drivers (SearchDriver::Native: the whitened quasi-Newton MAP fit, the
    device-resident acquisition maximiser). The NLopt-driven reference-faithful path is compared with the real reference in
    tests/test_gpu_step_parity.py."""
    previous = pkg.hostlib.get_search_driver()
    pkg.hostlib.set_search_driver(pkg.hostlib.NATIVE)
    yield
    pkg.hostlib.set_search_driver(previous)


CASES = [(S.SE, 6, 40, "uniform", "default"), (S.MATERN, 6, 40, "sls", "perturbed"), (S.SE, 16, 130, "sls", "perturbed"),
         (S.MATERN, 3, 1, "uniform", "default")]


@pytest.mark.parametrize("kt,D,N,kind,thk", CASES)
def test_gaussian_process_regressor_matches_reference(host, ref, kt, D, N, kind, thk):
    X, theta, b = S.make_X(N, D, kind), S.make_theta(D, thk), 0.005
    y = S.make_y(X)
    h, hr = host.gpr_create(kt, X, y, theta, b), ref.gpr_create(kt, X, y, theta, b)
    try:
        st = host.gpr_state(h, N, D)
        K_r, Kinv_r = ref.gpr_state(hr, N)
        assert S.rel_err(st["K"], K_r) < 1e-12          # public member m_K_y
        assert S.rel_err(st["Kinv"], Kinv_r) < RT       # public member m_K_y_inv
        np.testing.assert_array_equal(st["theta"], theta)
        reg, reg_r = host.gpr_regressor(h), ref.gpr_regressor(hr)
        np.testing.assert_array_equal(host.x_best(reg, D), ref.x_best(reg_r, D))
        Q = S.make_queries(12, D)
        got = [host.predict(reg, Q[:, m]) for m in range(Q.shape[1])]
        want = [ref.predict(reg_r, Q[:, m]) for m in range(Q.shape[1])]
        for i in range(4):
            assert S.rel_err([g[i] for g in got], [w[i] for w in want]) < RT, i
        # the batched form returns what the one-candidate calls return
        mu, sg, dmu, dsg = host.predict_batch(reg, Q)
        assert S.rel_err(mu, [w[0] for w in want]) < RT and S.rel_err(sg, [w[1] for w in want]) < RT
        assert S.rel_err(dmu.T, [w[2] for w in want]) < RT and S.rel_err(dsg.T, [w[3] for w in want]) < RT
        for acq, beta in ((S.EI, 1.0), (S.UCB, 2.5)):
            vals, grads = host.acq_values(reg, acq, beta, Q)
            want_v, want_g = zip(*[ref.acq(reg_r, acq, beta, Q[:, m]) for m in range(Q.shape[1])])
            assert S.rel_err(vals, want_v) < RT and S.rel_err(grads.T, want_g) < RT
            v1, g1 = host.acq(reg, acq, beta, Q[:, 0])   # CalcAcquisitionValue{,Derivative}, single point
            assert abs(v1 - want_v[0]) <= RT * max(np.max(np.abs(want_v)), 1e-300)
            assert S.rel_err(g1, want_g[0]) < 1e-4 or np.max(np.abs(g1 - want_g[0])) < RT * np.max(np.abs(want_g))
    finally:
        host.gpr_destroy(h)
        ref.gpr_destroy(hr)


@pytest.mark.parametrize("kt,D,N0,n_add", [(S.SE, 6, 40, 4), (S.MATERN, 5, 62, 5)])
def test_append_point_equals_the_reference_regressor_of_the_grown_data(host, ref, kt, D, N0, n_add):
    """GaussianProcessRegressor::AppendPoint (addition; the O(N^2) replacement of the rebuild at
    src/acquisition-function.cpp:281-296) against the reference regressor constructed on all the points."""
    N = N0 + n_add
    X, theta, b = S.make_X(N, D, "sls"), S.make_theta(D, "perturbed"), 0.005
    y = S.make_y(X)
    h, hr = host.gpr_create(kt, X[:, :N0], y[:N0], theta, b), ref.gpr_create(kt, X, y, theta, b)
    try:
        for n in range(N0, N):
            host.gpr_append_point(h, X[:, n], y[n])
        st = host.gpr_state(h, N, D)
        K_r, Kinv_r = ref.gpr_state(hr, N)
        assert S.rel_err(st["K"], K_r) < 1e-12 and S.rel_err(st["Kinv"], Kinv_r) < RT
        reg, reg_r = host.gpr_regressor(h), ref.gpr_regressor(hr)
        np.testing.assert_array_equal(host.x_best(reg, D), ref.x_best(reg_r, D))
        Q = S.make_queries(10, D)
        for m in range(Q.shape[1]):
            got, want = host.predict(reg, Q[:, m]), ref.predict(reg_r, Q[:, m])
            for i in range(4):
                assert np.max(np.abs(np.asarray(got[i]) - np.asarray(want[i]))) < RT * max(1.0, np.max(np.abs(want[i]))), (m, i)
    finally:
        host.gpr_destroy(h)
        ref.gpr_destroy(hr)


def test_calc_large_ky_free_function(host, ref):
    X, theta = S.make_X(33, 5, "sls"), S.make_theta(5, "perturbed")
    for kt in (S.SE, S.MATERN):
        assert S.rel_err(host.large_ky(kt, X, theta, 0.01), ref.large_ky(kt, X, theta, 0.01)) < 1e-12


def test_empty_regressor_has_zero_acquisition(host):
    """src/acquisition-function.cpp:176-179, 206-209: no data -> value 0, zero gradient."""
    h = host.gpr_create(S.SE, np.zeros((0, 0)), np.zeros(0), np.zeros(1), 0.005)
    try:
        reg = host.gpr_regressor(h)
        v = host.lib.b200_acq_value(reg, 3, S.EI, pkg.hostlib.C.c_double(1.0), pkg.hostlib._p(np.full(3, 0.5)))
        assert v == 0.0
    finally:
        host.gpr_destroy(h)


PREF_CASES = [(S.SE, 4, 30, False), (S.MATERN, 6, 45, False), (S.SE, 5, 36, True), (S.MATERN, 4, 24, True)]


@pytest.mark.parametrize("kt,D,N,use_map", PREF_CASES)
def test_preference_regressor_map_is_a_maximiser_of_the_reference_objective(host, ref, kt, D, N, use_map):
    X = S.make_X(N, D, "sls")
    offsets, idx = S.make_tuples(X)
    a, r, b, pv, btl = 0.5, 0.5, 0.005, 0.25, 0.01
    h = host.pref_create(kt, X, offsets, idx, use_map, a, r, b, pv, btl, num_iters=6000)  # a budget that lets the fit converge
    try:
        st = host.pref_state(h, N, D)
        sol = np.concatenate([st["y"], [st["theta"][0], st["b"]], st["theta"][1:]]) if use_map else st["y"]
        # the reference regressor put into OUR solution: same state, same predictions
        hr = ref.pref_create(kt, X, offsets, idx, use_map, a, r, b, pv, btl, sol)
        try:
            st_r = ref.pref_state(hr, N, D)
            assert S.rel_err(st["K"], st_r["K"]) < 1e-12 and S.rel_err(st["L"], st_r["L"]) < 1e-9
            f_r, g_r = ref.pref_objective(hr, sol)
            f_h, _ = host.pref_objective(h, sol)        # EvaluateMapObjective == the reference's NLopt callback
            assert abs(f_h - f_r) <= 1e-9 * abs(f_r)
            # gradient parity away from the optimum (at the optimum the gradient is pure cancellation)
            away = sol * (1.0 + 0.05 * np.random.default_rng(9).standard_normal(len(sol)))
            assert S.rel_err(host.pref_objective(h, away)[1], ref.pref_objective(hr, away)[1]) < 1e-6
            # first-order optimality of the reference objective at our solution, bounds respected
            lo = np.concatenate([np.full(N, -10.0), np.full(D + 2, 1e-8)]) if use_map else np.full(N, -10.0)  # the reference's box
            hi = np.full(len(sol), 10.0)
            pg = np.where(((sol <= lo * (1 + 1e-9)) & (g_r < 0)) | ((sol >= hi) & (g_r > 0)), 0.0, g_r)
            if use_map:   # hyper-parameters live on a log scale (b ~ 5e-3): d F / d log x = x dF/dx
                pg[N:] *= sol[N:]
            scale = max(1.0, abs(f_r))
            assert np.max(np.abs(pg)) < 2e-5 * scale, (np.max(np.abs(pg)), f_r)
            # no perturbed point does better under the reference objective
            rng = np.random.default_rng(3)
            for eps in (1e-3, 1e-2, 1e-1):
                for _ in range(4):
                    z = np.clip(sol + eps * rng.standard_normal(len(sol)) * np.maximum(np.abs(sol), 0.05), lo, hi)
                    assert ref.pref_objective(hr, z, want_grad=False)[0] <= f_r + 1e-9 * scale
            # and the zero initial point of the reference is worse
            x0 = np.concatenate([np.zeros(N), [a, b], np.full(D, r)]) if use_map else np.zeros(N)
            assert ref.pref_objective(hr, x0, want_grad=False)[0] < f_r
            reg, reg_r = host.pref_regressor(h), ref.pref_regressor(hr)
            Q = S.make_queries(6, D)
            for m in range(Q.shape[1]):
                got, want = host.predict(reg, Q[:, m]), ref.predict(reg_r, Q[:, m])
                assert abs(got[0] - want[0]) < RT * max(1.0, np.max(np.abs(st["y"])))
                assert abs(got[1] - want[1]) < RT
                assert S.rel_err(got[2], want[2]) < 1e-4 and S.rel_err(got[3], want[3]) < 1e-4
            np.testing.assert_array_equal(host.pref_find_arg_max(h, D), X[:, int(np.argmax(st["y"]))])
        finally:
            ref.pref_destroy(hr)
        assert 1 < host.pref_num_map_evaluations(h) <= 6000 + 700  # the budget, plus at most one inner solve started below it
    finally:
        host.pref_destroy(h)


def test_preference_regressor_damp_data(host, tmp_path):
    X = S.make_X(9, 3)
    offsets, idx = S.make_tuples(X)
    h = host.pref_create(S.SE, X, offsets, idx, False, 0.5, 0.5, 0.005, 0.25, 0.01)
    try:
        host.pref_damp_data(h, str(tmp_path), "t_")
        got = np.loadtxt(os.path.join(tmp_path, "t_X.csv"), delimiter=",")
        np.testing.assert_allclose(got, X, rtol=1e-5)
        rows = [list(map(int, line.split(","))) for line in open(os.path.join(tmp_path, "t_D.csv")).read().split()]
        assert rows == [list(idx[offsets[t]:offsets[t + 1]]) for t in range(len(offsets) - 1)]
    finally:
        host.pref_destroy(h)


@pytest.mark.parametrize("kt", [S.SE, S.MATERN])
def test_gpr_map_hyperparameters_maximise_the_reference_objective(host, ref, kt):
    D, N = 3, 40
    X = S.make_X(N, D)
    y = S.make_y(X)
    h = host.gpr_create(kt, X, y)                    # MAP estimation of (a, b, r)
    try:
        st = host.gpr_state(h, N, D)
        x = np.concatenate([[st["theta"][0], st["b"]], st["theta"][1:]])
        f, g = ref.gpr_objective(kt, X, y, x[None, :])
        lo, hi = 1e-8, 50.0
        pg = np.where(((x <= lo) & (g[0] < 0)) | ((x >= hi) & (g[0] > 0)), 0.0, g[0])
        assert np.max(np.abs(pg * np.maximum(x, 1e-3))) < 1e-3 * max(1.0, abs(f[0])), (pg, x, f)
        x0 = np.concatenate([[0.5, 1e-4], np.full(D, 0.5)])   # the reference's starting point (prior means)
        assert f[0] >= ref.gpr_objective(kt, X, y, x0[None, :], want_grad=False)[0][0]
    finally:
        host.gpr_destroy(h)


@pytest.mark.parametrize("kt,acq", [(S.SE, S.EI), (S.MATERN, S.EI), (S.SE, S.UCB)])
def test_find_next_point_beats_a_dense_random_search_of_the_reference_objective(host, ref, kt, acq):
    D, N = 4, 25
    X, theta = S.make_X(N, D), S.make_theta(D)
    y = S.make_y(X)
    h, hr = host.gpr_create(kt, X, y, theta, 0.005), ref.gpr_create(kt, X, y, theta, 0.005)
    try:
        reg, reg_r = host.gpr_regressor(h), ref.gpr_regressor(hr)
        x = host.find_next_point(reg, D, n_global=40, n_local=50, acq_type=acq, beta=1.0)
        assert x.shape == (D,) and np.all(x >= 0.0) and np.all(x <= 1.0)
        v = ref.acq(reg_r, acq, 1.0, x, want_grad=False)[0]
        Q = S.make_queries(1500, D, seed=11)
        best_random = max(ref.acq(reg_r, acq, 1.0, Q[:, m], want_grad=False)[0] for m in range(Q.shape[1]))
        assert v >= best_random, (v, best_random)
        # interior coordinates are stationary under the reference's own derivative
        g = ref.acq(reg_r, acq, 1.0, x)[1]
        interior = (x > 1e-9) & (x < 1 - 1e-9)
        assert np.all(np.abs(g[interior]) < 1e-3 * max(1.0, np.max(np.abs(g)))) or v > best_random
    finally:
        host.gpr_destroy(h)
        ref.gpr_destroy(hr)


def test_find_next_points_returns_distinct_points_that_lower_each_others_criterion(host):
    D, N = 3, 20
    X, theta = S.make_X(N, D), S.make_theta(D)
    h = host.gpr_create(S.SE, X, S.make_y(X), theta, 0.005)
    try:
        reg = host.gpr_regressor(h)
        P = host.find_next_points(reg, D, 3, n_global=20, n_local=30)
        assert P.shape == (D, 3) and np.all(P >= 0) and np.all(P <= 1)
        d = [np.linalg.norm(P[:, i] - P[:, j]) for i in range(3) for j in range(i)]
        assert min(d) > 1e-3           # the temporary regressor kills sigma at the points already chosen
        # the first point is the single-point maximiser: compare with FindNextPoint under the (Matern) pair criterion
        v_first = host.acq(reg, S.EI, 1.0, P[:, 0], want_grad=False)[0]
        assert v_first > 0
    finally:
        host.gpr_destroy(h)


def test_warm_started_map_reaches_the_same_optimum_in_fewer_evaluations(host):
    """the MAP fit of iteration t+1 starts from the goodness values of iteration t."""
    kt, D = S.SE, 5
    X = S.make_X(33, D, "sls")
    offsets, idx = S.make_tuples(X)
    args = (False, 0.5, 0.5, 0.005, 0.25, 0.01, 1000)
    prev = host.pref_create(kt, X[:, :30], offsets[:11], idx[:offsets[10]], *args)         # iteration t: 10 tuples
    cold = host.pref_create(kt, X, offsets, idx, *args)                                    # iteration t+1, cold start
    warm = host.pref_create(kt, X, offsets, idx, *args, warm_from=prev)                    # iteration t+1, warm start
    try:
        y_cold, y_warm = host.pref_state(cold, 33, D)["y"], host.pref_state(warm, 33, D)["y"]
        assert S.rel_err(y_warm, y_cold) < 1e-5                     # fixed hyper-parameters: the optimum is unique
        assert host.pref_num_map_evaluations(warm) < host.pref_num_map_evaluations(cold)
    finally:
        for h in (prev, cold, warm):
            host.pref_destroy(h)


def test_regressor_lifetimes_do_not_leak_device_memory(host):
    """The reference rebuilds its regressors every iteration (src/sequential-line-search.cpp:94-103); here each one borrows a
    device context from a pool. 300 lifetimes of alternating sizes must leave the free device memory where it was."""
    import torch
    rng = np.random.default_rng(0)

    def lifetime(i):
        N, D = (20, 4) if i % 2 else (150, 7)
        X = rng.random((D, N))
        h = host.gpr_create(S.SE, X, S.make_y(X), S.make_theta(D, "default"), 0.005)
        host.predict(host.gpr_regressor(h), X[:, 0])
        host.gpr_destroy(h)

    for i in range(20):          # fill the pool / reach the steady state
        lifetime(i)
    torch.cuda.synchronize()
    free0, _ = torch.cuda.mem_get_info(0)
    for i in range(300):
        lifetime(i)
    torch.cuda.synchronize()
    free1, _ = torch.cuda.mem_get_info(0)
    assert free0 - free1 < 8 << 20, (free0, free1)


def test_concurrent_predictions_on_one_regressor_are_serialised(host):
    """With the reference's parallel multi-start search, hardware_concurrency threads call Predict* on ONE shared const
    Regressor (src/acquisition-function.cpp:125-144). Calls on a context are serialised inside the host layer."""
    import threading
    X = S.make_X(120, 6, "sls")
    h = host.gpr_create(S.MATERN, X, S.make_y(X), S.make_theta(6, "perturbed"), 0.005)
    try:
        reg, Q = host.gpr_regressor(h), S.make_queries(16, 6)
        want = [host.predict(reg, Q[:, m]) for m in range(Q.shape[1])]
        errors = []

        def worker(t):
            try:
                for rep in range(20):
                    m = (t + rep) % Q.shape[1]
                    got = host.predict(reg, Q[:, m])
                    for a, b in zip(got, want[m]):
                        if not np.array_equal(np.asarray(a), np.asarray(b)):
                            errors.append((t, rep, m))
            except Exception as e:  # noqa: BLE001
                errors.append(repr(e))

        threads = [threading.Thread(target=worker, args=(t,)) for t in range(8)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        assert not errors, errors[:3]
    finally:
        host.gpr_destroy(h)
