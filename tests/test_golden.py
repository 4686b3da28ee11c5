"""Golden vectors generated from the reference itself (tests/golden/make_golden.py) vs.
  - the plain-C oracle   (CPU, not gpu) — keeps the oracle pinned on boxes without /root/reference,
  - the CUDA library     (gpu)          — parity of libslsgp through the C ABI.
Tolerance for the CUDA path: north_star's FP64 bar, 1e-5 relative (the observed error is ~1e-10)."""
import glob
import importlib
import os

import numpy as np
import pytest

import support as S

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))
IDS = [os.path.basename(p)[:-4] for p in GOLDEN]


def _load(path):
    return {k: v for k, v in np.load(path).items()}


def check(name, got, want, rtol, atol=0.0):
    got, want = np.asarray(got), np.asarray(want)
    scale = max(float(np.max(np.abs(want))), 1e-300) if want.size else 1.0
    err = float(np.max(np.abs(got - want))) if want.size else 0.0
    assert err <= rtol * scale + atol, f"{name}: max abs err {err:.3e} vs scale {scale:.3e} (rtol {rtol})"


def test_fixtures_present():
    assert len(GOLDEN) >= 6


# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("path", GOLDEN, ids=IDS)
def test_oracle_matches_golden(oracle, path):
    g = _load(path)
    kt, X, theta = int(g["kernel_type"]), g["X"], g["theta"]
    D, N = X.shape
    m = oracle.model(kt, X, theta, float(g["gpr_noise"]), g["gpr_y"])
    check("K", m._keep[4], g["ref_K"], 1e-13)
    check("Kinv", oracle.inverse(g["ref_K"]), g["ref_Kinv"], 1e-8)
    i_best, f_best = oracle.f_best(m)
    np.testing.assert_array_equal(X[:, i_best], g["ref_x_best"])
    check("f_best", f_best, g["ref_f_best"], 1e-9)
    out = oracle.acq_batch(m, S.EI, 1.0, f_best, g["Q"])
    check("mu", out["mu"], g["ref_mu"], 1e-9)
    check("sigma", out["sigma"], g["ref_sigma"], 1e-9)
    check("dmu", out["dmu"], g["ref_dmu"], 1e-8)
    finite = np.isfinite(g["ref_dsigma"]).all(axis=0)
    check("dsigma", out["dsigma"][:, finite], g["ref_dsigma"][:, finite], 1e-7)
    check("ei", out["val"], g["ref_ei"], 1e-8, atol=1e-300)
    check("dei", out["grad"], g["ref_dei"], 1e-7, atol=1e-300)
    out = oracle.acq_batch(m, S.UCB, float(g["ucb_beta"]), f_best, g["Q"])
    check("ucb", out["val"], g["ref_ucb"], 1e-9)
    if "gpr_map_points" in g:
        for p, x in enumerate(g["gpr_map_points"]):
            f, gr = oracle.map_objective_gpr(kt, X, g["gpr_y"], x)
            check("gpr map f", f, g["ref_gpr_map_f"][p], 1e-9)
            check("gpr map grad", gr, g["ref_gpr_map_grad"][p], 1e-7)
    if "pref_solution" in g:
        a, r, b, var, btl = g["defaults"]
        use_map = bool(g["use_map"])
        for p, x in enumerate(g["pref_map_points"]):
            f, gr = oracle.map_objective_pref(kt, X, g["pref_offsets"], g["pref_idx"], use_map, a, r, b, var, btl, x)
            check("pref map f", f, g["ref_pref_map_f"][p], 1e-10)
            check("pref map grad", gr, g["ref_pref_map_grad"][p], 1e-8)


# ------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def ctx():
    slsb = importlib.import_module("sequential-line-search_b200")
    c = slsb.Context(0)
    yield c
    c.close()


RT = 1e-5  # north_star: 1e-5 relative in FP64


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=IDS)
def test_cuda_gpr_matches_golden(ctx, path):
    g = _load(path)
    kt, X, theta = int(g["kernel_type"]), g["X"], g["theta"]
    D, N = X.shape
    ctx.set_data(X)
    K = ctx.gram(kt, theta, float(g["gpr_noise"]))
    check("K", K, g["ref_K"], 1e-12)
    logdet, L = ctx.factor(want_L=True)
    check("L L^T", L @ L.T, g["ref_K"], 1e-12)
    assert np.all(np.triu(L, 1) == 0.0)
    sign, ld_ref = np.linalg.slogdet(g["ref_K"])
    check("logdet", logdet, ld_ref, 1e-10, atol=1e-9)
    check("Kinv", ctx.inverse(), g["ref_Kinv"], 1e-8)
    ctx.solve_alpha(g["gpr_y"])
    f_best, i_best = ctx.f_best()
    np.testing.assert_array_equal(X[:, i_best], g["ref_x_best"])
    check("f_best", f_best, g["ref_f_best"], 1e-9)
    mu, sigma, dmu, dsigma = ctx.posterior_batch(g["Q"])
    check("mu", mu, g["ref_mu"], RT)
    check("sigma", sigma, g["ref_sigma"], RT)
    check("dmu", dmu, g["ref_dmu"], RT)
    finite = np.isfinite(g["ref_dsigma"]).all(axis=0)
    check("dsigma", dsigma[:, finite], g["ref_dsigma"][:, finite], RT)
    ei, dei = ctx.acq_batch(0, 1.0, g["Q"])
    check("ei", ei, g["ref_ei"], RT, atol=1e-300)
    check("dei", dei, g["ref_dei"], RT, atol=1e-300)
    ucb, ducb = ctx.acq_batch(1, float(g["ucb_beta"]), g["Q"])
    check("ucb", ucb, g["ref_ucb"], RT)
    check("ducb", ducb[:, finite], g["ref_ducb"][:, finite], RT)
    if "gpr_map_points" in g:
        for p, x in enumerate(g["gpr_map_points"]):
            f, gr = ctx.map_objective_gpr(kt, g["gpr_y"], x)
            check("gpr map f", f, g["ref_gpr_map_f"][p], 1e-9)
            check("gpr map grad", gr, g["ref_gpr_map_grad"][p], RT)
            f2, _ = ctx.map_objective_gpr(kt, g["gpr_y"], x, want_grad=False)
            assert f2 == f


@pytest.mark.gpu
@pytest.mark.parametrize("path", [p for p in GOLDEN if "n1" not in p], ids=[i for i in IDS if "n1" not in i])
def test_cuda_preference_matches_golden(ctx, path):
    g = _load(path)
    kt, X = int(g["kernel_type"]), g["X"]
    D, N = X.shape
    a, r, b, var, btl = g["defaults"]
    use_map = bool(g["use_map"])
    sol = g["pref_solution"]
    theta, noise = g["ref_pref_theta"], float(g["ref_pref_noise"])
    # the state the reference regressor ended up in: K, LLT(K), y
    ctx.set_data(X)
    check("K", ctx.gram(kt, theta, noise), g["ref_pref_K"], 1e-12)
    _, L = ctx.factor(want_L=True)
    check("L", L, g["ref_pref_L"], 1e-9)
    ctx.solve_alpha(sol[:N])
    mu, sigma, dmu, dsigma = ctx.posterior_batch(g["Q"])
    check("mu", mu, g["ref_pref_mu"], RT)
    check("sigma", sigma, g["ref_pref_sigma"], RT)
    check("dmu", dmu, g["ref_pref_dmu"], RT)
    finite = np.isfinite(g["ref_pref_dsigma"]).all(axis=0)
    check("dsigma", dsigma[:, finite], g["ref_pref_dsigma"][:, finite], RT)
    ei, dei = ctx.acq_batch(0, 1.0, g["Q"])
    check("ei", ei, g["ref_pref_ei"], RT, atol=1e-300)
    check("dei", dei, g["ref_pref_dei"], RT, atol=1e-300)
    # MAP objective + gradient
    ctx.set_preferences(g["pref_offsets"], g["pref_idx"])
    if not use_map:  # fixed hyper-parameters: the reference builds K from the defaults first (:356-364)
        ctx.gram(kt, np.concatenate([[a], np.full(D, r)]), b, want=False)
        ctx.factor()
    for p, x in enumerate(g["pref_map_points"]):
        f, gr = ctx.map_objective_pref(kt, x, use_map, a, r, b, var, btl)
        check("pref map f", f, g["ref_pref_map_f"][p], 1e-9)
        check("pref map grad", gr, g["ref_pref_map_grad"][p], RT)
        f2, _ = ctx.map_objective_pref(kt, x, use_map, a, r, b, var, btl, want_grad=False)
        assert abs(f2 - f) <= 1e-12 * abs(f)
