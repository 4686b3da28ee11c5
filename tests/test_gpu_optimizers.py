"""The public optimisers through the Python module (pySequentialLineSearch), end to end on the GPU: the headless demo of
the reference (demos/sequential_line_search_nd/main.cpp; README example D = 6, 15 iterations) and the preferential
Bayesian optimisation loop (demos/preferential_bayesian_optimization_1d)."""
import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pkg = importlib.import_module("sequential-line-search_b200")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sls():
    pkg.build_python_module()
    sys.path.insert(0, pkg.LIB_DIR)
    import pySequentialLineSearch
    return pySequentialLineSearch


@pytest.mark.parametrize("kernel", ["matern", "se"])
def test_sequential_line_search_nd_demo_converges(sls, kernel):
    sys.path.insert(0, os.path.join(ROOT, "demos"))
    import sequential_line_search_nd as demo
    history, x_star, opt = demo.run(dims=6, iters=15, kernel=kernel, seed=0, verbose=False)
    ys = [h["y"] for h in history]
    assert ys[-1] > 0.9 and ys[-1] >= ys[0]             # the synthetic user ends near the optimum f = 1
    assert history[-1]["residual"] < 0.35
    assert demo.objective(x_star) > 0.85
    X = opt.get_raw_data_points()
    assert X.shape[0] == 6 and 3 <= X.shape[1] <= 45      # 3 points per iteration, minus merged ones
    for h in history:                                      # every slider stays inside the unit box
        a, b = h["slider"]
        assert min(a.min(), b.min()) >= 0.0 and max(a.max(), b.max()) <= 1.0
    # the regressor behind the optimiser answers the single-point queries of the GUI demos
    mu, sd, acq = opt.get_preference_value_mean(x_star), opt.get_preference_value_stdev(x_star), opt.get_acquisition_func_value(x_star)
    assert np.isfinite(mu) and sd >= 0.0 and acq >= 0.0


def test_submit_feedback_with_explicit_effort_and_last_selection(sls, tmp_path):
    rng = np.random.default_rng(1)
    opt = sls.SequentialLineSearchOptimizer(num_dims=3, use_slider_enlargement=False, use_map_hyperparams=False,
                                            kernel_type=sls.KernelType.ArdSquaredExponentialKernel,
                                            acquisition_func_type=sls.AcquisitionFuncType.GaussianProcessUpperConfidenceBound,
                                            initial_query_generator=lambda n: (rng.random(n), rng.random(n)),
                                            current_best_selection_strategy=sls.CurrentBestSelectionStrategy.LastSelection)
    opt.set_gaussian_process_upper_confidence_bound_hyperparam(2.0)
    chosen = opt.calc_point_from_slider_position(0.25)
    opt.submit_feedback_data(0.25, 100, 20, 30)
    np.testing.assert_allclose(opt.get_maximizer(), chosen)          # LastSelection: x^chosen heads the next slider
    a, b = opt.get_slider_ends()
    np.testing.assert_allclose(a, chosen)
    assert opt.get_raw_data_points().shape == (3, 3)
    opt.damp_data(str(tmp_path))
    assert os.path.exists(tmp_path / "X.csv") and os.path.exists(tmp_path / "D.csv")


def test_preferential_bayesian_optimizer_loop(sls):
    rng = np.random.default_rng(4)
    f = lambda x: -float(np.sum((x - 0.3) ** 2))   # noqa: E731
    opt = sls.PreferentialBayesianOptimizer(num_dims=2, use_map_hyperparams=False, kernel_type=sls.KernelType.ArdMatern52Kernel,
                                            initial_query_generator=lambda n, k: [rng.random(n) for _ in range(k)], num_options=3)
    best = []
    for it in range(8):
        options = opt.get_current_options()
        assert len(options) == 3 and all(o.shape == (2,) and 0 <= o.min() and o.max() <= 1 for o in options)
        pick = int(np.argmax([f(o) for o in options]))
        best.append(f(options[pick]))
        opt.submit_feedback_data(pick)
        opt.determine_next_query(num_global_search_iters=20, num_local_search_iters=20)
        assert np.linalg.norm(opt.get_current_options()[1] - opt.get_current_options()[2]) > 1e-6   # Schonlau: distinct options
    assert max(best[-3:]) > -0.02 and max(best[-3:]) >= best[0]
    assert f(opt.get_maximizer()) > -0.05
    # custom feedback (points that were never options)
    opt.submit_custom_feedback_data(np.array([0.3, 0.3]), [np.array([0.9, 0.9]), np.array([0.1, 0.8])])
    assert opt.get_raw_data_points().shape[0] == 2


def test_the_reference_nd_demo_source_runs_unmodified_on_the_host_layer(tmp_path):
    """Source-level drop-in: oracle/_ref/sls_nd_demo_dropin is the reference's own demos/sequential_line_search_nd/main.cpp
    (D = 8, 3 trials x 10 iterations, MAP hyper-parameters), compiled WITHOUT changes against this repository's headers
    and libsls_b200_host.so by oracle/Makefile (`dropin`). It must run, converge and write its three CSV reports."""
    import re
    import subprocess
    exe = os.path.join(ROOT, "oracle", "_ref", "sls_nd_demo_dropin")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/sls_nd_demo_dropin was not built (the reference tree is needed at build time)")
    res = subprocess.run([exe], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    maxima = [float(v) for v in re.findall(r"Found maximum: ([0-9.eE+-]+)", res.stdout)]
    assert len(maxima) == 3 and min(maxima) > 0.7, maxima      # f(x) = exp(-|x - 0.4|^2), maximum 1
    for name in ("objective_values.csv", "residual_norms.csv", "elapsed_times.csv"):
        rows = (tmp_path / name).read_text().strip().split("\n")
        assert len(rows) == 10 and all(len(r.split(",")) == 3 for r in rows), name
    obj = np.array([[float(v) for v in r.split(",")] for r in (tmp_path / "objective_values.csv").read_text().strip().split("\n")])
    assert np.all(obj[-1] >= obj[0] - 1e-12)                    # the chosen slider positions improve over the run


def test_the_reference_bo_1d_demo_source_runs_unmodified_on_the_host_layer(tmp_path):
    """oracle/_ref/bo_1d_demo_dropin: the reference's demos/bayesian_optimization_1d/{main,core}.cpp, unmodified, on this
    repository's GaussianProcessRegressor (hyper-parameters by MAP of the marginal likelihood) + acquisition_func::FindNextPoint.
    f(x) = 1 - 1.5 x sin(13 x) on [0, 1]: global maximum 2.274 at x = 0.852, a local one of 1.556 at x = 0.378."""
    import re
    import subprocess
    exe = os.path.join(ROOT, "oracle", "_ref", "bo_1d_demo_dropin")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/bo_1d_demo_dropin was not built (the reference tree is needed at build time)")
    res = subprocess.run([exe], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    maxima = [float(v) for v in re.findall(r"Found maximum: ([0-9.eE+-]+)", res.stdout)]
    assert len(maxima) == 5 and min(maxima) > 1.5, maxima                 # 5 trials x 15 iterations, every one at a maximum
    assert sum(m > 2.25 for m in maxima) >= 3, maxima                       # most at the global one
