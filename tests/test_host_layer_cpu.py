"""CPU-side checks of the C++ host layer (sequential-line-search_b200/host): the library loads, exports the whole facade,
its kernel function pointers agree with the reference's, and it refuses to work without a GPU (no CPU fallback)."""
import importlib

import numpy as np
import pytest

import support as S

pkg = importlib.import_module("sequential-line-search_b200")


def test_host_library_exports_every_facade_symbol():
    lib = pkg.hostlib.load_host_library()
    for name in pkg.hostlib.HOST_SYMBOLS:
        assert hasattr(lib, name), name


@pytest.mark.parametrize("kt", [S.SE, S.MATERN])
def test_kernel_function_pointers_match_reference(ref, kt):
    """Regressor::GetKernel() / GetKernelThetaDerivative() / GetKernelFirstArgDerivative() (regressor.hpp:30-32)."""
    host = pkg.hostlib.Host()
    rng = np.random.default_rng(5)
    for D in (1, 3, 16):
        for _ in range(5):
            xa, xb = rng.random(D), rng.random(D)
            theta = np.concatenate([[rng.uniform(0.1, 2.0)], rng.uniform(0.2, 1.5, D)])
            k, dth, dx = host.kernel(kt, xa, xb, theta)
            k_r, dth_r, dx_r = ref.kernel(kt, xa, xb, theta)
            assert abs(k - k_r) <= 1e-14 * max(1.0, abs(k_r))
            np.testing.assert_allclose(dth, dth_r, rtol=1e-12, atol=1e-15)
            np.testing.assert_allclose(dx, dx_r, rtol=1e-12, atol=1e-15)
    # coincident points: the Matern x-derivative is defined as zero there (kernel-functions.cpp:198)
    x = rng.random(4)
    theta = np.array([0.5, 0.5, 0.5, 0.5, 0.5])
    _, _, dx = host.kernel(kt, x, x, theta)
    assert np.all(dx == 0.0)


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    host = pkg.hostlib.Host()
    X = S.make_X(12, 3)
    with pytest.raises(pkg.hostlib.HostError, match="no usable CUDA device"):
        host.gpr_create(S.SE, X, S.make_y(X), S.make_theta(3), 0.005)


def test_preference_data_manager_matches_reference(ref):
    """AddNewPoints + MergeClosePoints (src/preference-data-manager.cpp:14-141): same X, same re-indexed tuples."""
    host = pkg.hostlib.Host()
    rng = np.random.default_rng(8)
    D = 4
    batches = [rng.random((D, 3))]
    for it in range(7):
        b = rng.random((D, 3))
        if it % 2 == 0:      # the new slider's first end coincides with an old point (x^+), as in the real loop
            b[:, 1] = batches[rng.integers(len(batches))][:, 0] + 1e-6 * rng.standard_normal(D)
        if it == 3:          # two members of the same batch nearly coincide
            b[:, 2] = b[:, 0] + 1e-7
        batches.append(b)
    batches.append(rng.random((D, 2)))   # a pairwise comparison
    Xh, oh, ih = host.data_manager_run(batches)
    Xr, orf, ir = ref.data_manager_run(batches)
    assert Xh.shape == Xr.shape and Xh.shape[1] < sum(b.shape[1] for b in batches)   # something was merged
    np.testing.assert_array_equal(Xh, Xr)
    np.testing.assert_array_equal(oh, orf)
    np.testing.assert_array_equal(ih[:oh[-1]], ir[:orf[-1]])


def test_slider_enlargement_stays_in_the_box_and_on_the_line():
    """Slider (src/slider.cpp:74-153): enlarged ends lie on the original line, inside [0,1]^D, at most `scale` times
    longer; short sliders are stretched to the minimum length; without enlargement the ends are the inputs."""
    host = pkg.hostlib.Host()
    rng = np.random.default_rng(2)
    for D in (2, 6, 16):
        for _ in range(20):
            a, b = rng.random(D), rng.random(D)
            e0, e1 = host.slider(a, b, enlarge=False)
            np.testing.assert_array_equal(e0, a)
            np.testing.assert_array_equal(e1, b)
            e0, e1 = host.slider(a, b, enlarge=True)
            c, r = 0.5 * (a + b), a - 0.5 * (a + b)
            t0 = np.dot(e0 - c, r) / np.dot(r, r)
            t1 = np.dot(e1 - c, r) / np.dot(r, r)
            np.testing.assert_allclose(e0, c + t0 * r, atol=1e-12)
            np.testing.assert_allclose(e1, c + t1 * r, atol=1e-12)
            if np.linalg.norm(e0 - e1) > 0.25 + 1e-9:   # not a stretched (minimum-length) slider
                assert 1.0 - 1e-12 <= t0 <= 1.25 + 1e-12 and -1.25 - 1e-12 <= t1 <= -1.0 + 1e-12
                assert e0.min() >= 0 and e0.max() <= 1 and e1.min() >= 0 and e1.max() <= 1
                # an end stops early only because it hit the box
                if t0 < 1.25 - 1e-9:
                    assert min(e0.min(), 1 - e0.max()) < 1e-9
    # interior, short slider: stretched symmetrically to the minimum length
    a, b = np.full(3, 0.5), np.full(3, 0.5) + np.array([0.01, 0.0, 0.0])
    e0, e1 = host.slider(a, b, enlarge=True)
    assert abs(np.linalg.norm(e0 - e1) - 0.25) < 1e-9


def test_python_module_surface_matches_the_reference_binding():
    """python/pySequentialLineSearch.cpp:15-152: enums, classes, method names, keyword arguments."""
    import sys
    pkg.build_python_module()
    sys.path.insert(0, pkg.LIB_DIR)
    import pySequentialLineSearch as m
    assert int(m.KernelType.ArdSquaredExponentialKernel) == 0 and int(m.KernelType.ArdMatern52Kernel) == 1
    assert int(m.AcquisitionFuncType.GaussianProcessUpperConfidenceBound) == 1
    assert int(m.CurrentBestSelectionStrategy.LastSelection) == 1
    slso = ["set_hyperparams", "submit_feedback_data", "get_slider_ends", "calc_point_from_slider_position", "get_maximizer",
            "get_preference_value_mean", "get_preference_value_stdev", "get_acquisition_func_value", "get_raw_data_points",
            "damp_data", "set_gaussian_process_upper_confidence_bound_hyperparam"]
    pbo = ["set_hyperparams", "submit_feedback_data", "submit_custom_feedback_data", "determine_next_query", "get_current_options",
           "get_maximizer", "get_preference_value_mean", "get_preference_value_stdev", "get_acquisition_func_value",
           "get_raw_data_points", "damp_data", "set_gaussian_process_upper_confidence_bound_hyperparam"]
    for name in slso:
        assert hasattr(m.SequentialLineSearchOptimizer, name), name
    for name in pbo:
        assert hasattr(m.PreferentialBayesianOptimizer, name), name
    o = m.SequentialLineSearchOptimizer(num_dims=5, use_slider_enlargement=False, use_map_hyperparams=False,
                                        kernel_type=m.KernelType.ArdSquaredExponentialKernel,
                                        acquisition_func_type=m.AcquisitionFuncType.ExpectedImprovement,
                                        initial_query_generator=lambda n: (np.zeros(n), np.ones(n)),
                                        current_best_selection_strategy=m.CurrentBestSelectionStrategy.LastSelection)
    a, b = o.get_slider_ends()
    assert a.shape == (5,) and np.all(a == 0) and np.all(b == 1)
    np.testing.assert_allclose(o.calc_point_from_slider_position(0.3), np.full(5, 0.3))
    np.testing.assert_array_equal(o.get_maximizer(), a)
    assert o.get_preference_value_mean(np.zeros(5)) == 0.0 and o.get_acquisition_func_value(np.zeros(5)) == 0.0   # no data yet
    assert o.get_raw_data_points().size == 0
    o.set_hyperparams(kernel_signal_var=0.4, kernel_length_scale=0.3, noise_level=0.01, kernel_hyperparams_prior_var=0.2, btl_scale=0.02)
    q = m.PreferentialBayesianOptimizer(num_dims=3, num_options=4, initial_query_generator=lambda n, k: [np.full(n, i / k) for i in range(k)])
    opts = q.get_current_options()
    assert len(opts) == 4 and np.allclose(opts[2], 0.5)
    np.testing.assert_array_equal(q.get_maximizer(), opts[0])


def test_quasi_newton_driver_on_closed_form_problems():
    """host/src/optimizer.hpp (used where the reference calls NLopt): bounds are honoured, active-set solution of a box-constrained
    quadratic is exact, the Rosenbrock valley is descended to its minimum, the evaluation budget is respected."""
    host = pkg.hostlib.Host()
    n = 12
    x, f, evals = host.test_minimize(0, np.full(n, 0.3))
    want = np.clip(2.0 * np.arange(n) / n - 0.5, 0.0, 1.0)            # projection of the targets onto the box
    np.testing.assert_allclose(x, want, atol=1e-8)
    assert evals < 200
    x, f, evals = host.test_minimize(1, np.full(6, -1.2), max_evals=5000)
    np.testing.assert_allclose(x, np.ones(6), atol=1e-5)
    assert f < 1e-10 and evals < 5000
    x, f, evals = host.test_minimize(1, np.full(6, -1.2), max_evals=30)   # budget: stops early, still inside the box
    assert evals <= 30 + 8 and np.all(np.abs(x) <= 2.0) and f > 1e-10


def test_utils_match_the_reference(ref, tmp_path):
    """sequential_line_search::utils (include/sequential-line-search/utils.hpp:18-58, src/utils.cpp): BTL likelihood and its
    gradient against the reference's inline functions, uniform random vectors, CSV export."""
    host = pkg.hostlib.Host()
    rng = np.random.default_rng(5)
    for m, scale in ((2, 1.0), (3, 0.01), (5, 0.3), (1, 1.0)):
        f = rng.standard_normal(m) * 0.02
        v, d = host.btl(f, scale)
        v_r, d_r = ref.btl(f, scale)
        assert abs(v - v_r) <= 1e-13 * abs(v_r) and S.rel_err(d, d_r) < 1e-11
    v, d = host.btl(np.array([800.0, -800.0, 0.0]), 1.0)      # the reference's exp() overflows to inf / inf here
    assert v == 1.0 and np.all(np.isfinite(d))
    for n in (1, 7):
        x = host.random_vector(n)
        assert x.shape == (n,) and np.all(x >= 0.0) and np.all(x <= 1.0)
    assert not np.array_equal(host.random_vector(6), host.random_vector(6))
    X = np.array([[1.5, -2.0, 3.25], [1e-7, 123456.789, 0.0]])
    host.export_csv(tmp_path / "m.csv", X)
    text = (tmp_path / "m.csv").read_text()
    assert text == "1.5,-2,3.25\n1e-07,123457,0"             # Eigen's StreamPrecision (6 significant digits), no trailing newline
