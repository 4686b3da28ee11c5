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
