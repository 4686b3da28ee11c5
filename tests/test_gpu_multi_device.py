"""Multi-GPU inside the library (VERDICT r1 "Missing" #2): slsgp_ctx_create_multi groups several devices behind ONE context;
the model lives on the primary, sweeps replicate it peer to peer and split their candidates over the devices. Needs >= 2 GPUs
(`gpurun --gpus 2`); skipped on a single-GPU box."""
import importlib

import numpy as np
import pytest

import support as S

pkg = importlib.import_module("sequential-line-search_b200")
pytestmark = pytest.mark.gpu


def _n_gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.fixture(scope="module")
def group():
    n = _n_gpus()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    return list(range(min(n, 4)))


def _fit(ctx, kt=S.SE, N=300, D=8):
    X, theta = S.make_X(N, D, "sls"), S.make_theta(D, "perturbed")
    ctx.fit(X, kt, theta, 0.005, S.make_y(X))
    return X, theta


@pytest.mark.parametrize("mode", ["fp64", "tensor"])
def test_group_argmax_is_the_single_device_argmax(group, mode):
    single, multi = pkg.Context(0), pkg.Context(group)
    try:
        assert multi.device_count() == len(group)
        for c in (single, multi):
            _fit(c)
            c.set_sweep_mode(pkg.SWEEP_TENSOR if mode == "tensor" else pkg.SWEEP_FP64)
        seed, first, count = 5, 1000, (1 << 18) + 77  # above the 2^16 threshold, not divisible by the group size
        a = single.acq_argmax(0, 1.0, seed, first, count, want_grad=True)
        b = multi.acq_argmax(0, 1.0, seed, first, count, want_grad=True)
        assert a[2] == b[2] and a[1] == b[1]
        np.testing.assert_array_equal(a[0], b[0])
        np.testing.assert_allclose(a[3], b[3], rtol=1e-12, atol=1e-300)
        # a refit on the primary reaches the replicas (model version)
        for c in (single, multi):
            _fit(c, S.MATERN, 200, 8)
        a = single.acq_argmax(1, 2.0, seed, 0, 1 << 17)
        b = multi.acq_argmax(1, 2.0, seed, 0, 1 << 17)
        assert a[2] == b[2] and a[1] == b[1]
        # small ranges stay on the primary
        a = single.acq_argmax(0, 1.0, seed, 0, 5000)
        b = multi.acq_argmax(0, 1.0, seed, 0, 5000)
        assert a[2] == b[2] and a[1] == b[1]
    finally:
        single.close()
        multi.close()


def test_group_batches_equal_single_device_batches(group):
    single, multi = pkg.Context(0), pkg.Context(group)
    try:
        for c in (single, multi):
            _fit(c)
        Q = S.make_queries((1 << 16) + 1234, 8)
        v0, g0 = single.acq_batch(0, 1.0, Q)
        v1, g1 = multi.acq_batch(0, 1.0, Q)
        np.testing.assert_array_equal(v0, v1)
        np.testing.assert_array_equal(g0, g1)
        mu0, s0, dmu0, ds0 = single.posterior_batch(Q)
        mu1, s1, dmu1, ds1 = multi.posterior_batch(Q)
        np.testing.assert_array_equal(mu0, mu1)
        np.testing.assert_array_equal(ds0, ds1)
    finally:
        single.close()
        multi.close()


def test_group_maximize_is_at_least_as_good_as_the_sweep(group):
    single, multi = pkg.Context(0), pkg.Context(group)
    try:
        for c in (single, multi):
            _fit(c)
        seed, count = 11, 1 << 18
        _, v_sweep, _, _ = single.acq_argmax(0, 1.0, seed, 0, count)
        x, v, g, vs = multi.acq_maximize(0, 1.0, seed, 0, count, n_starts=256, n_iters=30)
        assert vs == v_sweep and v >= v_sweep
        assert np.all(x >= 0) and np.all(x <= 1)
        v_check, _ = single.acq_batch(0, 1.0, x[:, None])
        assert abs(v_check[0] - v) <= 1e-9 * max(abs(v), 1e-12)
    finally:
        single.close()
        multi.close()


def test_host_layer_find_next_point_on_a_group(group):
    """sequential_line_search::SetDevices: regressors built afterwards own a multi-GPU group; FindNextPoint (device maximiser)
    then searches on all of them. The point found must be as good as the single-device one under the same regressor."""
    host = pkg.hostlib.Host()
    previous = pkg.hostlib.get_search_driver()
    pkg.hostlib.set_search_driver(pkg.hostlib.NATIVE)
    X, theta = S.make_X(60, 6, "sls"), S.make_theta(6, "perturbed")
    y = S.make_y(X)
    try:
        pkg.hostlib.set_devices([0])
        h1 = host.gpr_create(S.MATERN, X, y, theta, 0.005)
        x1 = host.find_next_point(host.gpr_regressor(h1), 6, n_global=300, n_local=60)
        pkg.hostlib.set_devices(group)
        hg = host.gpr_create(S.MATERN, X, y, theta, 0.005)
        xg = host.find_next_point(host.gpr_regressor(hg), 6, n_global=300, n_local=60)
        v1 = host.acq(host.gpr_regressor(h1), S.EI, 1.0, x1, want_grad=False)[0]
        vg = host.acq(host.gpr_regressor(h1), S.EI, 1.0, xg, want_grad=False)[0]
        assert vg >= v1 * (1 - 1e-6)
        host.gpr_destroy(h1)
        host.gpr_destroy(hg)
    finally:
        pkg.hostlib.set_devices([0])
        pkg.hostlib.set_search_driver(previous)
