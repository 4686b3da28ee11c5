"""CPU-only checks of the reference-faithful search driver (SearchDriver::Reference) against the real reference loop library
(oracle/_ref/libsls_ref_loop.so): the parts of the Submit -> next-slider step that never touch the GPU."""
import importlib

import numpy as np
import pytest

import loop_support as LS

pkg = importlib.import_module("sequential-line-search_b200")


@pytest.fixture(scope="module")
def sides():
    if not LS.ref_loop_available():
        pytest.skip("oracle/_ref/libsls_ref_loop.so not built")
    if not pkg.hostlib.nlopt_available():
        pytest.skip("host layer built without NLopt")
    previous = pkg.hostlib.get_search_driver()
    yield LS.LoopLib("ref"), LS.LoopLib("b200")
    pkg.hostlib.set_search_driver(previous)


def test_driver_switch():
    if not pkg.hostlib.nlopt_available():
        with pytest.raises(RuntimeError):
            pkg.hostlib.set_search_driver(pkg.hostlib.REFERENCE)
        return
    previous = pkg.hostlib.get_search_driver()
    for mode in (pkg.hostlib.NATIVE, pkg.hostlib.HYBRID, pkg.hostlib.REFERENCE):
        pkg.hostlib.set_search_driver(mode)
        assert pkg.hostlib.get_search_driver() == mode
    pkg.hostlib.set_search_driver(previous)


def test_slider_enlargement_by_cobyla_is_the_reference_slider(sides):
    """src/slider.cpp:73-142 (two COBYLA solves) through the same NLopt: bit-identical ends, including the short-slider branch."""
    ref, b200 = sides
    pkg.hostlib.set_search_driver(pkg.hostlib.REFERENCE)
    rng = np.random.default_rng(0)
    for case in range(40):
        D = int(rng.integers(1, 12))
        e0, e1 = rng.random(D), rng.random(D)
        if case % 4 == 1:
            e1 = e0 + 0.01 * (rng.random(D) - 0.5)       # shorter than minimum_length
        if case % 4 == 2:
            e0[rng.integers(0, D)] = 0.0                  # an end on the boundary
        if case % 4 == 3:
            e1 = np.clip(e0 + 0.9 * (rng.random(D) - 0.5), 0.0, 1.0)
        r0, r1 = ref.slider(e0, e1, True)
        b0, b1 = b200.slider(e0, e1, True)
        np.testing.assert_array_equal(b0, r0)
        np.testing.assert_array_equal(b1, r1)


def test_closed_form_enlargement_agrees_with_cobyla(sides):
    """The Native / Hybrid drivers solve the enlargement in closed form; COBYLA stops within its xtol of the same ends."""
    ref, b200 = sides
    rng = np.random.default_rng(1)
    pkg.hostlib.set_search_driver(pkg.hostlib.HYBRID)
    for _ in range(20):
        D = int(rng.integers(2, 10))
        e0, e1 = rng.random(D), rng.random(D)
        if np.linalg.norm(e0 - e1) < 0.3:
            continue
        r0, r1 = ref.slider(e0, e1, True)
        b0, b1 = b200.slider(e0, e1, True)
        assert np.max(np.abs(b0 - r0)) < 2e-5 and np.max(np.abs(b1 - r1)) < 2e-5


def test_initial_queries_consume_the_same_rand_stream(sides):
    """GenerateRandomSliderEnds / GenerateRandomPoints (Eigen Random() over libc rand()) give the same first query on both sides."""
    ref, b200 = sides
    for D in (2, 6, 17):
        ends = []
        for L in (ref, b200):
            L.srand(123)
            opt = L.sls(D)
            ends.append(opt.slider_ends())
            opt.close()
        np.testing.assert_array_equal(ends[0][0], ends[1][0])
        np.testing.assert_array_equal(ends[0][1], ends[1][1])
        options = []
        for L in (ref, b200):
            L.srand(77)
            opt = L.pbo(D, num_options=3)
            options.append(opt.current_options())
            opt.close()
        np.testing.assert_array_equal(options[0], options[1])
