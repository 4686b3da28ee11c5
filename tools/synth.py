"""Seeded synthetic inputs of the benchmark / test configurations. Pure numpy, no checker code:
both the product bench and the tests draw their inputs from here so that CPU and GPU consume identical bytes."""
import numpy as np


def f64(a, order="F"):
    return np.require(np.asarray(a, dtype=np.float64), requirements=["F" if order == "F" else "C", "A", "W"])


def make_X(N, D, kind="uniform", seed=1):
    rng = np.random.default_rng(seed)
    if kind == "uniform":
        return f64(rng.random((D, N)))
    # "SLS-like": ceil(N/3) random segments in [0,1]^D, 3 points on each (clustered -> worse conditioning)
    X = np.empty((D, N))
    i = 0
    while i < N:
        a, b = rng.random(D), rng.random(D)
        for t in rng.random(3):
            if i < N:
                X[:, i] = (1 - t) * a + t * b
                i += 1
    return f64(X)


def make_theta(D, kind="default", seed=2):
    if kind == "default":
        return np.concatenate([[0.5], np.full(D, 0.5)])
    rng = np.random.default_rng(seed)
    return np.concatenate([[0.5], rng.uniform(0.2, 1.0, D)])


def nd_demo_objective(X):
    """exp(-||x - 0.4||^2): the synthetic oracle of demos/sequential_line_search_nd/main.cpp:26-34."""
    return np.exp(-((X - 0.4) ** 2).sum(axis=0))


def make_y(X, seed=3, noise=1e-2):
    rng = np.random.default_rng(seed)
    return nd_demo_objective(X) + noise * rng.standard_normal(X.shape[1])


def make_tuples(X):
    """P = ceil(N/3) preference triples over consecutive points, winner first (mimics AddNewPoints)."""
    N = X.shape[1]
    f = nd_demo_objective(X)
    offsets, idx = [0], []
    for s in range(0, N, 3):
        members = list(range(s, min(s + 3, N)))
        if len(members) < 2:
            members = [s - 1, s]
        w = max(members, key=lambda i: f[i])
        members.remove(w)
        idx += [w] + members
        offsets.append(len(idx))
    return np.asarray(offsets, dtype=np.uint32), np.asarray(idx, dtype=np.uint32)


def make_queries(M, D, seed=4):
    rng = np.random.default_rng(seed)
    return f64(rng.random((D, M)))


