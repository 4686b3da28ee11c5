"""Generate tests/golden/*.npz from the reference ITSELF.

Run in the build container (needs oracle/_ref/libsls_ref_probe.so, i.e. /root/reference compiled by oracle/Makefile):
    python tests/golden/make_golden.py
Every array stored under a `ref_` key was computed by the reference's unmodified C++ (src/*.cpp + mathtoolbox,
linked against include/eigen-lite). The inputs are stored alongside so the fixtures are self-contained on the GPU
box, where /root/reference does not exist.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import support as S  # noqa: E402

CASES = {
    # name: (kernel, D, N, X kind, theta kind, use_map)
    "se_d6_n24": (S.SE, 6, 24, "uniform", "default", True),
    "matern_d6_n24": (S.MATERN, 6, 24, "uniform", "perturbed", True),
    "se_d16_n70_sls": (S.SE, 16, 70, "sls", "perturbed", False),
    "matern_d8_n129_sls": (S.MATERN, 8, 129, "sls", "default", True),
    "se_d3_n1": (S.SE, 3, 1, "uniform", "perturbed", False),
    "se_d64_n65": (S.SE, 64, 65, "uniform", "default", False),
}


def make_case(ref, name, kt, D, N, xkind, tkind, use_map):
    a, r, b, var, btl = 0.5, 0.5, 0.005, 0.25, 0.01
    X = S.make_X(N, D, xkind)
    theta = S.make_theta(D, tkind)
    rng = np.random.default_rng(11)
    out = dict(kernel_type=kt, X=X, theta=theta, use_map=int(use_map), defaults=np.array([a, r, b, var, btl]))

    # --- GaussianProcessRegressor with given hyper-parameters
    y = S.make_y(X)
    noise = 0.005
    h = ref.gpr_create(kt, X, y, theta, noise)
    reg = ref.gpr_regressor(h)
    K, Kinv = ref.gpr_state(h, N)
    Q = np.concatenate([S.make_queries(24, D), X[:, :2], np.full((D, 1), 2.5)], axis=1)
    M = Q.shape[1]
    mu, sg = np.empty(M), np.empty(M)
    dmu, dsg = np.empty((D, M)), np.empty((D, M))
    ei, dei, ucb, ducb = np.empty(M), np.empty((D, M)), np.empty(M), np.empty((D, M))
    for q in range(M):
        mu[q], sg[q], dmu[:, q], dsg[:, q] = ref.predict(reg, Q[:, q])
        ei[q], dei[:, q] = ref.acq(reg, S.EI, 1.0, Q[:, q])
        ucb[q], ducb[:, q] = ref.acq(reg, S.UCB, 2.5, Q[:, q])
    x_best = ref.x_best(reg, D)
    out.update(gpr_y=y, gpr_noise=noise, ref_K=K, ref_Kinv=Kinv, Q=Q, ref_mu=mu, ref_sigma=sg, ref_dmu=dmu,
               ref_dsigma=dsg, ref_ei=ei, ref_dei=dei, ref_ucb=ucb, ref_ducb=ducb, ucb_beta=2.5, ref_x_best=x_best,
               ref_f_best=ref.predict(reg, x_best)[0])
    if N > 1:
        pts = np.stack([np.concatenate([[0.5, 1e-4], np.full(D, 0.5)]),
                        np.concatenate([[0.8, 0.01], rng.uniform(0.2, 1.0, D)])])
        f, g = ref.gpr_objective(kt, X, y, pts)
        out.update(gpr_map_points=pts, ref_gpr_map_f=f, ref_gpr_map_grad=g)
    ref.gpr_destroy(h)

    # --- PreferenceRegressor put into a given state
    if N > 1:
        offsets, idx = S.make_tuples(X)
        yp = 0.05 * rng.standard_normal(N)
        sol = np.concatenate([yp, [theta[0], 0.007], theta[1:]]) if use_map else yp
        h = ref.pref_create(kt, X, offsets, idx, use_map, a, r, b, var, btl, sol)
        st = ref.pref_state(h, N, D)
        reg = ref.pref_regressor(h)
        pmu, psg = np.empty(M), np.empty(M)
        pdmu, pdsg, pei, pdei = np.empty((D, M)), np.empty((D, M)), np.empty(M), np.empty((D, M))
        for q in range(M):
            pmu[q], psg[q], pdmu[:, q], pdsg[:, q] = ref.predict(reg, Q[:, q])
            pei[q], pdei[:, q] = ref.acq(reg, S.EI, 1.0, Q[:, q])
        x0 = np.concatenate([np.zeros(N), [a, b], np.full(D, r)]) if use_map else np.zeros(N)
        pts = np.stack([sol, sol * 1.1 + 0.01, x0])
        fs, gs = [], []
        for p in pts:
            f, g = ref.pref_objective(h, p)
            fs.append(f), gs.append(g)
        out.update(pref_offsets=offsets, pref_idx=idx, pref_solution=sol, ref_pref_theta=st["theta"],
                   ref_pref_noise=st["b"], ref_pref_K=st["K"], ref_pref_L=st["L"], ref_pref_mu=pmu,
                   ref_pref_sigma=psg, ref_pref_dmu=pdmu, ref_pref_dsigma=pdsg, ref_pref_ei=pei, ref_pref_dei=pdei,
                   pref_map_points=pts, ref_pref_map_f=np.array(fs), ref_pref_map_grad=np.stack(gs))
        ref.pref_destroy(h)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "ok", {k: np.asarray(v).shape for k, v in out.items() if k.startswith("ref_")})


if __name__ == "__main__":
    if not S.ref_available():
        S.build_oracle()
    ref = S.Ref()
    for name, args in CASES.items():
        make_case(ref, name, *args)
