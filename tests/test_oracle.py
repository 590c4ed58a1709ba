"""Pins the CPU oracle (oracle/) against golden vectors produced by the unmodified reference
(tests/golden/make_golden.py) and the reference's deterministic known-answer tests."""
import numpy as np
import pytest

from conftest import golden
from oracle import oracle as orc


def rel(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))


@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_gpr_fit_and_predict(name):
    g = golden("gpr")
    X = np.atleast_2d(g[f"gpr_{name}_X"]); y = g[f"gpr_{name}_y"]
    # effective hyper-parameters: the reference's setters round Python floats through fp32
    # (`src/gpr.py:59,72,85`), and the noise eye is fp32 (`src/gpr.py:170`)
    lam, sf, sn = g[f"gpr_{name}_lam_eff"], float(g[f"gpr_{name}_sf_eff"]), float(g[f"gpr_{name}_sn_eff"])
    sn = float(np.float32(sn ** 2)) ** 0.5
    if name == "c":
        X = X[:1]; y = y[:1]
    f = orc.fit(X, y, lam, sf, sn)
    np.testing.assert_allclose(f["Kf"], g[f"gpr_{name}_Kf"], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(f["Ky"], g[f"gpr_{name}_Ky"], rtol=1e-12, atol=1e-14)
    Kinv = g[f"gpr_{name}_Kyinv"]
    mean, cov = orc.predict(X, y, lam, sf, sn, g[f"gpr_{name}_Xp"], covar=True, Ky_inv=Kinv)
    np.testing.assert_allclose(mean, g[f"gpr_{name}_mean"], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(cov, g[f"gpr_{name}_cov"], rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(orc.se_gram(g[f"gpr_{name}_Xp"], X, lam, sf), g[f"gpr_{name}_Kpt"], rtol=1e-12)


@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_moment_matching_full_S(name):
    g = golden("moment_matching")
    X, y, u, S, sf = (g[f"mm_{name}_{k}"] for k in ("X", "y", "u", "S", "sf"))
    sf = float(sf)
    for tag in ("1", "2"):
        lam = g[f"mm_{name}_lam{tag}"]; Kinv = g[f"mm_{name}_Kinv{tag}"]
        m, beta, l = orc.mean_prop(Kinv, lam, u, S, X, y, sf)
        v = orc.variance_prop(Kinv, lam, u, S, X, m, beta, sf)
        assert abs(m - g[f"mm_{name}_mean{tag}"]) < 1e-10 * max(1, abs(m))
        assert abs(v - g[f"mm_{name}_var{tag}"]) < 1e-9 * sf ** 2
        np.testing.assert_allclose(beta, g[f"mm_{name}_beta{tag}"], rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(l, l)
    lam1, lam2 = g[f"mm_{name}_lam1"], g[f"mm_{name}_lam2"]
    m1, b1, _ = orc.mean_prop(g[f"mm_{name}_Kinv1"], lam1, u, S, X, y, sf)
    m2, b2, _ = orc.mean_prop(g[f"mm_{name}_Kinv2"], lam2, u, S, X, y, sf)
    c_bug = orc.covariance_prop(lam1, lam2, u, S, X, m1, m2, b1, b2, sf, sf, bugcompat=True)
    assert abs(c_bug - g[f"mm_{name}_cov12_torch"]) < 1e-9 * max(1.0, abs(c_bug))
    if name == "a":   # sigma_f == 1: the reference's NumPy twins apply
        c_ok = orc.covariance_prop(lam1, lam2, u, S, X, m1, m2, b1, b2, 1.0, 1.0, bugcompat=False)
        assert abs(c_ok - g["mm_a_cov12_numpy"]) < 1e-8 * max(1.0, abs(c_ok))
        assert abs(m1 - g["mm_a_mean1_numpy"]) < 1e-8


@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_c_oracle_moment_matching_diag(name):
    g = golden("moment_matching")
    X, y, u, S, sf = (g[f"mm_{name}_{k}"] for k in ("X", "y", "u", "S", "sf"))
    lam = g[f"mm_{name}_lam1"]; Kinv = g[f"mm_{name}_Kinv1"]; beta = g[f"mm_{name}_beta1"]
    m, v, part = orc.c_moment_match_diag(X, Kinv, beta, lam, float(sf), u, np.diag(S).copy())
    assert abs(m - g[f"mm_{name}_mean1_diag"]) < 1e-10 * max(1, abs(m))
    assert abs(v - g[f"mm_{name}_var1_diag"]) < 1e-9 * float(sf) ** 2
    # closed-form partials vs central differences of the NumPy oracle
    D = X.shape[1]
    s = np.diag(S).copy()
    def f(uu, ss):
        mm, bb, _ = orc.mean_prop(Kinv, lam, uu, np.diag(ss), X, y, float(sf))
        return mm, orc.variance_prop(Kinv, lam, uu, np.diag(ss), X, mm, bb, float(sf))
    h = 1e-6
    for k in range(D):
        e = np.zeros(D); e[k] = h
        mp, vp = f(u + e, s); mn, vn = f(u - e, s)
        assert abs((mp - mn) / (2 * h) - part[k]) < 1e-6 * max(1, abs(part[k]))
        assert abs((vp - vn) / (2 * h) - part[2 * D + k]) < 1e-5 * max(1, abs(part[2 * D + k]))
        mp, vp = f(u, s + e); mn, vn = f(u, s - e)
        assert abs((mp - mn) / (2 * h) - part[D + k]) < 1e-6 * max(1, abs(part[D + k]))
        assert abs((vp - vn) / (2 * h) - part[3 * D + k]) < 1e-5 * max(1, abs(part[3 * D + k]))


def _rollout_inputs(g, name):
    S, A, nxt = g[f"{name}_S"], g[f"{name}_A"], g[f"{name}_next"]
    X = np.concatenate([S, A], 1)
    lam, sf, sn = g[f"{name}_lam"], g[f"{name}_sf"], g[f"{name}_sn"]
    E = S.shape[1]
    fits = [orc.fit(X, nxt[:, a], lam[a], sf[a], float(np.float32(sn[a] ** 2)) ** 0.5) for a in range(E)]
    return X, nxt, lam, sf, fits


@pytest.mark.parametrize("name", ["r1", "r2", "r3", "r4", "r5"])
def test_rollout_cost_grad(name):
    g = golden("rollout")
    X, Y, lam, sf, fits = _rollout_inputs(g, name)
    E = Y.shape[1]
    Rd = g[f"{name}_Rd"]; Rd = None if Rd.size == 0 else Rd
    m = g[f"{name}_U"].shape[1]
    last_u = g[f"{name}_last"][:m]
    Kinvs = [f["Ky_inv"] for f in fits]; betas = [f["beta"] for f in fits]
    # NumPy oracle: forward only
    means, vars_ = orc.rollout(X, Kinvs, Y, lam, sf, g[f"{name}_x0"], g[f"{name}_U"])
    ref_means = g[f"{name}_means"]; ref_vars = np.stack([np.diag(c) for c in g[f"{name}_covs"]])
    np.testing.assert_allclose(means, ref_means, rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(vars_, ref_vars, rtol=1e-6, atol=1e-9)
    c = orc.cost(means, g[f"{name}_U"], vars_, g[f"{name}_xref"], g[f"{name}_uref"], float(g[f"{name}_gamma"]),
                 g[f"{name}_Q"], g[f"{name}_R"], Rd, last_u)
    assert abs(c - float(g[f"{name}_cost"])) < 1e-7 * max(1, abs(c))
    # C oracle: cost + closed-form adjoint vs the reference's autograd gradient
    cc, grad, cm, cv = orc.c_rollout_cost_grad(X, Kinvs, betas, lam, sf, g[f"{name}_x0"], g[f"{name}_U"],
                                               float(g[f"{name}_gamma"]), g[f"{name}_Q"], g[f"{name}_R"], Rd,
                                               last_u, g[f"{name}_xref"], g[f"{name}_uref"])
    assert abs(cc - float(g[f"{name}_cost"])) < 1e-7 * max(1, abs(cc))
    np.testing.assert_allclose(cm, ref_means, rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(cv, ref_vars, rtol=1e-6, atol=1e-9)
    gref = g[f"{name}_grad"]
    assert np.max(np.abs(grad - gref)) < 1e-6 * max(1.0, np.max(np.abs(gref)))


def test_shipped_data_config1():
    """BASELINE config 1 inputs (n=400, sigma_n=1e-5, cond(Ky)~2.6e6): looser bar, NaN mask identical."""
    g = golden("shipped")
    S, A, nxt = g["ship_S"], g["ship_A"], g["ship_next"]
    X = np.concatenate([S, A], 1)
    lam = g["ship_lam"]; sf = g["ship_sf"]
    sn2 = float(np.float32(float(g["ship_sn"][0]) ** 2))
    fits = [orc.fit(X, nxt[:, a], lam[a], sf[a], sn2 ** 0.5) for a in range(2)]
    for i in range(4):
        U = g[f"ship_U{i}"]
        cc, grad, cm, cv = orc.c_rollout_cost_grad(X, [f["Ky_inv"] for f in fits], [f["beta"] for f in fits], lam,
                                                   sf, g["ship_x0"], U, -1.0, 2 * np.eye(2), np.zeros((2, 2)))
        ref = float(g[f"ship_cost{i}"])
        assert np.isnan(cc) == np.isnan(ref)
        np.testing.assert_allclose(cm, g[f"ship_means{i}"], rtol=1e-6, atol=1e-8)
        if not np.isnan(ref):
            assert abs(cc - ref) < 1e-6 * abs(ref)
            gref = g[f"ship_grad{i}"]
            assert np.max(np.abs(grad - gref)) < 1e-5 * np.max(np.abs(gref))


def test_cost_known_answers():
    """src/test/test_mpc.py:15-57 (13.532174074852094) and :245-274 (158.2904623779527)."""
    g = golden("cost_kat")
    x = np.array([[1., 1.], [3., 3.]]); u = np.array([[2., 2.]])
    sig = np.array([[[1., 2.], [3., 4.]], [[5., 6.], [7., 8.]]])
    c = orc.cost(x, u, sig, np.array([.5, .5]), np.array([.6, .6]), 1.0, 2 * np.eye(2), np.ones((2, 2)))
    assert abs(c - 13.532174074852094) < 1e-10 and abs(c - float(g["kat_cost"])) < 1e-10
    xt = np.array([5., 4, 3, 2, 1, 0]).reshape(6, 1)
    st = np.array([1 / 6, 1 / 7, 1 / 8, 1 / 9, 1 / 10, 1 / 11]).reshape(6, 1)
    c = orc.cost(xt, np.zeros((5, 1)), st, np.zeros(1), np.zeros(1), -1.0, 2 * np.eye(1), np.zeros((1, 1)))
    assert abs(c - 158.2904623779527) < 1e-9 and abs(c - float(g["kat_state_cost"])) < 1e-9


def test_ref_port_matches_reference():
    """The torch port used for the CPU-baseline timing reproduces the reference's numbers."""
    from oracle.ref_port import RefPortProblem
    g = golden("rollout")
    for name in ("r1", "r3"):
        S, A, nxt = g[f"{name}_S"], g[f"{name}_A"], g[f"{name}_next"]
        Rd = g[f"{name}_Rd"]; Rd = None if Rd.size == 0 else Rd
        m = A.shape[1]
        p = RefPortProblem(np.concatenate([S, A], 1), nxt, g[f"{name}_lam"], g[f"{name}_sf"], g[f"{name}_sn"],
                           float(g[f"{name}_gamma"]), g[f"{name}_Q"], g[f"{name}_R"], Rd, g[f"{name}_last"][:m],
                           g[f"{name}_xref"], g[f"{name}_uref"])
        c, grad, means, vars_ = p.cost_and_grad(g[f"{name}_x0"], g[f"{name}_U"])
        assert abs(c - float(g[f"{name}_cost"])) < 1e-9 * max(1, abs(c))
        np.testing.assert_allclose(grad, g[f"{name}_grad"], rtol=1e-7, atol=1e-9)


def test_marginal_likelihood_and_reference_gradient():
    """ML value, and the reference's autograd gradient checked by central differences of the oracle."""
    g = golden("hyper")
    X, y = g["hy_X"], g["hy_y"]
    lam, sf, sn = g["hy_lam0"], float(g["hy_sf0"]), float(g["hy_sn0"])
    sn_eff = float(np.float32(sn ** 2)) ** 0.5
    ml = orc.marginal_likelihood(X, y, lam, sf, sn_eff)
    assert abs(ml - float(g["hy_ml"])) < 1e-8 * abs(ml)
    h = 1e-6
    for k in range(3):
        e = np.zeros(3); e[k] = h
        fd = (orc.marginal_likelihood(X, y, lam * np.exp(e), sf, sn_eff) - orc.marginal_likelihood(X, y, lam * np.exp(-e), sf, sn_eff)) / (2 * h)
        assert abs(fd - g["hy_dlam"][k]) < 1e-5 * max(1, abs(fd))
    fd = (orc.marginal_likelihood(X, y, lam, sf * np.exp(h), sn_eff) - orc.marginal_likelihood(X, y, lam, sf * np.exp(-h), sn_eff)) / (2 * h)
    assert abs(fd - float(g["hy_dsf"])) < 1e-5 * abs(fd)


# ---- round 2: full-covariance rollout and the staged reference ------------------------------------
def _fullcov_setup(g, name):
    X, y, lam = g[f"{name}_X"], g[f"{name}_y"], g[f"{name}_lam"]
    sn = g[f"{name}_sn"]
    fits = [orc.fit(X, y, lam[a], 1.0, float(np.float32(float(sn[a]) ** 2)) ** 0.5) for a in range(2)]
    return X, lam, [f["Ky_inv"] for f in fits], [f["beta"] for f in fits]


@pytest.mark.parametrize("name", ["fc1", "fc2"])
@pytest.mark.parametrize("use_c", [False, True])
def test_full_covariance_rollout_vs_reference_numpy_functions(name, use_c):
    """oracle.rollout_full / the C restatement against rollouts assembled from the reference's NumPy mean_prop /
    variance_prop / covariance_prop (shared and distinct length-scales), and the reference's NumPy cost on the full Sigma."""
    g = golden("fullcov")
    X, lam, Kis, betas = _fullcov_setup(g, name)
    means, covs = orc.rollout_full(X, Kis, betas, lam, np.ones(2), g[f"{name}_x0"], g[f"{name}_U"], use_c=use_c)
    np.testing.assert_allclose(means, g[f"{name}_means"], rtol=1e-8, atol=1e-11)
    np.testing.assert_allclose(covs, g[f"{name}_covs"], rtol=1e-6, atol=1e-10)
    for key, gamma in (("cost_gm1", -1.0), ("cost_gp07", 0.7)):
        c = orc.cost(means, g[f"{name}_U"], covs, g[f"{name}_xref"], g[f"{name}_uref"], gamma, g[f"{name}_Q"], g[f"{name}_R"])
        assert abs(c - float(g[f"{name}_{key}"])) <= 1e-8 * abs(float(g[f"{name}_{key}"]))


def test_full_covariance_step_equals_variance_only_step_on_the_diagonal():
    """With a diagonal input covariance the diagonal of the full-covariance step is the variance-only step."""
    rng = np.random.default_rng(3)
    n, E, m = 80, 3, 1
    D = E + m
    X = rng.uniform(-1, 1, (n, D)); Y = rng.normal(size=(n, E)) * 0.3
    lam = rng.uniform(0.8, 2.0, (E, D)); sf = np.array([1.0, 1.2, 0.9])
    fits = [orc.fit(X, Y[:, a], lam[a], sf[a], 0.2) for a in range(E)]
    u = rng.normal(size=D) * 0.3; s = rng.uniform(1e-3, 0.05, D)
    mean, cov = orc.moment_match_full(X, [f["Ky_inv"] for f in fits], [f["beta"] for f in fits], lam, sf, u, np.diag(s), use_c=True)
    for a in range(E):
        mo, vo, _ = orc.c_moment_match_diag(X, fits[a]["Ky_inv"], fits[a]["beta"], lam[a], sf[a], u, s)
        assert abs(mean[a] - mo) < 1e-12 and abs(cov[a, a] - vo) < 1e-11


def test_staged_reference_matches_the_oracle():
    """oracle/_ref (what `bench.py --impl reference` times) is the reference: its objective / gradient agree with the
    C oracle on a seeded problem.  Skipped where the reference checkout was never staged."""
    import os
    import subprocess
    import sys
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if not os.path.exists(os.path.join(root, "oracle", "_ref", "src", "mpc.py")):
        pytest.skip("oracle/_ref not staged")
    r = subprocess.run([sys.executable, os.path.join(root, "oracle", "ref_runner.py"), "--n", "96", "--H", "2", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-800:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    sys.path.insert(0, root)
    from bench import synth
    n, E, m, H = 96, 4, 1, 2
    S, A, nxt, rng = synth(n, E, m, 0)
    x0 = rng.uniform(-0.5, 0.5, E)
    U = np.random.default_rng(1).uniform(-0.3, 0.3, (1, H * m))[0].reshape(H, m)
    X = np.concatenate([S, A], 1)
    lam = np.full((E, E + m), 2.0)
    fits = [orc.fit(X, nxt[:, a], lam[a], 1.0, float(np.float32(0.1 ** 2)) ** 0.5) for a in range(E)]
    c, _, _, _ = orc.c_rollout_cost_grad(X, [f["Ky_inv"] for f in fits], [f["beta"] for f in fits], lam, np.ones(E),
                                         x0, U, -1.0, 2 * np.eye(E), 0.01 * np.eye(m))
    assert abs(out["runs"]["2"]["cost"] - c) <= 1e-8 * max(1.0, abs(c)), (out["runs"]["2"]["cost"], c)
