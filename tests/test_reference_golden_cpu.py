"""The oracle against outputs of the REFERENCE ITSELF (tests/golden/reference_v1.npz).

The fixture was produced by importing /root/reference/mocat -- the reference's own, unmodified source -- with a NumPy
stand-in for jax on sys.path (tests/golden/make_reference_golden.py, tests/golden/jaxshim/README.md) and evaluating its
deterministic functions.  This pins `oracle/` on numbers the reference code computed, beyond the known-answer tests the
reference ships; the reference is not needed (and not read) to run this file.
"""
import os

import numpy as np
import numpy.testing as npt
import pytest

from oracle import core, metrics as ometrics, models, online_smoothing as oos, pf as opf, svgd as osvgd, teki as oteki

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def R():
    return np.load(os.path.join(HERE, "golden", "reference_v1.npz"))


def test_gaussian_potential_variants(R):                                   # utils.py:49-106
    x, mean, prec, sq = R["gp_x"], R["gp_mean"], R["gp_prec"], R["gp_sqrt_prec"]
    npt.assert_allclose([models.gaussian_potential(v, mean, prec=2.5) for v in x], R["gp_scalar"], rtol=1e-12)
    npt.assert_allclose([models.gaussian_potential(v, mean, prec=np.array([0.5, 1.5, 2.0])) for v in x], R["gp_diag"],
                        rtol=1e-12)
    npt.assert_allclose([models.gaussian_potential(v, mean, prec=prec, det_prec=np.linalg.det(prec)) for v in x],
                        R["gp_full"], rtol=1e-12)
    npt.assert_allclose(models.gaussian_potential(x, mean, sqrt_prec=sq, det_prec=np.linalg.det(prec)), R["gp_sqrt"],
                        rtol=1e-12)


def test_ess_and_tempering_search(R):                                      # metrics.py:69-78, utils.py:205-237, smc.py:311-326
    lw = R["ess_lw"]
    assert core.log_ess_log_weight(lw) == pytest.approx(float(R["ess_log"]), rel=1e-12)
    assert core.ess_log_weight(lw) == pytest.approx(float(R["ess_lin"]), rel=1e-12)
    lik, lw0 = R["bis_lik"], R["bis_lw"]
    target = np.log(0.9 * core.ess_log_weight(lw0))
    b, e, it = core.bisect(lambda x: core.log_ess_log_weight(lw0 - (x - 0.2) * lik) - target, [0.2, 1.0], 1000, 1e-5)
    npt.assert_allclose(b, R["bis_bounds"], rtol=1e-12)
    npt.assert_allclose(e, R["bis_evals"], rtol=1e-7, atol=1e-12)
    assert it == int(R["bis_iters"])
    beta, _ = core.next_temperature_adaptive(lw0, lik, 0.2, 1.0, core.ess_log_weight(lw0), retain=0.9)
    assert beta == pytest.approx(float(R["smc_next_temperature"]), rel=1e-12)


def test_gaussian_kernel_bandwidths_and_svgd_interaction(R):              # kernels.py:82-116,220-229; svgd.py:18-32
    a, b = R["k_a"], R["k_b"]
    assert ometrics.gaussian_call(a, b, 1.3) == pytest.approx(float(R["k_val"]), rel=1e-12)
    npt.assert_allclose(ometrics.gaussian_grad_x(a, b, 1.3), R["k_grad_x"], rtol=1e-12)
    npt.assert_allclose(ometrics.gaussian_grad_y(a, b, 1.3), R["k_grad_y"], rtol=1e-12)
    npt.assert_allclose(ometrics.gaussian_diag_grad_xy(a, b, 1.3), R["k_diag_grad_xy"], rtol=1e-12)
    X, G = R["bw_X"], R["svgd_G"]
    assert osvgd.median_bandwidth(X) == pytest.approx(float(R["bw_median"]), rel=1e-9)
    assert osvgd.mean_bandwidth(X) == pytest.approx(float(R["bw_mean"]), rel=1e-9)
    npt.assert_allclose(osvgd.phi(X, G, 0.9), R["svgd_phi"], rtol=1e-9, atol=1e-12)
    npt.assert_allclose(osvgd.phi_double_loop(X, G, 0.9), R["svgd_phi"], rtol=1e-9, atol=1e-12)


def test_ksd(R):                                                           # metrics.py:88-130
    X, G = R["bw_X"], R["svgd_G"]
    assert ometrics.ksd(X, G, 1.1) == pytest.approx(float(R["ksd_plain"]), rel=1e-9)
    assert ometrics.ksd(X, G, 1.1, log_weight=R["ksd_lw"]) == pytest.approx(float(R["ksd_weighted"]), rel=1e-9)


def test_teki_covariances_and_adaptive_temperature(R):                    # teki.py:20-35,168-185
    cx, cxy, cy = oteki.calculate_covariances(R["teki_vals"], R["teki_sim"])
    npt.assert_allclose(cx, R["teki_cov_x"], rtol=1e-10, atol=1e-14)
    npt.assert_allclose(cxy, R["teki_cov_xy"], rtol=1e-10, atol=1e-14)
    npt.assert_allclose(cy, R["teki_cov_y"], rtol=1e-10, atol=1e-14)

    class Sc:
        data, dim = R["teki_data"], 4
    s = oteki.TemperedEKI(Sc, len(R["teki_vals"]), 0, adaptive=True, ess_threshold=0.8)
    temp = s._next_temperature(dict(sim=R["teki_sim"], temperature=0.1, iter=1), R["teki_prec"])
    assert temp == pytest.approx(float(R["teki_next_temperature"]), rel=1e-10)


def test_linear_gaussian_potentials_and_kalman_filter(R):                 # linear_gaussian.py:73-84,118-128; kalman.py:16-57
    lg = models.LinearGaussianSSM(np.zeros(3), np.eye(3), R["lg_F"], R["lg_Q"], R["lg_H"], R["lg_R"])
    npt.assert_allclose(oos.transition_potential(lg, R["lg_x0"], R["lg_x1"]), R["lg_transition_potential"], rtol=1e-10)
    npt.assert_allclose(lg.likelihood_potential(R["lg_x1"], R["lg_y"]), R["lg_likelihood_potential"], rtol=1e-10)
    means, covs, _ = opf.kalman_filter(lg, R["kalman_y"])
    npt.assert_allclose(means, R["kalman_mean"], rtol=1e-9, atol=1e-12)
    npt.assert_allclose(covs, R["kalman_cov"], rtol=1e-9, atol=1e-12)


def test_lorenz96_field_flow_and_potentials(R):                           # lorenz96.py:14-44; nonlinear_gaussian.py:98-121
    x = R["l96_x"]
    npt.assert_allclose(models.lorenz96_rhs(x, 8.0), R["l96_rhs"], rtol=1e-12)
    npt.assert_allclose(models.lorenz96_dopri(x, 0.05, 8.0), R["l96_flow"], rtol=1e-7)      # two adaptive integrations
    # the device convention (one classical RK4 step) against the reference's adaptive flow
    assert np.max(np.abs(models.lorenz96_rk4(x, 0.05, 8.0, 1) - R["l96_flow"])) < 2e-2
    assert np.max(np.abs(models.lorenz96_rk4(x, 0.05, 8.0, 5) - R["l96_flow"])) < 5e-5
    s = models.Lorenz96SSM(dim=8)
    npt.assert_allclose(s.likelihood_potential(R["l96_xnew"], R["l96_y"]), R["l96_likelihood_potential"], rtol=1e-10)
    # transition potential given the reference's own flow (the flows differ by the integrator, not by the density)
    r = R["l96_xnew"] - R["l96_flow"]
    npt.assert_allclose(0.5 * np.sum(r * r, axis=1) + 0.5 * 8 * np.log(2 * np.pi), R["l96_transition_potential"], rtol=1e-10)
    # scalars of the optimal proposal (nonlinear_gaussian.py:152-186) for Q = R = P0 = I
    npt.assert_allclose(R["opt_initial_kalman_gain"], 0.5 * np.eye(8), atol=1e-12)
    npt.assert_allclose(R["opt_proposal_kalman_gain"], 0.5 * np.eye(8), atol=1e-12)
    npt.assert_allclose(R["opt_proposal_covariance_sqrt"], np.sqrt(0.5) * np.eye(8), atol=1e-12)
    npt.assert_allclose(np.abs(R["opt_weight_precision_sqrt"]), np.sqrt(0.5) * np.eye(8), atol=1e-12)


def test_gk_simulator_and_abc_adaptation(R):                              # gk.py:68-96; abc/smc.py:94-98,163-166
    gk = models.GKTransformed(np.zeros(8))
    npt.assert_allclose(gk.constrain(R["gk_x"]), R["gk_constrain"], rtol=1e-12)
    npt.assert_allclose(gk.simulate(R["gk_x"], R["gk_u01"]), R["gk_summary"], rtol=1e-9, atol=1e-10)
    thr = core.quantile_linear(R["abc_dist"], 0.9 * 380.0 / 501)
    assert thr == pytest.approx(float(R["abc_next_threshold"]), rel=1e-12)
    _, var = core.colstats(R["abc_value"])
    npt.assert_allclose(var / 4 * 2.38 ** 2, R["abc_stepsize"], rtol=1e-10)


def test_rastrigin_likelihood_potential(R):                               # scenarios/toy_examples.py:135-149
    u, _ = models.Rastrigin(5, 1.3).potential_and_grad(R["ras_x"])
    npt.assert_allclose(u, R["ras_likelihood_potential"], rtol=1e-10)


def test_host_cdict_behaves_like_the_reference_container(R):             # core.py:20-84, the PRODUCT's host container
    """mocat_b200.core.cdict (NumPy-backed, written from the interface) against what the reference's own cdict returned
    for the same operations: indexing by int / index array / slice and `+`, with a nested cdict (indexed / appended),
    a static_cdict (left alone), the scalar `time` (added) and a plain python scalar (kept)"""
    from mocat_b200.core import cdict, static_cdict
    mk = lambda o: cdict(value=np.arange(15.0).reshape(5, 3) + o, potential=np.arange(5.0) * 2 + o, time=1.5 + o, label=3,  # noqa: E731
                         inner=cdict(alpha=np.arange(5.0) / 10 + o), frozen=static_cdict(beta=np.arange(5.0) + o))
    c1, c2 = mk(0.0), mk(100.0)
    for name, idx in {"int": 2, "arr": np.array([0, 3, 3]), "slice": slice(1, 4)}.items():
        r = c1[idx]
        npt.assert_array_equal(r.value, R[f"cdict_get_{name}_value"])
        npt.assert_array_equal(r.potential, R[f"cdict_get_{name}_potential"])
        npt.assert_array_equal(r.inner.alpha, R[f"cdict_get_{name}_inner_alpha"])
        npt.assert_array_equal(r.frozen.beta, R[f"cdict_get_{name}_frozen_beta"])
        assert r.time == float(R[f"cdict_get_{name}_time"]) and r.label == int(R[f"cdict_get_{name}_label"])
    a = c1 + c2
    npt.assert_array_equal(a.value, R["cdict_add_value"])
    npt.assert_array_equal(a.potential, R["cdict_add_potential"])
    npt.assert_array_equal(a.inner.alpha, R["cdict_add_inner_alpha"])
    npt.assert_array_equal(a.frozen.beta, R["cdict_add_frozen_beta"])
    assert a.time == float(R["cdict_add_time"]) and a.label == int(R["cdict_add_label"])
    assert (c1 + None).potential is c1.potential                          # core.py:61-62
