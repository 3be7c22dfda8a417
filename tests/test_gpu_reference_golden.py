"""CUDA path against outputs of the REFERENCE ITSELF (tests/golden/reference_v1.npz: /root/reference/mocat executed under a
NumPy stand-in for jax, tests/golden/make_reference_golden.py) -- neither `oracle/` nor the reference is imported here.
fp32 kernels against the reference's algorithms in fp64: tolerances are fp32 round-off of the respective formula."""
import os

import numpy as np
import numpy.testing as npt
import pytest

pytestmark = pytest.mark.gpu
R = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_v1.npz"))


@pytest.fixture(scope="module")
def mocat(lib):
    import mocat_b200
    return mocat_b200


def _t(a):
    import torch
    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float32), device="cuda")


def test_reference_ess(mocat):                                             # metrics.py:69-78
    assert mocat.log_ess_log_weight(R["ess_lw"].astype(np.float32)) == pytest.approx(float(R["ess_log"]), abs=5e-6)
    assert mocat.ess_log_weight(R["ess_lw"].astype(np.float32)) == pytest.approx(float(R["ess_lin"]), rel=1e-5)


def test_reference_gaussian_kernel_and_bandwidths(mocat):                  # kernels.py:82-116,220-229
    k = mocat.kernels.Gaussian(bandwidth=1.3)
    a, b = R["k_a"], R["k_b"]
    assert k(a, b) == pytest.approx(float(R["k_val"]), rel=1e-5)
    npt.assert_allclose(k.grad_x(a, b), R["k_grad_x"], rtol=2e-5)
    npt.assert_allclose(k.grad_y(a, b), R["k_grad_y"], rtol=2e-5)
    X = R["bw_X"].astype(np.float32)
    assert mocat.kernels.median_bandwidth_update(X, 0) == pytest.approx(float(R["bw_median"]), rel=2e-5)
    assert mocat.kernels.mean_bandwidth_update(X, 0) == pytest.approx(float(R["bw_mean"]), rel=2e-5)


def test_reference_svgd_interaction(mocat):                                # transport/svgd.py:18-32
    import torch
    from mocat_b200 import engine
    h = torch.tensor([0.9], dtype=torch.float32, device="cuda")
    phi = engine.svgd_phi(_t(R["bw_X"]), _t(R["svgd_G"]), h, 0).cpu().numpy()
    npt.assert_allclose(phi, R["svgd_phi"], atol=3e-5 * np.abs(R["svgd_phi"]).max(), rtol=1e-4)


def test_reference_ksd(mocat):                                             # metrics.py:88-130
    k = mocat.kernels.Gaussian(bandwidth=1.1)
    X, G = R["bw_X"].astype(np.float32), R["svgd_G"].astype(np.float32)
    assert mocat.metrics.ksd(X, k, grad_potential=G, bandwidth=1.1) == pytest.approx(float(R["ksd_plain"]), rel=3e-4)
    got = mocat.metrics.ksd(X, k, grad_potential=G, log_weight=R["ksd_lw"].astype(np.float32), bandwidth=1.1)
    assert got == pytest.approx(float(R["ksd_weighted"]), rel=3e-4)


def test_reference_linear_gaussian_potentials_and_kalman(mocat):          # linear_gaussian.py:73-84; kalman.py:16-57
    sc = mocat.ssm.TimeHomogenousLinearGaussian(np.zeros(3), np.eye(3), R["lg_F"], R["lg_Q"], R["lg_H"], R["lg_R"])
    pot = mocat.online_smoothing.transition_potential(sc, R["lg_x0"], 0.0, R["lg_x1"], 1.0)
    npt.assert_allclose(pot, R["lg_transition_potential"], rtol=2e-5, atol=2e-5)     # full Q: the reference's convention
    means, covs = mocat.ssm.run_kalman_filter_for_marginals(sc, R["kalman_y"], np.arange(15.0))[:2]
    npt.assert_allclose(means, R["kalman_mean"], rtol=5e-5, atol=5e-5)      # the model struct and y are fp32
    npt.assert_allclose(covs, R["kalman_cov"], rtol=5e-5, atol=5e-6)


def test_reference_lorenz96_transition_potential(mocat):                  # lorenz96.py:14-44; nonlinear_gaussian.py:98-105
    """five RK4 substeps are within 5e-5 of the reference's adaptive Dormand-Prince flow, so the
    potentials agree to |x' - flow| * 5e-5"""
    sc = mocat.ssm.Lorenz96(dim=8, substeps=5)
    pot = mocat.online_smoothing.transition_potential(sc, R["l96_x"], 0.0, R["l96_xnew"], 0.05)
    npt.assert_allclose(pot, R["l96_transition_potential"], atol=2e-3)


def test_reference_rastrigin_potential(mocat):                            # scenarios/toy_examples.py:135-149
    sc = mocat.scenarios.Rastrigin(dim=5, a=1.3, prior_std=3.0)
    got = np.asarray(sc.likelihood_potential(R["ras_x"].astype(np.float32)))
    npt.assert_allclose(got, R["ras_likelihood_potential"], rtol=2e-5, atol=2e-4)


def test_reference_run_lorenz96_particle_filter(mocat):
    """the reference's OWN bootstrap-filter run on Lorenz-96 d = 8 (tests/golden/reference_runs_pf_v1.npz: its adaptive
    flow, multinomial resampling, n = 1000) against the device filter (one RK4 step per interval, n = 20000) on the same
    observations: filter means within the reference run's Monte-Carlo error, ESS fractions step by step"""
    P = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_runs_pf_v1.npz"))
    sc = mocat.ssm.Lorenz96(dim=8, likelihood_std=2.0)
    n = 20000
    out = mocat.ssm.run_particle_filter_for_marginals(sc, mocat.ssm.BootstrapFilter(), P["pf_y"].astype(np.float32), P["pf_t"],
                                                      3, n=n, ess_threshold=0.5, resampling='multinomial')
    d = out.mean - P["pf_mean"]
    assert np.sqrt(np.mean(d ** 2)) < 0.35 and np.abs(d).max() < 1.0, (np.sqrt(np.mean(d ** 2)), np.abs(d).max())
    ratio = (out.ess / n) / (P["pf_ess"] / float(P["pf_n"]))
    assert np.all(ratio > 0.4) and np.all(ratio < 2.5), ratio
    err_d = np.sqrt(np.mean((out.mean - P["pf_x"]) ** 2, axis=1))
    err_r = np.sqrt(np.mean((P["pf_mean"] - P["pf_x"]) ** 2, axis=1))
    assert np.max(np.abs(err_d - err_r)) < 0.2


@pytest.mark.parametrize("name,n", [("opt", 20000), ("enkf", 4000)])
def test_reference_runs_optimal_proposal_and_enkf(mocat, name, n):
    """the reference's OWN optimal-proposal and ensemble-Kalman filter runs (ssm/nonlinear_gaussian.py:134-350; no test
    upstream) against the device filters on the same observations: filter means, spreads, ESS fractions"""
    P = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_runs_pf_v1.npz"))
    sc = mocat.ssm.Lorenz96(dim=8, likelihood_std=2.0)
    filt = mocat.ssm.OptimalNonLinearGaussianParticleFilter() if name == "opt" else mocat.ssm.EnsembleKalmanFilter()
    out = mocat.ssm.run_particle_filter_for_marginals(sc, filt, P["pf_y"].astype(np.float32), P["pf_t"], 5, n=n,
                                                      ess_threshold=0.5, resampling='multinomial')
    d = out.mean - P[name + "_mean"]
    assert np.sqrt(np.mean(d ** 2)) < 0.25 and np.abs(d).max() < 0.7, (np.sqrt(np.mean(d ** 2)), np.abs(d).max())
    vr = out.var.mean(1) / P[name + "_var"].mean(1)
    assert np.all(vr > 0.8) and np.all(vr < 1.25), vr
    ratio = (out.ess / n) / (P[name + "_ess"] / float(P[name + "_n"]))
    assert np.all(ratio > 0.35) and np.all(ratio < 2.5), ratio


def test_reference_run_tempered_smc(mocat):
    """config C2 in small: the reference's OWN run of MetropolisedSMCSampler + RandomWalk on Rastrigin d = 2 (n = 1000,
    tests/golden/reference_runs_smc_v1.npz) against the device sampler (n = 20000): the adaptive temperature ladder entry
    by entry, the ESS pattern with its resampling points, the log normalising constant"""
    S = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_runs_smc_v1.npz"))
    n = 20000
    sc = mocat.scenarios.Rastrigin(dim=2, a=1.0, prior_std=3.0)
    out = mocat.run(sc, mocat.MetropolisedSMCSampler(mocat.RandomWalk(stepsize=0.5), resampling='multinomial'), n, random_key=0)
    beta = np.asarray(out.temperature, np.float64)
    assert len(beta) == len(S["smc_temperature"])
    npt.assert_allclose(beta[:8], S["smc_temperature"][:8], rtol=0.04)
    npt.assert_allclose(beta[8:], S["smc_temperature"][8:], rtol=0.12)      # the reference run's Monte-Carlo state (n = 1000)
    ess_d, ess_r = np.asarray(out.ess) / n, S["smc_ess"] / float(S["smc_n"])
    npt.assert_allclose(ess_d[:-1], ess_r[:-1], atol=1e-3)
    assert abs(ess_d[-1] - ess_r[-1]) < 0.08
    npt.assert_allclose(out.log_norm_constant, S["smc_log_norm_constant"], atol=0.15)


def test_reference_run_svgd(mocat):
    """transport/svgd.py end to end: the reference's OWN 15 SVGD iterations from a given ensemble (deterministic:
    adagrad, mean bandwidth re-adapted every iteration, full-covariance Gaussian target) against the device run from the
    same ensemble -- fp32 round-off accumulated over the iterations"""
    V = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_runs_svgd_v1.npz"))
    sc = mocat.scenarios.Gaussian(mean=V["svgd_mean"], covariance=V["svgd_cov"], prior_std=2.0)
    out = mocat.run(sc, mocat.SVGD(stepsize=0.1, max_iter=15, keep_history=True), 100, random_key=0,
                    initial_state=mocat.cdict(value=V["svgd_X0"].astype(np.float32)))
    assert out.value.shape == V["svgd_value"].shape
    npt.assert_allclose(out.value[0], V["svgd_value"][0], atol=1e-6)
    npt.assert_allclose(out.value, V["svgd_value"], atol=2e-4)
    assert out.bandwidth == pytest.approx(float(V["svgd_bandwidth"][-1]), rel=1e-4)
    npt.assert_allclose(out.potential[-1], V["svgd_potential"][-1], rtol=1e-4, atol=1e-4)


def test_reference_run_smc_abc(mocat):
    """config C5 in small: the reference's OWN SMC-ABC run on the g-and-k model (n = 1000, 12 iterations) against the
    device sampler (n = 20000): ESS pattern with its resampling point, thresholds on the log scale, mean acceptance"""
    A = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_runs_abc_v1.npz"))
    n = 20000
    sc = mocat.abc.GKTransformedUniformPrior(data=A["abc_data"])
    out = mocat.run(sc, mocat.abc.MetropolisedABCSMCSampler(max_iter=12, keep_history=True), n, random_key=0)
    assert len(out.threshold) == len(A["abc_threshold"])
    npt.assert_allclose(np.asarray(out.ess) / n, A["abc_ess"] / float(A["abc_n"]), atol=2e-3)
    dlog = np.log(out.threshold) - np.log(A["abc_threshold"])
    assert np.all(np.abs(dlog[:7]) < 0.7) and np.all(np.abs(dlog[7:]) < 0.3), dlog
    npt.assert_allclose(out.alpha.mean(axis=-1)[1:], A["abc_alpha_mean"][1:], atol=0.05)


def test_reference_run_rm_metropolised_smc(mocat):
    """the reference's OWN run of RMMetropolisedSMCSampler + MALA (n = 500, tests/golden/reference_runs_rm_v1.npz) against
    the device sampler (n = 20000, stepsize adapted on the device): stepsize trajectory, ladder, evidence"""
    S = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_runs_rm_v1.npz"))
    sc = mocat.scenarios.Rastrigin(dim=2, a=1.0, prior_std=3.0)
    smp = mocat.RMMetropolisedSMCSampler(mocat.Underdamped(stepsize=0.3), rm_stepsize=1.0, resampling='multinomial')
    out = mocat.run(sc, smp, 20000, random_key=0)
    assert len(out.temperature) == len(S["rm_temperature"])
    npt.assert_allclose(out.stepsize[:9], S["rm_stepsize"][:9], rtol=0.03)
    npt.assert_allclose(out.stepsize[9:], S["rm_stepsize"][9:], rtol=0.12)
    npt.assert_allclose(out.temperature[:6], S["rm_temperature"][:6], rtol=0.04)
    npt.assert_allclose(out.temperature[6:], S["rm_temperature"][6:], rtol=0.1)
    assert abs(out.log_norm_constant[-1] - S["rm_log_norm_constant"][-1]) < 0.35
