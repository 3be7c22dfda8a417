"""The reference's own end-to-end tests, run through the drop-in host API on the device
(mocat/src/tests/test_transport.py, test_ssm.py, test_ssm_linear_gaussian.py, test_abc_gk.py,
test_kernels.py) -- same scenarios, sample sizes and tolerances."""
import numpy as np
import numpy.testing as npt
import pytest

pytestmark = pytest.mark.gpu

COV = np.array([[1., 0.9], [0.9, 2.]])
POST_COV = np.linalg.inv(np.linalg.inv(COV) + np.eye(2) / 49.0)


@pytest.fixture(scope="module")
def mocat(lib):
    import mocat_b200
    return mocat_b200


def _scenario(mocat):
    # tests/test_transport.py:26-28: Gaussian(covariance), prior_sample = 7 z, prior_potential = .5 (x/7^2)^2
    return mocat.scenarios.Gaussian(covariance=COV, prior_std=7.0, prior_pscale=1 / 49.0)


def _resample_final(sample, n, seed=1):
    from oracle import core
    cdf = core.cdf_from_log_weights(sample.log_weight[-1])
    return sample.value[-1][core.ancestors_multinomial(cdf, np.random.default_rng(seed).random(n))]


def _check_smc(mocat, sample, n):
    vals = _resample_final(sample, n)                                    # test_transport.py:57-62
    npt.assert_array_almost_equal(vals.mean(0), np.zeros(2), decimal=0)
    npt.assert_array_almost_equal(np.cov(vals.T), POST_COV, decimal=1)
    lik_prec = np.linalg.inv(COV)
    dets = np.array([np.linalg.det(np.linalg.inv(lik_prec * t + np.eye(2) / 49.0)) for t in sample.temperature])
    npt.assert_array_almost_equal(sample.log_norm_constant, 0.5 * (np.log(dets) - 4 * np.log(7)), 0)   # :49-55


PRESCHEDULE = np.arange(0., 1.1, 0.1)


def test_tempered_preschedule_RW(mocat):
    n = 10_000
    sample = mocat.run(_scenario(mocat), mocat.MetropolisedSMCSampler(mocat.RandomWalk(stepsize=1.0)), n,
                       random_key=np.array([0, 0], np.uint32), temperature_schedule=PRESCHEDULE)
    _check_smc(mocat, sample, n)
    npt.assert_array_equal(sample.temperature, PRESCHEDULE[1:])          # test_transport.py:134
    assert sample.value.shape == (len(PRESCHEDULE) - 1, n, 2)
    assert sample.time > 0 and sample.summary.sampler == "Metropolised SMC Sampler"


def test_tempered_preschedule_UD(mocat):
    n = 10_000
    sample = mocat.run(_scenario(mocat), mocat.MetropolisedSMCSampler(mocat.Underdamped(stepsize=1.0)), n,
                       random_key=0, temperature_schedule=PRESCHEDULE)
    _check_smc(mocat, sample, n)
    npt.assert_array_equal(sample.temperature, PRESCHEDULE[1:])


def test_tempered_adaptive_RW(mocat):
    n = 10_000
    sample = mocat.run(_scenario(mocat), mocat.MetropolisedSMCSampler(mocat.RandomWalk(stepsize=1.0)), n, random_key=0)
    _check_smc(mocat, sample, n)
    assert sample.temperature[-1] == 1.0


def test_tempered_adaptive_UD(mocat):
    n = 10_000
    sample = mocat.run(_scenario(mocat),
                       mocat.MetropolisedSMCSampler(mocat.Underdamped(stepsize=1.0, leapfrog_steps=10)), n, random_key=0)
    _check_smc(mocat, sample, n)


def test_host_buffers_in_and_history_off(mocat):
    n = 50_000
    x0 = (np.random.default_rng(0).standard_normal((n, 5)) * 3).astype(np.float32)
    sc = mocat.scenarios.Rastrigin(dim=5, a=1.0, prior_std=3.0)
    sample = mocat.run(sc, mocat.MetropolisedSMCSampler(mocat.Underdamped(stepsize=0.1), max_iter=5,
                                                        resampling='systematic', keep_history=False),
                       n, random_key=1, initial_state=mocat.cdict(value=x0))
    assert sample.value.shape == (1, n, 5) and len(sample.temperature) == 6
    assert np.all(np.diff(sample.temperature) > 0)
    # potential() through the device kernel equals prior + T * likelihood of the returned fields
    sc.temperature = float(sample.temperature[-1])
    U = sc.potential(sample.value[0][:100])
    npt.assert_allclose(U, sample.potential[0][:100], rtol=2e-5, atol=1e-4)


# ------------------------------------------------------------------------------------------------ SVGD
def _check_moments(vals):
    npt.assert_array_almost_equal(vals.mean(0), np.zeros(2), decimal=0)
    npt.assert_array_almost_equal(np.cov(vals.T), POST_COV, decimal=1)


def test_svgd_mean_bandwidth_default(mocat):                              # test_transport.py:69-74
    sample = mocat.run(_scenario(mocat), mocat.SVGD(max_iter=1000, stepsize=0.8), n=100, random_key=0)
    assert sample.value.shape == (1001, 100, 2)
    _check_moments(sample[-1].value)


def test_svgd_median_update_subclass(mocat):                              # test_transport.py:76-89
    class SVGD_median(mocat.SVGD):
        def adapt(self, ensemble_state, ensemble_extra):
            ensemble_extra.parameters.kernel_params.bandwidth = mocat.kernels.median_bandwidth_update(ensemble_state.value)
            ensemble_state.kernel_params = ensemble_extra.parameters.kernel_params
            return ensemble_state, ensemble_extra

    sample = mocat.run(_scenario(mocat), SVGD_median(max_iter=1000, stepsize=1.0), n=100, random_key=0)
    _check_moments(sample[-1].value)


def test_svgd_callable_stepsize(mocat):                                   # test_transport.py:106-111
    sample = mocat.run(_scenario(mocat), mocat.SVGD(max_iter=1000, stepsize=lambda i: 10 * i ** -0.5), n=100,
                       random_key=0)
    _check_moments(sample[-1].value)


def test_gaussian_kernel_kat(mocat):                                      # test_kernels.py:16-29
    k = mocat.kernels.Gaussian()
    z, o = np.zeros(5), np.ones(5)
    npt.assert_array_almost_equal(k(z, z), 1.)
    npt.assert_array_almost_equal(k(z, o), 0.082085006)
    npt.assert_array_almost_equal(k.grad_x(z, o), np.ones(5) * 0.082085006)
    npt.assert_array_almost_equal(k.grad_y(z, o), np.ones(5) * -0.082085006)


# ------------------------------------------------------------------------------------------------ SSM
@pytest.mark.parametrize("dim", [1, 5])
def test_bootstrap_pf_coverage_and_kalman(mocat, dim):
    """tests/test_ssm.py:33-44 (truth between particle min and max at every t, n=2e3, T=20) and
    tests/test_ssm_linear_gaussian.py:71-104 (Kalman mean within 3/4 sigma of the truth)."""
    ssm = mocat.ssm.TimeHomogenousLinearGaussian(initial_mean=np.zeros(dim), initial_covariance=np.eye(dim),
                                                 transition_matrix=np.eye(dim), transition_covariance=np.eye(dim),
                                                 likelihood_matrix=np.eye(dim), likelihood_covariance=np.eye(dim))
    t = np.arange(20, dtype=np.float64)
    sim = ssm.simulate(t, random_key=0)
    pf = mocat.ssm.run_particle_filter_for_marginals(ssm, mocat.ssm.BootstrapFilter(), sim.y, t, random_key=0, n=2000)
    assert pf.value.shape == (20, 2000, dim) and pf.log_weight.shape == (20, 2000) and pf.ess.shape == (20,)
    assert np.all(pf.value.min(1) <= sim.x) and np.all(sim.x <= pf.value.max(1))
    mus, covs, ll = mocat.ssm.run_kalman_filter_for_marginals(ssm, sim.y, t, return_log_likelihood=True)
    sd = np.sqrt(np.einsum('tii->ti', covs))
    k = 3 if dim == 1 else 4
    assert np.all(np.abs(mus - sim.x) < k * sd)
    npt.assert_allclose(pf.mean, mus, atol=0.35 if dim == 1 else 0.8)   # ESS ~ 100-250 in 5-D at n=2e3
    assert abs(pf.log_norm_constant[-1] - ll) < 1.0 + dim
    # online API (filtering.py:173-252): same result as the batch call when driven step by step
    p = mocat.ssm.initiate_particles(ssm, mocat.ssm.BootstrapFilter(), 2000, 0, sim.y[0], t[0])
    for i in range(1, 5):
        p = mocat.ssm.propagate_particle_filter(ssm, mocat.ssm.BootstrapFilter(), p, sim.y[i], t[i], 0)
    npt.assert_allclose(p.ess, pf.ess[:5], rtol=1e-6)
    npt.assert_array_equal(p.value, pf.value[:5])


def test_lorenz96_filter_runs(mocat):
    from oracle import models as omodels
    d = 40
    _, y = omodels.Lorenz96SSM(dim=d).simulate(10, np.random.default_rng(0), spinup=300)
    ssm = mocat.ssm.Lorenz96(dim=d)
    out = mocat.ssm.run_particle_filter_for_marginals(ssm, mocat.ssm.BootstrapFilter(), y, np.arange(10) * 0.05,
                                                      random_key=3, n=100_000, ess_threshold=2.0,
                                                      resampling='systematic', keep_history=False)
    assert out.value.shape == (1, 100_000, d) and np.all(np.isfinite(out.log_norm_constant))
    assert np.all(out.resampled[1:] == 1)


# ------------------------------------------------------------------------------------------------ ABC
def test_abc_gk(mocat):
    """tests/test_abc_gk.py:130-164 style: SMC-ABC recovers the g-and-k parameters: sum|mean - truth| < 3
    (n=1e3 there with a 100-order-statistic summary; here the m=8 sorted-draw summary of config C5)."""
    from oracle import models as omodels
    truth = np.array([3., 1., 2., .5])
    sc0 = mocat.abc.GKTransformedUniformPrior(n_unsummarised_data=8)
    x_true = sc0.unconstrain(truth)
    data = omodels.GKTransformed(np.zeros(8)).simulate(x_true[None], np.random.default_rng(1).random((1, 8)))[0]
    sc = mocat.abc.GKTransformedUniformPrior(data=data)
    sample = mocat.run(sc, mocat.abc.MetropolisedABCSMCSampler(max_iter=60), 5000, random_key=0)
    assert np.all(np.diff(sample.threshold[1:]) <= 1e-9)
    alive = sample.log_weight[-1] > -np.inf
    post_mean = sc.constrain(sample.value[-1][alive]).mean(0)
    assert np.abs(post_mean - truth).sum() < 6.0      # m=8 draws carry far less information than 1000 draws
    assert np.abs(post_mean[0] - truth[0]) < 1.5      # location is identified even from 8 draws


# ------------------------------------------------------------------------------------------------ config C4
def _logistic_data(N=257, d=11, seed=3):
    rng = np.random.default_rng(seed)
    A = rng.normal(size=(N, d)).astype(np.float32)
    w_true = rng.normal(size=d)
    t = (rng.random(N) < 1 / (1 + np.exp(-A @ w_true))).astype(np.float32)
    return A, t


@pytest.mark.parametrize("d", [3, 11, 50])
def test_logistic_regression_potential_grad(mocat, d):
    from oracle import models
    A, t = _logistic_data(N=257, d=d)
    sc = mocat.scenarios.LogisticRegression(A, t, prior_std=2.0)
    ref = models.LogisticRegression(A, t)
    prior = models.IsoGaussianPrior(d, 0.0, 2.0)
    X = np.random.default_rng(0).normal(size=(300, d)).astype(np.float32) * 0.7
    X[0] *= 30.0                                                            # saturated logits: stable softplus
    ul, gl = ref.potential_and_grad(X)
    up, gp = prior.potential_and_grad(X)
    for beta in (1.0, 0.3):
        U, G = sc.tempered_potential_and_grad(X, beta)
        npt.assert_allclose(U, up + beta * ul, rtol=2e-5, atol=1e-3)
        npt.assert_allclose(G, gp + beta * gl, rtol=2e-4, atol=2e-3)


def test_svgd_logistic_regression_matches_oracle(mocat):
    """C4 at test size: 40 SVGD iterations (mean bandwidth, adagrad) against the NumPy restatement"""
    from oracle import models, philox, svgd
    d, n = 11, 128
    A, t = _logistic_data(N=257, d=d)
    sc = mocat.scenarios.LogisticRegression(A, t)
    sample = mocat.run(sc, mocat.SVGD(max_iter=40, stepsize=0.05), n=n, random_key=5)
    x0 = philox.normals(5, np.arange(n, dtype=np.uint64), 0, philox.P_INIT, d, dtype=np.float32)
    npt.assert_allclose(sample.value[0], x0, atol=1e-5)
    ref, prior = models.LogisticRegression(A, t), models.IsoGaussianPrior(d, 0.0, 1.0)

    def pg(x):
        ul, gl = ref.potential_and_grad(x)
        up, gp = prior.potential_and_grad(x)
        return up + ul, gp + gl
    o = svgd.SVGD(pg, x0.astype(np.float64), 0.05, bandwidth='mean', max_iter=40)
    o.run()
    npt.assert_allclose(sample.value[-1], o.x, atol=5e-3)
    npt.assert_allclose(sample.bandwidth, o.h, rtol=1e-3)


def test_rm_metropolised_smc_stepsize_adaptation(mocat):
    """RMMetropolisedSMCSampler (transport/smc.py:376-428): the stepsize follows the oracle's Robbins-Monro recursion
    (same Philox streams; the weighted mean acceptance differs by O(1/n) through rare accept/reject flips)"""
    from oracle import models, smc as osmc
    n, d, seed = 4000, 2, 3
    sc = mocat.scenarios.Rastrigin(dim=d, a=1.0, prior_std=3.0)
    smp = mocat.RMMetropolisedSMCSampler(mocat.Underdamped(stepsize=0.3), rm_stepsize=1.0, resampling='systematic',
                                         keep_history=False)
    out = mocat.run(sc, smp, n, random_key=seed)
    orc = osmc.TemperedSMC(models.IsoGaussianPrior(d, 0.0, 3.0), models.Rastrigin(d, 1.0), n, seed, move='mala',
                           stepsize=0.3, resampling='systematic', rm_stepsize=1.0, rm_target=0.651)
    chain = orc.run()
    ref = np.array([0.3] + [c['stepsize'] for c in chain[1:]])
    assert out.stepsize.shape == out.temperature.shape
    assert out.stepsize[0] == pytest.approx(0.3)
    npt.assert_allclose(out.stepsize[1:4], ref[1:4], rtol=2e-2)        # before the particle systems decorrelate
    assert abs(len(out.stepsize) - len(ref)) <= 3
    assert np.all(np.diff(np.log(out.stepsize)) <= 1.0 * (1 - 0.651) + 1e-6)   # |log step| <= rm * (1 - target)
    assert out.temperature[-1] == 1.0


@pytest.mark.parametrize("d,dy", [(1, 1), (3, 2), (8, 8)])
def test_kalman_filter_device_matches_oracle(mocat, d, dy):
    """mb_kalman_filter (ssm/linear_gaussian/kalman.py:16-57 on the device, fp64) against the oracle's recursion:
    means, covariances and the innovation log-likelihood, for full and partial observation"""
    from oracle import models as omodels, pf as opf
    rng = np.random.default_rng(d * 10 + dy)
    A = rng.standard_normal((d, d)); F = 0.9 * A / np.max(np.abs(np.linalg.eigvals(A)))
    B = rng.standard_normal((d, d)); Q = B @ B.T / d + 0.1 * np.eye(d)
    Cm = rng.standard_normal((d, d)); P0 = Cm @ Cm.T / d + 0.5 * np.eye(d)
    H = rng.standard_normal((dy, d))
    E_ = rng.standard_normal((dy, dy)); R = E_ @ E_.T / dy + 0.2 * np.eye(dy)
    m0 = rng.standard_normal(d)
    sc = mocat.ssm.TimeHomogenousLinearGaussian(m0, P0, F, Q, H, R)
    sim = sc.simulate(np.arange(60.0), 5)
    mus, covs, ll = mocat.ssm.run_kalman_filter_for_marginals(sc, sim.y, sim.t, return_log_likelihood=True)
    y32 = sim.y.astype(np.float32).astype(np.float64)                   # the device reads the observations in fp32
    omus, ocovs, oll = opf.kalman_filter(omodels.LinearGaussianSSM(m0, P0, F, Q, H, R), y32)
    # the POD model struct carries fp32 matrices (Cholesky factors of P0, Q; precision root of R): 1e-7 relative inputs
    npt.assert_allclose(mus, omus, rtol=5e-5, atol=5e-5)
    npt.assert_allclose(covs, ocovs, rtol=5e-5, atol=5e-6)
    npt.assert_allclose(ll, oll, rtol=2e-5)
    hm, hc, hl = mocat.ssm.kalman_filter_host(sc, y32, sim.t, return_log_likelihood=True)
    npt.assert_allclose(mus, hm, rtol=5e-5, atol=5e-5)


def test_thinned_history_stream_and_checkpoint_resume(tmp_path):
    """(i) history_every streams a thinned history to pinned host memory and its records equal the
    corresponding records of the full stacked history; (ii) a filter continued from a checkpoint file is bit-identical
    to the uninterrupted run (Philox counters are keyed on particle and time index)"""
    import numpy as np
    import numpy.testing as npt
    from mocat_b200 import ssm, history
    sc = ssm.Lorenz96(dim=8)
    tt = np.arange(12) * 0.05
    sim = sc.simulate(tt, 3, spinup=200)
    kw = dict(n=3001, ess_threshold=0.5, resampling='systematic')
    full = ssm.run_particle_filter_for_marginals(sc, ssm.BootstrapFilter(), sim.y, sim.t, 9, keep_history=True, **kw)
    assert full.value.shape == (12, 3001, 8) and not hasattr(full, 'history_index')
    thin = ssm.run_particle_filter_for_marginals(sc, ssm.BootstrapFilter(), sim.y, sim.t, 9, history_every=5, **kw)
    npt.assert_array_equal(thin.history_index, [0, 5, 10, 11])
    npt.assert_array_equal(thin.value, full.value[[0, 5, 10, 11]])
    npt.assert_array_equal(thin.log_weight, full.log_weight[[0, 5, 10, 11]])
    npt.assert_array_equal(thin.ess, full.ess)
    # checkpoint after 7 observations, continue from the file
    part = ssm.run_particle_filter_for_marginals(sc, ssm.BootstrapFilter(), sim.y[:7], sim.t[:7], 9, keep_history=False, **kw)
    path = tmp_path / "pf.cdict"
    history.save_checkpoint(part, path)
    with pytest.raises(RuntimeError):
        history.save_checkpoint(part, path)                            # core.py:99-103: no silent overwrite
    # disturb the pooled engine, then resume
    ssm.run_particle_filter_for_marginals(sc, ssm.BootstrapFilter(), sim.y[:3], sim.t[:3], 1, keep_history=False, **kw)
    res = history.load_checkpoint(path)
    cont = ssm.run_particle_filter_for_marginals(sc, ssm.BootstrapFilter(), sim.y[7:], sim.t[7:], 9, initial_sample=res,
                                                 keep_history=False, ess_threshold=0.5, resampling='systematic')
    npt.assert_array_equal(cont.engine.values().cpu().numpy(), full.value[-1])
    npt.assert_array_equal(cont.engine.lw.cpu().numpy(), full.log_weight[-1])
    npt.assert_array_equal(cont.ess, full.ess)
    npt.assert_allclose(cont.log_norm_constant, full.log_norm_constant, rtol=0, atol=0)
