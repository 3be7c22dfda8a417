"""End-to-end runs of the REFERENCE's samplers (tests/golden/reference_runs_v1.npz, produced by
tests/golden/make_reference_runs.py: the reference's own source under a NumPy stand-in for jax) against the oracle and
against exact answers.  Statistical comparisons only -- the stand-in's random streams are neither jax's nor ours."""
import os

import numpy as np
import pytest

from oracle import models, teki as oteki

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def R():
    return np.load(os.path.join(HERE, "golden", "reference_runs_v1.npz"))


def test_teki_oracle_matches_the_reference_run(R):
    """transport/teki.py end to end on the g-and-k simulator (prior U(0, 2)^4, 8 sorted draws): temperature ladder and
    posterior moments of the constrained parameters, reference run (n = 1000) vs oracle run (n = 2000)"""
    sc = models.GKTransformed(R["teki_data"], prior_max=2.0)
    ad = oteki.TemperedEKI(sc, 2000, 1, adaptive=True, ess_threshold=0.9).run()
    t_ref, t_or = R["teki_adaptive_temperature"], ad["temperature_schedule"]
    assert t_ref[0] == 0.0 and t_ref[-1] == 1.0 and np.all(np.diff(t_ref) > 0)
    assert abs(len(t_ref) - len(t_or)) <= 3
    assert abs(t_ref[1] - t_or[1]) < 0.06                               # first adaptive temperature (ESS = 0.9 n)
    post = sc.constrain(ad["x"])
    np.testing.assert_allclose(post.mean(0), R["teki_adaptive_mean"], atol=0.08)
    np.testing.assert_allclose(post.std(0), R["teki_adaptive_std"], atol=0.06)
    fx = oteki.TemperedEKI(sc, 2000, 2, temperature_schedule=np.linspace(0.0, 1.0, 11)).run()
    np.testing.assert_allclose(R["teki_schedule_temperature"], np.linspace(0.0, 1.0, 11), atol=1e-12)
    np.testing.assert_allclose(fx["temperature_schedule"], R["teki_schedule_temperature"], atol=1e-12)
    post = sc.constrain(fx["x"])
    np.testing.assert_allclose(post.mean(0), R["teki_schedule_mean"], atol=0.08)
    np.testing.assert_allclose(post.std(0), R["teki_schedule_std"], atol=0.06)


def test_reference_smoothers_against_the_rts_smoother(R):
    """ssm/online_smoothing.py and ssm/backward.py as the reference runs them (n = 300, lag 6): FFBSi is exact to
    Monte-Carlo error; both fixed-lag branches re-sample the lag window at every step, so only a fraction of distinct
    values survives at interior times -- the path degeneracy the device tests (tests/test_gpu_backward.py) account for"""
    sm, var = R["lg_rts_mean"], R["lg_rts_var"]
    n = int(R["smoother_n"])
    se = np.sqrt(var.max() / n)
    assert np.max(np.abs(R["ffbsi_mean"] - sm)) < 5 * se
    assert np.median(R["ffbsi_unique_fraction"][:-1]) > 0.2
    for name, frac_hi in (("pf", 0.2), ("bs", 0.45)):
        err = np.abs(R[f"smoother_{name}_mean"] - sm)
        uf = R[f"smoother_{name}_unique_fraction"]
        assert err.max() < 0.8 and np.median(err) < 0.25, (name, err.max())
        assert np.median(uf[2:-7]) < frac_hi, (name, uf)                # interior times: degenerate
        assert uf[-1] > np.median(uf[2:-7])                             # the newest slice has been re-sampled once, not `lag` times
        assert abs(R[f"smoother_{name}_var"][-1] - var[-1]) < 0.5 * var[-1]


def test_lorenz96_particle_filter_oracle_matches_the_reference_run():
    """config C3 in small (tests/golden/reference_runs_pf_v1.npz): the reference's bootstrap filter on Lorenz-96 d = 8
    (its adaptive Dormand-Prince flow, multinomial resampling, n = 1000) against the oracle (one RK4 step per interval, the
    device convention, n = 20000) on the same observations: filter means within the reference run's Monte-Carlo error,
    the ESS fractions step by step, and the same error against the simulated truth"""
    from oracle import pf as opf
    P = np.load(os.path.join(HERE, "golden", "reference_runs_pf_v1.npz"))
    s = models.Lorenz96SSM(dim=8, r_std=2.0)
    n = 20000
    out = opf.BootstrapPF(s, n, 1, ess_threshold=0.5, resampling='multinomial').run(P["pf_y"])
    means = np.array([opf.weighted_moments(o['x'], o['lw'])[0] for o in out])
    d = means - P["pf_mean"]
    assert np.sqrt(np.mean(d ** 2)) < 0.35 and np.abs(d).max() < 1.0
    ratio = (np.array([o['ess'] for o in out]) / n) / (P["pf_ess"] / float(P["pf_n"]))
    assert np.all(ratio > 0.4) and np.all(ratio < 2.5), ratio
    err_o = np.sqrt(np.mean((means - P["pf_x"]) ** 2, axis=1))
    err_r = np.sqrt(np.mean((P["pf_mean"] - P["pf_x"]) ** 2, axis=1))
    assert np.max(np.abs(err_o - err_r)) < 0.2


@pytest.mark.parametrize("name,n", [("opt", 20000), ("enkf", 4000)])
def test_optimal_proposal_and_enkf_oracles_match_the_reference_runs(name, n):
    """ssm/nonlinear_gaussian.py:134-350 (no test upstream): the reference's optimal-proposal filter (n = 1000) and
    ensemble Kalman filter (n = 500) on the observations of the run above against oracle.pf.OptimalPF / EnKF: filter means,
    spreads and -- for the weighted filter -- the ESS fractions step by step"""
    from oracle import pf as opf
    P = np.load(os.path.join(HERE, "golden", "reference_runs_pf_v1.npz"))
    s = models.Lorenz96SSM(dim=8, r_std=2.0)
    cls = opf.OptimalPF if name == "opt" else opf.EnKF
    out = cls(s, n, 1, ess_threshold=0.5, resampling='multinomial').run(P["pf_y"])
    mom = [opf.weighted_moments(o['x'], o['lw']) for o in out]
    means, var = np.array([m[0] for m in mom]), np.array([m[1] for m in mom])
    d = means - P[name + "_mean"]
    assert np.sqrt(np.mean(d ** 2)) < 0.25 and np.abs(d).max() < 0.7
    vr = var.mean(1) / P[name + "_var"].mean(1)
    assert np.all(vr > 0.8) and np.all(vr < 1.25), vr
    ratio = (np.array([o['ess'] for o in out]) / n) / (P[name + "_ess"] / float(P[name + "_n"]))
    assert np.all(ratio > 0.35) and np.all(ratio < 2.5), ratio


def test_tempered_smc_oracle_matches_the_reference_run():
    """config C2 in small (tests/golden/reference_runs_smc_v1.npz): the reference's MetropolisedSMCSampler with a
    random-walk move on Rastrigin d = 2 (n = 1000) against oracle.smc.TemperedSMC (n = 20000): the adaptive temperature
    ladder entry by entry, the ESS pattern including the resampling points, the log normalising constant and the mean
    acceptance probabilities"""
    from oracle import smc as osmc
    S = np.load(os.path.join(HERE, "golden", "reference_runs_smc_v1.npz"))
    n = 20000
    chain = osmc.TemperedSMC(models.IsoGaussianPrior(2, 0.0, 3.0), models.Rastrigin(2, 1.0), n=n, seed=0, move='rw',
                             stepsize=0.5, resampling='multinomial').run()
    beta = np.array([c['beta'] for c in chain])
    assert len(beta) == len(S["smc_temperature"])
    np.testing.assert_allclose(beta[:8], S["smc_temperature"][:8], rtol=0.04)
    # later entries carry the Monte-Carlo state of the n = 1000 reference run (the oracle at n = 1000 spreads over
    # 0.76-0.82 at entry 13 across seeds; the reference run has 0.823)
    np.testing.assert_allclose(beta[8:], S["smc_temperature"][8:], rtol=0.12)
    ess_o, ess_r = np.array([c['ess'] for c in chain]) / n, S["smc_ess"] / float(S["smc_n"])
    np.testing.assert_allclose(ess_o[:-1], ess_r[:-1], atol=2e-4)       # pinned by the search: 0.9^k, reset when below 0.5
    assert abs(ess_o[-1] - ess_r[-1]) < 0.08                            # the last step stops at temperature 1
    np.testing.assert_allclose([c['log_norm_constant'] for c in chain], S["smc_log_norm_constant"], atol=0.15)
    np.testing.assert_allclose([float(np.mean(c['alpha'])) for c in chain][1:-1], S["smc_alpha_mean"][1:-1], atol=0.06)


def test_svgd_oracle_reproduces_the_reference_run():
    """transport/svgd.py end to end (tests/golden/reference_runs_svgd_v1.npz): SVGD is deterministic once the ensemble is
    given, so the oracle must reproduce the reference's own 15 iterations (adagrad, mean bandwidth re-adapted every
    iteration, full-covariance Gaussian likelihood + Gaussian prior) to the accuracy of the stand-in's finite-difference
    gradients"""
    from oracle import svgd as osvgd
    V = np.load(os.path.join(HERE, "golden", "reference_runs_svgd_v1.npz"))
    prior, lik = models.IsoGaussianPrior(2, 0.0, 2.0), models.GaussianTarget(V["svgd_mean"], V["svgd_cov"])

    def pg(x):
        up, gp = prior.potential_and_grad(x)
        ul, gl = lik.potential_and_grad(x)
        return up + ul, gp + gl
    s = osvgd.SVGD(pg, V["svgd_X0"], 0.1, bandwidth='mean', max_iter=15)
    hs, xs = [s.h], [s.x.copy()]
    while s.iter < 15:
        xs.append(s.update().copy())
        hs.append(s.h)
    np.testing.assert_allclose(hs, V["svgd_bandwidth"], rtol=1e-6)
    np.testing.assert_allclose(np.array(xs), V["svgd_value"], atol=2e-6)
    np.testing.assert_allclose(s.U, V["svgd_potential"][-1], rtol=1e-5, atol=1e-6)


def test_smc_abc_oracle_matches_the_reference_run():
    """config C5 in small (tests/golden/reference_runs_abc_v1.npz): the reference's MetropolisedABCSMCSampler on the
    g-and-k model (n = 1000, 12 iterations) against oracle.abc.SMCABC (n = 20000): the ESS / alive-fraction pattern with its
    resampling point exactly, the adaptive thresholds (quantiles of a heavy-tailed distance distribution: log scale) and
    the mean acceptance probability of the random-walk ABC move"""
    from oracle import abc as oabc
    A = np.load(os.path.join(HERE, "golden", "reference_runs_abc_v1.npz"))
    n = 20000
    chain = oabc.SMCABC(models.GKTransformed(A["abc_data"]), n, 0, max_iter=12).run()
    assert len(chain) == len(A["abc_threshold"])
    np.testing.assert_allclose(np.array([c['ess'] for c in chain]) / n, A["abc_ess"] / float(A["abc_n"]), atol=2e-3)
    dlog = np.log([c['threshold'] for c in chain]) - np.log(A["abc_threshold"])
    assert np.all(np.abs(dlog[:7]) < 0.7) and np.all(np.abs(dlog[7:]) < 0.3), dlog
    np.testing.assert_allclose([float(np.mean(c['alpha'])) for c in chain][1:], A["abc_alpha_mean"][1:], atol=0.05)


def test_rm_metropolised_smc_oracle_matches_the_reference_run():
    """RMMetropolisedSMCSampler with the MALA move (tests/golden/reference_runs_rm_v1.npz, n = 500) against the oracle
    (n = 20000): the Robbins-Monro stepsize trajectory -- rising from 0.3 to about 1.25 while almost every proposal is
    accepted, then falling back as the target sharpens -- the temperature ladder and the evidence"""
    from oracle import smc as osmc
    S = np.load(os.path.join(HERE, "golden", "reference_runs_rm_v1.npz"))
    chain = osmc.TemperedSMC(models.IsoGaussianPrior(2, 0.0, 3.0), models.Rastrigin(2, 1.0), n=20000, seed=0, move='mala',
                             stepsize=0.3, resampling='multinomial', rm_stepsize=1.0, rm_target=0.651).run()
    assert len(chain) == len(S["rm_temperature"])
    step = np.array([0.3] + [c['stepsize'] for c in chain[1:]])
    np.testing.assert_allclose(step[:9], S["rm_stepsize"][:9], rtol=0.03)
    np.testing.assert_allclose(step[9:], S["rm_stepsize"][9:], rtol=0.12)
    beta = np.array([c['beta'] for c in chain])
    np.testing.assert_allclose(beta[:6], S["rm_temperature"][:6], rtol=0.04)
    np.testing.assert_allclose(beta[6:], S["rm_temperature"][6:], rtol=0.1)
    assert abs(chain[-1]['log_norm_constant'] - S["rm_log_norm_constant"][-1]) < 0.35
