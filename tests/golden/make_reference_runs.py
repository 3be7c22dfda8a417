"""Generate tests/golden/reference_runs_v1.npz: END-TO-END runs of the reference's own samplers.

Same mechanism as make_reference_golden.py (the reference's unmodified source under the NumPy stand-in for jax,
tests/golden/jaxshim), but whole algorithms instead of single functions: tempered EKI on the g-and-k simulator and the
fixed-lag particle smoothers / FFBSi on a 1-d linear-Gaussian model -- the (f)-rows of SURVEY section 8 for which the
reference ships no test.  The random streams are NumPy's, not jax's threefry, so these are STATISTICAL references: the
tests compare summary statistics (temperature ladders, posterior moments, smoothing errors, path degeneracy) within
Monte-Carlo tolerances, never samples.  vmap is a Python loop here, hence the small ensembles (about 5 minutes in total).

Run from the repo root, in the build container:  python tests/golden/make_reference_runs.py        (-> reference_runs_v1.npz)
                                                  python tests/golden/make_reference_runs.py pf     (-> reference_runs_pf_v1.npz)
                                                  python tests/golden/make_reference_runs.py smc    (-> reference_runs_smc_v1.npz)
                                                  python tests/golden/make_reference_runs.py svgd   (-> reference_runs_svgd_v1.npz)
                                                  python tests/golden/make_reference_runs.py abc    (-> reference_runs_abc_v1.npz)
                                                  python tests/golden/make_reference_runs.py rm     (-> reference_runs_rm_v1.npz)
"""
import os
import sys

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "jaxshim"))
sys.path.insert(0, "/root/reference")

import numpy as np  # noqa: E402
from scipy.special import ndtri  # noqa: E402

import mocat  # noqa: E402  (the reference)
from mocat.src.abc.scenarios.gk import GKTransformedUniformPrior  # noqa: E402
from mocat.src.ssm.linear_gaussian.linear_gaussian import TimeHomogenousLinearGaussian  # noqa: E402
from jax import random  # noqa: E402  (the stand-in)


def gk_problem():
    truth = np.array([1.5, 1.0, 1.0, 0.5])
    z = ndtri(np.random.default_rng(0).random(8))
    e = np.exp(-truth[2] * z)
    data = np.sort(truth[0] + truth[1] * (1 + 0.8 * (1 - e) / (1 + e)) * z * (1 + z * z) ** truth[3])

    class GK(GKTransformedUniformPrior):
        n_unsummarised_data = 8
        prior_maxs = 2.0

        def summarise_data(self, d):
            return np.sort(d)
    sc = GK()
    sc.data = data
    return sc, data


def lg_problem(T=22):
    F, Q, R, P0 = 0.9, 0.5, 0.8, 1.0
    sc = TimeHomogenousLinearGaussian(initial_mean=np.zeros(1), initial_covariance=np.array([[P0]]),
                                      transition_matrix=np.array([[F]]), transition_covariance=np.array([[Q]]),
                                      likelihood_matrix=np.array([[1.0]]), likelihood_covariance=np.array([[R]]))
    rng = np.random.default_rng(4)
    x = np.zeros(T)
    x[0] = rng.normal() * np.sqrt(P0)
    for k in range(1, T):
        x[k] = F * x[k - 1] + rng.normal() * np.sqrt(Q)
    y = (x + rng.normal(size=T) * np.sqrt(R))[:, None]
    mu, P, mus, Ps, mup, Pp = 0.0, P0, [], [], [], []
    for k in range(T):
        if k > 0:
            mu, P = F * mu, F * P * F + Q
        mup.append(mu); Pp.append(P)
        K = P / (P + R)
        mu, P = mu + K * (y[k, 0] - mu), (1 - K) * P
        mus.append(mu); Ps.append(P)
    sm, Pk = mus[:], Ps[:]
    for k in range(T - 2, -1, -1):
        G = Ps[k] * F / Pp[k + 1]
        sm[k] = mus[k] + G * (sm[k + 1] - mup[k + 1])
        Pk[k] = Ps[k] + G * (Pk[k + 1] - Pp[k + 1]) * G
    return sc, y, np.arange(float(T)), np.array(sm), np.array(Pk)


def build():
    g = {}
    # ---- tempered EKI (transport/teki.py) on the g-and-k simulator, prior U(0, 2)^4, 8 sorted draws
    sc, data = gk_problem()
    g["teki_data"] = data
    for name, smp, n in (("adaptive", mocat.AdaptiveTemperedEKI(ess_threshold=0.9), 1000),
                         ("schedule", mocat.TemperedEKI(temperature_schedule=np.linspace(0.0, 1.0, 11)), 1000)):
        out = mocat.run(sc, smp, n, random.PRNGKey(1))
        post = np.asarray(sc.constrain(out.value[-1]))
        g[f"teki_{name}_temperature"] = np.asarray(out.temperature, np.float64)
        g[f"teki_{name}_mean"], g[f"teki_{name}_std"] = post.mean(0), post.std(0)
        g[f"teki_{name}_n"] = np.int64(n)
    # ---- fixed-lag smoothers (ssm/online_smoothing.py) and FFBSi (ssm/backward.py) on the 1-d linear-Gaussian model
    lg, y, t, sm, Pk = lg_problem()
    g["lg_y"], g["lg_rts_mean"], g["lg_rts_var"] = y, sm, Pk
    n, lag = 300, 6
    pf = mocat.ssm.BootstrapFilter()
    for name, bs in (("pf", False), ("bs", True)):
        p = mocat.ssm.initiate_particles(lg, pf, n, random.PRNGKey(9), y[0], t[0])
        for k in range(1, len(t)):
            p = mocat.ssm.propagate_particle_smoother(lg, pf, p, y[k], t[k], random.PRNGKey(900 + k), lag, backward_sim=bs)
        v = np.asarray(p.value)[:, :, 0]
        g[f"smoother_{name}_mean"] = v.mean(1)
        g[f"smoother_{name}_var"] = v.var(1)
        g[f"smoother_{name}_unique_fraction"] = np.array([len(np.unique(r)) / n for r in v])
    g["smoother_n"], g["smoother_lag"] = np.int64(n), np.int64(lag)
    # forward_filtering_backward_simulation (backward.py:303-350) = the two calls below plus timing
    filt = mocat.ssm.run_particle_filter_for_marginals(lg, pf, y, t, random.PRNGKey(5), n=n)
    out = mocat.ssm.backward_simulation(lg, filt, random.PRNGKey(6), n)
    v = np.asarray(out.value)[:, :, 0]
    g["ffbsi_mean"], g["ffbsi_var"] = v.mean(1), v.var(1)
    g["ffbsi_unique_fraction"] = np.array([len(np.unique(r)) / n for r in v])
    return g


def build_pf():
    """the bootstrap particle filter of config C3 in small (ssm/filtering.py:255-324 on ssm/scenarios/lorenz96.py): d = 8,
    R = 4 I (an ensemble of 1000 keeps an ESS of 10-250), trajectory started from the prior; the transition is the
    reference's adaptive Dormand-Prince flow, resampling its multinomial `random.categorical` at ess < 0.5 n"""
    from mocat.src.ssm.scenarios.lorenz96 import Lorenz96
    d, T, n = 8, 10, 1000
    sc = Lorenz96(dim=d, likelihood_covariance=4.0 * np.eye(d))
    rng = np.random.default_rng(3)
    xs = np.empty((T, d))
    xs[0] = rng.standard_normal(d)
    for k in range(1, T):
        xs[k] = np.asarray(sc.transition_function(xs[k - 1], 0.0, 0.05)) + rng.standard_normal(d)
    ys = xs + 2.0 * rng.standard_normal((T, d))
    t = np.arange(T) * 0.05
    out = mocat.ssm.run_particle_filter_for_marginals(sc, mocat.ssm.BootstrapFilter(), ys, t, random.PRNGKey(2), n=n,
                                                      ess_threshold=0.5)
    lw = np.asarray(out.log_weight)
    w = np.exp(lw - lw.max(1, keepdims=True))
    w /= w.sum(1, keepdims=True)
    v = np.asarray(out.value)
    mean = np.einsum('tn,tnd->td', w, v)
    var = np.einsum('tn,tnd->td', w, (v - mean[:, None, :]) ** 2)
    g = {"pf_x": xs, "pf_y": ys, "pf_t": t, "pf_n": np.int64(n), "pf_ess": np.asarray(out.ess, np.float64),
         "pf_mean": mean, "pf_var": var}
    # the same observations through the optimal-proposal filter and the ensemble Kalman filter
    # (ssm/nonlinear_gaussian.py:134-350) -- the reference ships no test for either
    from mocat.src.ssm.nonlinear_gaussian import OptimalNonLinearGaussianParticleFilter, EnsembleKalmanFilter
    for name, filt, m in (("opt", OptimalNonLinearGaussianParticleFilter(), 1000), ("enkf", EnsembleKalmanFilter(), 500)):
        o = mocat.ssm.run_particle_filter_for_marginals(sc, filt, ys, t, random.PRNGKey(4), n=m, ess_threshold=0.5)
        lw = np.asarray(o.log_weight)
        w = np.exp(lw - lw.max(1, keepdims=True))
        w /= w.sum(1, keepdims=True)
        v = np.asarray(o.value)
        mu = np.einsum('tn,tnd->td', w, v)
        g[f"{name}_n"], g[f"{name}_ess"] = np.int64(m), np.asarray(o.ess, np.float64)
        g[f"{name}_mean"], g[f"{name}_var"] = mu, np.einsum('tn,tnd->td', w, (v - mu[:, None, :]) ** 2)
    return g


def build_smc():
    """config C2 in small: MetropolisedSMCSampler with a random-walk move (transport/smc.py:228-373) on the Rastrigin
    target (scenarios/toy_examples.py:135-149, a = 1, d = 2) under an N(0, 3^2 I) prior, adaptive tempering
    (ESS retain 0.9, resample below 0.5 n), n = 1000"""
    from mocat.src.scenarios import toy_examples
    import jax.numpy as jnp

    class Ras(toy_examples.Rastrigin):
        def prior_potential(self, x, random_key=None):
            return 0.5 * jnp.sum(jnp.square(x)) / 9.0

        def prior_sample(self, random_key):
            return 3.0 * random.normal(random_key, (self.dim,))
    sc = Ras(dim=2, a=1.0)
    out = mocat.run(sc, mocat.MetropolisedSMCSampler(mocat.RandomWalk(stepsize=0.5)), 1000, random.PRNGKey(0))
    v = np.asarray(out.value[-1])
    lw = np.asarray(out.log_weight[-1])
    w = np.exp(lw - lw.max()); w /= w.sum()
    return {"smc_n": np.int64(1000), "smc_temperature": np.asarray(out.temperature, np.float64),
            "smc_log_norm_constant": np.asarray(out.log_norm_constant, np.float64), "smc_ess": np.asarray(out.ess, np.float64),
            "smc_final_mean": w @ v, "smc_final_second_moment": w @ (v * v),
            "smc_alpha_mean": np.asarray(out.alpha, np.float64).mean(axis=-1)}


def build_svgd():
    """transport/svgd.py:35-146 end to end: SVGD is deterministic once the ensemble is given, so this run is an EXACT
    reference (up to the finite-difference gradients of the stand-in, 1e-8): 15 iterations, adagrad(0.1, momentum 0.9),
    mean-distance bandwidth re-adapted every iteration (the default), full-covariance Gaussian likelihood
    (scenarios/toy_examples.py:17-51) under an N(0, 2^2 I) prior, 100 particles"""
    from mocat.src.scenarios import toy_examples
    import jax.numpy as jnp

    class G(toy_examples.Gaussian):
        def prior_potential(self, x, random_key=None):
            return 0.5 * jnp.sum(jnp.square(x)) / 4.0
    mean, cov = np.array([1.0, -0.5]), np.array([[1.0, 0.6], [0.6, 2.0]])
    sc = G(mean=mean, covariance=cov)
    X0 = np.random.default_rng(0).standard_normal((100, 2)) * 2.0
    out = mocat.run(sc, mocat.SVGD(stepsize=0.1, max_iter=15), 100, random.PRNGKey(0), initial_state=mocat.cdict(value=X0))
    return {"svgd_X0": X0, "svgd_mean": mean, "svgd_cov": cov, "svgd_value": np.asarray(out.value, np.float64),
            "svgd_potential": np.asarray(out.potential, np.float64),
            "svgd_bandwidth": np.asarray(out.kernel_params.bandwidth, np.float64)}


def build_abc():
    """config C5 in small: MetropolisedABCSMCSampler with the default random-walk ABC move (abc/smc.py:100-245) on the
    g-and-k model (prior U(0, 10)^4 through the probit transform, 8 sorted draws, data simulated at (3, 1, 2, 0.5)),
    12 iterations, n = 1000: adaptive thresholds (quantile of the distances at 0.9 ESS / n), ESS, acceptance"""
    truth = np.array([3.0, 1.0, 2.0, 0.5])

    class GK(GKTransformedUniformPrior):
        n_unsummarised_data = 8

        def summarise_data(self, d):
            return np.sort(d)
    sc = GK()
    z = ndtri(1e-5 + np.random.default_rng(1).random(8) * (1 - 2e-5))
    e = np.exp(-truth[2] * z)
    sc.data = np.sort(truth[0] + truth[1] * (1 + 0.8 * (1 - e) / (1 + e)) * z * (1 + z * z) ** truth[3])
    out = mocat.run(sc, mocat.abc.MetropolisedABCSMCSampler(max_iter=12), 1000, random.PRNGKey(0))
    return {"abc_data": np.asarray(sc.data), "abc_n": np.int64(1000), "abc_threshold": np.asarray(out.threshold, np.float64),
            "abc_ess": np.asarray(out.ess, np.float64), "abc_alpha_mean": np.asarray(out.alpha, np.float64).mean(axis=-1),
            "abc_alive_fraction": (np.asarray(out.log_weight) > -np.inf).mean(axis=-1)}


def build_rm():
    """RMMetropolisedSMCSampler (transport/smc.py:376-428) with the MALA move Underdamped(stepsize = 0.3) (friction = inf,
    one leapfrog step: mcmc/standard_mcmc.py:72-153) on the Rastrigin / Gaussian-prior target of build_smc, n = 500:
    Robbins-Monro adaptation of the stepsize towards the 0.651 acceptance target (gradients of the potentials by the
    stand-in's central differences)"""
    from mocat.src.scenarios import toy_examples
    import jax.numpy as jnp

    class Ras(toy_examples.Rastrigin):
        def prior_potential(self, x, random_key=None):
            return 0.5 * jnp.sum(jnp.square(x)) / 9.0

        def prior_sample(self, random_key):
            return 3.0 * random.normal(random_key, (self.dim,))
    out = mocat.run(Ras(dim=2, a=1.0), mocat.RMMetropolisedSMCSampler(mocat.Underdamped(stepsize=0.3), rm_stepsize=1.0), 500,
                    random.PRNGKey(0))
    return {"rm_n": np.int64(500), "rm_temperature": np.asarray(out.temperature, np.float64),
            "rm_stepsize": np.asarray(out.stepsize, np.float64), "rm_ess": np.asarray(out.ess, np.float64),
            "rm_log_norm_constant": np.asarray(out.log_norm_constant, np.float64),
            "rm_alpha_mean": np.asarray(out.alpha, np.float64).mean(axis=-1)}


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "rm":
        g = build_rm()
        np.savez_compressed(os.path.join(HERE, "reference_runs_rm_v1.npz"), **g)
        for k, v in g.items():
            print(k, np.round(v, 4))
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "abc":
        g = build_abc()
        np.savez_compressed(os.path.join(HERE, "reference_runs_abc_v1.npz"), **g)
        for k, v in g.items():
            print(k, np.round(v, 3))
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "svgd":
        g = build_svgd()
        np.savez_compressed(os.path.join(HERE, "reference_runs_svgd_v1.npz"), **g)
        print(g["svgd_value"].shape, np.round(g["svgd_bandwidth"], 4), g["svgd_value"][-1].mean(0))
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "smc":
        g = build_smc()
        np.savez_compressed(os.path.join(HERE, "reference_runs_smc_v1.npz"), **g)
        for k, v in g.items():
            print(k, np.round(v, 4))
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "pf":
        g = build_pf()
        np.savez_compressed(os.path.join(HERE, "reference_runs_pf_v1.npz"), **g)
        print(np.round(g["pf_ess"], 1), np.sqrt(np.mean((g["pf_mean"] - g["pf_x"]) ** 2, axis=1)))
        sys.exit(0)
    g = build()
    np.savez_compressed(os.path.join(HERE, "reference_runs_v1.npz"), **g)
    for k, v in g.items():
        print(k, np.round(np.asarray(v, np.float64), 3) if np.size(v) <= 22 else np.shape(v))
