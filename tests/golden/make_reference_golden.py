"""Generate tests/golden/reference_v1.npz by running the REFERENCE'S OWN SOURCE.

`/root/reference/mocat` (SamDuffield/mocat v0.2.6, unmodified, never copied) is imported with a NumPy stand-in for jax
on sys.path (tests/golden/jaxshim: jax is not installed and not installable in the build container).  Only RNG-free
functions are evaluated -- or functions whose random inputs are produced here and stored next to the outputs -- because
the stand-in's random streams are not jax's threefry streams.  The numbers are those of the reference's algorithms in
fp64, not of XLA's fp32 kernels.  The fixture pins `oracle/` (tests/test_reference_golden_cpu.py) and the CUDA path
(tests/test_gpu_reference_golden.py) on outputs of the reference itself; /root/reference does not exist on the GPU box,
so only the stored numbers travel.

Run from the repo root, in the build container:  python tests/golden/make_reference_golden.py
"""
import os
import sys

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "jaxshim"))
sys.path.insert(0, "/root/reference")

import numpy as np  # noqa: E402

import mocat  # noqa: E402  (the reference)
from mocat.src import utils, metrics, kernels  # noqa: E402
from mocat.src.core import cdict  # noqa: E402
from mocat.src.transport import svgd as ref_svgd, teki as ref_teki, smc as ref_smc  # noqa: E402
from mocat.src.ssm.linear_gaussian.linear_gaussian import TimeHomogenousLinearGaussian  # noqa: E402
from mocat.src.ssm.linear_gaussian.kalman import run_kalman_filter_for_marginals  # noqa: E402
from mocat.src.ssm.scenarios.lorenz96 import Lorenz96, lorenz96_dynamics  # noqa: E402
from mocat.src.ssm.nonlinear_gaussian import OptimalNonLinearGaussianParticleFilter  # noqa: E402
from mocat.src.abc.scenarios.gk import GKTransformedUniformPrior  # noqa: E402
from mocat.src.abc import smc as ref_abc_smc  # noqa: E402
from mocat.src.scenarios import toy_examples  # noqa: E402
from jax import random  # noqa: E402  (the stand-in)


def build():
    g = {}
    rng = np.random.default_rng(20261017)

    # ---- utils.gaussian_potential (utils.py:49-106): scalar / diagonal / full precision, sqrt-precision, vectorised
    x = rng.standard_normal((7, 3))
    mean = rng.standard_normal(3)
    B = rng.standard_normal((3, 3))
    prec = B @ B.T + 0.5 * np.eye(3)
    sq = np.linalg.cholesky(prec)
    g["gp_x"], g["gp_mean"], g["gp_prec"], g["gp_sqrt_prec"] = x, mean, prec, sq
    g["gp_scalar"] = np.array([utils.gaussian_potential(xi, mean, prec=2.5) for xi in x])
    g["gp_diag"] = np.array([utils.gaussian_potential(xi, mean, prec=np.array([0.5, 1.5, 2.0])) for xi in x])
    g["gp_full"] = np.array([utils.gaussian_potential(xi, mean, prec=prec, det_prec=np.linalg.det(prec)) for xi in x])
    g["gp_sqrt"] = np.asarray(utils.gaussian_potential(x, mean, sqrt_prec=sq, det_prec=np.linalg.det(prec)))

    # ---- metrics.log_ess_log_weight / ess_log_weight (metrics.py:69-78)
    lw = rng.standard_normal(500) * 3.0
    lw[::17] = -np.inf
    g["ess_lw"] = lw
    g["ess_log"] = np.float64(metrics.log_ess_log_weight(lw))
    g["ess_lin"] = np.float64(metrics.ess_log_weight(lw))

    # ---- utils.bisect (utils.py:205-237) on the tempering objective of transport/smc.py:311-326
    lik = np.abs(rng.standard_normal(400)) * 4.0
    lw0 = rng.standard_normal(400) * 0.1
    target = np.log(0.9 * metrics.ess_log_weight(lw0))
    fun = lambda b: metrics.log_ess_log_weight(lw0 - (b - 0.2) * lik) - target  # noqa: E731
    bounds, evals, iters = utils.bisect(fun, np.array([0.2, 1.0]), max_iter=1000, tol=1e-5)
    g["bis_lik"], g["bis_lw"] = lik, lw0
    g["bis_bounds"], g["bis_evals"], g["bis_iters"] = np.asarray(bounds), np.asarray(evals), np.int64(iters)

    # MetropolisedSMCSampler.next_temperature_adaptive (smc.py:311-326) on a hand-made ensemble state
    smp = ref_smc.MetropolisedSMCSampler(mocat.RandomWalk(stepsize=0.1))
    st = cdict(temperature=np.full(400, 0.2), likelihood_potential=lik, log_weight=lw0,
               ess=np.full(400, metrics.ess_log_weight(lw0)))
    g["smc_next_temperature"] = np.float64(smp.next_temperature_adaptive(st, cdict()))

    # ---- kernels.Gaussian (kernels.py:82-116) and the bandwidth heuristics (kernels.py:220-229)
    k = kernels.Gaussian(bandwidth=1.3)
    a, b = rng.standard_normal(4), rng.standard_normal(4)
    g["k_a"], g["k_b"] = a, b
    g["k_val"] = np.float64(k(a, b))
    g["k_grad_x"], g["k_grad_y"], g["k_diag_grad_xy"] = k.grad_x(a, b), k.grad_y(a, b), k.diag_grad_xy(a, b)
    X = rng.standard_normal((64, 5)) * 0.8 + 0.3
    g["bw_X"] = X
    g["bw_median"] = np.float64(kernels.median_bandwidth_update(X))
    g["bw_mean"] = np.float64(kernels.mean_bandwidth_update(X))

    # ---- transport/svgd.py:18-32 kernelised_grad_matrix (full batch)
    G = rng.standard_normal((64, 5))
    n = len(X)
    g["svgd_G"] = G
    g["svgd_phi"] = np.asarray(ref_svgd.kernelised_grad_matrix(X, G, k, cdict(bandwidth=0.9),
                                                               np.tile(np.arange(n), (n, 1))))

    # ---- metrics.ksd (metrics.py:88-130), unweighted and weighted
    lwk = rng.standard_normal(n) * 0.5
    g["ksd_lw"] = lwk
    g["ksd_plain"] = np.float64(metrics.ksd(X, k, grad_potential=G, bandwidth=1.1))
    g["ksd_weighted"] = np.float64(metrics.ksd(X, k, grad_potential=G, log_weight=lwk, bandwidth=1.1))

    # ---- transport/teki.py:20-35 calculate_covariances and :168-185 AdaptiveTemperedEKI.next_temperature
    vals, sim = rng.standard_normal((300, 4)), rng.standard_normal((300, 6)) * 2.0 + 1.0
    sim[:, :2] += vals[:, :2]
    cx, cxy, cy = ref_teki.calculate_covariances(vals, sim)
    g["teki_vals"], g["teki_sim"] = vals, sim
    g["teki_cov_x"], g["teki_cov_xy"], g["teki_cov_y"] = cx, cxy, cy
    data = rng.standard_normal(6) + 1.0
    cyx = cy - cxy.T @ np.linalg.inv(cx + 1e-5 * np.eye(4)) @ cxy
    prec_yx = np.linalg.inv(cyx + 1e-5 * np.eye(6))
    ad = ref_teki.AdaptiveTemperedEKI(ess_threshold=0.8)
    g["teki_data"], g["teki_prec"] = data, prec_yx
    g["teki_next_temperature"] = np.float64(ad.next_temperature(cdict(temperature=0.1, simulated_data=sim),
                                                                cdict(data=data, prec_y_given_x=prec_yx)))

    # ---- linear-Gaussian state-space model: potentials (linear_gaussian.py:73-84,104-113) and the Kalman filter
    #      (kalman.py:16-57; its line 20 passes the Cholesky factor of P0 as covariance, identical for P0 = I)
    d = 3
    A = rng.standard_normal((d, d)); F = 0.8 * A / max(1.0, np.max(np.abs(np.linalg.eigvals(A))))
    Bq = rng.standard_normal((d, d)); Q = Bq @ Bq.T / d + 0.3 * np.eye(d)
    H = rng.standard_normal((2, d))
    R = np.array([[0.5, 0.1], [0.1, 0.4]])
    lg = TimeHomogenousLinearGaussian(initial_mean=np.zeros(d), initial_covariance=np.eye(d), transition_matrix=F,
                                      transition_covariance=Q, likelihood_matrix=H, likelihood_covariance=R)
    x0s, x1s = rng.standard_normal((9, d)), rng.standard_normal((9, d))
    ys = rng.standard_normal((9, 2))
    g["lg_F"], g["lg_Q"], g["lg_H"], g["lg_R"] = F, Q, H, R
    g["lg_x0"], g["lg_x1"], g["lg_y"] = x0s, x1s, ys
    g["lg_transition_potential"] = np.array([lg.transition_potential(x0s[i], 0.0, x1s[i], 1.0) for i in range(9)])
    g["lg_likelihood_potential"] = np.array([lg.likelihood_potential(x1s[i], ys[i], 1.0) for i in range(9)])
    yk = rng.standard_normal((15, 2))
    mus, covs = run_kalman_filter_for_marginals(lg, yk, np.arange(15.0))
    g["kalman_y"], g["kalman_mean"], g["kalman_cov"] = yk, np.asarray(mus), np.asarray(covs)

    # ---- Lorenz 96 (ssm/scenarios/lorenz96.py:14-44): vector field, adaptive Dormand-Prince flow over dt = 0.05, and
    #      the potentials of NonLinearGaussian (nonlinear_gaussian.py:98-121); the optimal proposal's matrices (:152-186)
    l96 = Lorenz96(dim=8)
    xl = rng.standard_normal((5, 8)) * 2.0 + 3.0
    g["l96_x"] = xl
    g["l96_rhs"] = np.stack([lorenz96_dynamics(v, 0.0, 8.0) for v in xl])
    g["l96_flow"] = np.stack([l96.transition_function(v, 0.0, 0.05) for v in xl])
    xn = g["l96_flow"] + rng.standard_normal((5, 8))
    yl = xn + rng.standard_normal((5, 8))
    g["l96_xnew"], g["l96_y"] = xn, yl
    g["l96_transition_potential"] = np.array([l96.transition_potential(xl[i], 0.0, xn[i], 0.05) for i in range(5)])
    g["l96_likelihood_potential"] = np.array([l96.likelihood_potential(xn[i], yl[i], 0.05) for i in range(5)])
    opt = OptimalNonLinearGaussianParticleFilter()
    opt.startup(l96)
    g["opt_initial_kalman_gain"] = np.asarray(opt.initial_kalman_gain)
    g["opt_proposal_kalman_gain"] = np.asarray(opt.proposal_kalman_gain)
    g["opt_proposal_covariance_sqrt"] = np.asarray(opt.proposal_covariance_sqrt)
    g["opt_weight_precision_sqrt"] = np.asarray(opt.weight_precision_sqrt)

    # ---- g-and-k simulator (abc/scenarios/gk.py:68-96): constrain and full_data_sample for given uniforms
    class GK(GKTransformedUniformPrior):
        n_unsummarised_data = 8

        def summarise_data(self, data):
            return np.sort(data)
    gk = GK()
    xs = rng.standard_normal((6, 4))
    keys = [random.PRNGKey(100 + i) for i in range(6)]
    g["gk_x"] = xs
    g["gk_u01"] = np.stack([random.uniform(kk, shape=(8,)) for kk in keys])          # the stand-in's raw draws of the key
    g["gk_constrain"] = np.stack([gk.constrain(v) for v in xs])
    g["gk_summary"] = np.stack([gk.likelihood_sample(xs[i], keys[i]) for i in range(6)])

    # ---- toy targets (scenarios/toy_examples.py): Rastrigin and Gaussian likelihood potentials
    ras = toy_examples.Rastrigin(dim=5, a=1.3)
    xr = rng.standard_normal((6, 5)) * 2.0
    g["ras_x"] = xr
    g["ras_likelihood_potential"] = np.array([ras.likelihood_potential(v) for v in xr])

    # ---- SMC-ABC adaptation (abc/smc.py:94-98 adapt_stepsize_scaled_diag_cov, :163-166 next_threshold_adaptive)
    dist = np.abs(rng.standard_normal(501))
    ab = ref_abc_smc.MetropolisedABCSMCSampler()
    ex = cdict(parameters=cdict(ess_threshold_retain=0.9))
    g["abc_dist"] = dist
    g["abc_next_threshold"] = np.float64(ab.next_threshold_adaptive(cdict(distance=dist, ess=np.full(501, 380.0)), ex))
    xv = rng.standard_normal((501, 4)) * np.array([0.5, 1.0, 2.0, 0.1])
    _, ex2 = ref_abc_smc.adapt_stepsize_scaled_diag_cov(cdict(value=xv), cdict(parameters=cdict()))
    g["abc_value"], g["abc_stepsize"] = xv, np.asarray(ex2.parameters.stepsize)
    # ---- core.cdict (core.py:20-84): integer / array / slice indexing and `+`, with a nested cdict, a static_cdict,
    #      a scalar `time` and a python scalar -- recorded field by field (the container itself does not travel)
    from mocat.src.core import static_cdict
    mk = lambda o: cdict(value=np.arange(15.0).reshape(5, 3) + o, potential=np.arange(5.0) * 2 + o, time=1.5 + o, label=3,  # noqa: E731
                         inner=cdict(alpha=np.arange(5.0) / 10 + o), frozen=static_cdict(beta=np.arange(5.0) + o))
    c1, c2 = mk(0.0), mk(100.0)
    picks = {"int": 2, "arr": np.array([0, 3, 3]), "slice": slice(1, 4)}
    for name, idx in picks.items():
        r = c1[idx]
        g[f"cdict_get_{name}_value"], g[f"cdict_get_{name}_potential"] = np.asarray(r.value), np.asarray(r.potential)
        g[f"cdict_get_{name}_inner_alpha"], g[f"cdict_get_{name}_frozen_beta"] = np.asarray(r.inner.alpha), np.asarray(r.frozen.beta)
        g[f"cdict_get_{name}_time"], g[f"cdict_get_{name}_label"] = np.float64(r.time), np.int64(r.label)
    a = c1 + c2
    g["cdict_add_value"], g["cdict_add_potential"] = np.asarray(a.value), np.asarray(a.potential)
    g["cdict_add_inner_alpha"], g["cdict_add_frozen_beta"] = np.asarray(a.inner.alpha), np.asarray(a.frozen.beta)
    g["cdict_add_time"], g["cdict_add_label"] = np.float64(a.time), np.int64(a.label)
    return g


if __name__ == "__main__":
    g = build()
    np.savez_compressed(os.path.join(HERE, "reference_v1.npz"), **g)
    print({k: (np.shape(v), str(np.asarray(v).dtype)) for k, v in g.items()})
