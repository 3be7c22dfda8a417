"""Pin the oracle's samplers on the reference's statistical end-to-end tests."""
import numpy as np
import numpy.testing as npt

from oracle import core, models, smc, pf, abc, svgd

COV = np.array([[1., 0.9], [0.9, 2.]])
POST_COV = np.linalg.inv(np.linalg.inv(COV) + np.eye(2) / 49.0)


def _fixture():
    # tests/test_transport.py:24-29 (incl. its prior-potential quirk 0.5*(x/7**2)**2, pscale=1/49)
    return models.IsoGaussianPrior(2, 0.0, 7.0, pscale=1 / 49.0), models.GaussianTarget(np.zeros(2), COV)


def _check_chain(chain, n):
    temps = np.array([s['beta'] for s in chain])
    lnc = np.array([s['log_norm_constant'] for s in chain])
    lik_prec = np.linalg.inv(COV)
    dets = np.array([np.linalg.det(np.linalg.inv(lik_prec * t + np.eye(2) / 49.0)) for t in temps])
    npt.assert_array_almost_equal(lnc, 0.5 * (np.log(dets) - 4 * np.log(7)), 0)     # test_transport.py:49-55
    last = chain[-1]
    cdf = core.cdf_from_log_weights(last['lw'])
    rng = np.random.default_rng(1)
    vals = last['x'][core.ancestors_multinomial(cdf, rng.random(n))]                # :57-62
    npt.assert_array_almost_equal(vals.mean(0), np.zeros(2), decimal=0)
    npt.assert_array_almost_equal(np.cov(vals.T), POST_COV, decimal=1)
    return temps


def test_tempered_smc_preschedule_rw():
    prior, lik = _fixture()
    n = 10000
    s = smc.TemperedSMC(prior, lik, n, seed=0, move='rw', stepsize=1.0,
                        temperature_schedule=np.arange(0., 1.1, 0.1))
    temps = _check_chain(s.run(), n)
    npt.assert_allclose(temps, np.arange(0., 1.1, 0.1)[1:], atol=1e-12)             # :134


def test_tempered_smc_adaptive_mala():
    prior, lik = _fixture()
    n = 10000
    s = smc.TemperedSMC(prior, lik, n, seed=0, move='mala', stepsize=1.0, leapfrog_steps=10)
    temps = _check_chain(s.run(), n)
    assert temps[-1] == 1.0 and np.all(np.diff(temps) > 0)


def test_pf_vs_kalman_c1():
    # config C1: d=1, F=Q=H=R=P0=1, T=100, n=1e4; PF log Z vs Kalman log-likelihood
    lg = models.LinearGaussianSSM(np.zeros(1), np.eye(1), np.eye(1), np.eye(1), np.eye(1), np.eye(1))
    _, y = lg.simulate(100, np.random.default_rng(0))
    means, covs, ll = pf.kalman_filter(lg, y)
    m2, c2, _ = pf.kalman_filter(lg, y, reproduce_cov0_bug=True)
    npt.assert_allclose(means, m2)            # P0 = I: the kalman.py:20 quirk is invisible
    out = pf.BootstrapPF(lg, 10000, seed=0, ess_threshold=0.5).run(y)
    assert abs(out[-1]['log_z'] - ll) < 0.6
    for t, st in enumerate(out):
        mean, _ = pf.weighted_moments(st['x'], st['lw'])
        assert abs(mean[0] - means[t, 0]) < 0.15
        assert st['x'].min() <= means[t, 0] <= st['x'].max()       # tests/test_ssm.py:33-44 style coverage


def test_smc_abc_linear_gaussian_like():
    # tests/test_abc_gk.py:130-164 style: SMC-ABC on g-and-k recovers parameters loosely
    rng = np.random.default_rng(0)
    true_x = np.array([-0.524, -1.28, -0.84, -1.645])        # ~ Phi^-1((3,1,2,.5)/10)
    sc0 = models.GKTransformed(np.zeros(8))
    data = sc0.simulate(true_x[None], rng.random((1, 8)))[0]
    sc = models.GKTransformed(data)
    s = abc.SMCABC(sc, 2000, seed=0, max_iter=30)
    chain = s.run()
    thr = np.array([c['threshold'] for c in chain])
    assert np.all(np.diff(thr) <= 1e-12)                      # thresholds decrease
    assert chain[-1]['ess'] > 100


def test_svgd_gaussian_moments():
    # tests/test_transport.py:65-89: n=100, 1000 iters, mean/median bandwidth: mean decimal=0, cov decimal=1
    prior, lik = _fixture()
    pg = lambda x: tuple(a + b for a, b in zip(prior.potential_and_grad(x), lik.potential_and_grad(x)))
    x0 = prior.sample(np.random.default_rng(0).standard_normal((100, 2)))
    s = svgd.SVGD(pg, x0, stepsize=1.0, bandwidth='median', max_iter=1000)
    x = s.run()
    npt.assert_array_almost_equal(x.mean(0), np.zeros(2), decimal=0)
    npt.assert_array_almost_equal(np.cov(x.T), POST_COV, decimal=1)


def test_exact_cdf_is_order_and_sharding_independent():
    """the claim behind bit-exact ancestors on any tile order / any number of GPUs: the quantised
    weights are multiples of 2^-52 that sum below 2, so every fp64 partial sum is exact and ANY association of the
    additions gives the same bits -- sequential cumsum, reversed, pairwise tree, per-shard scans + offsets"""
    from fractions import Fraction
    rng = np.random.default_rng(5)
    for n in (1, 7, 1000, 65_537):
        w = rng.random(n).astype(np.float32) ** 5
        w[rng.random(n) < 0.2] = 0.0
        w[0] = max(w[0], 1e-3)
        w = (w / w.astype(np.float64).sum()).astype(np.float32)
        q = core.quantise_weights(w)
        assert np.all(q * 2.0 ** 52 == np.rint(q * 2.0 ** 52))              # multiples of 2^-52
        total = np.cumsum(q)[-1]
        assert total < 2.0
        if n <= 1000:                                                       # exact rational check of the fp64 sum
            assert Fraction(float(total)) == sum(Fraction(float(v)) for v in q)
        assert np.cumsum(q[::-1])[-1] == total                              # reversed order
        tree = q.copy()                                                     # pairwise tree reduction
        while tree.size > 1:
            if tree.size % 2:
                tree = np.append(tree, 0.0)
            tree = tree[0::2] + tree[1::2]
        assert tree[0] == total
        ref = np.cumsum(q)
        for shards in (2, 3, 8):                                            # per-shard scans + exclusive offsets
            parts = np.array_split(q, shards)
            offs = np.concatenate([[0.0], np.cumsum([p.sum() for p in parts])])
            glued = np.concatenate([o + np.cumsum(p) for o, p in zip(offs, parts) if p.size])
            assert np.array_equal(glued, ref)
        perm = rng.permutation(n)                                           # any permutation: same total bits
        assert np.cumsum(q[perm])[-1] == total


def test_optimal_proposal_oracle_agrees_with_bootstrap_and_reference_matrices():
    """oracle.pf.OptimalPF (scalar form of OptimalNonLinearGaussianParticleFilter, ssm/nonlinear_gaussian.py:134-276):
    (i) its scalars equal the matrices the reference's `startup` (:152-186) builds for Q = q^2 I, R = r^2 I, H = I;
    (ii) it estimates the same log-evidence / filtering mean as the bootstrap filter, with a larger ESS"""
    import numpy as np
    from oracle import models as om, pf as opf
    q, r, p0, d = 0.7, 1.3, 2.0, 4
    Q, R, P0, H = q * q * np.eye(d), r * r * np.eye(d), p0 * p0 * np.eye(d), np.eye(d)
    Wp = np.linalg.inv(H @ Q @ H.T + R)                                   # weight_precision :170-171
    Kp = Q @ H.T @ Wp                                                     # proposal_kalman_gain (utils kalman_gain)
    Pc = Q - Kp @ H @ Q                                                   # proposal_cov :174-176
    K0 = P0 @ H.T @ np.linalg.inv(H @ P0 @ H.T + R)
    P0c = np.linalg.inv(np.linalg.inv(P0) + H.T @ np.linalg.inv(R) @ H)   # inverse of init_cond_prec :163-164
    v = q * q + r * r
    np.testing.assert_allclose(Kp, q * q / v * np.eye(d), atol=1e-12)
    np.testing.assert_allclose(Pc, q * q * r * r / v * np.eye(d), atol=1e-12)
    np.testing.assert_allclose(Wp, np.eye(d) / v, atol=1e-12)
    np.testing.assert_allclose(K0, p0 * p0 / (p0 * p0 + r * r) * np.eye(d), atol=1e-12)
    np.testing.assert_allclose(P0c, np.eye(d) / (1 / p0 ** 2 + 1 / r ** 2), atol=1e-12)
    ssm = om.Lorenz96SSM(dim=8, init_mean=3.0)
    rng = np.random.default_rng(0)                                        # data consistent with the prior (x_0 ~ N(m0, P0)):
    x = ssm.initial_sample(rng.standard_normal(ssm.dim))                  # the bootstrap estimate is then reliable too
    ys = []
    for t in range(5):
        if t > 0:
            x = ssm.transition_sample(x, rng.standard_normal(ssm.dim))
        ys.append(x + ssm.r_std * rng.standard_normal(ssm.dim))
    ys = np.array(ys)
    res = {}
    for name, cls in (("boot", opf.BootstrapPF), ("opt", opf.OptimalPF)):
        lz, ess, mean = [], [], []
        for seed in range(3):
            out = cls(ssm, 60000, seed, ess_threshold=0.5, resampling='systematic', normal_dtype=np.float32).run(ys)
            lz.append(out[-1]['log_z']); ess.append(np.mean([s['ess'] for s in out[1:]]))
            mean.append(opf.weighted_moments(out[-1]['x'], out[-1]['lw'])[0])
        res[name] = (np.mean(lz), np.std(lz), np.mean(ess), np.mean(mean, axis=0))
    assert res["opt"][2] > 2.0 * res["boot"][2]
    # the reference's optimal filter starts from ZERO log-weights (:209-214), i.e. its running evidence lacks the factor
    # p(y_0) = N(y_0; m_0, (p0^2 + r^2) I) the bootstrap weights carry
    v0 = ssm.init_std ** 2 + ssm.r_std ** 2
    lp_y0 = -0.5 * np.sum((ys[0] - ssm.init_mean) ** 2) / v0 - 0.5 * ssm.dim * np.log(2 * np.pi * v0)
    assert abs(res["opt"][0] + lp_y0 - res["boot"][0]) < 0.5 + 3 * (res["opt"][1] + res["boot"][1])
    np.testing.assert_allclose(res["opt"][3], res["boot"][3], atol=0.25)
