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
