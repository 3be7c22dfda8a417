"""CPU checks of oracle/online_smoothing.py and oracle/backward.py: transition potentials against SciPy's Gaussian
log-density (normalising constant of utils.py:49-79 included) and the law of the Gumbel-max stitching draw."""
import numpy as np
from scipy.stats import multivariate_normal as mvn

from oracle import online_smoothing as oos, models as om


def _lg(d, rng):
    A = rng.standard_normal((d, d)); F = 0.5 * A
    B = rng.standard_normal((d, d)); Q = B @ B.T / d + 0.3 * np.eye(d)
    return om.LinearGaussianSSM(np.zeros(d), np.eye(d), F, Q, np.eye(d), np.eye(d)), F, Q


def test_transition_potential_is_negative_log_density():
    """diagonal process noise: the reference's potential (utils.py:26-30,49-79) is the negative log transition density;
    FULL covariance: the reference multiplies the ROW vector by inv(chol(Q)) (reset_covariance, utils.py:257-258), which
    is the density of the precision (L^T L)^-1, not Q^-1 -- mirrored exactly by the oracle and the device (and pinned on
    the reference's own output by tests/test_reference_golden_cpu.py)"""
    rng = np.random.default_rng(0)
    F = 0.5 * rng.standard_normal((3, 3))
    Qd = np.diag([0.4, 1.1, 2.0])
    o = om.LinearGaussianSSM(np.zeros(3), np.eye(3), F, Qd, np.eye(3), np.eye(3))
    x0, x1 = rng.standard_normal((6, 3)), rng.standard_normal((6, 3))
    ref = [-mvn.logpdf(x1[i], F @ x0[i], Qd) for i in range(6)]
    np.testing.assert_allclose(oos.transition_potential(o, x0, x1), ref, rtol=1e-12)
    o, F, Q = _lg(3, rng)                                              # full Q
    L = np.linalg.cholesky(Q)
    quirk = np.linalg.inv(L.T @ L)                                     # the precision the reference's formula implies
    ref = [0.5 * (x1[i] - F @ x0[i]) @ quirk @ (x1[i] - F @ x0[i]) + 1.5 * np.log(2 * np.pi) + np.log(np.diag(L)).sum()
           for i in range(6)]
    np.testing.assert_allclose(oos.transition_potential(o, x0, x1), ref, rtol=1e-12)
    l96 = om.Lorenz96SSM(dim=8, q_std=0.7)
    x0 = rng.standard_normal((4, 8)) + 3
    x1 = l96.transition_function(x0) + 0.7 * rng.standard_normal((4, 8))
    ref = [-mvn.logpdf(x1[i], l96.transition_function(x0[i:i + 1])[0], 0.49 * np.eye(8)) for i in range(4)]
    np.testing.assert_allclose(oos.transition_potential(l96, x0, x1), ref, rtol=1e-12)


def test_full_stitch_draws_from_the_stitching_law():
    """online_smoothing.py:21-31: P(j | x0_i) proportional to exp(lw1_j - transition_potential(x0_i -> x1_j)); one fixed
    end replicated 4000 times must reproduce those probabilities (independent Gumbel streams per row)"""
    rng = np.random.default_rng(1)
    o, F, Q = _lg(2, rng)
    n_c, reps = 7, 4000
    x0 = np.repeat(rng.standard_normal((1, 2)), reps, axis=0)
    x1 = rng.standard_normal((n_c, 2))
    lw1 = rng.standard_normal(n_c)
    idx = oos.full_stitch(o, x0, x1, lw1, 3, 2)
    logits = lw1 - oos.transition_potential(o, np.repeat(x0[:1], n_c, axis=0), x1)
    p = np.exp(logits - logits.max()); p /= p.sum()
    freq = np.bincount(idx, minlength=n_c) / reps
    assert np.max(np.abs(freq - p)) < 4.0 * np.sqrt(0.25 / reps)


def test_fixed_lag_stitching_shapes_and_gather():
    rng = np.random.default_rng(2)
    o, F, Q = _lg(2, rng)
    early, recent = rng.standard_normal((3, 50, 2)), rng.standard_normal((4, 50, 2))
    out, inds = oos.fixed_lag_stitching(o, early, recent, np.zeros(50), 5, 6)
    assert out.shape == (6, 50, 2)
    np.testing.assert_array_equal(out[:3], early)
    np.testing.assert_array_equal(out[3:], recent[1:, inds])
