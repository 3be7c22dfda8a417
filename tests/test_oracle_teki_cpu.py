"""CPU check of oracle/teki.py (transport/teki.py:20-185 has no test upstream): on a linear-Gaussian simulator the
tempered ensemble Kalman inversion reproduces the conjugate posterior, with a fixed schedule, with the default geometric
rule and with the adaptive (ESS) rule; the temperature bookkeeping follows the reference's index conventions."""
import numpy as np
from scipy.special import ndtri

from oracle import teki


class LinearGaussianSimulator:
    dim = 2

    def __init__(self):
        self.H = np.array([[1.0, 0.5], [-0.3, 1.2], [0.8, -0.7]])
        self.r = 0.6
        self.data = np.array([0.9, -0.4, 1.1])

    def simulate(self, x, u01):
        u = np.clip(np.asarray(u01, np.float64), 1e-7, 1 - 1e-7)
        return x @ self.H.T + self.r * ndtri(u)

    def posterior(self):
        prec = np.eye(2) + self.H.T @ self.H / self.r ** 2
        cov = np.linalg.inv(prec)
        return cov @ self.H.T @ self.data / self.r ** 2, cov


def _check(st, sc, tol_mean=0.06, tol_cov=0.06):
    mean, cov = sc.posterior()
    np.testing.assert_allclose(st['x'].mean(0), mean, atol=tol_mean)
    np.testing.assert_allclose(np.cov(st['x'].T), cov, atol=tol_cov)


def test_fixed_schedule_reaches_the_conjugate_posterior():
    sc = LinearGaussianSimulator()
    sched = np.linspace(0.0, 1.0, 11)
    st = teki.TemperedEKI(sc, 4000, 3, temperature_schedule=sched).run()
    np.testing.assert_allclose(st['temperature_schedule'], sched)          # schedule[iter], first entry never used
    assert st['iter'] == 10 and st['perturb_nan'] == 0
    _check(st, sc)


def test_default_geometric_rule_and_adaptive_rule():
    sc = LinearGaussianSimulator()
    st = teki.TemperedEKI(sc, 4000, 4).run()
    temps = st['temperature_schedule']
    np.testing.assert_allclose(temps[1:4], np.round(2.0 ** (np.arange(1, 4) / 50.0) - 1.0, 4))
    assert temps[-1] >= 1.0 and temps[-2] < 1.0 and st['iter'] == 50       # 2^(50/50) - 1 = 1
    ad = teki.TemperedEKI(sc, 4000, 5, adaptive=True, ess_threshold=0.8).run()
    t = ad['temperature_schedule']
    assert t[-1] == 1.0 and np.all(np.diff(t) > 0) and 3 <= ad['iter'] <= 60
    _check(ad, sc, 0.08, 0.08)


def test_termination_on_ensemble_collapse():
    sc = LinearGaussianSimulator()
    s = teki.TemperedEKI(sc, 500, 6, temperature_schedule=np.linspace(0, 1, 6), term_std=10.0)
    st = s.startup()
    assert s.terminated(st)                                                # std < 10 prior_stds at once
