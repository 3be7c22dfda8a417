"""Tempered ensemble Kalman inversion -- NumPy (fp64) restatement.  TEST INFRASTRUCTURE.

Follows /root/reference/mocat/src/transport/teki.py:
  :20-35    calculate_covariances (spread matrices / sqrt(n - 1))
  :74-102   TemperedEKI.startup: simulated_data = likelihood_sample(value), prior_stds = std(value, ddof=1),
            temperature = 0, prec_y_given_x = I; default next_temperature = round(2^(iter/50) - 1, 4) (:91-92);
            a schedule gives schedule[iter] (:62-66, index clamped by jnp)
  :104-111  termination_criterion: temperature >= max, iter >= max_iter, all(ensemble_std(constrain(value)) <
            term_std * prior_stds), any NaN in value
  :113-150  update
  :153-185  AdaptiveTemperedEKI.next_temperature: regula falsi (utils.bisect) on
            log_ess(-(x - temperature) * pseudo_likelihood_potential) - log(n * ess_threshold)
The reference ships no test for this sampler.  Pinned on the reference's own source run under the NumPy stand-in for
jax: calculate_covariances and the adaptive temperature to fp64 round-off (tests/test_reference_golden_cpu.py), whole
runs on the g-and-k simulator through ladder and posterior moments (tests/test_reference_runs_cpu.py); and on the
defining property of the update -- on a linear-Gaussian simulator the tempered ensemble reproduces the conjugate
posterior (tests/test_oracle_teki_cpu.py).
Randomness (convention of csrc/teki.cu): perturbation normals = philox.normals(seed, gid, step = iter, P_MOVE, d_y);
simulator uniforms = philox.uniforms24(seed, gid, step = iter, P_SIM, m) (step 0 at startup); prior sample =
philox.normals(seed, gid, 0, P_INIT, d_x).
"""
import numpy as np
from . import core, philox


def calculate_covariances(vals, sim):
    """teki.py:20-35"""
    n = vals.shape[0]
    s_x = (vals - vals.mean(0)) / np.sqrt(n - 1)
    s_y = (sim - sim.mean(0)) / np.sqrt(n - 1)
    return s_x.T @ s_x, s_x.T @ s_y, s_y.T @ s_y


class TemperedEKI:
    """scenario: object with .data (d_y,), .dim, .simulate(x, u01) -> (n, d_y), optional .constrain(x);
    simulate_with(x, step) may be overridden for simulators that are not driven by uniforms."""

    def __init__(self, scenario, n, seed, temperature_schedule=None, max_temperature=1.0, max_iter=10000, term_std=0.0,
                 nugget=1e-5, adaptive=False, ess_threshold=0.9, bisection_tol=1e-5, max_bisection_iter=1000,
                 normal_dtype=np.float64):
        self.sc, self.n, self.seed = scenario, int(n), int(seed)
        self.schedule = None if temperature_schedule is None else np.asarray(temperature_schedule, np.float64)
        self.max_temperature, self.max_iter = float(max_temperature), int(max_iter)
        if self.schedule is not None:                                  # :62-66
            self.max_temperature, self.max_iter = float(self.schedule[-1]), len(self.schedule)
        self.term_std, self.nugget = float(term_std), float(nugget)
        self.adaptive, self.ess_threshold = bool(adaptive), float(ess_threshold)
        self.tol, self.max_bis = float(bisection_tol), int(max_bisection_iter)
        self.normal_dtype = normal_dtype
        self.gid = np.arange(self.n, dtype=np.uint64)
        self.d_y = len(scenario.data)

    # -- pieces
    def _simulate(self, x, step):
        u = philox.uniforms24(self.seed, self.gid, step, philox.P_SIM, self.d_y)
        return self.sc.simulate(x, u)

    def _ensemble_std(self, x):                                        # :81-86
        c = self.sc.constrain(x) if hasattr(self.sc, 'constrain') else x
        return np.std(c, axis=0, ddof=1)

    def _next_temperature(self, st, prec):
        it = st['iter']
        if self.adaptive:                                              # :168-185
            diff = st['sim'] - self.sc.data
            ppot = 0.5 * np.einsum('ni,ij,nj->n', diff, prec, diff)
            temp, its = core.next_temperature_adaptive(np.zeros(self.n), ppot, st['temperature'], self.max_temperature,
                                                       float(self.n), retain=self.ess_threshold, tol=self.tol,
                                                       max_iter=self.max_bis)
            st['search_iters'] = its
            return temp
        if self.schedule is not None:
            return float(self.schedule[min(it, len(self.schedule) - 1)])
        return float(np.round(2.0 ** (it / 50.0) - 1.0, 4))            # :91-92

    # -- sampler protocol
    def startup(self, x0=None):
        if x0 is None:
            x = philox.normals(self.seed, self.gid, 0, philox.P_INIT, self.sc.dim, dtype=self.normal_dtype).astype(np.float64)
        else:
            x = np.asarray(x0, np.float64).copy()
        st = dict(x=x, sim=self._simulate(x, 0), temperature=0.0, iter=0, perturb_nan=0)
        self.prior_stds = np.std(x, axis=0, ddof=1)                    # :99
        return st

    def terminated(self, st):                                          # :104-111
        return bool(st['temperature'] >= self.max_temperature or st['iter'] >= self.max_iter
                    or np.all(self._ensemble_std(st['x']) < self.term_std * self.prior_stds)
                    or np.any(np.isnan(st['x'])))

    def update(self, st):                                              # :113-150
        st = dict(st)
        st['iter'] += 1
        x, sim = st['x'], st['sim']
        d_x, d_y = x.shape[1], sim.shape[1]
        cov_x, cov_xy, cov_y = calculate_covariances(x, sim)
        cov_y_given_x = cov_y - cov_xy.T @ np.linalg.inv(cov_x + self.nugget * np.eye(d_x)) @ cov_xy
        with np.errstate(invalid='ignore'):
            try:
                chol = np.linalg.cholesky(cov_y_given_x + self.nugget * np.eye(d_y))
            except np.linalg.LinAlgError:                              # jnp.linalg.cholesky returns NaN instead
                chol = np.full((d_y, d_y), np.nan)
        prec = np.linalg.inv(cov_y_given_x + self.nugget * np.eye(d_y))
        prev = st['temperature']
        new = self._next_temperature(st, prec)
        with np.errstate(divide='ignore'):
            alph = 1.0 / (new - prev)
        st['temperature'] = new
        cov_alph = cov_y + (alph - 1.0) * cov_y_given_x
        gain = cov_xy @ np.linalg.inv(cov_alph + self.nugget * np.eye(d_y))
        z = philox.normals(self.seed, self.gid, st['iter'], philox.P_MOVE, d_y, dtype=self.normal_dtype).astype(np.float64)
        with np.errstate(invalid='ignore'):
            perturbs = np.sqrt(alph - 1.0) * z @ chol.T
        st['perturb_nan'] = int(np.isnan(perturbs).sum())
        perturbs = np.where(np.isnan(perturbs), 0.0, perturbs)
        st['x'] = x + (self.sc.data - sim + perturbs) @ gain.T
        st['sim'] = self._simulate(st['x'], st['iter'])
        st['gain'], st['cov_y_given_x'], st['alph'] = gain, cov_y_given_x, alph
        return st

    def run(self, x0=None):
        st = self.startup(x0)
        temps = [0.0]
        while not self.terminated(st):
            st = self.update(st)
            temps.append(st['temperature'])
        st['temperature_schedule'] = np.array(temps)
        return st
