"""Backward simulation (FFBSi; csrc/backward.cu, ssm/backward.py:20-40,241-315 upstream) against oracle/backward.py and
against the exact Rauch-Tung-Striebel smoother of a linear-Gaussian model."""
import ctypes as C

import numpy as np
import numpy.testing as npt
import pytest

from oracle import backward as obw, models as omodels

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def E(lib):
    import torch
    import mocat_b200 as mocat
    from mocat_b200 import _lib, models
    return torch, _lib, models, mocat, lib


def _sample(E, ssm_struct, dt, x0, lw0, x1, seed, step):
    torch, l, m, mocat, lib = E
    n_pf, d = x0.shape
    n_s = len(x1) if x1 is not None else 700
    x0d = torch.as_tensor(x0.astype(np.float32), device="cuda")
    lwd = torch.as_tensor(lw0.astype(np.float32), device="cuda")
    x1d = None if x1 is None else torch.as_tensor(x1.astype(np.float32), device="cuda")
    work = torch.empty((n_pf, d), dtype=torch.float32, device="cuda")
    idx = torch.empty(n_s, dtype=torch.int32, device="cuda")
    out = torch.empty((n_s, d), dtype=torch.float32, device="cuda")
    lib.call("mb_backward_sample", lib.ctx(), C.byref(ssm_struct), dt, l.ptr(x0d), l.ptr(lwd), n_pf, l.ptr(x1d), n_s,
             l.ptr(work), seed, step, l.ptr(idx), l.ptr(out), l.stream())
    return idx.cpu().numpy().astype(np.int64), out.cpu().numpy()


@pytest.mark.parametrize("kind,d", [("lg", 1), ("lg", 3), ("lg", 8), ("l96", 8), ("l96", 40)])
def test_backward_step_matches_oracle(E, kind, d):
    """one backward step: the same Gumbel noise, fp32 logits on the device vs fp64 in the oracle -- the arg-max can only
    differ for near-ties, and then between two particles of nearly equal posterior weight"""
    torch, l, m, mocat, lib = E
    rng = np.random.default_rng(d + len(kind))
    n_pf, n_s = 1003, 777                                  # ragged tiles (n_pf % 128, n_pf % 4 != 0)
    if kind == "lg":
        A = rng.standard_normal((d, d)); F = 0.8 * A / max(1.0, np.max(np.abs(np.linalg.eigvals(A))))
        B = rng.standard_normal((d, d)); Q = B @ B.T / d + 0.3 * np.eye(d)
        o = omodels.LinearGaussianSSM(np.zeros(d), np.eye(d), F, Q, np.eye(d), np.eye(d))
        s = m.make_lg_ssm(np.zeros(d), np.eye(d), F, Q, np.eye(d), np.eye(d))
        x0 = rng.standard_normal((n_pf, d)) * 2.0
        x1 = x0[rng.integers(n_pf, size=n_s)] @ F.T + rng.standard_normal((n_s, d)) @ np.linalg.cholesky(Q).T
        dt = 1.0
    else:
        o = omodels.Lorenz96SSM(dim=d, q_std=0.7)
        s = m.make_lorenz96(dim=d, q_std=0.7)
        x0 = rng.standard_normal((n_pf, d)) * 2.0 + 3.0
        x1 = o.transition_function(x0[rng.integers(n_pf, size=n_s)]) + 0.7 * rng.standard_normal((n_s, d))
        dt = 0.05
    x0 = x0.astype(np.float32).astype(np.float64)
    x1 = x1.astype(np.float32).astype(np.float64)
    lw0 = (rng.standard_normal(n_pf) * 2.0).astype(np.float32).astype(np.float64)
    lw0[rng.integers(n_pf, size=20)] = -np.inf              # dead particles are never chosen
    idx, xs = _sample(E, s, dt, x0, lw0, x1, 11, 6)
    ref_idx, ref_x = obw.full_resampling(o, x0, lw0, x1, 11, 6)
    assert np.all(np.isfinite(lw0[idx]))
    same = np.mean(idx == ref_idx)
    assert same > 0.995, same
    npt.assert_array_equal(xs, x0[idx].astype(np.float32))
    # final-time draw (no transition term): a categorical from the weights
    idx_f, _ = _sample(E, s, dt, x0, lw0, None, 11, 9)
    ref_f, _ = obw.final_draw(x0, lw0, 700, 11, 9)
    assert np.mean(idx_f == ref_f) > 0.995


def test_ffbsi_matches_rts_smoother(E):
    """forward filtering / backward simulation through the public API on a 1-d linear-Gaussian model: the smoothed means
    of the backward trajectories against the exact RTS smoother, and against the oracle's backward simulation run on the
    same filter output"""
    torch, l, m, mocat, lib = E
    F, Q, R, P0 = 0.9, 0.5, 0.8, 1.0
    sc = mocat.ssm.TimeHomogenousLinearGaussian(np.zeros(1), [[P0]], [[F]], [[Q]], [[1.0]], [[R]])
    T, n = 25, 4000
    sim = sc.simulate(np.arange(float(T)), 3)
    out = mocat.ssm.forward_filtering_backward_simulation(sc, mocat.ssm.BootstrapFilter(), sim.y, sim.t, n, 5, n_pf=n,
                                                          ess_threshold=0.5)
    assert out.value.shape == (T, n, 1) and not hasattr(out, 'log_weight')
    assert out.num_transition_evals[0] == 0 and np.all(out.num_transition_evals[1:] == n * n)
    # Kalman filter + RTS smoother
    mu, P, mus, Ps, mup, Pp = 0.0, P0, [], [], [], []
    for k in range(T):
        if k > 0:
            mu, P = F * mu, F * P * F + Q
        mup.append(mu); Pp.append(P)
        K = P / (P + R)
        mu, P = mu + K * (sim.y[k, 0] - mu), (1 - K) * P
        mus.append(mu); Ps.append(P)
    sm = mus[:]
    for k in range(T - 2, -1, -1):
        G = Ps[k] * F / Pp[k + 1]
        sm[k] = mus[k] + G * (sm[k + 1] - mup[k + 1])
    est = out.value[:, :, 0].mean(axis=1)
    assert np.max(np.abs(est - np.array(sm))) < 0.12        # Monte-Carlo error of 4000 trajectories (sd ~ 0.7 / sqrt(ESS))
    # the oracle's backward pass on the SAME filter history and seed reproduces the trajectories
    pf = mocat.ssm.run_particle_filter_for_marginals(sc, mocat.ssm.BootstrapFilter(), sim.y, sim.t, 5, n=n, ess_threshold=0.5,
                                                     keep_history=True)
    bs = mocat.ssm.backward_simulation(sc, pf, 6, n_samps=500)
    o = omodels.LinearGaussianSSM(np.zeros(1), [[P0]], [[F]], [[Q]], [[1.0]], [[R]])
    ref = obw.backward_simulation(o, pf.value.astype(np.float64), pf.log_weight.astype(np.float64), 500, 6)
    assert np.mean(bs.value == ref.astype(np.float32)) > 0.97   # a flipped arg-max at step t changes the path before t
