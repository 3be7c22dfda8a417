"""GPU parity tests of the individual kernels (through the C-ABI) against the NumPy oracle."""
import ctypes as C

import numpy as np
import numpy.testing as npt
import pytest

from oracle import core, philox

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng(lib):
    import mocat_b200.engine as e
    return e


def _t(a, dtype=None):
    import torch
    return torch.as_tensor(np.ascontiguousarray(a), device="cuda") if dtype is None else \
        torch.as_tensor(np.ascontiguousarray(a), dtype=dtype, device="cuda")


# ------------------------------------------------------------------------------------------------ K2
@pytest.mark.parametrize("n", [1, 31, 1000, 10_000, 1_000_003])
def test_lse_ess_parity(eng, n):
    rng = np.random.default_rng(n)
    lw = (rng.standard_normal(n) * 3.0).astype(np.float32)
    out = eng.lse_ess(_t(lw)).cpu().numpy()
    ref = core.lse_ess(lw)
    # fp32 exp on the device (2 ulp) + fp64 accumulation: tolerance 2e-6 absolute on the logs
    npt.assert_allclose(out[3:], ref, atol=2e-6, rtol=0)
    assert out[0] == np.float64(lw.max())


def test_lse_ess_inf_and_tempered(eng):
    rng = np.random.default_rng(0)
    n = 50_000
    lw = rng.standard_normal(n).astype(np.float32)
    lw[rng.random(n) < 0.3] = -np.inf                       # ABC-style dead particles
    out = eng.lse_ess(_t(lw)).cpu().numpy()
    npt.assert_allclose(out[3:], core.lse_ess(lw), atol=2e-6)
    w01 = np.where(rng.random(n) < 0.4, 0.0, -np.inf).astype(np.float32)
    out = eng.lse_ess(_t(w01)).cpu().numpy()
    npt.assert_allclose(np.exp(out[5]), np.sum(w01 == 0), rtol=1e-9)      # ess = #alive (Appendix A.1)
    dead = np.full(1000, -np.inf, np.float32)
    out = eng.lse_ess(_t(dead)).cpu().numpy()
    assert out[3] == -np.inf and np.isnan(out[5])                          # all -inf: LSE -inf, log-ESS NaN
    lik = (rng.random(n) * 20).astype(np.float32)
    out = eng.lse_ess(_t(lw), _t(lik), 0.37).cpu().numpy()
    npt.assert_allclose(out[3:], core.lse_ess_tempered(lw, lik, np.float32(0.37)), atol=5e-6)


# ------------------------------------------------------------------------------------------------ K4
@pytest.mark.parametrize("n", [1, 5, 4095, 4096, 4097, 100_000, 3_000_017])
def test_cumsum_bit_exact(eng, n):
    rng = np.random.default_rng(n)
    w = rng.random(n).astype(np.float32) ** 4
    w[rng.random(n) < 0.1] = 0.0
    w = (w / w.astype(np.float64).sum()).astype(np.float32)
    cdf = eng.cumsum_f32(_t(w), 2.0 ** 52).cpu().numpy()
    ref = core.cdf_from_weights(w, normalised=True)
    npt.assert_array_equal(cdf, ref)                       # bit-exact, any n / tile order
    assert cdf[-1] == 1.0 and np.all(np.diff(cdf) >= 0)
    # run-to-run determinism of the decoupled look-back
    npt.assert_array_equal(eng.cumsum_f32(_t(w), 2.0 ** 52).cpu().numpy(), cdf)


def test_cumsum_from_log_weights(eng):
    import torch
    rng = np.random.default_rng(1)
    n = 200_000
    lw = (rng.standard_normal(n) * 2).astype(np.float32)
    lw[::7] = -np.inf
    ctl = eng.ControlBlock()
    rec = np.zeros(1, dtype=eng._lib.CONTROL_DTYPE)[0]
    o = eng.lse_ess(_t(lw)).cpu().numpy()
    rec["wmax"], rec["s1"], rec["s2"], rec["resample"] = o[0], o[1], o[2], 1
    ctl.write(rec)
    cdf = eng.cumsum_lw(_t(lw), ctl).cpu().numpy()
    ref = core.cdf_from_log_weights(lw)
    npt.assert_allclose(cdf, ref, atol=5e-7)               # device __expf vs numpy exp: tolerance, not bits
    assert cdf[-1] == 1.0 and np.all(np.diff(cdf) >= 0)
    dead = np.where(np.isinf(lw))[0]
    dead = dead[dead > 0]
    npt.assert_array_equal(cdf[dead], cdf[dead - 1])       # dead particles own no mass -> never selected


# ------------------------------------------------------------------------------------------------ K5
@pytest.mark.parametrize("n,n_out", [(1, 1), (7, 7), (5000, 5000), (100_003, 100_003), (10_000, 3_000), (3_000, 50_000)])
def test_ancestors_bit_exact(eng, n, n_out):
    rng = np.random.default_rng(n + n_out)
    w = rng.random(n).astype(np.float32) ** 8              # skewed: long runs and long gaps
    w[rng.random(n) < 0.3] = 0.0
    w[0] = max(w[0], 1e-3)
    w = (w / w.astype(np.float64).sum()).astype(np.float32)
    cdf = core.cdf_from_weights(w)
    cdf_d = _t(cdf)
    u0 = 0.37123456789
    a = eng.ancestors(cdf_d, 0, n_out, u=_t(np.array([u0]))).cpu().numpy()
    npt.assert_array_equal(a, core.ancestors_systematic(cdf, u0, n_out))
    u = rng.random(n_out)
    a = eng.ancestors(cdf_d, 1, n_out, u=_t(u)).cpu().numpy()
    npt.assert_array_equal(a, core.ancestors_multinomial(cdf, u))


def test_ancestors_philox_and_degenerate(eng):
    n = 20_000
    cdf = core.cdf_from_weights(np.full(n, 1.0 / n, np.float32))
    a = eng.ancestors(_t(cdf), 0, n, seed=5, step=3).cpu().numpy()
    u0 = philox.uniform53(5, np.zeros(1, np.uint64), 3, philox.P_RESAMPLE)[0]
    npt.assert_array_equal(a, core.ancestors_systematic(cdf, u0))
    a = eng.ancestors(_t(cdf), 1, n, seed=5, step=3, gid0=100).cpu().numpy()
    u = philox.uniform53(5, np.arange(100, 100 + n, dtype=np.uint64), 3, philox.P_RESAMPLE)
    npt.assert_array_equal(a, core.ancestors_multinomial(cdf, u))
    # all mass on one particle
    w = np.zeros(n, np.float32); w[1234] = 1.0
    cdf = core.cdf_from_weights(w)
    a = eng.ancestors(_t(cdf), 0, n, u=_t(np.array([0.5]))).cpu().numpy()
    assert np.all(a == 1234)


# ------------------------------------------------------------------------------------------------ K6
def test_gather_state(eng):
    rng = np.random.default_rng(2)
    n, d = 10_001, 7
    ld = (n + 31) // 32 * 32
    src = np.zeros((d, ld), np.float32)
    src[:, :n] = rng.standard_normal((d, n))
    anc = np.sort(rng.integers(0, n, n)).astype(np.int32)
    dst = eng.gather_state(_t(anc), _t(src)).cpu().numpy()
    npt.assert_array_equal(dst[:, :n], src[:, anc])


# ------------------------------------------------------------------------------------------------ K7
def test_quantile_and_colstats(eng):
    rng = np.random.default_rng(3)
    n = 100_001
    v = np.abs(rng.standard_normal(n)).astype(np.float32) * 10
    v[::11] = v[5]                                          # duplicates
    for q in (0.0, 0.1234, 0.45, 0.9, 1.0):
        out = eng.quantile(_t(v), q).cpu().numpy()
        npt.assert_allclose(out[0], core.quantile_linear(v, q), rtol=1e-12)
    v2 = rng.standard_normal(1000).astype(np.float32)       # negatives
    npt.assert_allclose(eng.quantile(_t(v2), 0.3).cpu().numpy()[0], core.quantile_linear(v2, 0.3), rtol=1e-12)
    d = 4
    ld = (n + 31) // 32 * 32
    x = np.zeros((d, ld), np.float32)
    x[:, :n] = rng.standard_normal((d, n)) * np.array([[1.], [2.], [0.1], [5.]]) + 100.0
    mean, var = eng.colstats(_t(x), n)
    rm, rv = core.colstats(x[:, :n].T)
    npt.assert_allclose(mean.cpu().numpy(), rm, rtol=1e-9)
    npt.assert_allclose(var.cpu().numpy(), rv, rtol=1e-6)


# ------------------------------------------------------------------------------------------------ sorted-uniform path
def _sorted_ancestors(eng, cdf_d, n, n_out, mode, seed, step):
    import torch
    from mocat_b200 import _lib
    L = _lib.get()
    B = int(L.dll.mb_strata_count(n_out))
    hist = torch.zeros(B, dtype=torch.int32, device="cuda")
    offs = torch.zeros(B + 1, dtype=torch.int32, device="cuda")
    anc = torch.empty(n_out, dtype=torch.int32, device="cuda")
    counts = None
    if mode == 1:
        L.call("mb_strata_hist", L.ctx(), n_out, 0, B, seed, step, None, _lib.ptr(hist), 1, _lib.stream())
        counts = hist.cpu().numpy().copy()
    L.call("mb_ancestors_sorted", L.ctx(), _lib.ptr(cdf_d), n, None, mode, _lib.ptr(hist), _lib.ptr(offs), B, seed, step,
           0, n_out, _lib.ptr(anc), n_out, None, _lib.stream())
    assert int(hist.sum().item()) == 0                      # the consumed counts are left zeroed for the next resampling
    return anc.cpu().numpy(), counts, offs.cpu().numpy(), B


@pytest.mark.parametrize("n", [1, 17, 5000, 100_003, 1_000_000])
def test_sorted_ancestors_bit_exact(eng, n):
    """production resampling path (systematic and stratified-exact multinomial) == oracle, bit for bit"""
    rng = np.random.default_rng(n)
    w = rng.random(n).astype(np.float32) ** 6
    w[rng.random(n) < 0.3] = 0.0
    w[0] = max(w[0], 1e-3)
    w = (w / w.astype(np.float64).sum()).astype(np.float32)
    cdf = core.cdf_from_weights(w)
    cdf_d = _t(cdf)
    seed, step = 77, 5
    a_sys, _, _, _ = _sorted_ancestors(eng, cdf_d, n, n, 0, seed, step)
    u0 = philox.uniform53(seed, np.zeros(1, np.uint64), step, philox.P_RESAMPLE)[0]
    npt.assert_array_equal(a_sys, core.ancestors_systematic(cdf, u0))
    a_mul, hist, offs, B = _sorted_ancestors(eng, cdf_d, n, n, 1, seed, step)
    ref, u = core.ancestors_multinomial_stratified(cdf, seed, step)
    assert B == core.strata_count(n) and hist.sum() == n and offs[-1] == n
    npt.assert_array_equal(a_mul, ref)
    assert np.all(w[a_mul] > 0)                             # zero-weight particles are never selected


def test_stratified_multinomial_is_multinomial(eng):
    """offspring counts of the stratified-exact scheme follow Multinomial(n, w): chi-square on 50 weight bins and
    the variance of the counts (systematic resampling would have ~zero variance here)"""
    n = 200_000
    rng = np.random.default_rng(0)
    w = rng.random(n).astype(np.float32)
    w = (w / w.astype(np.float64).sum()).astype(np.float32)
    cdf = core.cdf_from_weights(w)
    a, _, _, _ = _sorted_ancestors(eng, _t(cdf), n, n, 1, 3, 1)
    counts = np.bincount(a, minlength=n)
    bins = np.array_split(np.arange(n), 50)
    obs = np.array([counts[b].sum() for b in bins], dtype=np.float64)
    exp = np.array([w[b].astype(np.float64).sum() for b in bins]) * n
    chi2 = np.sum((obs - exp) ** 2 / exp)
    assert chi2 < 100.0                                     # 49 dof: mean 49, sd 9.9
    # per-particle counts ~ Poisson-like: Var ~ E for multinomial (for systematic Var << E)
    big = w > np.median(w)
    ratio = counts[big].var() / counts[big].mean()
    assert 0.8 < ratio < 1.4


@pytest.mark.parametrize("n,d,h", [(700, 3, 1.0), (300, 50, 7.0), (65, 2, 10.0), (2049, 5, 2.5)])
def test_ksd_parity(lib, n, d, h):
    """mb_ksd (metrics.ksd, metrics.py:88-130) against the fp64 restatement: unweighted / weighted, reference sign /
    score sign, ragged tile edges"""
    import torch
    import mocat_b200 as mocat
    from mocat_b200 import metrics, kernels
    from mocat_b200.core import cdict
    from oracle import metrics as om
    rng = np.random.default_rng(n)
    x = rng.standard_normal((n, d)).astype(np.float32) * 1.3
    g = (x + 0.3 * rng.standard_normal((n, d))).astype(np.float32)
    lw = (0.5 * rng.standard_normal(n)).astype(np.float32)
    kern = kernels.Gaussian(bandwidth=h)
    samp = cdict(value=x, grad_potential=g)
    for ref in (True, False):
        sign = 'reference' if ref else 'score'
        npt.assert_allclose(metrics.ksd(samp, kern, stein_sign=sign), om.ksd(x, g, h, reference_sign=ref), rtol=3e-4)
        npt.assert_allclose(metrics.ksd(torch.as_tensor(x, device="cuda"), kern, grad_potential=g, log_weight=lw,
                                        stein_sign=sign),
                            om.ksd(x, g, h, log_weight=lw, reference_sign=ref), rtol=3e-4)
    npt.assert_allclose(metrics.ksd(samp, kern, bandwidth=2 * h), om.ksd(x, g, 2 * h), rtol=3e-4)
    with pytest.raises(TypeError):
        metrics.ksd(x, kern)
