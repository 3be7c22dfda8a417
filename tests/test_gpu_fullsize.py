"""Parity at the sizes BASELINE.json quotes ("no parity test at any BASELINE config size").

n = 1e8 for the reduction / scan / ancestor kernels (64-bit index paths, multi-chunk strata scan, streaming tempering
search), a full C2 population step at n = 1e6 against oracle.smc, and the tcgen05 SVGD interaction at n = 32768, d = 50
against the fp64 oracle on a subset of rows.  The oracle side runs in chunks so that every test stays within seconds
of CPU time and a few GB of host memory."""
import numpy as np
import numpy.testing as npt
import pytest

from oracle import core, models as omodels, philox, smc as osmc, svgd as osvgd

pytestmark = pytest.mark.gpu
N8 = 100_000_000


@pytest.fixture(scope="module")
def E(lib):
    import torch
    from mocat_b200 import _lib, engine, models
    return torch, _lib, engine, models, lib


def _chunks(n, step=10_000_000):
    for lo in range(0, n, step):
        yield lo, min(n, lo + step)


def _lse3_chunked(w_of, n):
    """(max, sum e^{w-max}, sum e^{2(w-max)}) in fp64, two passes over chunks"""
    m = -np.inf
    for lo, hi in _chunks(n):
        m = max(m, float(np.max(w_of(lo, hi))))
    s1 = s2 = 0.0
    for lo, hi in _chunks(n):
        e = np.exp(w_of(lo, hi).astype(np.float64) - m)
        s1 += float(e.sum())
        s2 += float((e * e).sum())
    return m, s1, s2


def test_lse_ess_1e8(E):
    """K2 at the headline size, plain and tempered (lw - dbeta * lik), with dead (-inf) particles"""
    torch, l, e, m, lib = E
    g = torch.Generator(device="cuda").manual_seed(1)
    lw = torch.randn(N8, device="cuda", generator=g) * 3.0
    lw[::1_000_003] = float("-inf")
    lik = torch.rand(N8, device="cuda", generator=g) * 20.0
    lw_h, lik_h = lw.cpu().numpy(), lik.cpu().numpy()
    out = e.lse_ess(lw).cpu().numpy()
    mx, s1, s2 = _lse3_chunked(lambda lo, hi: lw_h[lo:hi], N8)
    ref = np.array([np.log(s1) + mx, np.log(s2) + 2 * mx, 2 * (np.log(s1) + mx) - (np.log(s2) + 2 * mx)])
    assert out[0] == mx
    npt.assert_allclose(out[3:], ref, atol=2e-6, rtol=0)               # fp32 ex2 + fp64 accumulation
    db = np.float32(0.37)
    out = e.lse_ess(lw, lik, float(db)).cpu().numpy()
    w_of = lambda lo, hi: (lw_h[lo:hi] - db * lik_h[lo:hi]).astype(np.float32)   # one fp32 FMA on the device: <= 1 ulp apart
    mx, s1, s2 = _lse3_chunked(w_of, N8)
    ref = np.array([np.log(s1) + mx, np.log(s2) + 2 * mx, 2 * (np.log(s1) + mx) - (np.log(s2) + 2 * mx)])
    npt.assert_allclose(out[3:], ref, atol=5e-6, rtol=0)


def test_cumsum_and_sorted_ancestors_1e8(E):
    """K4/K5 legacy chain (materialised fp64 CDF) at n = 1e8: the decoupled look-back scan is bit-exact against np.cumsum
    of the quantised weights, systematic ancestors are bit-exact, the stratified-exact multinomial is stratum-sorted, selects no
    zero-weight particle and reproduces the expected offspring mass over 1000 blocks of particles"""
    torch, l, e, m, lib = E
    rng = np.random.default_rng(7)
    w = rng.random(N8, dtype=np.float32) ** 4
    w[rng.integers(N8, size=N8 // 10)] = 0.0
    w = (w / w.astype(np.float64).sum()).astype(np.float32)
    wd = torch.as_tensor(w, device="cuda")
    cdf_d = e.cumsum_f32(wd, 2.0 ** 52)
    ref = core.cdf_from_weights(w, normalised=True)
    cdf = cdf_d.cpu().numpy()
    assert np.array_equal(cdf, ref)
    del ref
    # sorted-uniform ancestor search over the 1e8-entry CDF (multi-chunk strata offsets: B = 2^22 > 131072)
    L = lib
    B = int(L.dll.mb_strata_count(N8))
    assert B > 131072
    hist = torch.zeros(B, dtype=torch.int32, device="cuda")
    offs = torch.zeros(B + 1, dtype=torch.int32, device="cuda")
    anc = torch.empty(N8, dtype=torch.int32, device="cuda")
    seed, step = 5, 9
    L.call("mb_ancestors_sorted", L.ctx(), l.ptr(cdf_d), N8, None, 0, l.ptr(hist), l.ptr(offs), B, seed, step, 0, N8,
           l.ptr(anc), N8, None, l.stream())
    u0 = philox.uniform53(seed, np.zeros(1, np.uint64), step, philox.P_RESAMPLE)[0]
    a = anc.cpu().numpy()
    for lo, hi in _chunks(N8):
        u = (np.arange(lo, hi, dtype=np.float64) + u0) / float(N8)
        assert np.array_equal(a[lo:hi], np.minimum(np.searchsorted(cdf, u, side='right'), N8 - 1)), lo
    L.call("mb_strata_hist", L.ctx(), N8, 0, B, seed, step, None, l.ptr(hist), 1, l.stream())
    assert int(hist.sum(dtype=torch.int64).item()) == N8
    L.call("mb_ancestors_sorted", L.ctx(), l.ptr(cdf_d), N8, None, 1, l.ptr(hist), l.ptr(offs), B, seed, step, 0, N8,
           l.ptr(anc), N8, None, l.stream())
    a = anc.cpu().numpy()
    assert int(offs[-1].item()) == N8 and int(hist.sum(dtype=torch.int64).item()) == 0
    assert a.min() >= 0 and a.max() < N8
    # sorted at the granularity of the strata: output g of stratum s has u_g in [s/B, (s+1)/B), so its ancestor's CDF
    # interval must meet that range
    gs = np.arange(0, N8, 97)
    st = np.searchsorted(offs.cpu().numpy(), gs, side='right') - 1
    ag = a[gs].astype(np.int64)
    assert np.all(w[ag] > 0)                                            # zero-weight particles are never selected
    assert np.all(cdf[ag] > st / float(B))
    assert np.all(np.where(ag > 0, cdf[np.maximum(ag - 1, 0)], 0.0) <= (st + 1) / float(B))
    edges = np.linspace(0, N8, 1001).astype(np.int64)
    obs = np.bincount(a // (N8 // 1000), minlength=1000).astype(np.float64)   # offspring of each block of 1e5 particles
    mass = np.diff(np.concatenate([[0.0], cdf[edges[1:] - 1]]))
    chi2 = np.sum((obs - N8 * mass) ** 2 / (N8 * mass))
    assert chi2 < 1000 + 6 * np.sqrt(2000.0)                           # 999 dof


def test_stratified_multinomial_bit_exact_multichunk(E):
    """the multi-chunk strata scan (B > 131072 strata, i.e. n > 4M) against the oracle, bit for bit"""
    torch, l, e, m, lib = E
    n = 8_400_000
    rng = np.random.default_rng(3)
    w = rng.random(n, dtype=np.float32) ** 6
    w[rng.random(n) < 0.3] = 0.0
    w = (w / w.astype(np.float64).sum()).astype(np.float32)
    cdf = core.cdf_from_weights(w)
    L = lib
    B = int(L.dll.mb_strata_count(n))
    assert B == core.strata_count(n) and B > 131072
    hist = torch.zeros(B, dtype=torch.int32, device="cuda")
    offs = torch.zeros(B + 1, dtype=torch.int32, device="cuda")
    anc = torch.empty(n, dtype=torch.int32, device="cuda")
    cdf_d = torch.as_tensor(cdf, device="cuda")
    L.call("mb_strata_hist", L.ctx(), n, 0, B, 77, 5, None, l.ptr(hist), 1, l.stream())
    L.call("mb_ancestors_sorted", L.ctx(), l.ptr(cdf_d), n, None, 1, l.ptr(hist), l.ptr(offs), B, 77, 5, 0, n,
           l.ptr(anc), n, None, l.stream())
    ref, _ = core.ancestors_multinomial_stratified(cdf, 77, 5)
    assert np.array_equal(anc.cpu().numpy(), ref)


def test_temper_adapt_streaming_1e8(E):
    """K3 in its STREAMING form (n > 1.2 M: the data pass per evaluation, grid-sync merge) at n = 1e8: temperature,
    ESS and log-evidence of one adaptive step against the regula falsi of the oracle evaluated on chunks"""
    torch, l, e, m, lib = E
    g = torch.Generator(device="cuda").manual_seed(3)
    lik = torch.rand(N8, device="cuda", generator=g) * 40.0 + torch.randn(N8, device="cuda", generator=g).abs() * 5.0
    lw = torch.zeros(N8, device="cuda")
    lik_h = lik.cpu().numpy()
    ctl = e.ControlBlock()
    rec = np.zeros(1, dtype=l.CONTROL_DTYPE)[0]
    rec['wmax'], rec['s1'], rec['s2'] = 0.0, float(N8), float(N8)
    rec['lse'] = rec['lse2'] = rec['log_ess'] = np.log(float(N8))
    rec['ess'] = float(N8)
    ctl.write(rec)
    tp = m.make_temper(max_iter=100)
    lib.call("mb_temper_adapt", lib.ctx(), l.ptr(lw), l.ptr(lik), N8, __import__("ctypes").byref(tp), 1, N8, N8,
             l.ptr(ctl.t), l.ptr(ctl.hist), None, l.stream())
    c = ctl.read()

    def log_ess(b):
        db = np.float32(b)
        mx, s1, s2 = _lse3_chunked(lambda lo, hi: (-db * lik_h[lo:hi]).astype(np.float32), N8)
        return 2 * (np.log(s1) + mx) - (np.log(s2) + 2 * mx), np.log(s1) + mx
    log_target = np.log(0.9 * N8)
    bnd, ev, it = core.bisect(lambda b: log_ess(b)[0] - log_target, [0.0, 1.0], max_iter=1000, tol=1e-5)
    b_ref = bnd[int(np.argmin(np.abs(ev)))]
    assert 0 < c['beta'] < 1 and c['iter'] == 1
    npt.assert_allclose(c['beta'], b_ref, rtol=2e-4)                    # same bracket walk, fp32 exponentials
    le, lse = log_ess(c['beta'])
    npt.assert_allclose(c['log_ess'], le, atol=5e-6)                    # the reported ESS is the ESS at the device's beta
    npt.assert_allclose(c['lse'], lse, atol=5e-6)
    npt.assert_allclose(c['log_z'], lse - np.log(float(N8)), atol=5e-6)
    assert abs(c['ess'] - 0.9 * N8) < 2e-5 * N8                         # tol 1e-5 on log-ESS
    lw_h = lw[:1000].cpu().numpy()
    npt.assert_allclose(lw_h, -np.float32(c['beta']) * lik_h[:1000], rtol=1e-6, atol=1e-6)   # weights updated in place


def test_c2_population_step_1e6_vs_oracle(E):
    """config C2 at its own size: startup + one full update (MALA move, adaptive temperature) of n = 1e6 particles
    against oracle.smc with the same Philox streams"""
    torch, l, e, m, lib = E
    n, d, seed = 1_000_000, 5, 4
    tgt = m.make_target(l.LIK_RASTRIGIN, d, prior_std=3.0, a=1.0)
    eng = e.SMCEngine(tgt, m.make_move(l.MOVE_MALA, 0.1), m.make_temper(max_iter=50), n, seed,
                      resampling=l.RESAMPLE_SYSTEMATIC)
    orc = osmc.TemperedSMC(omodels.IsoGaussianPrior(d, 0.0, 3.0), omodels.Rastrigin(d, 1.0), n, seed,
                           move='mala', stepsize=0.1, resampling='systematic', max_iter=50)
    eng.startup()
    st = orc.startup()
    c = eng.ctl.read()
    x = eng.values().cpu().numpy().astype(np.float64)
    assert float(np.mean(np.abs(x - st['x']) > 3e-5)) < 1e-4
    npt.assert_allclose(c['beta'], st['beta'], rtol=2e-4)
    npt.assert_allclose(c['ess'], st['ess'], rtol=2e-4)
    npt.assert_allclose(c['log_z'], st['log_norm_constant'], atol=2e-4)
    eng.update()
    st2 = orc.update(st)
    c2 = eng.ctl.read()
    assert c2['iter'] == 1 and c2['resampled'] == int(st2['resampled'])
    x2 = eng.values().cpu().numpy().astype(np.float64)
    assert np.any(np.abs(x2 - st2['x']) > 1e-4, axis=1).mean() < 2e-3   # accept/reject flips at |u - alpha| ~ 1e-6
    npt.assert_allclose(c2['beta'], st2['beta'], rtol=2e-3)
    npt.assert_allclose(c2['ess'], st2['ess'], rtol=2e-3)
    npt.assert_allclose(c2['alpha_mean'], st2['alpha'].mean(), atol=5e-4)


def test_svgd_phi_tcgen05_c4_size(E):
    """config C4's interaction (n = 32768, d = 50) on tcgen05 against the fp64 oracle on 96 rows spread over the tiles;
    tolerance as in test_svgd_phi_tcgen05_parity (bf16 operands: 8e-3 of max|phi|)"""
    torch, l, e, m, lib = E
    n, d = 32768, 50
    rng = np.random.default_rng(11)
    X = (rng.standard_normal((n, d)) * 0.6 + 0.5).astype(np.float32)
    G = rng.standard_normal((n, d)).astype(np.float32)
    Xd, Gd = torch.as_tensor(X, device="cuda"), torch.as_tensor(G, device="cuda")
    h = e.pairdist_bandwidth(Xd, "median")
    hv = float(h.item())
    phi = e.svgd_phi(Xd, Gd, h, 1).cpu().numpy()
    assert np.all(np.isfinite(phi))
    rows = np.unique(np.concatenate([np.arange(0, n, 512), [1, 127, 128, n - 129, n - 1], rng.integers(n, size=27)]))
    X64, G64 = X.astype(np.float64), G.astype(np.float64)
    ref = np.empty((len(rows), d))
    for k, i in enumerate(rows):                                        # svgd.py:18-32 for one row: mean_j [ -K_ij G_j + grad_xj K ]
        diff = X64[i] - X64                                             # (n, d)
        K = np.exp(-0.5 * np.sum(diff * diff, axis=1) / hv ** 2)
        ref[k] = (-(K[:, None] * G64).sum(0) + (K[:, None] * diff).sum(0) / hv ** 2) / n
    full = osvgd.phi(X64[:300], G64[:300], hv)                          # the row formula above IS the oracle's phi
    chk = np.empty((5, d))
    for k in range(5):
        diff = X64[k] - X64[:300]
        K = np.exp(-0.5 * np.sum(diff * diff, axis=1) / hv ** 2)
        chk[k] = (-(K[:, None] * G64[:300]).sum(0) + (K[:, None] * diff).sum(0) / hv ** 2) / 300
    npt.assert_allclose(chk, full[:5], rtol=1e-10, atol=1e-14)
    scale = np.abs(ref).max()
    assert np.abs(phi[rows] - ref).max() <= 8e-3 * scale
    # the median heuristic itself at this size, against the exact fp64 median of a 2048-row block's distances to all
    sub = X64[:2048]
    d2 = np.maximum((sub * sub).sum(1)[:, None] + (X64 * X64).sum(1)[None, :] - 2.0 * sub @ X64.T, 0.0)
    med = np.median(np.sqrt(d2))
    npt.assert_allclose(hv, med / np.sqrt(2.0 * np.log(n)), rtol=5e-3)  # a 2048-row sample of the n^2 distances
