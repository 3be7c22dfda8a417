"""CUDA path against the committed golden vectors (tests/golden/oracle_v3.npz) -- no import of `oracle/` here.
Tolerances as in DESIGN.md section 2: bit-exact for integer / CDF work, fp32 transcendental tolerances otherwise."""
import os

import numpy as np
import numpy.testing as npt
import pytest

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_v3.npz"))


@pytest.fixture(scope="module")
def E(lib):
    import mocat_b200.engine as e
    import mocat_b200.models as m
    from mocat_b200 import _lib
    return e, m, _lib


def _t(a):
    import torch
    return torch.as_tensor(np.ascontiguousarray(a), device="cuda")


def test_golden_lse_ess(E):
    e, m, l = E
    out = e.lse_ess(_t(G["lse_lw"])).cpu().numpy()
    npt.assert_allclose(out[3:], G["lse_out"], atol=2e-6, rtol=0)


def test_golden_cdf_and_ancestors_bit_exact(E):
    e, m, l = E
    cdf = e.cumsum_f32(_t(G["cdf_w"]), 2.0 ** 52)
    npt.assert_array_equal(cdf.cpu().numpy(), G["cdf"])
    anc = e.ancestors(cdf, l.RESAMPLE_MULTINOMIAL, u=_t(G["anc_u"]))
    npt.assert_array_equal(anc.cpu().numpy(), G["anc_multinomial"])
    anc = e.ancestors(cdf, l.RESAMPLE_SYSTEMATIC, u=_t(np.array([0.37])))
    npt.assert_array_equal(anc.cpu().numpy(), G["anc_systematic"])


def test_golden_quantile(E):
    e, m, l = E
    v = _t(G["quant_v"])
    got = [e.quantile(v, float(q)).cpu().numpy()[0] for q in G["quant_q"]]
    npt.assert_allclose(got, G["quant_out"], rtol=0, atol=1e-7)        # exact order statistics, fp32 data


def test_golden_svgd(E):
    import torch
    e, m, l = E
    X, Gr = _t(G["svgd_X"]), _t(G["svgd_G"])
    h = torch.tensor([float(G["svgd_h"])], dtype=torch.float32, device="cuda")
    phi = e.svgd_phi(X, Gr, h, 0).cpu().numpy()
    npt.assert_allclose(phi, G["svgd_phi"], atol=3e-5 * np.abs(G["svgd_phi"]).max(), rtol=1e-4)
    npt.assert_allclose(e.pairdist_bandwidth(X, "median", 0).item(), G["svgd_median_h"], rtol=2e-5)
    npt.assert_allclose(e.pairdist_bandwidth(X, "mean", 0).item(), G["svgd_mean_h"], rtol=2e-5)


def test_golden_logistic_regression(E):
    e, m, l = E
    U, Gd = e.logistic_potential_grad(_t(G["lr_A"]), _t(G["lr_t"]), 0.0, 0.5, 0.7, _t(G["lr_W"]))
    npt.assert_allclose(U.cpu().numpy(), G["lr_U"], rtol=2e-5, atol=1e-4)
    npt.assert_allclose(Gd.cpu().numpy(), G["lr_G"], rtol=2e-4, atol=1e-4)


def test_golden_tempered_smc_schedule(E):
    """same Philox streams as the stored oracle run (n = 512, seed 7): the adaptive schedule agrees closely for the
    first searches (before accept/reject flips decorrelate the two particle systems), the evidence statistically"""
    e, m, l = E
    tgt = m.make_target(l.LIK_RASTRIGIN, 2, prior_std=3.0, a=1.0)
    eng = e.SMCEngine(tgt, m.make_move(l.MOVE_MALA, 0.1), m.make_temper(), 512, 7, resampling=l.RESAMPLE_SYSTEMATIC)
    eng.startup()
    for _ in range(64):
        eng.update()
    c = eng.ctl.read()
    assert c["done"] == 1
    hist = eng.ctl.read_hist(int(c["iter"]) + 1)
    gb = G["smc_beta"]
    npt.assert_allclose(hist["beta"][:4], gb[:4], rtol=2e-3)
    assert abs(len(hist["beta"]) - len(gb)) <= 3
    assert abs(hist["log_z"][-1] - G["smc_log_z"][-1]) < 0.5
    assert hist["beta"][-1] == 1.0


def test_golden_particle_filter_c1(E):
    import torch
    e, m, l = E
    s = m.make_lg_ssm([0.0], [[1.0]], [[0.9]], [[0.5]], [[1.0]], [[0.3]])
    y = torch.as_tensor(G["pf_y"].astype(np.float32), device="cuda")
    eng = e.PFEngine(s, 2000, 5, ess_threshold=0.5, resampling=l.RESAMPLE_SYSTEMATIC)
    eng.init(y[0])
    for t in range(1, len(y)):
        eng.step(y[t])
    hist = eng.ctl.read_hist(len(y))
    npt.assert_allclose(hist["log_z"][:3], G["pf_log_z"][:3], atol=2e-4)      # identical streams before ancestor flips
    assert abs(hist["log_z"][-1] - G["pf_log_z"][-1]) < 0.15
    assert abs(hist["log_z"][-1] - float(G["kalman_loglik"])) < 0.5


def test_golden_exact_systematic_ancestors(E):
    """fused resampler (linear mode, caller's u0 bits) against the stored exact-rational ancestors"""
    import torch
    e, m, l = E
    L = l.get()
    w = _t(G["cdf_w"])
    n = w.numel()
    ws = torch.zeros((int(L.dll.mb_rs_workspace_bytes(n)) + 7) // 8, dtype=torch.int64, device="cuda")
    for k0, ref in zip(G["anc_exact_k0"], G["anc_systematic_exact"]):
        anc = torch.empty(n, dtype=torch.int32, device="cuda")
        L.call("mb_rs_tile_sums", L.ctx(), l.ptr(ws), l.ptr(w), n, n, 0, None, 1, l.stream())
        L.call("mb_rs_ancestors", L.ctx(), l.ptr(ws), l.ptr(w), n, n, 0, None, 1, int(k0), None, None, l.ptr(anc), l.stream())
        npt.assert_array_equal(anc.cpu().numpy(), ref)


def test_golden_lorenz96_step(E):
    import torch
    e, m, l = E
    s = m.make_lorenz96(dim=8)
    y = torch.as_tensor(G["l96_y"].astype(np.float32), device="cuda")
    eng = e.PFEngine(s, 96, 13, ess_threshold=2.0, resampling=l.RESAMPLE_SYSTEMATIC)
    eng.init(y[0])
    npt.assert_allclose(eng.values().cpu().numpy(), G["l96_x0"], atol=2e-5)
    npt.assert_allclose(eng.lw.cpu().numpy(), G["l96_lw0"], rtol=2e-5, atol=1e-3)
    eng.step(y[1])
    same = eng.anc.cpu().numpy() == G["l96_anc"]
    assert same.mean() > 0.97
    npt.assert_allclose(eng.values().cpu().numpy()[same], G["l96_x1"][same], atol=6e-5, rtol=1e-5)
    npt.assert_allclose(eng.lw.cpu().numpy()[same], G["l96_lw1"][same], rtol=3e-5, atol=2e-3)
    npt.assert_allclose(eng.ctl.read()["log_z"], G["l96_log_z"][1], atol=5e-3)
