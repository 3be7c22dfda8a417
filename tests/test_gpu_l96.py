"""Lorenz-96 bootstrap-filter kernel (csrc/pf_l96.cu: row-major layout moved by TMA, lane-split particles, packed fp32x2) against the
fp64 oracle (oracle/pf.py, oracle/models.py), the reference-flow (Dormand-Prince) cross-check of SURVEY 8c, the row-major
diagnostics / gather kernels and the sharded gather path exercised through single-GPU virtual ranks."""
import ctypes as C
import pickle

import numpy as np
import numpy.testing as npt
import pytest

from oracle import models as omodels, pf as opf

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def E(lib):
    import torch
    from mocat_b200 import _lib, engine, models
    return torch, _lib, engine, models, lib


@pytest.mark.parametrize("d,substeps,n", [(8, 1, 4096), (16, 1, 3000), (40, 1, 4096), (40, 2, 2050), (40, 5, 1999)])
def test_l96_step_parity(E, d, substeps, n):
    """init + two filter steps with resampling every step: values / weights / log-evidence against the fp64 oracle fed
    the same Philox streams (ragged n: partial tiles, odd n: half-empty pair)"""
    torch, l, e, m, lib = E
    seed = 7
    ssm_o = omodels.Lorenz96SSM(dim=d, dt=0.05, substeps=substeps)
    _, y = ssm_o.simulate(3, np.random.default_rng(0), spinup=200)
    s = m.make_lorenz96(dim=d, dt=0.05, substeps=substeps)
    eng = e.PFEngine(s, n, seed, ess_threshold=2.0, resampling=l.RESAMPLE_SYSTEMATIC)
    orc = opf.BootstrapPF(ssm_o, n, seed, ess_threshold=2.0, resampling='systematic')
    yd = torch.as_tensor(y.astype(np.float32), device="cuda")
    eng.init(yd[0])
    st = orc.init(y[0])
    npt.assert_allclose(eng.values().cpu().numpy(), st['x'], atol=2e-5)
    npt.assert_allclose(eng.lw.cpu().numpy(), st['lw'], rtol=2e-5, atol=1e-3)
    npt.assert_allclose(eng.ctl.read()['log_z'], st['log_z'], atol=2e-3)
    for t in (1, 2):
        eng.step(yd[t])
        st_new = orc.step(st, y[t])
        a_dev = eng.anc.cpu().numpy()
        same = a_dev == st_new['ancestors']
        assert np.mean(~same) < 5e-3                 # exp in fp32 (MUFU) vs NumPy: rare boundary flips
        x1 = eng.values().cpu().numpy()
        # fp32 RK4 vs fp64: 4.4e-7 per step on the attractor; fast-math normals dominate
        npt.assert_allclose(x1[same], st_new['x'][same], atol=6e-5, rtol=1e-5)
        c1 = eng.ctl.read()
        npt.assert_allclose(c1['log_z'], st_new['log_z'], atol=6e-3)
        npt.assert_allclose(c1['ess'], st_new['ess'], rtol=3e-2)
        # continue the oracle from the device population so that rare flips do not accumulate
        st = dict(st_new, x=x1.astype(np.float64), lw=eng.lw.cpu().numpy().astype(np.float64),
                  ess=float(c1['ess']), log_z=float(c1['log_z']))


def test_l96_weights_carried_without_resampling(E):
    """ess_threshold = 0: never resample, log-weights accumulate (filtering.py:292,303)"""
    torch, l, e, m, lib = E
    d, n, seed = 40, 1000, 3
    ssm_o = omodels.Lorenz96SSM(dim=d, r_std=6.0)
    _, y = ssm_o.simulate(3, np.random.default_rng(1), spinup=200)
    s = m.make_lorenz96(dim=d, r_std=6.0)
    eng = e.PFEngine(s, n, seed, ess_threshold=0.0, resampling=l.RESAMPLE_SYSTEMATIC)
    orc = opf.BootstrapPF(ssm_o, n, seed, ess_threshold=0.0, resampling='systematic')
    yd = torch.as_tensor(y.astype(np.float32), device="cuda")
    eng.init(yd[0])
    st = orc.init(y[0])
    for t in (1, 2):
        eng.step(yd[t])
        st = orc.step(st, y[t])
        assert not st['resampled'] and eng.ctl.read()['resampled'] == 0
    npt.assert_allclose(eng.values().cpu().numpy(), st['x'], atol=1e-4, rtol=1e-5)
    npt.assert_allclose(eng.lw.cpu().numpy(), st['lw'], rtol=3e-5, atol=2e-3)
    c = eng.ctl.read()
    npt.assert_allclose(c['log_z'], st['log_z'], atol=3e-3)
    npt.assert_allclose(c['ess'], st['ess'], rtol=1e-3)


def test_l96_reference_flow_cross_check(E):
    """the device DEFINES the transition map as `substeps` RK4 steps; the reference integrates with adaptive
    Dormand-Prince (lorenz96.py:23-26).  Filter log-evidence and weighted means of the device (substeps 1 / 5) against
    the oracle filter run on the Dormand-Prince flow with the same random numbers."""
    torch, l, e, m, lib = E
    d, n, seed, T = 40, 1024, 12, 4

    class DopriSSM(omodels.Lorenz96SSM):
        def transition_function(self, x):
            return omodels.lorenz96_dopri(x, self.dt, self.forcing)

    ssm_ref = DopriSSM(dim=d, r_std=2.0)
    _, y = omodels.Lorenz96SSM(dim=d, r_std=2.0).simulate(T, np.random.default_rng(3), spinup=300)
    ref = opf.BootstrapPF(ssm_ref, n, seed, ess_threshold=2.0, resampling='systematic').run(y)
    ref_mean = opf.weighted_moments(ref[-1]['x'], ref[-1]['lw'])[0]
    yd = torch.as_tensor(y.astype(np.float32), device="cuda")
    dev = {}
    for substeps in (1, 5):
        s = m.make_lorenz96(dim=d, substeps=substeps, r_std=2.0)
        eng = e.PFEngine(s, n, seed, ess_threshold=2.0, resampling=l.RESAMPLE_SYSTEMATIC)
        eng.init(yd[0])
        for t in range(1, T):
            eng.step(yd[t])
        mean = eng.moments()[0].cpu().numpy()
        dev[substeps] = (float(eng.ctl.read()['log_z']), mean)
        print(f"substeps={substeps}: log_z {dev[substeps][0]:.4f} vs Dormand-Prince flow {ref[-1]['log_z']:.4f}; "
              f"max |mean diff| {np.max(np.abs(mean - ref_mean)):.4f}")
    # 5 substeps: 6.6e-6 flow deviation -> same particle system up to rare ancestor flips
    assert abs(dev[5][0] - ref[-1]['log_z']) < 0.05
    assert np.max(np.abs(dev[5][1] - ref_mean)) < 0.05
    # 1 substep (throughput setting): 4.5e-3 flow deviation, well inside the Monte-Carlo error of the filter
    assert abs(dev[1][0] - ref[-1]['log_z']) < 0.5
    assert np.max(np.abs(dev[1][1] - ref_mean)) < 0.3


def test_rows_moments_and_skip(E):
    torch, l, e, m, lib = E
    d, n = 40, 5000
    s = m.make_lorenz96(dim=d)
    eng = e.PFEngine(s, n, 1, ess_threshold=2.0, resampling=l.RESAMPLE_SYSTEMATIC)
    y = torch.full((d,), 0.3, device="cuda")
    eng.init(y)
    x, lw = eng.values().cpu().numpy().astype(np.float64), eng.lw.cpu().numpy().astype(np.float64)
    mean, var = (t.cpu().numpy() for t in eng.moments())
    rm, rv = opf.weighted_moments(x, lw)
    npt.assert_allclose(mean, rm, atol=1e-5)
    npt.assert_allclose(var, rv, rtol=1e-4, atol=1e-6)
    # collapsed weights: rows of weightless particles are skipped, the result is still exact
    eng._lw_full[:n] = torch.as_tensor(np.where(np.arange(n) % 977 == 5, 0.0, -500.0).astype(np.float32), device="cuda")
    c = eng.ctl.read()
    c['wmax'] = 0.0
    eng.ctl.write(c)
    mean, _ = (t.cpu().numpy() for t in eng.moments())
    npt.assert_allclose(mean, x[np.arange(n) % 977 == 5].mean(0), atol=1e-5)


@pytest.mark.parametrize("d", [8, 40])
def test_gather_rows_staged_and_direct(E, d):
    """TMA gather (cp.async.bulk of the span of source rows, or of one row per ancestor; bulk store of the result) ==
    per-element gather == NumPy, for sorted ancestors (short spans), collapsed ancestors and random ancestors (one copy
    per row), with a ragged last group of outputs"""
    torch, l, e, m, lib = E
    n = 10_007
    ld = (n + 31) // 32 * 32
    rng = np.random.default_rng(d)
    src = torch.randn((ld, d), device="cuda")
    rows = src[:n].cpu().numpy()
    for kind in ("sorted", "collapsed", "random"):
        if kind == "sorted":
            anc = np.sort(rng.integers(n, size=n))
        elif kind == "collapsed":
            anc = np.sort(rng.choice(rng.integers(n, size=7), size=n))
        else:
            anc = rng.integers(n, size=n)
        ad = torch.as_tensor(anc.astype(np.int32), device="cuda")
        for staged in (1, 0):
            dst = torch.zeros_like(src)
            lib.call("mb_gather_rows", lib.ctx(), l.ptr(ad), n, d, l.ptr(src), n, l.ptr(dst), staged, l.stream())
            got = dst[:n].cpu().numpy()
            assert np.array_equal(got, rows[anc]), (kind, staged)
            assert float(dst[n:].abs().sum()) == 0.0                   # nothing written past the last output


def test_l96_sharded_step_virtual_ranks(E):
    """the whole sharded filter step as 4 virtual ranks on one GPU -- per-rank tile sums, totals, pass B (ancestors pushed
    to the rank that owns the output), pass C (every rank pulls its share of the heavy tiles), then the step kernel
    (ancestor rows fetched by TMA from whichever rank's buffer owns them) -- reproduces the single-population step bit
    for bit, for a spread-out and for a collapsed weight profile"""
    torch, l, e, m, lib = E
    d, world, nl, seed = 40, 4, 32 * 320, 9
    n = world * nl
    s = m.make_lorenz96(dim=d)
    _, y = omodels.Lorenz96SSM(dim=d).simulate(2, np.random.default_rng(0), spinup=100)
    yd = torch.as_tensor(y.astype(np.float32), device="cuda")
    for collapse in (False, True):
        eng = e.PFEngine(s, n, seed, ess_threshold=2.0, resampling=l.RESAMPLE_SYSTEMATIC)
        eng.init(yd[0])
        if not collapse:                                   # flatten the weights: every particle keeps ~1 offspring, the
            eng._lw_full.mul_(0.02)                        # ancestors near the shard boundaries live on the neighbour
            o = e.lse_ess(eng.lw).cpu().numpy()
            c = eng.ctl.read(); c['wmax'], c['s1'], c['s2'] = o[0], o[1], o[2]; eng.ctl.write(c)
        x0, lw0, ctl0 = eng.x.clone(), eng._lw_full.clone(), eng.ctl.t.clone()
        eng.step(yd[1])
        x_ref, anc_ref, lw_ref = eng.x.clone(), eng.anc.clone(), eng._lw_full.clone()
        anc = torch.full((n,), -7, dtype=torch.int32, device="cuda")
        wss = [torch.zeros((int(lib.dll.mb_rs_workspace_bytes(nl)) + 7) // 8, dtype=torch.int64, device="cuda") for _ in range(world)]
        ctls = []
        for r in range(world):
            ctl = e.ControlBlock(); ctl.t.copy_(ctl0); ctls.append(ctl)
            lib.call("mb_rs_tile_sums", lib.ctx(), l.ptr(wss[r]), l.ptr(lw0[r * nl:]), nl, n, 1, l.ptr(ctl.t), 0, l.stream())
        totals = torch.stack([ws[0] for ws in wss]).contiguous()
        shards = []
        for r in range(world):
            sh = l.Shard()
            sh.rank, sh.world, sh.n_local, sh.n_total = r, world, nl, n
            for q in range(world):
                sh.x_peers[q] = x0[q * nl:].data_ptr()
                sh.anc_peers[q] = anc[q * nl:].data_ptr()
                sh.lw_peers[q] = lw0[q * nl:].data_ptr()
                sh.ws_peers[q] = wss[q].data_ptr()
            shards.append(sh)
        for r in range(world):
            lib.call("mb_rs_ancestors", lib.ctx(), l.ptr(wss[r]), l.ptr(lw0[r * nl:]), nl, n, 1, l.ptr(ctls[r].t), 0, -1,
                     l.ptr(totals), C.byref(shards[r]), l.ptr(anc[r * nl:]), l.stream())
        for r in range(world):
            lib.call("mb_rs_heavy", lib.ctx(), l.ptr(wss[r]), l.ptr(lw0[r * nl:]), nl, n, 1, l.ptr(ctls[r].t), 0, -1,
                     l.ptr(totals), C.byref(shards[r]), l.ptr(anc[r * nl:]), l.stream())
        assert torch.equal(anc, anc_ref)
        owner = (anc.cpu().numpy().astype(np.int64) // nl)
        assert (owner != (np.arange(n) // nl)).any()       # some ancestors do live on another (virtual) rank
        x_out = torch.zeros_like(x0)
        lw = lw0.clone()
        for r in range(world):
            lib.call("mb_pf_l96_step", lib.ctx(), C.byref(s), l.ptr(x0[r * nl:]), l.ptr(x_out[r * nl:]), nl, n,
                     l.ptr(anc[r * nl:]), l.ptr(yd[1]), l.ptr(lw[r * nl:]), seed, 1, r * nl, 2.0, l.ptr(ctls[r].t), None,
                     C.byref(shards[r]), None, l.stream())
        assert torch.equal(lw[:n], lw_ref[:n])
        assert torch.equal(x_out[:n], x_ref[:n])


def test_pf_api_resample_continue_and_pickle(E, tmp_path):
    """resample_particles (filtering.py:202-217), initial_sample continuation (:266-276) and cdict persistence"""
    import mocat_b200 as mocat
    torch, l, e, m, lib = E
    d, n = 8, 3000
    _, y = omodels.Lorenz96SSM(dim=d).simulate(9, np.random.default_rng(0), spinup=300)
    ssm, pf = mocat.ssm.Lorenz96(dim=d), mocat.ssm.BootstrapFilter()
    t = np.arange(9) * 0.05
    full = mocat.ssm.run_particle_filter_for_marginals(ssm, pf, y, t, 5, n=n, resampling='systematic')
    part = mocat.ssm.run_particle_filter_for_marginals(ssm, pf, y[:4], t[:4], 5, n=n, resampling='systematic')
    cont = mocat.ssm.run_particle_filter_for_marginals(ssm, pf, y[4:], t[4:], 5, initial_sample=part,
                                                       resampling='systematic')
    assert cont.value.shape == full.value.shape == (9, n, d)
    npt.assert_array_equal(cont.value, full.value)                  # same engine state, same Philox steps
    npt.assert_array_equal(cont.ess, full.ess)
    npt.assert_allclose(cont.log_norm_constant, full.log_norm_constant, rtol=0, atol=0)
    # resample_particles: weights reset, population drawn from the weighted one
    p0 = mocat.ssm.initiate_particles(ssm, pf, n, 3, y[0], 0.0, resampling='systematic')
    mean_w = p0.mean[-1]
    p1 = mocat.ssm.resample_particles(p0, 3)
    assert np.all(p1.log_weight[-1] == 0.0) and p1.ess[-1] == n
    se = np.sqrt(p0.var[-1] / max(p0.ess[-1], 1.0))
    assert np.all(np.abs(p1.value[-1].mean(0) - mean_w) < 6 * se + 1e-3)
    p2 = mocat.ssm.propagate_particle_filter(ssm, pf, p1, y[1], 0.05, 3)
    assert p2.value.shape == (2, n, d) and np.isfinite(p2.log_norm_constant[-1])
    # persistence: the live engine is dropped, the arrays survive
    p2.save(tmp_path / "pf.cdict")
    back = mocat.load_cdict(tmp_path / "pf.cdict")
    assert not hasattr(back, 'engine')
    npt.assert_array_equal(back.value, p2.value)
    assert isinstance(pickle.dumps(full), bytes)
    with pytest.raises(mocat.MocatB200Error):
        mocat.ssm.propagate_particle_filter(ssm, pf, back, y[2], 0.10, 3)


@pytest.mark.parametrize("d,n,q,r,p0", [(8, 4096, 1.0, 1.0, 1.0), (40, 2050, 0.7, 1.3, 2.0), (16, 1999, 1.0, 0.5, 1.0)])
def test_l96_optimal_proposal_parity(E, d, n, q, r, p0):
    """OptimalNonLinearGaussianParticleFilter (ssm/nonlinear_gaussian.py:134-276) compiled into the Lorenz-96 kernels:
    conditioned initial sample with zero weights, proposal mx + Kp (y - mx) + sd_p z, weights from the prediction --
    values / weights / log-evidence against oracle.pf.OptimalPF on the same Philox streams"""
    torch, l, e, m, lib = E
    seed = 11
    ssm_o = omodels.Lorenz96SSM(dim=d, dt=0.05, q_std=q, r_std=r, init_std=p0, init_mean=0.5)
    _, y = ssm_o.simulate(4, np.random.default_rng(0), spinup=200)
    s = m.make_lorenz96(dim=d, dt=0.05, q_std=q, r_std=r, init_std=p0, init_mean=0.5)
    s.proposal = l.PROPOSAL_OPTIMAL
    eng = e.PFEngine(s, n, seed, ess_threshold=0.5, resampling=l.RESAMPLE_SYSTEMATIC)
    orc = opf.OptimalPF(ssm_o, n, seed, ess_threshold=0.5, resampling='systematic')
    yd = torch.as_tensor(y.astype(np.float32), device="cuda")
    eng.init(yd[0])
    st = orc.init(y[0])
    npt.assert_allclose(eng.values().cpu().numpy(), st['x'], atol=3e-5)
    assert np.all(eng.lw.cpu().numpy() == 0.0)                       # initial_log_weight = 0 (:209-214)
    c0 = eng.ctl.read()
    npt.assert_allclose(c0['log_z'], 0.0, atol=1e-6)
    npt.assert_allclose(c0['ess'], n, rtol=1e-6)
    for t in (1, 2, 3):
        eng.step(yd[t])
        st_new = orc.step(st, y[t])
        c1 = eng.ctl.read()
        assert bool(c1['resampled']) == st_new['resampled']
        same = np.ones(n, bool)
        if st_new['resampled']:
            same = eng.anc.cpu().numpy() == st_new['ancestors']
            assert np.mean(~same) < 5e-3
        x1 = eng.values().cpu().numpy()
        npt.assert_allclose(x1[same], st_new['x'][same], atol=8e-5, rtol=1e-5)
        npt.assert_allclose(eng.lw.cpu().numpy()[same], st_new['lw'][same], rtol=3e-5, atol=2e-3)
        npt.assert_allclose(c1['log_z'], st_new['log_z'], atol=6e-3)
        npt.assert_allclose(c1['ess'], st_new['ess'], rtol=3e-2)
        st = dict(st_new, x=x1.astype(np.float64), lw=eng.lw.cpu().numpy().astype(np.float64),
                  ess=float(c1['ess']), log_z=float(c1['log_z']))


def test_l96_optimal_proposal_api_beats_bootstrap(E):
    """through the reference API: the optimal proposal keeps a far larger ESS than the bootstrap filter on the same
    data and both estimate the same log-evidence (n large enough for the bootstrap estimate to settle)"""
    import mocat_b200 as mocat
    from mocat_b200 import ssm
    sc = ssm.Lorenz96(dim=8, initial_mean=3.0)
    rng = np.random.default_rng(0)                   # data consistent with the prior, so that the bootstrap estimate is reliable
    x = 3.0 + rng.standard_normal(8)
    ys = []
    for t in range(6):
        if t > 0:
            x = sc._flow(x, 0.05) + rng.standard_normal(8)
        ys.append(x + rng.standard_normal(8))
    from mocat_b200.core import cdict
    sim = cdict(y=np.array(ys), t=np.arange(6) * 0.05)
    out = {}
    for name, pf in (("boot", ssm.BootstrapFilter()), ("opt", ssm.OptimalNonLinearGaussianParticleFilter())):
        out[name] = ssm.run_particle_filter_for_marginals(sc, pf, sim.y, sim.t, 5, n=1 << 20, ess_threshold=0.5,
                                                          resampling='systematic')
    assert np.all(out["opt"].ess[1:] > 2.0 * out["boot"].ess[1:])
    # zero initial log-weights (nonlinear_gaussian.py:209-214): the optimal filter's evidence lacks p(y_0) = N(y_0; 0, 2 I)
    lp_y0 = -0.5 * np.sum((sim.y[0] - 3.0) ** 2) / 2.0 - 0.5 * 8 * np.log(2 * np.pi * 2.0)
    npt.assert_allclose(out["opt"].log_norm_constant[-1] + lp_y0, out["boot"].log_norm_constant[-1], atol=0.3)
    npt.assert_allclose(out["opt"].mean[-1], out["boot"].mean[-1], atol=0.15)
    with pytest.raises(Exception):
        ssm.run_particle_filter_for_marginals(ssm.TimeHomogenousLinearGaussian(dim=1), ssm.OptimalNonLinearGaussianParticleFilter(),
                                              sim.y[:, :1], sim.t, 5, n=1000)


@pytest.mark.parametrize("d,n", [(8, 3000), (40, 4097), (16, 501)])
def test_enkf_step_parity(E, d, n):
    """EnsembleKalmanFilter (ssm/nonlinear_gaussian.py:279-350) on the device: ensemble mean / covariance / gain and the
    analysed ensemble of two steps against oracle.pf.EnKF on the same Philox streams"""
    torch, l, e, m, lib = E
    seed = 5
    ssm_o = omodels.Lorenz96SSM(dim=d, dt=0.05, r_std=0.8, init_mean=1.0, init_std=2.0)
    _, y = ssm_o.simulate(3, np.random.default_rng(0), spinup=200)
    s = m.make_lorenz96(dim=d, dt=0.05, r_std=0.8, init_mean=1.0, init_std=2.0)
    s.proposal = l.PROPOSAL_ENKF
    eng = e.PFEngine(s, n, seed, ess_threshold=0.5, resampling=l.RESAMPLE_SYSTEMATIC)
    orc = opf.EnKF(ssm_o, n, seed)
    yd = torch.as_tensor(y.astype(np.float32), device="cuda")
    eng.init(yd[0])
    st = orc.init(y[0])
    npt.assert_allclose(eng.values().cpu().numpy(), st['x'], atol=3e-5)
    for t in (1, 2):
        eng.step(yd[t])
        st = orc.step(st, y[t])
        npt.assert_allclose(eng._enkf_mean.cpu().numpy(), st['mean'], atol=2e-5 * (1 + np.abs(st['mean']).max()))
        npt.assert_allclose(eng._enkf_cov.cpu().numpy(), st['cov'], atol=2e-5 * np.abs(st['cov']).max())
        npt.assert_allclose(eng._enkf_gain.cpu().numpy(), st['gain'], atol=2e-5)
        x1 = eng.values().cpu().numpy()
        npt.assert_allclose(x1, st['x'], atol=2e-4, rtol=1e-5)
        assert np.all(eng.lw.cpu().numpy() == 0.0)
        c = eng.ctl.read()
        assert c['ess'] == n and c['resample'] == 0 and c['log_z'] == 0.0
        st = dict(st, x=x1.astype(np.float64))


def test_enkf_api_tracks_lorenz96(E):
    """through the reference API: the ensemble Kalman filter with 2000 members tracks a 40-dimensional Lorenz-96
    trajectory (where the bootstrap filter of the same size degenerates): the ensemble-mean error stays below the
    observation noise and the ensemble spread is commensurate with it"""
    from mocat_b200 import ssm
    sc = ssm.Lorenz96(dim=40)
    tt = np.arange(40) * 0.05
    sim = sc.simulate(tt, 1, spinup=500)
    out = ssm.run_particle_filter_for_marginals(sc, ssm.EnsembleKalmanFilter(), sim.y, sim.t, 2, n=2000)
    rmse = np.sqrt(np.mean((out.mean[10:] - sim.x[10:]) ** 2))
    assert rmse < 0.8, rmse                                            # observation noise std is 1
    spread = np.sqrt(np.mean(out.var[10:]))
    assert 0.3 * rmse < spread < 3.0 * rmse, (spread, rmse)
    npt.assert_allclose(out.ess, 2000.0)
    assert not np.any(out.resampled)
