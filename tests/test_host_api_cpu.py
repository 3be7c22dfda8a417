"""Host-side mirror of the reference API, checked on CPU the way the reference's own tests do
(mocat/src/tests/test_core.py: Testcdict, TestSampler) plus the small pieces of host logic that decide what runs."""
import numpy as np
import numpy.testing as npt
import pytest

import mocat_b200 as mocat
from mocat_b200 import core, engine, sample


class TestCdict:                                                        # test_core.py:17-76
    def _make(self):
        return core.cdict(test_arr=np.ones((10, 3)), test_float=3.)

    def test_init(self):
        c = self._make()
        assert hasattr(c, 'test_arr') and hasattr(c, 'test_float')
        npt.assert_array_equal(c.test_arr, np.ones((10, 3)))
        assert c.test_float == 3.

    def test_copy(self):
        c = self._make()
        c2 = c.copy()
        assert isinstance(c2, core.cdict) and isinstance(c2.test_float, float)
        c2.test_arr = np.zeros(5)
        c2.test_float = 9.
        npt.assert_array_equal(c.test_arr, np.ones((10, 3)))
        assert c.test_float == 3.

    def test_getitem(self):
        c0 = self._make()[0]
        assert isinstance(c0, core.cdict)
        npt.assert_array_equal(c0.test_arr, np.ones(3))
        assert c0.test_float == 3.
        idx = self._make()[np.array([1, 1, 4])]                         # the ancestor gather of core.py:46-56
        assert idx.test_arr.shape == (3, 3)

    def test_additem(self):
        c = self._make()
        other = core.cdict(test_arr=np.ones((2, 3)), test_float=7., time=25.)
        c.time = 10.
        s = c + other
        assert isinstance(s, core.cdict)
        npt.assert_array_equal(s.test_arr, np.ones((12, 3)))
        assert s.time == 35. and s.test_float == 3.
        npt.assert_array_equal(c.test_arr, np.ones((10, 3)))
        assert c.time == 10.

    def test_save_load(self, tmp_path):                                 # core.py:91-121
        c = self._make()
        c.save(tmp_path / "state")
        back = core.load_cdict(tmp_path / "state.cdict")
        npt.assert_array_equal(back.test_arr, c.test_arr)
        with pytest.raises(RuntimeError):
            c.save(tmp_path / "state")


class TestSampler:                                                      # test_core.py:79-99
    def test_init_and_deepcopy(self):
        s = sample.Sampler(name='test', other=np.zeros(2))
        assert s.name == 'test' and hasattr(s, 'parameters')
        npt.assert_array_equal(s.parameters.other, np.zeros(2))
        s2 = s.deepcopy()
        assert isinstance(s2, sample.Sampler)
        s2.name = 'other'
        s2.parameters.other = 10.
        assert s.name == 'test'
        npt.assert_array_equal(s.parameters.other, np.zeros(2))


def test_key_to_seed_accepts_jax_style_keys():
    assert core.key_to_seed(None) == 0
    assert core.key_to_seed(7) == 7
    assert core.key_to_seed(np.array([1, 2], dtype=np.uint32)) == (1 << 32) | 2      # jax PRNGKey layout uint32[2]
    assert core.key_to_seed(np.array([5], dtype=np.uint32)) == 5


def test_interaction_variant_policy():
    # exact fp32 kernels for small ensembles, tcgen05 (bf16 operands) from n = 2048 when d fits one K panel
    assert engine.interaction_variant(100, 2) == 0
    assert engine.interaction_variant(2048, 50) == 1
    assert engine.interaction_variant(32768, 64) == 0
    assert engine.interaction_variant(32768, 50, 0) == 0 and engine.interaction_variant(10, 2, 1) == 1


def test_samplers_reject_what_the_device_cannot_run():
    with pytest.raises(mocat._lib.MocatB200Error):
        mocat.MetropolisedSMCSampler(object())                          # not a compiled move
    with pytest.raises(TypeError):
        class MyScenario(mocat.Scenario):                               # per-particle Python potential: no CPU fallback
            def likelihood_potential(self, x, random_key=None):
                return 0.0
        MyScenario()
    s = mocat.RMMetropolisedSMCSampler(mocat.Underdamped(stepsize=0.1), rm_stepsize=0.5)
    assert s.parameters.rm_stepsize == 0.5 and s.check_every == 8 and s.mcmc_sampler.tuning.target == 0.651   # adapted on the device: no per-iteration poll
    assert mocat.RandomWalk(stepsize=0.1).tuning.target == 0.234        # standard_mcmc.py:29,84


class _Probe(sample.Sampler):
    name = 'probe'
    flag = 1

    def __init__(self, **kw):
        super().__init__(**kw)
        self.tuning = core.cdict(target=0.5)

    def _run_device(self, scenario, initial_state, initial_extra):
        return core.cdict(value=np.zeros((1, 2)), extra_iter=initial_extra.iter)


def test_sampler_option_routing_and_startup_merge():
    """sample.py:24-72: options named like attributes set attributes, the others become tunable parameters; startup
    kwargs update both, and fill the run's `extra.parameters` only where the caller left a gap"""
    s = _Probe(flag=2, stepsize=0.3, max_iter=50, unset=None)
    assert s.flag == 2 and s.max_iter == 50 and not hasattr(s.parameters, 'flag')
    assert s.parameters.stepsize == 0.3 and s.parameters.unset is None
    extra = core.cdict(parameters=core.cdict(stepsize=None, other=7))
    st, ex = s.startup(None, 10, None, extra, flag=3, stepsize=0.4, not_known=1)
    assert s.flag == 3 and s.parameters.stepsize == 0.4 and not hasattr(s.parameters, 'not_known')
    assert ex is extra and ex.iter == 0
    assert ex.parameters.stepsize == 0.4 and ex.parameters.other == 7 and ex.parameters.unset is None
    ex.iter = 5
    ex.parameters.stepsize = 9.0
    _, ex2 = s.startup(None, 10, None, ex)
    assert ex2.iter == 5 and ex2.parameters.stepsize == 9.0               # caller's values win
    s.max_iter = 2.5
    with pytest.raises(AttributeError):
        s.startup(None, 10, None, core.cdict())


def test_run_driver_protocol():
    """sample.py:110-148: class or instance, n and random_key recorded, summary and wall time attached"""
    sc = core.cdict(name='toy')
    out = sample.run(sc, _Probe, 4, 11, flag=9)
    assert out.summary.sampler == 'probe' and out.summary.scenario == 'toy'
    assert out.summary.parameters is not None and out.summary.tuning.target == 0.5
    assert out.time >= 0.0 and out.extra_iter == 0
    s = _Probe(stepsize=1.0)
    ex = core.cdict(iter=3)
    out = sample.run(core.cdict(), s, 8, None, initial_extra=ex)
    assert s.n == 8 and not hasattr(ex, 'random_key') and out.extra_iter == 3
    assert not hasattr(out.summary, 'scenario') and ex.parameters.stepsize == 1.0
    with pytest.raises(NotImplementedError):
        sample.run(core.cdict(), sample.Sampler(name='serial'), 1, 0)


def test_smc_sampler_startup_chain_without_device(monkeypatch):
    """the whole host start-up path of a tempered SMC run (Sampler -> TransportSampler -> TemperedSMCSampler ->
    MetropolisedSMCSampler) with the device engine replaced by a recorder: options routed, engine configured from
    the sampler's parameters, temperature reset, chain cleaned"""
    from mocat_b200 import transport, _lib

    calls = {}

    class FakeEngine:
        def __init__(self, target, move, temper, n, seed, resampling, schedule):
            calls.update(target=target, move=move, temper=temper, n=n, seed=seed, resampling=resampling, schedule=schedule)

        def startup(self, x0):
            calls['x0'] = x0

    monkeypatch.setattr(engine.SMCEngine, 'acquire',
                        classmethod(lambda cls, target, move, temper, n, seed, resampling=0, schedule=None:
                                    FakeEngine(target, move, temper, n, seed, resampling, schedule)))

    def fake_loop(self, scenario, state, extra):
        assert extra.engine is state.engine and extra.iter == 0
        return core.cdict(temperature=np.array([0.0, 0.4, 1.0]), value=np.zeros((3, 5, 2)))

    monkeypatch.setattr(transport.MetropolisedSMCSampler, '_run_device', fake_loop)
    sc = mocat.scenarios.Rastrigin(dim=2, a=1.0, prior_std=3.0)
    smp = mocat.MetropolisedSMCSampler(mocat.Underdamped(stepsize=0.25, leapfrog_steps=3), mcmc_steps=2,
                                       ess_threshold_retain=0.8, max_iter=77)
    x0 = np.ones((5, 2), np.float32)
    out = mocat.run(sc, smp, 5, np.array([0, 9], dtype=np.uint32), initial_state=mocat.cdict(value=x0),
                    resampling='systematic', ess_threshold_resample=0.3)
    assert calls['n'] == 5 and calls['seed'] == 9 and calls['x0'] is x0
    assert calls['resampling'] == _lib.RESAMPLE_SYSTEMATIC and smp.resampling == 'systematic'        # attribute option
    mv, tp = calls['move'], calls['temper']
    assert mv.kind == _lib.MOVE_MALA and mv.mcmc_steps == 2 and mv.leapfrog_steps == 3
    assert abs(mv.stepsize - 0.25) < 1e-7
    assert tp.ess_retain == 0.8 and tp.ess_resample == 0.3 and tp.max_iter == 77                     # parameter option
    assert calls['target'].dim == 2 and calls['target'].kind == _lib.LIK_RASTRIGIN
    assert sc.temperature == 1.0                                         # clean_chain: smc.py:177-184
    assert out.summary.sampler == smp.name and out.summary.scenario == 'Rastrigin'
    assert out.summary.parameters.mcmc_steps == 2


def test_kalman_filter_host_matches_oracle_and_struct_mirror():
    """kalman_filter_host (ssm/linear_gaussian/kalman.py:16-57 in NumPy; the device version is tested on the GPU) against the
    oracle recursion, and the POD mirror of the model that the filter kernels receive"""
    from oracle import models as omodels, pf as opf
    F = np.array([[0.9, 0.1], [0.0, 0.8]])
    Q = np.array([[0.5, 0.1], [0.1, 0.4]])
    H = np.array([[1.0, 0.5]])
    R = np.array([[0.3]])
    P0 = np.array([[1.5, 0.2], [0.2, 0.7]])
    m0 = np.array([0.3, -0.2])
    sc = mocat.ssm.TimeHomogenousLinearGaussian(initial_mean=m0, initial_covariance=P0, transition_matrix=F,
                                                transition_covariance=Q, likelihood_matrix=H, likelihood_covariance=R)
    assert sc.dim == 2 and sc.dim_obs == 1
    sim = sc.simulate(np.arange(15.0), 4)
    assert sim.x.shape == (15, 2) and sim.y.shape == (15, 1)
    mus, covs, ll = mocat.ssm.kalman_filter_host(sc, sim.y, sim.t, return_log_likelihood=True)
    omus, ocovs, oll = opf.kalman_filter(omodels.LinearGaussianSSM(m0, P0, F, Q, H, R), sim.y)
    npt.assert_allclose(mus, omus, rtol=1e-12, atol=1e-12)
    npt.assert_allclose(covs, ocovs, rtol=1e-12, atol=1e-12)
    npt.assert_allclose(ll, oll, rtol=1e-12)
    s = sc._ssm()                                                        # what mb_pf_init / mb_pf_step receive
    assert s.kind == mocat._lib.SSM_LINEAR_GAUSSIAN and s.dim == 2 and s.dim_obs == 1
    stride = mocat._lib.MB_MAX_SMALL_DIM
    npt.assert_allclose([s.F[0], s.F[1], s.F[stride], s.F[stride + 1]], F.ravel(), rtol=1e-7)
    LQ = np.linalg.cholesky(Q)
    npt.assert_allclose([s.LQ[0], s.LQ[stride], s.LQ[stride + 1]], [LQ[0, 0], LQ[1, 0], LQ[1, 1]], rtol=1e-6)


def test_teki_samplers_route_options_like_the_reference():
    """transport/teki.py:41-72,155-166: a schedule fixes max_temperature / max_iter, the adaptive sampler keeps its search
    parameters in `parameters`; a Python next_temperature callable cannot run on the device"""
    s = mocat.TemperedEKI(temperature_schedule=np.linspace(0.0, 0.8, 9), nugget=1e-4)
    assert s.max_iter == 9 and s.max_temperature == pytest.approx(0.8) and s._mode() == 0
    assert s.parameters.nugget == 1e-4 and s.parameters.term_std == 0.0
    d = mocat.TemperedEKI()
    assert d._mode() == 1 and d.max_iter == 10000 and d.max_temperature == 1.0
    a = mocat.AdaptiveTemperedEKI(ess_threshold=0.7, bisection_tol=1e-4, max_bisection_iter=50, max_iter=30)
    assert a._mode() == 2 and a.max_iter == 30
    assert a._engine_kwargs() == dict(ess_threshold=0.7, tol=1e-4, max_search_iter=50)
    with pytest.raises(mocat._lib.MocatB200Error):
        mocat.TemperedEKI(next_temperature=lambda state, extra: 0.5)
    with pytest.raises(AttributeError):                                     # teki.py:77-78: the scenario needs data
        class NoData:
            data = None
        s.startup(NoData(), 10, None, mocat.cdict())


def test_online_smoothing_host_helpers():
    """the smoother's ESS test (online_smoothing.py:227: resample iff ess < n - 1e-3) and the namespace of the reference
    (mocat/ssm.py exports propagate_particle_smoother{,_pf,_bs})"""
    from mocat_b200 import online_smoothing as osm
    assert osm._ess(np.zeros(100)) == pytest.approx(100.0)
    assert osm._ess(np.log(np.r_[1.0, np.zeros(99) + 1e-300])) == pytest.approx(1.0)
    for name in ("propagate_particle_smoother", "propagate_particle_smoother_pf", "propagate_particle_smoother_bs",
                 "backward_simulation", "forward_filtering_backward_simulation"):
        assert callable(getattr(mocat.ssm, name))
    with pytest.raises(mocat._lib.MocatB200Error):                          # no live engine: loaded from disk
        osm.propagate_particle_smoother_pf(mocat.ssm.TimeHomogenousLinearGaussian(dim=1), mocat.ssm.BootstrapFilter(),
                                           mocat.cdict(value=np.zeros((1, 4, 1)), log_weight=np.zeros((1, 4)), t=np.zeros(1)),
                                           np.zeros(1), 1.0, 0, 2)
