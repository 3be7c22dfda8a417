"""Tempered ensemble Kalman inversion on the device (csrc/teki.cu; transport/teki.py:38-185 upstream) against
oracle/teki.py: update-by-update parity of the ensemble, the temperature rules, and a run through the public API."""
import numpy as np
import numpy.testing as npt
import pytest

from oracle import models as omodels, teki as oteki

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mocat(lib):
    import mocat_b200
    return mocat_b200


def _data(m, prior_max, seed=0):
    from scipy.special import ndtri
    truth = np.array([1.5, 1.0, 1.0, 0.5])
    z = ndtri(np.random.default_rng(seed).random(m))
    e = np.exp(-truth[2] * z)
    return np.sort(truth[0] + truth[1] * (1 + 0.8 * (1 - e) / (1 + e)) * z * (1 + z * z) ** truth[3]), truth


@pytest.mark.parametrize("m,mode", [(8, "schedule"), (4, "adaptive"), (16, "adaptive"), (8, "default")])
def test_teki_updates_match_oracle(mocat, m, mode):
    """startup + updates, compared after every update: value / simulated_data within fp32 simulator accuracy (fast-math
    exp / pow in the g-and-k quantile function), temperatures to 1e-4 (adaptive: regula falsi on fp32 potentials)"""
    from mocat_b200 import teki
    data, _ = _data(m, 2.0)
    n, seed = 3001, 5                                        # ragged: n % 32, n % 256 != 0
    sc = mocat.abc.GKTransformedUniformPrior(data=data, prior_maxs=2.0)
    osc = omodels.GKTransformed(data, prior_max=2.0)
    sched = np.linspace(0.0, 1.0, 6)
    kw = dict(temperature_schedule=sched) if mode == "schedule" else (dict(adaptive=True, ess_threshold=0.8) if mode == "adaptive" else {})
    orc = oteki.TemperedEKI(osc, n, seed, normal_dtype=np.float32, **kw)
    dev_mode = {"schedule": 0, "default": 1, "adaptive": 2}[mode]
    eng = teki.TEKIEngine(sc._device(), n, seed, dev_mode, orc.max_temperature, orc.max_iter, 0.0, 1e-5,
                          schedule=sched if mode == "schedule" else None, ess_threshold=0.8)
    eng.startup()
    st = orc.startup()
    npt.assert_allclose(eng.x.cpu().numpy(), st['x'], atol=2e-5)
    sim0 = eng.sim.cpu().numpy()
    assert np.mean(np.abs(sim0 - st['sim']) <= 2e-3 * (1 + np.abs(st['sim']))) > 0.999
    s = eng.read()
    npt.assert_allclose(np.ctypeslib.as_array(s.prior_stds), orc.prior_stds, rtol=1e-5)
    for it in range(1, 5):
        if orc.terminated(st):
            break
        eng.update()
        st = orc.update(st)
        s = eng.read()
        assert s.iter == it and s.done == 0
        assert abs(s.temperature - st['temperature']) < 2e-4, (s.temperature, st['temperature'])
        gain = np.ctypeslib.as_array(s.gain).reshape(4, 16)[:, :m]
        npt.assert_allclose(gain, st['gain'], rtol=2e-2, atol=2e-3 * np.max(np.abs(st['gain'])))
        x = eng.x.cpu().numpy()
        assert np.mean(np.abs(x - st['x']) < 5e-3) > 0.995, np.max(np.abs(x - st['x']))
        assert s.perturb_nan == st['perturb_nan']
        # carry the device ensemble into the oracle so that every update is compared on its own
        st['x'] = x.astype(np.float64)
        st['sim'] = eng.sim.cpu().numpy().astype(np.float64)
        st['temperature'] = float(s.temperature)


def test_teki_run_api(mocat):
    """mocat.run with AdaptiveTemperedEKI / TemperedEKI on the g-and-k scenario: the temperature ladder is increasing and
    ends at 1, the history is stacked, the ensemble moves towards the data (the oracle's run gives constrained means
    (1.35, 0.95, 0.97, 0.92) at n = 2000) and a collapsed ensemble terminates through term_std"""
    data, truth = _data(8, 2.0)
    sc = mocat.abc.GKTransformedUniformPrior(data=data, prior_maxs=2.0)
    osc = omodels.GKTransformed(data, prior_max=2.0)
    out = mocat.run(sc, mocat.AdaptiveTemperedEKI(ess_threshold=0.9), 4000, random_key=1)
    t = out.temperature
    assert t[0] == 0.0 and t[-1] == 1.0 and np.all(np.diff(t) > 0) and 3 <= len(t) - 1 <= 40
    assert out.value.shape == (len(t), 4000, 4) and out.simulated_data.shape == (len(t), 4000, 8)
    ref = oteki.TemperedEKI(osc, 4000, 1, adaptive=True, ess_threshold=0.9, normal_dtype=np.float32).run()
    post, post_ref = sc.constrain(out.value[-1]), osc.constrain(ref['x'])
    npt.assert_allclose(post.mean(0), post_ref.mean(0), atol=0.05)
    npt.assert_allclose(post.std(0), post_ref.std(0), atol=0.05)
    assert abs(len(t) - len(ref['temperature_schedule'])) <= 1
    assert abs(post.mean(0)[0] - truth[0]) < abs(1.0 - truth[0])          # A moved from the prior mean towards the truth
    # the REFERENCE'S OWN run of the same problem (its source under the NumPy stand-in for jax, n = 1000,
    # tests/golden/make_reference_runs.py): temperature ladder and posterior moments, Monte-Carlo tolerances
    import os
    RR = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_runs_v1.npz"))
    npt.assert_allclose(data, RR["teki_data"], rtol=1e-12)
    assert abs(len(t) - len(RR["teki_adaptive_temperature"])) <= 3
    assert abs(t[1] - RR["teki_adaptive_temperature"][1]) < 0.05
    npt.assert_allclose(post.mean(0), RR["teki_adaptive_mean"], atol=0.08)
    npt.assert_allclose(post.std(0), RR["teki_adaptive_std"], atol=0.06)
    fixed = mocat.run(sc, mocat.TemperedEKI(temperature_schedule=np.linspace(0, 1, 11)), 2000, random_key=2)
    npt.assert_allclose(fixed.temperature, np.linspace(0, 1, 11), atol=1e-12)
    npt.assert_allclose(sc.constrain(fixed.value[-1]).mean(0), post_ref.mean(0), atol=0.08)
    npt.assert_allclose(sc.constrain(fixed.value[-1]).mean(0), RR["teki_schedule_mean"], atol=0.1)
    npt.assert_allclose(sc.constrain(fixed.value[-1]).std(0), RR["teki_schedule_std"], atol=0.08)
    stop = mocat.run(sc, mocat.TemperedEKI(temperature_schedule=np.linspace(0, 1, 11), term_std=10.0), 500, random_key=3)
    assert len(stop.temperature) == 1 and stop.value.shape[0] == 1
    with pytest.raises(mocat.MocatB200Error):
        mocat.TemperedEKI(next_temperature=lambda s, e: 0.5)
