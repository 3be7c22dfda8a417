"""CPU checks of the exact-rational systematic resampling convention (oracle/core.py, mirrored by
csrc/resample_fused.cu) and of the pairwise normal streams (oracle/philox.py, mirrored by csrc/pf_l96.cu)."""
from fractions import Fraction

import numpy as np
import numpy.testing as npt

from oracle import core, philox


def _brute(e, k0, n_out):
    """definition: a_i = min{ j : (i + k0/2^32)/n_out < C_j / S } with rationals"""
    C = np.cumsum(np.asarray(e, dtype=object))
    S = int(C[-1])
    out = []
    for i in range(n_out):
        u = Fraction(i * 2 ** 32 + k0, n_out * 2 ** 32)
        out.append(next(j for j in range(len(e)) if u < Fraction(int(C[j]), S)))
    return np.array(out)


def test_exact_systematic_matches_rational_definition():
    rng = np.random.default_rng(0)
    for n in (1, 2, 5, 64, 257):
        w = rng.random(n).astype(np.float32)
        w[rng.random(n) < 0.2] = 0.0
        if not w.any():
            w[0] = 1.0
        e = core.integer_weights(w)
        for k0 in (0, 1, 2 ** 31, 2 ** 32 - 1):
            npt.assert_array_equal(core.ancestors_systematic_exact(e, k0), _brute(e, k0, n))


def test_exact_systematic_boundary_cases_are_settled_with_integers():
    """weights chosen so that many (i + u0)/n fall EXACTLY on a CDF boundary: u < C/S is strict"""
    n = 1024
    e = np.full(n, 2 ** 30, dtype=np.uint64)                # C_j / S = (j + 1)/n exactly
    a = core.ancestors_systematic_exact(e, 0)               # u_i = i/n = C_{i-1}/S  -> not below C_{i-1}: ancestor i
    npt.assert_array_equal(a, np.arange(n))
    a = core.ancestors_systematic_exact(e, 2 ** 32 - 1)
    npt.assert_array_equal(a, np.arange(n))
    assert np.all(core.ancestors_systematic_exact(np.zeros(7, np.uint64), 5) == 6)     # all-zero: last particle


def test_integer_weights_resolution_and_range():
    assert core.rs_scale_bits(100_000_000) == 36 and core.rs_scale_bits(4096) == 40 and core.rs_scale_bits(2 ** 31 - 1) == 32
    w = np.array([1.0, 2.0 ** -36, 2.0 ** -38, np.nan, -1.0, 0.0], np.float32)
    e = core.integer_weights(w, n_total=100_000_000)
    npt.assert_array_equal(e, np.array([2 ** 36, 1, 0, 0, 0, 0], dtype=np.uint64))
    # the total of n_total weights of size 1 fits 63 bits
    assert 100_000_000 * 2 ** 36 < 2 ** 63


def test_exact_vs_floating_point_evaluation_at_1e7():
    """how far the exact-rational ancestors are from the literal fp64 `cumsum` + `searchsorted` evaluation of the same
    integer weights (expected ~ n^2 2^-53 boundary flips, each by one index)"""
    n = 10_000_000
    rng = np.random.default_rng(1)
    w = rng.random(n, dtype=np.float32)
    e = core.integer_weights(w)
    k0 = 0x12345678
    a = core.ancestors_systematic_exact(e, k0)
    cdf = np.cumsum(e.astype(np.float64)) / float(e.sum(dtype=np.uint64))
    f = np.minimum(np.searchsorted(cdf, (np.arange(n) + k0 / 2.0 ** 32) / n, side='right'), n - 1)
    diff = np.nonzero(a != f)[0]
    assert len(diff) <= 5 and (len(diff) == 0 or np.max(np.abs(a[diff] - f[diff])) == 1)


def test_pairwise_normals_layout_and_law():
    gid = np.arange(2000, dtype=np.uint64)
    z = philox.normals_pairwise(3, gid, 7, philox.P_MOVE, 40)
    assert z.shape == (2000, 40)
    # particle 2m gets the cos branch, 2m+1 the sin branch of the same (u1, u2): equal radius
    x0, x1, x2, x3 = philox.raw(3, gid[::2] >> np.uint64(1), 7, philox.P_MOVE, 0)
    zc, zs = philox.box_muller(x0, x1)
    npt.assert_array_equal(z[0::2, 0], zc)
    npt.assert_array_equal(z[1::2, 0], zs)
    zc, zs = philox.box_muller(x2, x3)
    npt.assert_array_equal(z[0::2, 1], zc)
    npt.assert_array_equal(z[1::2, 1], zs)
    # sharding independent: any subset of global ids gives the same numbers
    npt.assert_array_equal(philox.normals_pairwise(3, gid[37:91], 7, philox.P_MOVE, 40), z[37:91])
    assert abs(z.mean()) < 0.02 and abs(z.std() - 1.0) < 0.02
    assert abs(np.corrcoef(z[0::2, 0], z[1::2, 0])[0, 1]) < 0.1


def test_rk4_flow_deviation_from_reference_flow():
    """SURVEY 8c numbers: one-step deviation of `substeps` RK4 steps from the Dormand-Prince flow on the attractor"""
    from oracle import models
    rng = np.random.default_rng(0)
    x = models.Lorenz96SSM(dim=40).simulate(6, rng, spinup=500)[0]
    ref = models.lorenz96_dopri(x, 0.05)
    dev = {s: float(np.max(np.abs(models.lorenz96_rk4(x, 0.05, substeps=s) - ref))) for s in (1, 2, 5)}
    assert dev[1] < 1.5e-2 and dev[2] < 1e-3 and dev[5] < 3e-5 and dev[1] > dev[2] > dev[5]
