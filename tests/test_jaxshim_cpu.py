"""The NumPy stand-in for jax that the reference-run fixtures were generated with (tests/golden/jaxshim): its vmap / scan /
while_loop / cond / grad / random must behave like the jax functions they stand in for, or the fixtures mean nothing.
Loaded under a private module name so that no `jax` module leaks into the test session."""
import importlib.util
import os
import sys

import numpy as np
import numpy.testing as npt
import pytest

SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "jaxshim")


@pytest.fixture(scope="module")
def jx():
    saved = {k: v for k, v in sys.modules.items() if k == "jax" or k.startswith("jax.")}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, SHIM)
    try:
        import jax                                                      # the stand-in
        assert os.path.dirname(os.path.dirname(jax.__file__)) == SHIM
        yield jax
    finally:
        sys.path.remove(SHIM)
        for k in [k for k in sys.modules if k == "jax" or k.startswith("jax.")]:
            del sys.modules[k]
        sys.modules.update(saved)


def test_vmap_is_a_loop_over_the_mapped_axes(jx):
    x, y = np.arange(12.0).reshape(4, 3), np.arange(3.0)
    npt.assert_allclose(jx.vmap(lambda a, b: a @ b, (0, None))(x, y), x @ y)
    npt.assert_allclose(jx.vmap(lambda a: a.sum(), (1,))(x), x.sum(0))
    out = jx.vmap(lambda a: (a * 2, a.sum()))(x)
    npt.assert_allclose(out[0], 2 * x)
    npt.assert_allclose(out[1], x.sum(1))


def test_scan_while_cond_are_functional(jx):
    carry, ys = jx.lax.scan(lambda c, v: (c + v, c * v), 0.0, np.arange(5.0))
    assert carry == 10.0
    npt.assert_allclose(ys, [0, 0, 2, 9, 24])

    class Box:                                                          # a mutable carried object, like the reference's cdict
        def __init__(self, v):
            self.v = v

    def body(b, _):
        b.v = b.v + 1                                                   # mutates its argument, as the reference does
        return b, b.v
    start = Box(0)
    final, hist = jx.lax.scan(body, start, None, length=3)
    assert final.v == 3 and start.v == 0                                # the caller's object is untouched
    npt.assert_array_equal(hist, [1, 2, 3])                             # every step's value, not three views of the last
    out = jx.lax.while_loop(lambda s: s[0] < 5, lambda s: (s[0] + 1, s[1] * 2), (0, 1))
    assert out == (5, 32)
    assert jx.lax.cond(True, lambda a: a + 1, lambda a: a - 1, 1) == 2
    assert jx.lax.cond(False, lambda a: a + 1, lambda a: a - 1, 1) == 0
    # carried arrays are copied: the caller's array is not mutated by a body that writes in place
    a0 = np.zeros(3)

    def inplace(s):
        s[0][...] += 1
        return (s[0], s[1] + 1)
    jx.lax.while_loop(lambda s: s[1] < 3, inplace, (a0, 0))
    npt.assert_array_equal(a0, np.zeros(3))


def test_grad_by_central_differences(jx):
    f = lambda v: np.sum(v ** 3) + np.sin(v[0])                         # noqa: E731
    x = np.array([0.3, -1.2, 2.0])
    val, g = jx.value_and_grad(f)(x)
    assert val == pytest.approx(f(x))
    npt.assert_allclose(g, 3 * x ** 2 + np.array([np.cos(0.3), 0, 0]), rtol=1e-7)
    npt.assert_allclose(jx.grad(f)(x), g)


def test_random_is_deterministic_per_key_and_arrays_stay_arrays(jx):
    k = jx.random.PRNGKey(3)
    npt.assert_array_equal(jx.random.normal(k, (4,)), jx.random.normal(k, (4,)))
    k1, k2 = jx.random.split(k)
    assert not np.array_equal(jx.random.normal(k1, (4,)), jx.random.normal(k2, (4,)))
    u = jx.random.uniform(k, (1000,), minval=2.0, maxval=3.0)
    assert u.min() >= 2.0 and u.max() < 3.0
    idx = jx.random.categorical(k, np.log(np.array([0.1, 0.0, 0.9]) + 1e-300), shape=(2000,))
    assert set(np.unique(idx)) <= {0, 2} and abs(np.mean(idx == 2) - 0.9) < 0.03
    a = jx.numpy.zeros(5)
    b = a.at[2].set(7.0)
    assert a[2] == 0.0 and b[2] == 7.0 and b.block_until_ready() is b
    stacked = jx.vmap(lambda i: i + 1)(np.arange(3))
    assert stacked.dtype == np.int32 and isinstance(stacked[0] - 1, jx.numpy.ndarray)       # jax's int32 arrays


def test_logsumexp_convention(jx):
    lse = jx.scipy.special.logsumexp
    assert lse(np.array([0.0, np.log(3.0)])) == pytest.approx(np.log(4.0))
    assert lse(np.array([-np.inf, -np.inf])) == -np.inf                 # non-finite maximum counts as 0
    assert lse(np.array([1.0, 2.0]), b=0.5) == pytest.approx(np.log(0.5 * (np.e + np.e ** 2)))
