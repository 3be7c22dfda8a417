"""jax.random on NumPy generators keyed by the PRNG key.  NOT threefry: streams differ from real jax."""
import numpy as np


def PRNGKey(seed):
    return np.array([0, int(seed) & 0xFFFFFFFF], dtype=np.uint32)


def _rng(key):
    return np.random.default_rng([int(v) for v in np.asarray(key).ravel()])


def split(key, num=2):
    return _rng(key).integers(0, 2 ** 32, size=(int(num), 2), dtype=np.uint32)


def normal(key, shape=(), dtype=None):
    return _rng(key).standard_normal(shape)


def uniform(key, shape=(), dtype=None, minval=0.0, maxval=1.0):
    return minval + (maxval - minval) * _rng(key).random(shape)


def categorical(key, logits, axis=-1, shape=None):
    logits = np.asarray(logits, np.float64)
    if shape is None:
        g = _rng(key).gumbel(size=logits.shape)
        return np.argmax(logits + g, axis=axis)
    g = _rng(key).gumbel(size=tuple(shape) + logits.shape)
    return np.argmax(logits + g, axis=-1)


def choice(key, a, shape=(), replace=True, p=None):
    return _rng(key).choice(a, size=shape, replace=replace, p=p)


def poisson(key, lam, shape=None):
    return _rng(key).poisson(lam, size=shape)
