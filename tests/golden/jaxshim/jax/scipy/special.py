import numpy as np
from scipy.special import ndtr, ndtri, gammaln  # noqa: F401


def logsumexp(a, axis=None, b=None, keepdims=False):
    """jax.scipy.special.logsumexp: the maximum is taken out, a non-finite maximum counts as 0"""
    a = np.asarray(a, np.float64)
    amax = np.max(a, axis=axis, keepdims=True)
    amax = np.where(np.isfinite(amax), amax, 0.0)
    e = np.exp(a - amax) if b is None else np.asarray(b) * np.exp(a - amax)
    with np.errstate(divide="ignore"):
        out = np.log(np.sum(e, axis=axis, keepdims=True)) + amax
    return out if keepdims else np.squeeze(out, axis=axis) if axis is not None else out.reshape(())[()]
