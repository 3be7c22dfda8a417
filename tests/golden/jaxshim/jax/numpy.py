"""jax.numpy -> NumPy (fp64)."""
from numpy import *  # noqa: F401,F403
import numpy as _np
from numpy import linalg, fft, ndarray, newaxis, inf, pi  # noqa: F401

DeviceArray = _np.ndarray
float32, float64, int32 = _np.float32, _np.float64, _np.int32


def array(x, dtype=None, **kw):
    return _np.array(x, dtype=dtype)


def asarray(x, dtype=None, **kw):
    return _np.asarray(x, dtype=dtype)
