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


class _At:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        arr = self.arr

        class _Setter:
            def set(self, value):
                out = _np.array(arr, copy=True).view(ShimArray)
                out[idx] = value
                return out

            def add(self, value):
                out = _np.array(arr, copy=True).view(ShimArray)
                out[idx] += value
                return out
        return _Setter()


class ShimArray(_np.ndarray):
    """ndarray with the two jax.Array members the reference touches: .at[...].set() and .block_until_ready()"""

    @property
    def at(self):
        return _At(self)

    # jax arrays stay arrays under indexing and arithmetic (no NumPy scalars): the reference tests `isinstance(x,
    # jnp.ndarray) and x.dtype == 'int32'` on values that went through vmap / scan
    def __getitem__(self, idx):
        out = super().__getitem__(idx)
        return out if isinstance(out, _np.ndarray) else _np.asarray(out).view(ShimArray)

    def __array_wrap__(self, obj, context=None, return_scalar=False):
        return _np.asarray(obj).view(ShimArray)

    def block_until_ready(self):
        return self


def array(x, dtype=None, **kw):  # noqa: F811
    return _np.array(x, dtype=dtype).view(ShimArray)


def asarray(x, dtype=None, **kw):  # noqa: F811
    return _np.asarray(x, dtype=dtype).view(ShimArray)


def zeros(shape, dtype=float):  # noqa: F811
    return _np.zeros(shape, dtype=dtype).view(ShimArray)


def ones(shape, dtype=float):  # noqa: F811
    return _np.ones(shape, dtype=dtype).view(ShimArray)


def append(arr, values, axis=None):  # noqa: F811
    return _np.append(arr, values, axis=axis).view(ShimArray)
