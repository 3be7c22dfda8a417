"""NumPy stand-in for jax (see ../README.md): eager, fp64, test-fixture generation only."""
from . import numpy, random, lax, scipy, tree_util, util, experimental, example_libraries  # noqa: F401
from .api import vmap, jit, grad, value_and_grad  # noqa: F401


class Array:            # scipy's array-API helpers look this name up when a module called `jax` is loaded
    pass
