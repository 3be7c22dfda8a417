"""jax.lax control flow as Python control flow.  jax is functional: a traced function cannot alias or mutate its
inputs, so the carried values are deep-copied around every call (the reference mutates cdict fields in place)."""
import copy

import numpy as np
from .api import _stack, _take, _axis_size


def scan(f, init, xs, length=None):
    carry, ys = init, []
    n = length if xs is None else _axis_size(xs, 0)
    for i in range(n):
        carry, y = f(copy.deepcopy(carry), None if xs is None else _take(xs, 0, i))
        ys.append(copy.deepcopy(y))
    return carry, (_stack(ys) if ys and ys[0] is not None else None)


def while_loop(cond_fun, body_fun, init_val):
    val = init_val
    while bool(np.all(cond_fun(val))):
        val = body_fun(copy.deepcopy(val))
    return val


def cond(pred, true_fun, false_fun, *operands):
    operands = copy.deepcopy(operands)
    return true_fun(*operands) if bool(pred) else false_fun(*operands)


def map(f, xs):  # noqa: A001
    return _stack([f(_take(xs, 0, i)) for i in range(_axis_size(xs, 0))])
