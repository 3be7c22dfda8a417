"""vmap / jit / grad as plain Python."""
import numpy as np


def _is_tree(x):
    return hasattr(x, "tree_flatten") and hasattr(type(x), "tree_unflatten")


def _map_leaves(f, x):
    if isinstance(x, (tuple, list)):
        return type(x)(_map_leaves(f, v) for v in x)
    if isinstance(x, dict):
        return {k: _map_leaves(f, v) for k, v in x.items()}
    if _is_tree(x):
        leaves, aux = x.tree_flatten()
        return type(x).tree_unflatten(aux, [_map_leaves(f, v) for v in leaves])
    return f(x)


def _stack(outs):
    first = outs[0]
    if isinstance(first, (tuple, list)):
        return type(first)(_stack([o[k] for o in outs]) for k in range(len(first)))
    if isinstance(first, dict):
        return {k: _stack([o[k] for o in outs]) for k in first}
    if _is_tree(first):
        flat = [o.tree_flatten() for o in outs]
        aux = flat[0][1]
        n = len(flat[0][0])
        return type(first).tree_unflatten(aux, [_stack([f[0][k] for f in flat]) for k in range(n)])
    from .numpy import ShimArray
    out = np.stack([np.asarray(o) for o in outs])
    if out.dtype.kind in 'iu' and out.dtype != np.uint32:           # jax's default integer type
        out = out.astype(np.int32)
    return out.view(ShimArray)


def _axis_size(arg, ax):
    if isinstance(arg, (tuple, list)):
        return _axis_size(arg[0], ax if not isinstance(ax, (tuple, list)) else ax[0])
    if _is_tree(arg):
        leaves = [v for v in arg.tree_flatten()[0] if np.ndim(v) > 0]
        return np.shape(leaves[0])[ax]
    return np.shape(arg)[ax]


def _take(arg, ax, i):
    if ax is None:
        return arg
    if isinstance(ax, (tuple, list)):
        return type(arg)(_take(a, x, i) for a, x in zip(arg, ax))
    return _map_leaves(lambda v: np.take(np.asarray(v), i, axis=ax) if np.ndim(v) > 0 else v, arg)


def vmap(fun, in_axes=0, out_axes=0):
    def mapped(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        n = None
        for a, ax in zip(args, axes):
            if ax is not None:
                n = _axis_size(a, ax)
                break
        outs = [fun(*[_take(a, ax, i) for a, ax in zip(args, axes)]) for i in range(n)]
        return _stack(outs)
    return mapped


def jit(fun=None, static_argnums=None, **kwargs):
    if fun is None:
        return lambda f: f
    return fun


def value_and_grad(fun, argnums=0, has_aux=False):
    def vg(*args):
        x = np.asarray(args[argnums], np.float64)
        val = fun(*args)
        g = np.zeros_like(x)
        it = np.nditer(x, flags=["multi_index"])
        for _ in it:
            idx = it.multi_index
            h = 1e-6 * (1.0 + abs(x[idx]))
            xp, xm = x.copy(), x.copy()
            xp[idx] += h
            xm[idx] -= h
            ap = list(args); ap[argnums] = xp
            am = list(args); am[argnums] = xm
            g[idx] = (float(fun(*ap)) - float(fun(*am))) / (2 * h)
        return val, g
    return vg


def grad(fun, argnums=0):
    vg = value_and_grad(fun, argnums)
    return lambda *args: vg(*args)[1]
