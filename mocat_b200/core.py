"""cdict and Scenario: host-side mirror of mocat/src/core.py.

`cdict` keeps mocat's semantics (attribute dict; integer/array indexing applies to every array field,
core.py:46-56; `+` concatenates along axis 0, core.py:58-74; pickle save/load, core.py:91-121) with
NumPy arrays in place of jax arrays.  `Scenario` keeps the public methods (potential, prior_potential,
likelihood_potential, grad_potential, potential_and_grad, tempered_potential, prior_sample, temperature,
dim, name; core.py:146-261) but evaluates them with the CUDA kernels: a scenario must therefore belong to
one of the built-in device families (mocat_b200.scenarios); arbitrary per-particle Python callables
cannot run on the device and raise -- there is no CPU fallback.
"""
import copy
import pickle
from pathlib import Path

import numpy as np


def _stacked(value):
    """array-like state fields take part in indexing / concatenation (core.py:46-56, 58-74)"""
    return isinstance(value, np.ndarray) and value.ndim > 0


def _nested(value):
    return isinstance(value, cdict) and not isinstance(value, static_cdict)


class cdict:
    """Attribute container for sampler state and results (interface of mocat/src/core.py:20-84, NumPy-backed).

    c.field / c['field'] read a field; c[index] applies `index` to every stacked array field (and to nested cdicts);
    a + b appends b's stacked fields to a's along axis 0 (scalars named `time` add up); save/load go through pickle.
    Fields flagged `_mocat_transient` (device engines, CUDA handles) belong to the live session only: they are shared
    by copy() and dropped when the container is pickled."""

    def __init__(self, **fields):
        vars(self).update(fields)

    # -- mapping-like access
    def keys(self):
        return vars(self).keys()

    def __iter__(self):
        return iter(vars(self))

    @property
    def is_empty(self):
        return len(vars(self)) == 0

    def __repr__(self):
        return f"mocat.cdict({vars(self)!r})"

    def __getitem__(self, item):
        if isinstance(item, str):
            return vars(self)[item]
        picked = {k: (v[item] if (_stacked(v) or _nested(v)) else v) for k, v in vars(self).items()}
        return type(self)(**picked) if type(self) is cdict else cdict(**picked)

    # -- copies
    def copy(self):
        return cdict(**vars(self))

    def deepcopy(self):
        keep = {k: v for k, v in vars(self).items() if getattr(v, '_mocat_transient', False)}
        rest = copy.deepcopy({k: v for k, v in vars(self).items() if k not in keep})
        return cdict(**rest, **keep)

    # -- concatenation along the leading (iteration) axis
    def __add__(self, other):
        merged = dict(vars(self))
        if other is None:
            return cdict(**merged)
        theirs = vars(other)
        for key in merged.keys() & theirs.keys():
            mine, new = merged[key], theirs[key]
            if isinstance(mine, np.ndarray) or isinstance(new, np.ndarray):
                merged[key] = np.concatenate([np.atleast_1d(mine), np.atleast_1d(new)], axis=0)
            elif (_nested(mine) and _nested(new)) or key == 'time':
                merged[key] = mine + new
        return cdict(**merged)

    # -- persistence
    def __getstate__(self):
        return {k: v for k, v in vars(self).items() if not getattr(v, '_mocat_transient', False)}

    def __setstate__(self, state):
        vars(self).update(state)

    def save(self, path, overwrite=False):
        save_cdict(self, path, overwrite)


class static_cdict(cdict):
    """a cdict whose fields are NOT indexed or concatenated with the state (parameters, summaries)"""


def _cdict_path(path):
    path = Path(path)
    return path if path.suffix == '.cdict' else path.with_suffix('.cdict')


def save_cdict(in_cdict, path, overwrite=False):
    """pickle to `<path>.cdict` (core.py:91-108); refuses to clobber an existing file unless overwrite"""
    target = _cdict_path(path)
    if target.exists() and not overwrite:
        raise RuntimeError(f"{target} exists; pass overwrite=True to replace it")
    target.parent.mkdir(parents=True, exist_ok=True)
    target.write_bytes(pickle.dumps(in_cdict))


def load_cdict(path):
    """inverse of save_cdict (core.py:111-121)"""
    source = Path(path)
    if source.suffix != '.cdict' or not source.is_file():
        raise ValueError(f"{source}: expected an existing .cdict file")
    return pickle.loads(source.read_bytes())


def key_to_seed(random_key):
    """mocat passes a jax PRNGKey (uint32[2]); accept that, an int, or None -> 64-bit Philox key."""
    if random_key is None:
        return 0
    if isinstance(random_key, (int, np.integer)):
        return int(random_key) & 0xFFFFFFFFFFFFFFFF
    k = np.asarray(random_key).astype(np.uint64).ravel()
    if k.size == 1:
        return int(k[0])
    return int((k[0] << np.uint64(32)) | (k[1] & np.uint64(0xFFFFFFFF)))


class Scenario:
    """Target distribution U(x) = U_prior(x) + temperature * U_lik(x) (core.py:190-194) from a built-in
    device family.  Subclasses set `lik_kind` and fill `_target()`."""
    name = None
    dim = None
    temperature = 1.0
    lik_kind = None

    def __init__(self, name=None, prior_mean=0.0, prior_std=1.0, prior_pscale=None, **kwargs):
        if name is not None:
            self.name = name
        self.prior_mean, self.prior_std = float(prior_mean), float(prior_std)
        self.prior_pscale = prior_pscale
        for key, value in kwargs.items():
            if hasattr(self, key):
                self.__dict__[key] = value
        if self.lik_kind is None:
            raise TypeError(
                f"{type(self).__name__}: only the built-in device scenario families (mocat_b200.scenarios) can be "
                "evaluated; per-particle Python potentials cannot run on the GPU and there is no CPU fallback")

    def __repr__(self):
        return f"mocat.Scenario.{self.__class__.__name__}({self.__dict__.__repr__()})"

    # -- device description -------------------------------------------------------------------------
    def _target(self):
        raise NotImplementedError

    def _potential_grad_device(self, X, temperature):
        """(U (n,), G (n, d)) device tensors for X (n, d) row-major float32 on the device"""
        from . import engine
        return engine.target_potential_grad(self._target(), temperature, X)

    def _eval(self, x, temperature):
        import torch
        x = np.asarray(x, dtype=np.float32)
        single = x.ndim == 1
        X = torch.as_tensor(np.atleast_2d(x), device="cuda").contiguous()
        U, G = self._potential_grad_device(X, temperature)
        U, G = U.cpu().numpy(), G.cpu().numpy()
        return (U[0], G[0]) if single else (U, G)

    # -- mocat API (random_key accepted and ignored: the built-in potentials are deterministic) -------
    def tempered_potential(self, x, temperature, random_key=None):
        return self._eval(x, temperature)[0]

    def tempered_potential_and_grad(self, x, temperature, random_key=None):
        return self._eval(x, temperature)

    def grad_tempered_potential(self, x, temperature, random_key=None):
        return self._eval(x, temperature)[1]

    def potential(self, x, random_key=None):
        return self._eval(x, self.temperature)[0]

    def grad_potential(self, x, random_key=None):
        return self._eval(x, self.temperature)[1]

    def potential_and_grad(self, x, random_key=None):
        return self._eval(x, self.temperature)

    def prior_potential(self, x, random_key=None):
        return self._eval(x, 0.0)[0]

    def grad_prior_potential(self, x, random_key=None):
        return self._eval(x, 0.0)[1]

    def prior_potential_and_grad(self, x, random_key=None):
        return self._eval(x, 0.0)

    def likelihood_potential(self, x, random_key=None):
        return self._eval(x, 1.0)[0] - self._eval(x, 0.0)[0]

    def grad_likelihood_potential(self, x, random_key=None):
        return self._eval(x, 1.0)[1] - self._eval(x, 0.0)[1]

    def likelihood_potential_and_grad(self, x, random_key=None):
        u1, g1 = self._eval(x, 1.0)
        u0, g0 = self._eval(x, 0.0)
        return u1 - u0, g1 - g0

    def prior_sample(self, random_key=None):
        """one draw mean + std * z (host; the samplers draw their initial populations on the device)"""
        rng = np.random.default_rng(key_to_seed(random_key))
        return (self.prior_mean + self.prior_std * rng.standard_normal(self.dim)).astype(np.float32)
