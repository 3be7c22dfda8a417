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


class cdict:
    def __init__(self, **kwargs):
        self.__dict__.update(kwargs)

    def copy(self):
        return cdict(**self.__dict__)

    def deepcopy(self):
        return copy.deepcopy(self)

    def __repr__(self):
        return f"mocat.cdict({self.__dict__.__repr__()})"

    def save(self, path, overwrite=False):
        save_cdict(self, path, overwrite)

    def __getitem__(self, item):
        if isinstance(item, str):
            return self.__dict__[item]
        out = self.copy()
        for key, attr in out.__dict__.items():
            if (isinstance(attr, np.ndarray) and attr.ndim > 0) \
                    or (isinstance(attr, cdict) and not isinstance(attr, static_cdict)):
                out.__setattr__(key, attr[item])
        return out

    def __add__(self, other):
        out = self.copy()
        if other is None:
            return out
        for key, attr in out.__dict__.items():
            if hasattr(other, key):
                o = other.__dict__[key]
                if isinstance(attr, np.ndarray) or isinstance(o, np.ndarray):
                    out.__setattr__(key, np.append(np.atleast_1d(attr), np.atleast_1d(o), axis=0))
                elif (isinstance(attr, cdict) and not isinstance(attr, static_cdict)
                      and isinstance(o, cdict) and not isinstance(o, static_cdict)) or key == 'time':
                    out.__setattr__(key, attr + o)
        return out

    @property
    def is_empty(self):
        return self.__dict__ == {}

    def keys(self):
        return self.__dict__.keys()

    def __iter__(self):
        return self.__dict__.__iter__()


class static_cdict(cdict):
    pass


def save_cdict(in_cdict, path, overwrite=False):
    path = Path(path)
    if path.suffix != '.cdict':
        path = path.with_suffix('.cdict')
    path.parent.mkdir(parents=True, exist_ok=True)
    if path.exists():
        if overwrite:
            path.unlink()
        else:
            raise RuntimeError(f'File {path} already exists.')
    with open(path, 'wb') as file:
        pickle.dump(in_cdict, file)


def load_cdict(path):
    path = Path(path)
    if not path.is_file():
        raise ValueError(f'Not a file: {path}')
    if path.suffix != '.cdict':
        raise ValueError(f'Not a .cdict file: {path}')
    with open(path, 'rb') as file:
        return pickle.load(file)


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
