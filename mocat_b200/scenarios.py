"""Built-in device scenario families (mocat/src/scenarios/toy_examples.py).

The reference attaches the prior by assigning Python callables (`scenario.prior_sample = lambda rk: ...`,
`scenario.prior_potential = lambda x, rk: ...`, e.g. tests/test_transport.py:27-28); here the prior is an
isotropic Gaussian given by `prior_mean`, `prior_std` (sampling) and `prior_pscale` (potential
0.5*sum(((x-mean)*pscale)^2), default 1/prior_std) so that it can be compiled into the kernels.
"""
import numpy as np

from . import _lib, models
from .core import Scenario


class Gaussian(Scenario):
    """toy_examples.py:17-51: U_lik = 0.5 |(x - mean) precision_sqrt^T|^2 (no normalising constant)."""
    name = "Gaussian"
    lik_kind = _lib.LIK_GAUSSIAN

    def __init__(self, dim=1, mean=None, covariance=None, **kwargs):
        if mean is not None:
            self.dim = int(np.asarray(mean).shape[-1])
        elif covariance is not None:
            self.dim = int(np.asarray(covariance).shape[-1])
        else:
            self.dim = int(dim)
        self.mean = np.zeros(self.dim) if mean is None else np.asarray(mean, np.float64)
        self.covariance = np.ones(self.dim) if covariance is None else np.asarray(covariance, np.float64)
        super().__init__(**kwargs)

    def _target(self):
        return models.make_target(self.lik_kind, self.dim, self.prior_mean, self.prior_std, self.prior_pscale,
                                  mean=self.mean, covariance=self.covariance)


class Rastrigin(Scenario):
    """toy_examples.py:135-149: U_lik = a d + sum(x^2 - a cos(2 pi x))."""
    name = "Rastrigin"
    lik_kind = _lib.LIK_RASTRIGIN

    def __init__(self, dim=1, a=1., **kwargs):
        self.dim = int(dim)
        self.a = float(a)
        super().__init__(**kwargs)

    def _target(self):
        return models.make_target(self.lik_kind, self.dim, self.prior_mean, self.prior_std, self.prior_pscale, a=self.a)


class LogisticRegression(Scenario):
    """Bayesian logistic regression of config C4 (BASELINE.json configs[3]; the reference ships no such Scenario,
    SURVEY 8d): U_lik(w) = sum_k softplus(a_k.w) - t_k a_k.w for features a_k (N x d) and labels t_k in {0,1}.
    Evaluated for whole ensembles by mb_logistic_potential_grad (SVGD's grad_potential); it is not one of the
    targets compiled into the SMC move kernels."""
    name = "LogisticRegression"
    lik_kind = _lib.LIK_NONE

    def __init__(self, features, labels, variant=None, **kwargs):
        import torch
        self.variant = variant          # None: tensor cores (bf16 operands) for ensembles of n >= 2048, exact fp32 below
        f = np.ascontiguousarray(np.asarray(features, np.float32))
        t = np.ascontiguousarray(np.asarray(labels, np.float32))
        if f.ndim != 2 or t.shape != (f.shape[0],):
            raise _lib.MocatB200Error("LogisticRegression: features (N, d) and labels (N,) expected")
        self.dim = int(f.shape[1])
        self.features = torch.as_tensor(f, device="cuda")
        self.labels = torch.as_tensor(t, device="cuda")
        super().__init__(**kwargs)

    def _target(self):                                                 # prior only (prior sampling in SVGD.startup)
        return models.make_target(_lib.LIK_NONE, self.dim, self.prior_mean, self.prior_std, self.prior_pscale)

    def _potential_grad_device(self, X, temperature):
        from . import engine
        pscale = 1.0 / self.prior_std if self.prior_pscale is None else self.prior_pscale
        variant = engine.interaction_variant(X.shape[0], min(self.dim, 60), self.variant)
        return engine.logistic_potential_grad(self.features, self.labels, self.prior_mean, pscale, temperature, X, variant)
