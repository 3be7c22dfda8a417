"""metrics: host mirror of mocat/src/metrics.py:69-78 (the ESS reduction every sampler uses) and :88-130 (kernelised
Stein discrepancy)."""
import numpy as np

from . import engine
from .core import cdict


def _lw_tensor(log_weight):
    import torch
    if isinstance(log_weight, cdict) and hasattr(log_weight, 'log_weight'):
        log_weight = log_weight.log_weight
    if isinstance(log_weight, torch.Tensor):
        return log_weight.to(device="cuda", dtype=torch.float32).contiguous()
    if not isinstance(log_weight, np.ndarray):
        raise TypeError('log_weight must be an array or cdict with log_weight attribute')
    return torch.as_tensor(np.ascontiguousarray(log_weight, dtype=np.float32), device="cuda")


def log_ess_log_weight(log_weight):
    """2 * logsumexp(w) - logsumexp(2 w), computed by mb_lse_ess."""
    return float(engine.lse_ess(_lw_tensor(log_weight))[5].item())


def ess_log_weight(log_weight):
    return float(np.exp(log_ess_log_weight(log_weight)))


def logsumexp(log_weight):
    return float(engine.lse_ess(_lw_tensor(log_weight))[3].item())


def ksd(sample, kernel, grad_potential=None, log_weight=None, ensemble_batchsize=None, random_key=None,
        stein_sign='reference', **kernel_params):
    """metrics.py:88-130: sqrt(sum_ij k0(x_i, x_j) w_i w_j) / sum w under the Gaussian kernel, one device contraction
    (mb_ksd).  stein_sign='reference' contracts the kernel gradients with grad_potential exactly as metrics.py:116-124
    does; 'score' uses -grad_potential (the Stein kernel whose discrepancy vanishes for an exact sample)."""
    import torch
    from . import _lib
    from .kernels import Gaussian
    if not isinstance(kernel, Gaussian):
        raise _lib.MocatB200Error("ksd: only the Gaussian kernel is compiled for the device (no CPU fallback)")
    if ensemble_batchsize is not None:
        raise _lib.MocatB200Error("ksd: ensemble minibatching is not built for the device (the full contraction is used)")
    vals = sample.value if isinstance(sample, cdict) else sample
    if grad_potential is None and isinstance(sample, cdict) and hasattr(sample, 'grad_potential'):
        grad_potential = sample.grad_potential                        # metrics.py:98-101
    elif grad_potential is None:
        raise TypeError('grad_potential not found')

    def dev(a):
        if isinstance(a, torch.Tensor):
            return a.to(device="cuda", dtype=torch.float32).contiguous()
        return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float32), device="cuda")
    X, G = dev(vals), dev(grad_potential)
    if X.ndim != 2 or G.shape != X.shape:
        raise TypeError('ksd: value and grad_potential must both be (n, d)')
    lw = dev(log_weight) if log_weight is not None else None
    h = float(kernel_params.get('bandwidth', kernel.parameters.bandwidth))
    out = torch.empty(3, dtype=torch.float64, device="cuda")
    L = _lib.get()
    L.call("mb_ksd", L.ctx(), _lib.ptr(X), _lib.ptr(G), _lib.ptr(lw), X.shape[0], X.shape[1], h,
           1 if stein_sign == 'reference' else 0, _lib.ptr(out), _lib.stream())
    return float(out[0].item())
