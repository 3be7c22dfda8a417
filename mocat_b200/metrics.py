"""metrics: host mirror of mocat/src/metrics.py:69-78 (the ESS reduction every sampler uses)."""
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
