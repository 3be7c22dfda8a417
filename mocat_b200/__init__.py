"""mocat_b200 -- B200-native particle-population hot path behind mocat's API.

Python host shell (mirrors mocat's Scenario / Sampler / run / ssm / abc API: mocat/__init__.py:1-63) over
hand-written sm_100a CUDA kernels reached through the C-ABI in include/mocat_b200.h.  No CPU fallback:
without the built library and a B200 every compute call raises.
"""
from . import _lib, engine, models  # noqa: F401
from .core import cdict, static_cdict, save_cdict, load_cdict, Scenario  # noqa: F401
from .sample import Sampler, run  # noqa: F401
from .mcmc import MCMCSampler, RandomWalk, Underdamped, Overdamped, Metropolis  # noqa: F401
from .transport import (TransportSampler, SMCSampler, TemperedSMCSampler, MetropolisedSMCSampler,  # noqa: F401
                        RMMetropolisedSMCSampler, SVGD, adagrad)
from .teki import TemperedEKI, AdaptiveTemperedEKI  # noqa: F401
from . import scenarios, kernels, metrics, ssm, abc, transport, history, online_smoothing  # noqa: F401
from .metrics import log_ess_log_weight, ess_log_weight  # noqa: F401
from ._lib import MocatB200Error  # noqa: F401

__version__ = "0.1.0"
