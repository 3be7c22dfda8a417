"""mocat_b200 -- B200-native particle-population hot path behind mocat's API.

Python host shell (mirrors mocat's Scenario / Sampler / run / ssm / abc API) over hand-written
sm_100a CUDA kernels reached through the C-ABI in include/mocat_b200.h.  No CPU fallback.
"""
from . import _lib, engine  # noqa: F401

__version__ = "0.1.0"
