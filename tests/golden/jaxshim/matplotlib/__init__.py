"""stub: the reference imports matplotlib for its plot helpers only"""
from . import pyplot, figure  # noqa: F401
