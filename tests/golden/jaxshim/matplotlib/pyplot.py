"""stub: only the names the reference touches at import time (type annotations)"""


class Axes:
    pass


def subplots(*a, **k):
    raise RuntimeError("matplotlib is stubbed in the fixture generator")


class Figure:
    pass
