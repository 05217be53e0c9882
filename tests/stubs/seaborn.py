"""seaborn stand-in -- TEST INFRASTRUCTURE: every call is a no-op; ``color_palette()`` returns RGB tuples."""
from _noop import Noop as _Noop


def color_palette(*a, **k):
    return [(0.1 * i, 0.5, 1.0 - 0.1 * i) for i in range(10)]


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    return _Noop()
