"""matplotlib.pyplot stand-in -- TEST INFRASTRUCTURE: every call is a no-op."""
from _noop import Noop as _Noop


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    return _Noop()
