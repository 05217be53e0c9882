"""matplotlib stand-in -- TEST INFRASTRUCTURE: every call is a no-op."""
from _noop import Noop as _Noop


def use(*a, **k):
    return None


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    return _Noop()
