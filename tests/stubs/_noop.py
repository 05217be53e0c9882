"""A module-like object that swallows every attribute access and call (plotting stubs)."""
import types


class Noop(object):
    def __call__(self, *a, **k):
        return Noop()

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return Noop()

    def __getitem__(self, k):
        return Noop()

    def __iter__(self):
        return iter(())


def noop_module(name, **fixed):
    m = types.ModuleType(name)
    m.__dict__.update(fixed)
    m.__getattr__ = lambda attr: Noop()  # PEP 562
    return m
