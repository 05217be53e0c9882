"""Put the stub modules first on ``sys.path`` and (optionally) import the reference package by path.

TEST INFRASTRUCTURE.  ``/root/reference`` exists only in the build container: callers skip when it is absent.
"""
import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
STUBBED = ("GPy", "pyDOE", "emcee", "seaborn", "matplotlib", "matplotlib.pyplot")


def install():
    if HERE not in sys.path:
        sys.path.insert(0, HERE)
    for name in STUBBED:
        mod = sys.modules.get(name)
        if mod is not None and not str(getattr(mod, "__file__", "")).startswith(HERE):
            raise RuntimeError("a real %s is already imported; the stubs must come first" % name)


def import_reference(name="bode_ref"):
    """The reference's own ``bode`` package (unmodified source, read in place) under a private module name."""
    install()
    if name in sys.modules:
        return sys.modules[name]
    pkg_dir = os.path.join(REF, "bode")
    spec = importlib.util.spec_from_file_location(name, os.path.join(pkg_dir, "__init__.py"),
                                                  submodule_search_locations=[pkg_dir])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod
