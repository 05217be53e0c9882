"""bode_b200 -- B200-native EKLD acquisition for Bayesian optimal design of experiments (Q = E[f]).

Public surface mirrors the reference package (``bode/__init__.py:1-2`` star-imports ``_core`` and ``_sampler``):
``KLSampler`` plus ``xik, bk, JeffreysPrior, BetaPrior, GammaPrior, UniformPrior, ExponentialPrior, ei``.
Added: ``EkldEngine`` / ``ekld_score`` (device-resident dense scoring), ``lhs`` (pyDOE-compatible stand-in).
Importing the package does not need a GPU; constructing an engine / sampler does (there is no CPU fallback).
"""
from ._core import *  # noqa: F401,F403
from ._core import __all__ as _core_all
from ._lhs import lhs  # noqa: F401
from ._sampler import *  # noqa: F401,F403
from ._sampler import __all__ as _sampler_all


def __getattr__(name):
    if name in ("EkldEngine", "ekld_score"):
        from . import engine
        return getattr(engine, name)
    raise AttributeError(name)


__all__ = list(_core_all) + list(_sampler_all) + ["lhs", "EkldEngine", "ekld_score"]
__version__ = "0.1.0"
