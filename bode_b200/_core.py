"""Host-side helpers that keep the reference's public names (``bode/_core.py:10`` ``__all__``).

These stay on the CPU by design (north star: priors, designs and plotting are host work).  ``xik`` / ``bk`` are the
closed-form integrals of the ARD-RBF kernel over the unit cube; the CUDA kernels evaluate the same expressions on
device (``csrc/prepare.cu``, ``csrc/score.cu``), these NumPy versions exist for API parity and for host-side use.
"""
from __future__ import annotations

import math
import sys

import numpy as np
from scipy import stats
from scipy.special import erf

__all__ = ['xik', 'bk', 'JeffreysPrior', 'BetaPrior', 'GammaPrior', 'UniformPrior', 'ExponentialPrior', 'ei']

LOWER, UPPER = 0.0, 1.0  # integration box per dimension (reference ``_core.py:13-14``)
_ROOT_PI = math.sqrt(math.pi)
_ROOT_2 = math.sqrt(2.0)


def xik(x, ell, ss):
    """Kernel mean embedding  int_[0,1]^d k(x, x') dx'  for every row of ``x`` (reference ``_core.py:86-92``)."""
    x = np.asarray(x, dtype=float)
    ell = np.asarray(ell, dtype=float)
    scale = _ROOT_2 * ell
    one_dim = (_ROOT_PI / _ROOT_2) * ell * (erf((UPPER - x) / scale) - erf((LOWER - x) / scale))
    return ss * np.prod(one_dim, axis=1)


def bk(ell, ss):
    """Double integral of the kernel over the unit cube (reference ``_core.py:95-101``)."""
    ell = np.asarray(ell, dtype=float)
    z = 1.0 / (_ROOT_2 * ell)
    one_dim = 2 * _ROOT_PI * (ell ** 2) * (z * erf(z) + np.exp(-(z ** 2)) / _ROOT_PI - 1.0 / _ROOT_PI)
    return ss * np.prod(one_dim)


def ei(mu, sigma, y_star, mode="min"):
    """Expected improvement with ``sigma`` a *variance* (reference ``_core.py:103-117``)."""
    sd = np.sqrt(sigma)
    if mode == "min":
        gap = y_star - mu
    elif mode == "max":
        gap = mu - y_star
    else:
        return None
    z = gap / sd
    return sd * stats.norm.pdf(z) + gap * stats.norm.cdf(z)


class _Prior(object):
    """log-density ``__call__`` that is -inf on negative support, plus ``sample(size)``."""

    dist = None

    def __call__(self, param):
        if np.any(np.asarray(param) < 0):
            return -np.inf
        return self._logpdf(param)

    def _logpdf(self, param):
        return self.dist.logpdf(param)

    def sample(self, size=1):
        return self.dist.rvs(size=size)


class JeffreysPrior(_Prior):
    """p(theta) ~ 1/theta (improper; cannot be sampled -- the reference exits the process, ``_core.py:25-27``)."""

    def _logpdf(self, param):
        return -1 * np.log(param)

    def sample(self, size=1):
        print("Cannot sample from JeffreysPrior! Choose another prior over this parameter.")
        sys.exit(1)


class BetaPrior(_Prior):
    def __init__(self, a=2, b=5):
        self.a, self.b = a, b
        self.dist = stats.beta(a=a, b=b)


class GammaPrior(_Prior):
    def __init__(self, a=8, scale=1):
        self.a, self.scale = a, scale
        self.dist = stats.gamma(a=a, scale=scale)


class ExponentialPrior(_Prior):
    def __init__(self, scale=1):
        self.scale = scale
        self.dist = stats.expon(scale=scale)


class UniformPrior(_Prior):
    """Uniform on [a, b].  (The reference's ``sample`` calls a non-existent ``beta.uniform``; fixed here.)"""

    def __init__(self, a=0, b=1):
        self.a, self.b = a, b
        self.dist = stats.uniform(loc=a, scale=b - a)

    def __call__(self, param):
        p = np.asarray(param)
        if np.any(p < 0) or np.any(p > self.b) or np.any(p < self.a):
            return -np.inf
        return -1 * np.log(self.b - self.a)
