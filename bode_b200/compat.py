"""Name-compatibility shims for the reference drivers' third-party imports that are absent in this image.

``GPy.kern.RBF`` is only ever passed as ``model_kern=`` (reference ``tests/samp_ex1.py:81``) -- a sentinel is enough
because the B200 path implements exactly the ARD-RBF kernel.  ``lhs`` follows pyDOE (see ``_lhs.py``).
"""
from ._lhs import lhs  # noqa: F401


class _Kern(object):
    class RBF(object):
        name = "rbf"

        def __init__(self, input_dim=1, ARD=True, **kw):
            self.input_dim, self.ARD = input_dim, ARD

    class Matern32(RBF):
        name = "Mat32 (unsupported: xik/bk are RBF-only, reference _core.py:86-101)"


class GPy(object):
    """``from bode_b200.compat import GPy`` then ``GPy.kern.RBF`` as in the reference drivers."""
    kern = _Kern
