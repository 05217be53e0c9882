"""GPy stand-in -- TEST INFRASTRUCTURE.  The slice of GPy 1.9.2 that the reference's EKLD path touches
(``bode/_sampler.py:391-492``): ``GPy.kern.RBF(input_dim, ARD=True)``, ``GPy.models.GPRegression(X, Y, kern)``,
``model.kern.{variance,lengthscale}`` / ``model.likelihood.variance`` as fixable parameter arrays,
``model.kern.K``, ``model.predict``, ``model.posterior.woodbury_inv`` / ``woodbury_vector``, ``model.log_likelihood``.

Arithmetic follows GPy's published implementation, function by function (names of the GPy routines in brackets):
  [Stationary._unscaled_dist]   r2 = -2 X X2^T + |X|^2 + |X2|^2, clipped at 0 (diagonal forced to 0 if X2 is None)
  [RBF.K_of_r]                  variance * exp(-0.5 r^2)
  [ExactGaussianInference]      Ky = K + (likelihood variance + 1e-8) I ; LW = jitchol(Ky) ; alpha = dpotrs(LW, Y)
  [Posterior.woodbury_inv]      dpotri(LW), lower triangle mirrored to the upper
  [PosteriorExact._raw_predict] mu = Kx^T alpha ; full_cov: Kxx - tmp^T tmp ; else Kdiag - sum(tmp^2), tmp = dtrtrs(LW, Kx)
  [Gaussian.predictive_values]  include_likelihood adds the likelihood variance (not the jitter)
"""
import numpy as np
from scipy.linalg import lapack

__version__ = "1.9.2-stub"


class Param(np.ndarray):
    """A parameter array that can be fixed to a value (paramz ``Param.fix`` / ``constrain_fixed``)."""

    def __new__(cls, value, owner=None):
        obj = np.array(value, dtype=np.float64).reshape(-1).view(cls)
        obj._owner = owner
        return obj

    def __array_finalize__(self, obj):
        self._owner = getattr(obj, "_owner", None)

    def _set(self, value):
        if value is not None:
            self[...] = np.asarray(value, dtype=np.float64).reshape(-1)
        if self._owner is not None:
            self._owner._changed()

    def fix(self, value=None, warning=True):
        self._set(value)

    constrain_fixed = fix

    def __setitem__(self, k, v):
        np.ndarray.__setitem__(self, k, v)
        if getattr(self, "_owner", None) is not None:
            self._owner._changed()


class _RBF(object):
    name = "rbf"

    def __init__(self, input_dim, variance=1.0, lengthscale=None, ARD=False, **kw):
        self.input_dim = int(input_dim)
        self.ARD = bool(ARD)
        self._model = None
        self.variance = Param([variance], self)
        nl = self.input_dim if ARD else 1
        self.lengthscale = Param(np.ones(nl) if lengthscale is None else lengthscale, self)

    def _changed(self):
        if self._model is not None:
            self._model._dirty = True

    def _dist(self, X, X2):
        ell = np.asarray(self.lengthscale)
        Xs = X / ell
        if X2 is None:
            Xsq = np.sum(np.square(Xs), 1)
            r2 = -2.0 * Xs.dot(Xs.T) + (Xsq[:, None] + Xsq[None, :])
            r2[np.diag_indices_from(r2)] = 0.0
        else:
            X2s = X2 / ell
            X1sq = np.sum(np.square(Xs), 1)
            X2sq = np.sum(np.square(X2s), 1)
            r2 = -2.0 * np.dot(Xs, X2s.T) + (X1sq[:, None] + X2sq[None, :])
        return np.sqrt(np.clip(r2, 0, np.inf))

    def K(self, X, X2=None):
        r = self._dist(np.asarray(X, dtype=np.float64), None if X2 is None else np.asarray(X2, dtype=np.float64))
        return float(self.variance[0]) * np.exp(-0.5 * r ** 2)

    def Kdiag(self, X):
        return np.full(X.shape[0], float(self.variance[0]))


class _Matern32(_RBF):
    name = "Mat32"

    def K(self, X, X2=None):
        r = self._dist(X, X2)
        return float(self.variance[0]) * (1.0 + np.sqrt(3.0) * r) * np.exp(-np.sqrt(3.0) * r)


class _Gaussian(object):
    def __init__(self, model, variance=1.0):
        self._model = model
        self.variance = Param([variance], self)

    def _changed(self):
        self._model._dirty = True


def _jitchol(A, maxtries=5):
    L, info = lapack.dpotrf(np.asfortranarray(A), lower=1)
    if info == 0:
        return L
    diagA = np.diag(A)
    if np.any(diagA <= 0.0):
        raise np.linalg.LinAlgError("not pd: non-positive diagonal elements")
    jitter = diagA.mean() * 1e-6
    for _ in range(maxtries):
        L, info = lapack.dpotrf(np.asfortranarray(A + np.eye(A.shape[0]) * jitter), lower=1)
        if info == 0:
            return L
        jitter *= 10
    raise np.linalg.LinAlgError("not positive definite, even with jitter.")


class _Posterior(object):
    def __init__(self, LW, alpha):
        self.woodbury_chol = LW
        self.woodbury_vector = alpha
        self._winv = None

    @property
    def woodbury_inv(self):
        if self._winv is None:
            Wi, _ = lapack.dpotri(np.asfortranarray(self.woodbury_chol), lower=1)
            Wi = np.tril(Wi)
            self._winv = Wi + np.tril(Wi, -1).T
        return self._winv


class _GPRegression(object):
    def __init__(self, X, Y, kernel=None, Y_metadata=None, normalizer=None, noise_var=1.0, mean_function=None):
        self.X = np.asarray(X, dtype=np.float64)
        self.Y = np.asarray(Y, dtype=np.float64)
        assert self.X.ndim == 2 and self.Y.ndim == 2
        self.kern = kernel if kernel is not None else _RBF(self.X.shape[1])
        self.kern._model = self
        self.likelihood = _Gaussian(self, noise_var)
        self._dirty = True
        self._post = None
        self._ll = None

    def _infer(self):
        if not self._dirty:
            return
        K = self.kern.K(self.X)
        Ky = K.copy()
        Ky[np.diag_indices_from(Ky)] += float(self.likelihood.variance[0]) + 1e-8
        LW = _jitchol(Ky)
        alpha, _ = lapack.dpotrs(LW, self.Y, lower=1)
        n, D = self.Y.shape
        logdet = 2.0 * np.sum(np.log(np.diag(LW)))
        self._ll = 0.5 * (-self.Y.size * np.log(2 * np.pi) - D * logdet - np.sum(alpha * self.Y))
        self._post = _Posterior(np.tril(LW), alpha)
        self._dirty = False

    @property
    def posterior(self):
        self._infer()
        return self._post

    def log_likelihood(self):
        self._infer()
        return float(self._ll)

    def predict(self, Xnew, full_cov=False, Y_metadata=None, kern=None, likelihood=None, include_likelihood=True):
        self._infer()
        Xnew = np.asarray(Xnew, dtype=np.float64)
        Kx = self.kern.K(self.X, Xnew)
        mu = np.dot(Kx.T, self._post.woodbury_vector)
        tmp, _ = lapack.dtrtrs(self._post.woodbury_chol, Kx, lower=1, trans=0, unitdiag=0)
        nv = float(self.likelihood.variance[0])
        if full_cov:
            var = self.kern.K(Xnew) - np.dot(tmp.T, tmp)
            if include_likelihood:
                var = var + np.eye(var.shape[0]) * nv
        else:
            var = (self.kern.Kdiag(Xnew) - np.square(tmp).sum(0))[:, None]
            if include_likelihood:
                var = var + nv
        return mu, var

    def optimize_restarts(self, *a, **k):
        raise NotImplementedError("GPy stub: MLE fitting is outside the EKLD hot path")

    optimize = optimize_restarts


class kern(object):
    RBF = _RBF
    Matern32 = _Matern32


class models(object):
    GPRegression = _GPRegression
