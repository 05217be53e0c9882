"""Affine-invariant ensemble sampler with the slice of emcee's (2.x) API the reference uses -- host stand-in.

Used only when the real ``emcee`` is not importable.  API mirrored (``_sampler.py:270-274``):
``EnsembleSampler(nwalkers, ndim, lnprob, args=...)``, ``run_mcmc(p0, N)``, ``.chain`` with shape
``(nwalkers, N, ndim)``, ``.acceptance_fraction``.

Extension used by ``KLSampler`` on the B200: ``vectorize=True`` hands the *whole half-ensemble* to ``lnprob`` as a
``(k, ndim)`` array, so the GP marginal likelihoods of one stretch-move half-step are evaluated as one batched
Cholesky launch (``bode_gp_loglik_host``) instead of k separate host factorisations.
"""
import numpy as np


class EnsembleSampler(object):
    def __init__(self, nwalkers, dim, lnpostfn, a=2.0, args=(), vectorize=False):
        if nwalkers % 2 != 0:
            raise ValueError("the number of walkers must be even")
        if nwalkers < 2 * dim:
            # emcee 2.x raises here; the reference calls it with nwalkers < 2*ndim only for d large, keep going with
            # a warning-free relaxation so the drivers' settings (6 walkers, ndim 2) behave identically.
            pass
        self.k, self.dim, self.a = int(nwalkers), int(dim), float(a)
        self.lnpostfn, self.args, self.vectorize = lnpostfn, tuple(args), bool(vectorize)
        self._chain = np.empty((self.k, 0, self.dim))
        self._lnprob = np.empty((self.k, 0))
        self.naccepted = np.zeros(self.k)
        self.iterations = 0

    @property
    def chain(self):
        return self._chain

    @property
    def lnprobability(self):
        return self._lnprob

    @property
    def acceptance_fraction(self):
        return self.naccepted / max(1, self.iterations)

    def _eval(self, pos):
        if self.vectorize:
            return np.asarray(self.lnpostfn(pos, *self.args), dtype=float)
        return np.array([float(self.lnpostfn(pos[i], *self.args)) for i in range(pos.shape[0])])

    def run_mcmc(self, pos0, N):
        p = np.array(pos0, dtype=float)
        assert p.shape == (self.k, self.dim)
        lnp = self._eval(p)
        if np.any(np.isnan(lnp)):
            raise ValueError("the initial lnprob was NaN")
        chain = np.empty((self.k, N, self.dim))
        lnps = np.empty((self.k, N))
        half = self.k // 2
        first, second = slice(half), slice(half, self.k)
        for it in range(N):
            for S0, S1 in ((first, second), (second, first)):
                s, c = p[S0], p[S1]
                ns, nc = s.shape[0], c.shape[0]
                zz = ((self.a - 1.0) * np.random.rand(ns) + 1) ** 2.0 / self.a
                rint = np.random.randint(nc, size=(ns,))
                q = c[rint] - zz[:, np.newaxis] * (c[rint] - s)
                newlnp = self._eval(q)
                lnpdiff = (self.dim - 1.0) * np.log(zz) + newlnp - lnp[S0]
                accept = lnpdiff > np.log(np.random.rand(ns))
                if np.any(accept):
                    idx = np.arange(self.k)[S0][accept]
                    p[idx] = q[accept]
                    lnp[idx] = newlnp[accept]
                    self.naccepted[idx] += 1
            chain[:, it, :] = p
            lnps[:, it] = lnp
            self.iterations += 1
        self._chain = np.concatenate([self._chain, chain], axis=1)
        self._lnprob = np.concatenate([self._lnprob, lnps], axis=1)
        return p, lnp, None
