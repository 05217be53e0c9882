#!/usr/bin/env python
"""Bimodal 1-D problem of the reference driver ``tests/samp_ex2.py`` (15 sequential designs) on the B200 path.

    python examples/samp_ex2.py [--quick] [--out DIR]
"""
import argparse
import os
import pickle
import shutil
import sys

import numpy as np
from scipy.stats import norm

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from bode import *  # noqa: E402,F401,F403
from bode_b200.compat import GPy, lhs  # noqa: E402


class Ex2Func(object):
    """Sum of two Gaussian bumps (reference tests/samp_ex2.py:28-37); parameters are attributes, not module globals."""

    def __init__(self, mu1=0.2, sigma1=0.05, mu2=0.8, sigma2=0.05):
        self.mu1, self.sigma1, self.mu2, self.sigma2 = mu1, sigma1, mu2, sigma2

    def __call__(self, x):
        return norm.pdf(x[0], loc=self.mu1, scale=self.sigma1) + norm.pdf(x[0], loc=self.mu2, scale=self.sigma2)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    np.random.seed(1263)
    n, dim = 3, 1
    num_it = 15 if not args.quick else 3
    objective = Ex2Func()
    X_true = lhs(dim, samples=100000, criterion='center')
    print('true E[f(x)]: ', np.mean([objective(x) for x in X_true]))
    X_init = lhs(dim, samples=n, criterion='center')
    Y_init = np.array([objective(x) for x in X_init])[:, None]
    out_dir = args.out or 'mcmc_ex2_n={0:d}_it={1:d}'.format(n, num_it)
    if os.path.isdir(out_dir):
        shutil.rmtree(out_dir)
    os.makedirs(out_dir)
    num_quad_points = 500
    quad_points = lhs(dim, num_quad_points)
    steps = 5000 if not args.quick else 400
    kls = KLSampler(X_init, Y_init, np.array([[.6]]),
                    model_kern=GPy.kern.RBF, bounds=[(0, 1)] * dim,
                    obj_func=objective, true_func=objective, noisy=False, energy=0.95,
                    nugget=1e-3, lengthscale=.2, variance=1., kld_tol=1e-3,
                    func_name=os.path.join(out_dir, 'ex2'),
                    num_quad_points=num_quad_points, quad_points=quad_points, quad_points_weight=np.ones(num_quad_points),
                    max_it=num_it, per_sampled=30,
                    mcmc_model=True, mcmc_chains=8, mcmc_steps=steps, mcmc_burn=100, mcmc_model_avg=80, mcmc_thin=20,
                    ego_iter=20, mcmc_parallel=8, initialize_from_prior=True,
                    variance_prior=GammaPrior(a=1, scale=2), lengthscale_prior=ExponentialPrior(scale=0.7),
                    noise_prior=JeffreysPrior())
    X, Y, X_u, kld, X_design, mu_qoi, sigma_qoi, comp_log = kls.optimize(num_designs=1000, verbose=1, plots=0, comp=True)
    for name, arr in (('X', X), ('Y', Y), ('X_u', X_u), ('kld', kld), ('mu_qoi', mu_qoi), ('sigma_qoi', sigma_qoi)):
        np.save(os.path.join(out_dir, name + '.npy'), arr)
    with open(os.path.join(out_dir, "comp.pkl"), "wb") as f:
        pickle.dump(comp_log, f)
    print('designs added:', X[n:, 0], '\nQ mean per iteration:', np.round(mu_qoi, 4))


if __name__ == '__main__':
    main()
