#!/usr/bin/env python
"""1-D toy problem of the reference driver ``tests/samp_ex1.py`` on the B200 path (BASELINE config 1).

Same problem, priors, MCMC settings, seed and output files as the reference driver; the acquisition is scored densely
over the 1000 centred-LHS designs on the GPU instead of through the inner EGO loop.

    python examples/samp_ex1.py [--quick] [--out DIR]
"""
import argparse
import os
import pickle
import shutil
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from bode import *  # noqa: E402,F401,F403  (KLSampler, priors, xik, bk, ei)
from bode_b200.compat import GPy, lhs  # noqa: E402


class Ex1Func(object):
    """f(x) = (4 (1 - sin(6x + 8 exp(6x - 7))) - 10) / 5 + sigma(x) eps   (reference tests/samp_ex1.py:28-35)"""

    def __init__(self, sigma=lambda x: 0.5):
        self.sigma = sigma

    def __call__(self, x):
        x = 6 * x
        return (4 * (1. - np.sin(x[0] + 8 * np.exp(x[0] - 7.))) - 10.) / 5. + self.sigma(x[0]) * np.random.randn()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true", help="short MCMC chains (smoke run)")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    np.random.seed(1333)
    n, dim, num_it = 3, 1, 1
    objective = Ex1Func(sigma=lambda x: 0)
    X_init = lhs(dim, samples=n, criterion='center')
    Y_init = np.array([objective(x) for x in X_init])[:, None]
    X_true = lhs(dim, samples=100000, criterion='center')
    true_mean = np.mean([objective(x) for x in X_true])
    print('true E[f(x)]: ', true_mean)
    num_quad_points = 500
    quad_points = lhs(dim, num_quad_points)
    out_dir = args.out or 'mcmc_ex1_n={0:d}_it={1:d}'.format(n, num_it)
    if os.path.isdir(out_dir):
        shutil.rmtree(out_dir)
    os.makedirs(out_dir)
    steps, burn, thin = (1000, 200, 20) if not args.quick else (260, 60, 20)
    kls = KLSampler(X_init, Y_init, x_hyp=False,
                    model_kern=GPy.kern.RBF,
                    bounds=[(0, 1)] * X_init.shape[1],
                    obj_func=objective, true_func=objective, noisy=False,
                    nugget=1e-3, lengthscale=0.2, variance=1., kld_tol=1e-6,
                    func_name=os.path.join(out_dir, 'ex1'), energy=0.95,
                    num_quad_points=num_quad_points, quad_points=quad_points, quad_points_weight=np.ones(num_quad_points),
                    max_it=num_it, per_sampled=20,
                    mcmc_model=True, mcmc_chains=6, mcmc_model_avg=60, mcmc_steps=steps, mcmc_burn=burn, mcmc_thin=thin,
                    mcmc_parallel=8, ego_iter=20, initialize_from_prior=True,
                    variance_prior=GammaPrior(a=1, scale=2), lengthscale_prior=ExponentialPrior(scale=0.7),
                    noise_prior=JeffreysPrior())
    X, Y, X_u, kld, X_design, mu_qoi, sigma_qoi, comp_log = kls.optimize(num_designs=1000, verbose=1, plots=1, comp=True)
    # on-disk artefacts of the reference driver (tests/samp_ex1.py:110-117)
    np.save(os.path.join(out_dir, 'X.npy'), X)
    np.save(os.path.join(out_dir, 'Y.npy'), Y)
    np.save(os.path.join(out_dir, 'X_u.npy'), X_u)
    np.save(os.path.join(out_dir, 'kld.npy'), kld)
    np.save(os.path.join(out_dir, 'mu_qoi.npy'), mu_qoi)
    np.save(os.path.join(out_dir, 'sigma_qoi.npy'), sigma_qoi)
    with open(os.path.join(out_dir, "comp.pkl"), "wb") as f:
        pickle.dump(comp_log, f)
    print('chosen design:', X[-1], ' Q mean/var before and after:', mu_qoi, sigma_qoi, ' true mean', true_mean)


if __name__ == '__main__':
    main()
