#!/usr/bin/env python
"""6-D sequential design loop on the B200 path (BASELINE config 5).

Python-3 port of the reference driver ``tests/samp_ex4.py:54-117`` (Python 2; it imports a non-existent ``sampler`` module
and runs its 5-D ``Ex2Func``): same priors, nugget, chains (14 x 1000, burn 200, thin 20, 56 hyper-samples), 10000 designs
per iteration and output files, on the driver's 6-D ``Ex1Func`` (``:28-39``), 20 initial observations, max_it = 50.  Every
iteration refits the hyper-samples on the host (batched device likelihoods), re-factorises, scores all designs and adds
the argmax.

    python examples/samp_ex4.py [--quick] [--iters N] [--out DIR]
"""
import argparse
import os
import pickle
import shutil
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from bode import *  # noqa: E402,F401,F403
from bode_b200.compat import GPy, lhs  # noqa: E402


class Ex1Func(object):
    """Reference tests/samp_ex4.py:28-39 (from Knowles 2006, ref. 16 of the BODE paper)."""

    def __call__(self, x):
        g = 100. * (((x[1:6] - 0.5) ** 2 - np.cos(2. * np.pi * (x[1:6] - 0.5))).sum() + 5.)
        k = 0.5 * (1 - x[0]) * (g + 1.)
        return (k - 150.) / 80.


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true", help="short MCMC chains, 3 iterations (smoke run)")
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    np.random.seed(1223)
    n, dim = 20, 6
    objective = Ex1Func()
    X_true = lhs(dim, 100000)
    true_mean = np.mean([objective(x) for x in X_true])
    print('true E[f(x)]: ', true_mean)
    X_init = lhs(dim, n)
    Y_init = np.array([objective(x) for x in X_init])[:, None]
    num_quad_points = 500
    quad_points = lhs(dim, num_quad_points)
    num_it = 3 if args.quick else args.iters
    out_dir = args.out or 'mcmc_ex4_n={0:d}_it={1:d}'.format(n, num_it)
    if os.path.isdir(out_dir):
        shutil.rmtree(out_dir)
    os.makedirs(out_dir)
    steps, burn, thin = (1000, 200, 20) if not args.quick else (280, 80, 20)
    kls = KLSampler(X_init, Y_init, np.array([[0.6]]),
                    model_kern=GPy.kern.RBF,
                    bounds=[(0, 1)] * X_init.shape[1],
                    obj_func=objective, true_func=objective, noisy=False,
                    nugget=1e-3, lengthscale=0.3, variance=1., kld_tol=1e-5,
                    func_name=os.path.join(out_dir, 'ex4'), energy=0.95,
                    num_quad_points=num_quad_points, quad_points=quad_points, quad_points_weight=np.ones(num_quad_points),
                    max_it=num_it, per_sampled=20, num_opt_restarts=100,
                    mcmc_model=True, mcmc_chains=14, mcmc_model_avg=56, mcmc_steps=steps, mcmc_burn=burn, mcmc_thin=thin,
                    mcmc_parallel=8, ego_iter=50,
                    variance_prior=GammaPrior(a=1, scale=2), lengthscale_prior=ExponentialPrior(scale=0.7),
                    noise_prior=JeffreysPrior())
    X, Y, X_u, kld, X_design, mu_qoi, sigma_qoi, comp_log = kls.optimize(num_designs=10000, verbose=0, plots=0, comp=True)
    np.save(os.path.join(out_dir, 'X.npy'), X)
    np.save(os.path.join(out_dir, 'Y.npy'), Y)
    np.save(os.path.join(out_dir, 'X_u.npy'), X_u)
    np.save(os.path.join(out_dir, 'kld.npy'), kld)
    np.save(os.path.join(out_dir, 'mu_qoi.npy'), mu_qoi)
    np.save(os.path.join(out_dir, 'sigma_qoi.npy'), sigma_qoi)
    with open(os.path.join(out_dir, "comp.pkl"), "wb") as f:
        pickle.dump(comp_log, f)
    print('Q mean per iteration:', mu_qoi, ' true mean', true_mean)


if __name__ == '__main__':
    main()
