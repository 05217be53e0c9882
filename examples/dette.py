#!/usr/bin/env python
"""8-D Dette-Pepelyshev problem on the B200 path (BASELINE config 3): 200 observations, |X_design| = 1e5, 64 hyper-samples.

Python-3 port of the reference driver ``tests/dette.py:98-160`` (which is Python 2 and truncates the function to its first
three terms, ``:82-94``): same priors, nugget, MCMC thinning and output files; the function is the full 8-D curved function
of Dette & Pepelyshev (2010) with the driver's ``(y - 50) / 50`` scaling, and the sizes are those BASELINE.json names for
this configuration.  The acquisition is scored densely over all 1e5 designs for all 64 hyper-samples on the GPU.

    python examples/dette.py [--quick] [--iters N] [--out DIR]
    torchrun --nproc-per-node 8 examples/dette.py ...      # candidate blocks sharded over the GPUs of one box
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


class DetteFunc(object):
    """Dette & Pepelyshev (2010), "Generalized Latin hypercube design for computer experiments", 8-D curved function,
    scaled like reference tests/dette.py:93-94."""

    def __call__(self, x):
        y = 4 * ((x[0] - 2 + 8 * x[1] - 8 * (x[1] ** 2)) ** 2) + (3 - 4 * x[1]) ** 2 + 16 * np.sqrt(x[2] + 1) * ((2 * x[2] - 1) ** 2)
        for i in range(4, 9):
            y = y + i * np.log(1 + np.sum(x[2:i]))
        return (y - 50) / 50.


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true", help="short MCMC chains, 2 iterations (smoke run)")
    ap.add_argument("--iters", type=int, default=30)          # reference: num_it = 30
    ap.add_argument("--designs", type=int, default=100000)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    group = None
    if "RANK" in os.environ and int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch
        import torch.distributed as tdist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        tdist.init_process_group("nccl")
        group = tdist.group.WORLD
    np.random.seed(1223)
    n, dim = 200, 8
    objective = DetteFunc()
    X_true = lhs(dim, 100000)
    true_mean = np.mean([objective(x) for x in X_true])
    print('true E[f(x)]: ', true_mean)
    X_init = lhs(dim, n)
    Y_init = np.array([objective(x) for x in X_init])[:, None]
    num_quad_points = 500
    quad_points = lhs(dim, num_quad_points)
    num_it = 2 if args.quick else args.iters
    out_dir = args.out or 'mcmc_dette_n={0:d}_it={1:d}'.format(n, num_it)
    if os.path.isdir(out_dir):
        shutil.rmtree(out_dir)
    os.makedirs(out_dir, exist_ok=True)
    steps, burn, thin = (5000, 1000, 20) if not args.quick else (300, 100, 20)
    kls = KLSampler(X_init, Y_init, np.array([[.6]]),
                    model_kern=GPy.kern.RBF,
                    bounds=[(0, 1)] * X_init.shape[1],
                    obj_func=objective, true_func=objective, noisy=False,
                    nugget=1e-3, lengthscale=0.3, variance=1., kld_tol=1e-6,
                    func_name=os.path.join(out_dir, 'dette'), energy=0.95,
                    num_quad_points=num_quad_points, quad_points=quad_points, quad_points_weight=np.ones(num_quad_points),
                    max_it=num_it, per_sampled=5, num_opt_restarts=100,
                    mcmc_model=True, mcmc_chains=16, mcmc_steps=steps, mcmc_burn=burn, mcmc_thin=thin, mcmc_model_avg=64,
                    ego_iter=40, mcmc_parallel=8,
                    variance_prior=GammaPrior(a=1, scale=2), lengthscale_prior=ExponentialPrior(scale=0.7),
                    noise_prior=JeffreysPrior(), process_group=group)
    X, Y, X_u, kld, X_design, mu_qoi, sigma_qoi, comp_log = kls.optimize(num_designs=args.designs, verbose=0, plots=0, comp=True)
    if group is None or int(os.environ.get("RANK", "0")) == 0:
        # on-disk artefacts of the reference driver (tests/dette.py:155-162)
        np.save(os.path.join(out_dir, 'X.npy'), X)
        np.save(os.path.join(out_dir, 'Y.npy'), Y)
        np.save(os.path.join(out_dir, 'X_u.npy'), X_u)
        np.save(os.path.join(out_dir, 'kld.npy'), kld)
        np.save(os.path.join(out_dir, 'mu_qoi.npy'), mu_qoi)
        np.save(os.path.join(out_dir, 'sigma_qoi.npy'), sigma_qoi)
        with open(os.path.join(out_dir, "comp.pkl"), "wb") as f:
            pickle.dump(comp_log, f)
        print('Q mean per iteration:', mu_qoi, ' true mean', true_mean)
    if group is not None:
        import torch.distributed as tdist
        tdist.destroy_process_group()


if __name__ == '__main__':
    main()
