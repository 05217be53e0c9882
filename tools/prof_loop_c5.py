import os, sys, io, contextlib, cProfile, pstats, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench, bode_b200
from bode_b200.engine import EkldEngine
cfg = bench.CONFIGS["c5"]
X0, Y0, Xd, f = bench.c5_problem(cfg)
S, d, Nd, iters = cfg["S"], cfg["d"], cfg["Nd"], cfg["iters"]
hyper = bench.hyper_c5(S, d)
eng = EkldEngine()
def one_loop():
    with contextlib.redirect_stdout(io.StringIO()):
        kls = bode_b200.KLSampler(X0, Y0, False, f, False, [(0, 1)] * d, mcmc_model=True, max_it=iters, kld_tol=0.0,
                                  mcmc_chains=2, mcmc_model_avg=S, mcmc_steps=100, mcmc_burn=10, mcmc_thin=5, X_design=Xd,
                                  engine=eng, hyper_samples=hyper, ekld_form=cfg["form"], process_group=None)
        out = kls.optimize(num_designs=Nd)
for _ in range(2): one_loop()
torch.cuda.synchronize(); t0 = time.perf_counter(); one_loop(); torch.cuda.synchronize(); print("loop %.2f ms" % ((time.perf_counter() - t0) * 1e3))
pr = cProfile.Profile(); pr.enable(); one_loop(); torch.cuda.synchronize(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28); print(s.getvalue()[:6000])
