"""Developer smoke: CUDA path vs oracle on a few small shapes (run on the GPU box)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bode_b200 import _ext
from oracle import ekld_oracle as o

def case(d, n, S, Nd, seed, ellmin=0.05):
    rng = np.random.default_rng(seed)
    X = rng.random((n, d)); Y = np.sin(3 * X.sum(1))[:, None]
    Xd = rng.random((Nd, d))
    hyp = np.column_stack([rng.gamma(1.0, 2.0, S) + 0.1] + [np.maximum(rng.exponential(0.7, S), ellmin) for _ in range(d)])
    t0 = time.time()
    r = _ext.ekld_score_host(X, Y, hyp, Xd, 1e-6, want_kld=True)
    t1 = time.time()
    ref = o.score(X, Y, Xd, hyp)
    rel = np.abs(r['kld'] - ref['kld']) / np.abs(ref['kld'])
    print(f"d={d} n={n} S={S} Nd={Nd}: gpu {t1-t0:.3f}s info={r['info'].max()} kld rel med {np.median(rel):.2e} max {np.nanmax(rel):.2e} "
          f"score maxabs {np.nanmax(np.abs(r['score']-ref['score'])):.2e} argmax {r['argmax']} vs {ref['argmax']} "
          f"muQ {abs(r['mu_Q']-ref['mu_Q']):.2e} sigQ {abs(r['sigma_Q']-ref['sigma_Q'])/abs(ref['sigma_Q']):.2e} nan {np.isnan(r['score']).sum()}")

if __name__ == "__main__":
    print(_ext.load().bode_version())
    case(1, 3, 4, 50, 0)
    case(2, 30, 8, 300, 1)
    case(3, 9, 5, 77, 2)
    case(8, 200, 8, 1000, 3)
    case(6, 60, 16, 500, 4)
    case(10, 500, 4, 200, 5)
    case(5, 100, 64, 3000, 6)
    print(_ext.fp64_peak())
