"""Developer probe: time prepare / score at a named shape on one GPU (CUDA events)."""
import sys, os, time, argparse
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bode_b200.engine import EkldEngine

ap = argparse.ArgumentParser()
ap.add_argument("--d", type=int, default=8); ap.add_argument("--n", type=int, default=200)
ap.add_argument("--S", type=int, default=64); ap.add_argument("--Nd", type=int, default=100000)
ap.add_argument("--reps", type=int, default=5)
a = ap.parse_args()
rng = np.random.default_rng(1223)
X = rng.random((a.n, a.d)); Y = np.sin(3 * X.sum(1))[:, None]
hyp = np.column_stack([rng.gamma(1.0, 2.0, a.S) + 0.1] + [np.maximum(rng.exponential(0.7, a.S), 0.05) for _ in range(a.d)])
Xd = torch.from_numpy(rng.random((a.Nd, a.d))).cuda()
eng = EkldEngine()
def ev(): return torch.cuda.Event(enable_timing=True)
for rep in range(a.reps):
    e0, e1, e2 = ev(), ev(), ev()
    e0.record(); eng.prepare(X, Y, hyp, 1e-6); e1.record(); r = eng.score(Xd, timed=True); e2.record()
    torch.cuda.synchronize()
    F = a.n**2 + a.n*(3*a.d+5) + 8*a.d + 20
    evals = a.Nd * a.S
    print(f"rep {rep}: prepare {e0.elapsed_time(e1):.3f} ms  score(all) {e1.elapsed_time(e2):.3f} ms  k_score {eng.last_kernel_ms:.3f} ms  "
          f"-> {evals/ (e0.elapsed_time(e2)*1e-3):.3e} evals/s ; k_score {evals*F/(eng.last_kernel_ms*1e-3)*1e-12:.2f} TFLOP/s (F={F})  info {int(eng.info.max())} nan {int(torch.isnan(r['score']).sum())}")
