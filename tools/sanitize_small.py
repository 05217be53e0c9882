"""Small end-to-end invocation of every kernel for compute-sanitizer (memcheck / racecheck / synccheck), one tool per run:
    compute-sanitizer --tool racecheck python tools/sanitize_small.py
Closed-form and quadrature mode, tile and global-memory prepare kernels, scoring, EKLD/reduction, comparator, likelihoods."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bode_b200 import _ext  # noqa: E402
from bode_b200.engine import EkldEngine  # noqa: E402
from tests import cases  # noqa: E402

eng = EkldEngine()
for (d, n, S, Nd) in ((3, 20, 3, 130), (6, 60, 4, 200), (5, 230, 2, 70)):   # n = 230: global-memory prepare kernel
    X, Y, Xd, hyp = cases.make_problem(d, n, S, Nd, seed=n, ell_lo=0.3, ell_hi=0.9)
    eng.prepare(X, Y, hyp, 1e-6)
    r = eng.score(Xd)
    print(d, n, "closed", int(eng.info.abs().max()), float(r["best"].item()), int(r["best_idx"].item()))
    eng.us_variance(Xd[:50])
    eng.loglik(X, Y, hyp, 1e-6)
X, Y, Xd, hyp = cases.make_problem(3, 24, 3, 100, seed=61, ell_lo=0.1, ell_hi=0.3)
Xq = np.random.default_rng(8).random((150, 3))
eng.prepare(X, Y, hyp, 1e-6, quad=(Xq, np.random.default_rng(9).random(150) + 0.5))
r = eng.score(Xd)
print("quad", float(r["best"].item()), eng.mu_sigma())
os.environ["BODE_ROWMEAN_VARIANT"] = "0"
_ext.reload_tuning()
eng.prepare(X, Y, hyp, 1e-6, quad=(Xq, None))
print("quad (thread-per-point)", float(eng.score(Xd)["best"].item()))
print("done")
