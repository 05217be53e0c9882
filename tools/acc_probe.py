"""Developer probe: accuracy of the two prepare kernels on the ill-conditioned prior-draw problems (error vs the 80-bit truth)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bode_b200 import _ext  # noqa: E402
from oracle import ekld_oracle as orc, truth  # noqa: E402
from tests import cases  # noqa: E402
from tests.tolerance import per_sample_err  # noqa: E402

for (d, n, S, Nd, seed) in [(2, 30, 32, 1000, 1263), (1, 8, 24, 300, 4), (6, 60, 24, 300, 1), (8, 200, 16, 300, 1223)]:
    X, Y, Xd, hyp = cases.prior_problem(d, n, S, Nd, seed)
    t = truth.ekld_truth(X, Y, Xd, hyp)
    ref = orc.score(X, Y, Xd, hyp)
    eo = per_sample_err(ref["kld"], t["kld"])
    for variant in ("-1", "0"):
        os.environ["BODE_PREP_VARIANT"] = variant
        _ext.reload_tuning()
        r = _ext.ekld_score_host(X, Y, hyp, Xd, 1e-6, want_kld=True, raise_not_pd=False)
        eg = per_sample_err(r["kld"], t["kld"])
        print("d=%d n=%d variant %s: gpu max %.3e median %.3e | oracle max %.3e median %.3e | worst ratio gpu/oracle %.3f" % (
            d, n, variant, eg.max(), np.median(eg), eo.max(), np.median(eo), (eg / np.maximum(eo, 1e-16)).max()))
