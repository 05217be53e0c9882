"""Developer probe: time of the K4 + reduction kernel (k_ekld_reduce) at config 3 under different options.
Whole-call time between CUDA events minus the scoring kernel's own events."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from bode_b200.engine import EkldEngine  # noqa: E402


def main():
    w = bench.CONFIGS["c3"]
    X, Y, hyp, Xd = bench.make_workload(w["d"], w["n"], w["S"], w["Nd"], w["seed"])
    eng = EkldEngine()
    eng.prepare(X, Y, hyp, w["nugget"] ** 2)
    xd = torch.from_numpy(Xd).to(eng.device)
    for want_var in (True, False):
        for form in (0, 1):
            ts = []
            for i in range(6):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                k0.record(); k1.record()
                torch.cuda.synchronize()
                e0.record()
                eng.score(xd, want_var=want_var, form=form, events=(k0, k1))
                e1.record()
                torch.cuda.synchronize()
                ts.append((e0.elapsed_time(e1), k0.elapsed_time(k1)))
            tot = np.median([t[0] for t in ts[2:]]); ks = np.median([t[1] for t in ts[2:]])
            print("want_var=%s form=%d: call %.3f ms, k_score %.3f ms, rest %.3f ms" % (want_var, form, tot, ks, tot - ks))


if __name__ == "__main__":
    main()
