"""Developer probe: duration of the scoring kernel (k_score) at a named config under the current tuning environment
(BODE_WAVES, BODE_CHUNK_KB, BODE_STAGES, BODE_SCORE_VARIANT ...).  Prints the median of the kernel's own CUDA events."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from bode_b200.engine import EkldEngine  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c3"
    nd = int(sys.argv[2]) if len(sys.argv) > 2 else None
    w = bench.CONFIGS[name]
    X, Y, hyp, Xd = bench.make_workload(w["d"], w["n"], w["S"], nd or w["Nd"], w["seed"])
    eng = EkldEngine()
    eng.prepare(X, Y, hyp, w["nugget"] ** 2)
    xd = torch.from_numpy(Xd).to(eng.device)
    ks = []
    for i in range(8):
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record(); k1.record()  # (creates the CUDA events: torch makes them lazily)
        r = eng.score(xd, want_var=False, events=(k0, k1))
        torch.cuda.synchronize()
        ks.append(k0.elapsed_time(k1))
    env = {k: v for k, v in os.environ.items() if k.startswith("BODE_")}
    print("%s Nd=%d %s: k_score %.4f ms (min %.4f)  best_idx %d  best %.17g" % (
        name, Xd.shape[0], env, float(np.median(ks[3:])), min(ks[3:]), int(r["best_idx"].item()), float(r["best"].item())))


if __name__ == "__main__":
    main()
