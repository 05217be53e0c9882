"""Developer probe: replay time of the captured step (bode_b200.engine.StepGraph) at config 3 with the design block on the
device and in pinned host memory (the latter adds the host->device copy, on a forked branch of the graph)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from bode_b200.engine import EkldEngine, StepGraph  # noqa: E402


def main():
    nd = int(sys.argv[1]) if len(sys.argv) > 1 else None
    w = bench.CONFIGS["c3"]
    X, Y, hyp, Xd = bench.make_workload(w["d"], w["n"], w["S"], nd or w["Nd"], w["seed"])
    eng = EkldEngine()
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    hX, hY, hH, hXd = pin(X), pin(Y), pin(hyp), pin(Xd)
    for name, xd in (("device design", hXd.to(eng.device)), ("pinned host design", hXd)):
        sg = StepGraph(eng, hX, hY, hH, xd, w["nugget"] ** 2)
        ts = []
        for i in range(8):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record(); sg.launch(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        print("%s: replay %.4f ms (min %.4f), winner %s" % (name, float(np.median(ts[3:])), min(ts[3:]), sg.winner()))


if __name__ == "__main__":
    main()
