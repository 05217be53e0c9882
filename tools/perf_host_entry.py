import sys, time, numpy as np
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from bode_b200 import _ext as ext
w = bench.CONFIGS["c3"]
X, Y, hyp, Xd = bench.make_workload(w["d"], w["n"], w["S"], w["Nd"], w["seed"])
ctx = ext.HostContext()
for i in range(6):
    t0 = time.perf_counter()
    r = ext.ekld_score_host(X, Y, hyp, Xd, w["nugget"] ** 2, ctx=ctx)
    t1 = time.perf_counter()
    print("host entry call %d: %.3f ms  argmax %d" % (i, (t1 - t0) * 1e3, int(r["argmax"]) if "argmax" in r else -1))
