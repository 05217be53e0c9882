"""Developer: run the scoring kernel from a stamped debug build (tools/dbg/libbode_dbg.so, built out of tree from
csrc/ with clock64 stamps) and print the per-sample, per-warp timeline."""
import os, sys, ctypes
import numpy as np
os.environ["BODE_B200_LIB"] = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dbg", "libbode_dbg.so")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bode_b200 import _ext
from bode_b200.engine import EkldEngine
rng = np.random.default_rng(1223)
n, d, S, Nd = 200, 8, 64, 100000
X = rng.random((n, d)); Y = np.sin(3 * X.sum(1))[:, None]
hyp = np.column_stack([rng.gamma(1.0, 2.0, S) + 0.1] + [np.maximum(rng.exponential(0.7, S), 0.05) for _ in range(d)])
Xd = torch.from_numpy(rng.random((Nd, d))).cuda()
eng = EkldEngine()
eng.prepare(X, Y, hyp, 1e-6); eng.score(Xd, timed=True); print("k_score ms", eng.last_kernel_ms)
lib = _ext.load()
lib.bode_debug_dump(b"/tmp/ts.bin")
a = np.fromfile("/tmp/ts.bin", dtype=np.int64).reshape(8, 16, 16, 8)
t = a[2]
ns = int((t[0, :, 0] > 0).sum())
base = t[:, 1, 0].min()
for ls in (2, 3):
    print("sample", ls)
    for w in range(16):
        r = t[w, ls] - base
        print("  warp %2d rg %d cg %d: start %6d A_done %6d postA %6d B_done %6d postB %6d epi %6d | A %5d  B %5d" % (w, w // 4, w % 4, r[0], r[2], r[3], r[4], r[6], r[7], r[2] - r[1], r[4] - r[3]))
per = (t[:, 2:ns, 0] - t[:, 1:ns - 1, 0]).mean()
print("mean clk/sample %.0f | A span (first start -> postA barrier) %.0f | B span (postA -> postB barrier) %.0f | mean B per warp %.0f | epilogue tail %.0f" % (
    per, (t[:, 1:ns, 3].max(axis=0) - t[:, 1:ns, 0].min(axis=0)).mean(), (t[:, 1:ns, 6].max(axis=0) - t[:, 1:ns, 3].max(axis=0)).mean(),
    (t[:, 1:ns, 4] - t[:, 1:ns, 3]).mean(), (t[:, 1:ns, 7].max(axis=0) - t[:, 1:ns, 6].max(axis=0)).mean()))
