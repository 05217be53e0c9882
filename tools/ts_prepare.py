"""Developer: clock64 timeline of the prepare kernel (stamped debug build, `make -C bode_b200/csrc dbg`)."""
import os, sys, ctypes, argparse
import numpy as np
os.environ["BODE_B200_LIB"] = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dbg", "libbode_dbg.so")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bode_b200 import _ext
from bode_b200.engine import EkldEngine
ap = argparse.ArgumentParser()
ap.add_argument("--d", type=int, default=8); ap.add_argument("--n", type=int, default=200); ap.add_argument("--S", type=int, default=64)
a = ap.parse_args()
rng = np.random.default_rng(1223)
X = rng.random((a.n, a.d)); Y = np.sin(3 * X.sum(1))[:, None]
hyp = np.column_stack([rng.gamma(1.0, 2.0, a.S) + 0.1] + [np.maximum(rng.exponential(0.7, a.S), 0.05) for _ in range(a.d)])
eng = EkldEngine()
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); eng.prepare(X, Y, hyp, 1e-6); e1.record(); torch.cuda.synchronize()
print("prepare ms", e0.elapsed_time(e1))
if a.n <= 224 and os.environ.get("BODE_PREP_VARIANT") != "0":
    buf = (ctypes.c_longlong * 1024)()
    assert _ext.load().bode_debug_prepare_tiles(buf) == 0
    t = np.array(buf[:], dtype=np.int64).reshape(64, 16)[: min(64, a.S)]
    names = ["setup+U+e", "gram", "cholesky", "scalars", "inverse", "w", "copy-out"]
    dt = np.diff(t[:, :8], axis=1)
    print("tile kernel, mean clk per section:", {n_: int(v) for n_, v in zip(names, dt.mean(axis=0))}, "total", int((t[:, 7] - t[:, 0]).mean()))
    print("cholesky detail (sum over panels): update warp1 %d, update warp0 %d, diag factor+inverse warp0 %d, warp1 update+wait %d" % (
        t[:, 8].mean(), t[:, 13].mean(), (t[:, 9] - t[:, 13]).mean(), t[:, 14].mean()))
    sys.exit(0)
buf = (ctypes.c_longlong * 512)()
assert _ext.load().bode_debug_prepare(buf) == 0
t = np.array(buf[:], dtype=np.int64).reshape(64, 8)[: min(64, a.S)]
names = ["setup+U+e", "gram", "cholesky", "scalars", "w backsub", "inverse", "retile"]
dt = np.diff(t[:, :7], axis=1)
print("mean clk per section:", {n_: int(v) for n_, v in zip(names[1:], dt.mean(axis=0))}, "total", int((t[:, 6] - t[:, 0]).mean()))
