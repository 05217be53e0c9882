"""Developer: run the scoring kernel from the stamped debug build (`make -C bode_b200/csrc dbg` -> tools/dbg/libbode_dbg.so,
compiled with -DBODE_TIMELINE) and print the per-sample, per-warp timeline of a few steady-state work items."""
import os, sys, argparse
import numpy as np
os.environ["BODE_B200_LIB"] = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dbg", "libbode_dbg.so")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bode_b200 import _ext
from bode_b200.engine import EkldEngine
ap = argparse.ArgumentParser()
ap.add_argument("--d", type=int, default=8); ap.add_argument("--n", type=int, default=200)
ap.add_argument("--S", type=int, default=64); ap.add_argument("--Nd", type=int, default=100000)
ap.add_argument("--verbose", type=int, default=1)
ap.add_argument("--dbg-a", type=int, default=0, help="phase-A switches: 8 no DMMA, 16 no exp, 32 no Kc store, 64 no loads")
ap.add_argument("--out", default="/tmp/ts.bin")
a_ = ap.parse_args()
rng = np.random.default_rng(1223)
n, d, S, Nd = a_.n, a_.d, a_.S, a_.Nd
X = rng.random((n, d)); Y = np.sin(3 * X.sum(1))[:, None]
hyp = np.column_stack([rng.gamma(1.0, 2.0, S) + 0.1] + [np.maximum(rng.exponential(0.7, S), 0.05) for _ in range(d)])
Xd = torch.from_numpy(rng.random((Nd, d))).cuda()
eng = EkldEngine()
_ext.load().bode_debug_phase_a(a_.dbg_a)
for _ in range(2):
    eng.prepare(X, Y, hyp, 1e-6); eng.score(Xd, timed=True)
print("k_score ms", eng.last_kernel_ms)
lib = _ext.load()
assert lib.bode_debug_dump(a_.out.encode()) == 0
raw = np.fromfile(a_.out, dtype=np.int64)
a = raw[: 8 * 16 * 16 * 8].reshape(8, 16, 16, 8)
tw = raw[8 * 16 * 16 * 8: 8 * 16 * 16 * 10].reshape(8, 16, 16, 2)
ta = raw[8 * 16 * 16 * 10:].reshape(8, 16, 16, 4)
names = ["start", "aux_ready", "A_done", "postA_bar", "B_done", "red_done", "postB_bar", "epi_done"]
for item in range(8):
    t = a[item]
    nw = int((t[:, 1, 0] > 0).sum())
    ns = int((t[0, :, 0] > 0).sum())
    if nw == 0 or ns < 3: continue
    t = t[:nw]
    if a_.verbose and item == 2:
        base = t[:, 1, 0].min()
        for ls in (2,):
            print("item", item, "sample", ls)
            for w in range(nw):
                r = t[w, ls] - base
                print("  warp %2d: start %6d A_done %6d postA %6d B_done %6d postB %6d epi %6d | A %5d (kc %5d, xik %5d)  B %5d  wait_full %5d wait_refill %5d" % (
                    w, r[0], r[2], r[3], r[4], r[6], r[7], r[2] - r[1], r[5] - r[1], r[2] - r[5], r[4] - r[3], tw[item, w, ls, 0], tw[item, w, ls, 1]) + "  | aux_wait %5d setup %5d loop %5d reduce %4d" % (r[1] - r[0], ta[item, w, ls, 0] - base - r[1], ta[item, w, ls, 1] - ta[item, w, ls, 0], ta[item, w, ls, 2] - ta[item, w, ls, 1]))
    per = (t[:, 2:ns, 0] - t[:, 1:ns - 1, 0]).mean()
    A = ta[item, :nw, 1:ns]
    print('   phase A inner (mean clk over warps): setup %.0f loop %.0f reduce %.0f | xik %.0f' % ((A[:, :, 0] - t[:, 1:ns, 1]).mean(), (A[:, :, 1] - A[:, :, 0]).mean(), (A[:, :, 2] - A[:, :, 1]).mean(), (t[:, 1:ns, 2] - t[:, 1:ns, 5]).mean()))
    print("item %d: clk/sample %.0f | A span (first start -> last postA) %.0f, mean A per warp %.0f (max %.0f) | B span (postA -> postB) %.0f, mean B per warp %.0f (max %.0f) | epi tail %.0f | ring wait full %.0f refill %.0f" % (
        item, per, (t[:, 1:ns, 3].max(axis=0) - t[:, 1:ns, 0].min(axis=0)).mean(), (t[:, 1:ns, 2] - t[:, 1:ns, 1]).mean(), (t[:, 1:ns, 2] - t[:, 1:ns, 1]).max(axis=0).mean(),
        (t[:, 1:ns, 6].max(axis=0) - t[:, 1:ns, 3].max(axis=0)).mean(), (t[:, 1:ns, 4] - t[:, 1:ns, 3]).mean(), (t[:, 1:ns, 4] - t[:, 1:ns, 3]).max(axis=0).mean(),
        (t[:, 1:ns, 7].max(axis=0) - t[:, 1:ns, 6].max(axis=0)).mean(), tw[item, :nw, 1:ns, 0].mean(), tw[item, :nw, 1:ns, 1].sum(axis=0).mean()))
