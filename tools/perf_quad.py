"""Developer probe: throughput of the quadrature row-mean kernel (K5)."""
import sys, os, ctypes, argparse
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bode_b200 import _ext
ap = argparse.ArgumentParser()
ap.add_argument("--d", type=int, default=10); ap.add_argument("--Np", type=int, default=100000)
ap.add_argument("--Nq", type=int, default=10000); ap.add_argument("--S", type=int, default=8)
a = ap.parse_args()
lib = _ext.load()
rng = np.random.default_rng(0)
hyp = torch.from_numpy(np.column_stack([rng.gamma(2, 1, a.S) + 0.2] + [rng.uniform(0.2, 0.9, a.S) for _ in range(a.d)])).cuda()
P = torch.from_numpy(rng.random((a.Np, a.d))).cuda(); Xq = torch.from_numpy(rng.random((a.Nq, a.d))).cuda()
out = torch.empty((a.S, a.Np), dtype=torch.float64, device="cuda"); scr = torch.empty(4096, dtype=torch.uint8, device="cuda")
p = lambda t: ctypes.c_void_p(t.data_ptr())
for rep in range(4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _ext.check(lib.bode_quad_rowmean_dev(p(hyp), a.S, a.d + 1, a.d, p(P), a.Np, p(Xq), None, a.Nq, p(out), p(scr), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    pairs = a.S * a.Np * a.Nq
    print(f"rep {rep}: {ms:.3f} ms  {pairs/ms*1e-6:.1f} G pair-evals/s  ({pairs*(2*a.d+12)/ms*1e-9:.2f} T FP64-op/s of ~18.6 issue peak)")
