import numpy as np, sys
a = np.fromfile(sys.argv[1], dtype=np.int64)[: 2 * 16 * 4 * 16 * 4].reshape(2, 16, 4, 16, 4)
item, ls = 0, 2
t = a[item, :, ls]
nq = int((t[0, :, 0] > 0).sum())
base = t[:, 0, 0].min()
print("chunks", nq)
for w in (0, 1, 2, 3, 7, 12):
    print("warp %2d rg %d: " % (w, w % 4) + " | ".join("q%d s%d w%d c%d r%d" % (q, t[w, q, 0] - base, t[w, q, 1] - t[w, q, 0], t[w, q, 2] - t[w, q, 1], t[w, q, 3] - t[w, q, 2]) for q in range(nq)))
tot = (t[:, :nq, 3] - t[:, :nq, 0])
print("mean per-chunk clk: wait %.0f compute %.0f release %.0f ; B total mean %.0f" % ((t[:, :nq, 1] - t[:, :nq, 0]).mean(), (t[:, :nq, 2] - t[:, :nq, 1]).mean(), (t[:, :nq, 3] - t[:, :nq, 2]).mean(), (t[:, nq - 1, 3] - t[:, 0, 0]).mean()))
print("per-chunk compute clk by rg (mean over cg):")
for rg in range(4):
    print("  rg", rg, [(int((t[rg::4, q, 2] - t[rg::4, q, 1]).mean())) for q in range(nq)], "sum", int((t[rg::4, :nq, 2] - t[rg::4, :nq, 1]).sum(axis=1).mean()))
