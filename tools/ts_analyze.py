import numpy as np, sys
a = np.fromfile(sys.argv[1], dtype=np.int64).reshape(8, 16, 16, 8)
names = ["start", "aux_ready", "A_done", "postA_bar", "B_done", "red_done", "postB_bar", "epi_done"]
for item in (0, 3):
    t = a[item]
    if t[0, 0, 0] == 0: continue
    base = t[:, 0, 0].min()
    print("item", item, "samples recorded:", int((t[0, :, 0] > 0).sum()))
    for ls in range(1, 4):
        if t[0, ls, 0] == 0: break
        print(" sample", ls)
        for w in (0, 1, 5, 10, 15):
            r = t[w, ls] - t[w, ls, 0]
            print("   warp %2d: " % w + "  ".join("%s+%d" % (names[k], r[k]) for k in range(1, 8)) + "   (start at %d)" % (t[w, ls, 0] - base))
    # per-phase average over warps and samples 1..N-1
    ns = int((t[0, :, 0] > 0).sum())
    d = np.diff(t[:, 1:ns, :], axis=2).astype(float)
    print(" mean clk per phase over warps/samples:", {names[k + 1]: int(d[:, :, k].mean()) for k in range(7)}, " total per sample", int((t[:, 2:ns, 0] - t[:, 1:ns - 1, 0]).mean()))
