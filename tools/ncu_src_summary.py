"""Summarise an `ncu --page source --csv` dump: executed instructions and stall samples per opcode / hot regions."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
ops = collections.Counter(); samp = collections.Counter()
tot_inst = 0; tot_samp = 0
body = rows[2:]
for r in body:
    if len(r) < len(hdr): continue
    src = r[ix["Source"]].strip()
    op = src.split()[0] if src else "?"
    if op.startswith("@"): op = src.split()[1]
    op = op.split(".")[0] + ("." + op.split(".")[1] if op.startswith(("LDS","STS","LDG","STG","DMMA","BAR","SYNCS","LDL","STL")) and "." in op else "")
    ie = int(float(r[ix["Instructions Executed"]] or 0)); s = int(float(r[ix["# Samples"]] or 0))
    ops[op] += ie; samp[op] += s; tot_inst += ie; tot_samp += s
print("total warp-instructions %d, samples %d" % (tot_inst, tot_samp))
print("%-14s %14s %7s %9s %7s" % ("opcode", "inst", "inst%", "samples", "samp%"))
for op, c in sorted(ops.items(), key=lambda kv: -samp[kv[0]])[:28]:
    print("%-14s %14d %6.2f%% %9d %6.2f%%" % (op, c, 100.0 * c / tot_inst, samp[op], 100.0 * samp[op] / max(1, tot_samp)))
if len(sys.argv) > 2:
    # top-N hottest individual instructions with their stall breakdown
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    top = sorted(body, key=lambda r: -int(float(r[ix["# Samples"]] or 0)) if len(r) >= len(hdr) else 0)[: int(sys.argv[2])]
    for r in top:
        st = sorted(((int(float(r[ix[c]] or 0)), c) for c in stall_cols), reverse=True)[:3]
        print("%s  samples=%s exec=%s  %s  | %s" % (r[ix["Address"]][-5:], r[ix["# Samples"]], r[ix["Instructions Executed"]],
              r[ix["Source"]].strip()[:70], ", ".join("%s=%d" % (c[6:], v) for v, c in st)))

# ---- split at the CTA-wide barriers: (setup) | phase A ... BAR | phase B ... BAR | combine -----------------------------
full = [r for r in body if len(r) >= len(hdr)]
bars = [i for i, r in enumerate(full) if r[ix['Source']].strip().split()[0].startswith('BAR') or ' BAR.SYNC' in r[ix['Source']]]
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
def S(rs): return sum(int(float(r[ix['# Samples']] or 0)) for r in rs)
def I(rs): return sum(int(float(r[ix['Instructions Executed']] or 0)) for r in rs)
tot = max(1, S(full))
cuts = [0] + [b + 1 for b in bars] + [len(full)]
print("regions between CTA-wide barriers (SASS order): samples%, inst%, DMMA share of the region's instructions, top stalls")
for lo, hi in zip(cuts, cuts[1:]):
    rs = full[lo:hi]
    if S(rs) < 0.005 * tot: continue
    nd = sum(int(float(r[ix['Instructions Executed']] or 0)) for r in rs if 'DMMA' in r[ix['Source']])
    d = {c: sum(int(float(r[ix[c]] or 0)) for r in rs) for c in stall_cols}
    t = max(1, sum(d.values()))
    print("  sass[%6d:%6d]  samples %5.1f%%  inst %5.1f%%  dmma/inst %4.1f%%  stalls: %s" % (lo, hi, 100.0 * S(rs) / tot, 100.0 * I(rs) / max(1, tot_inst),
          100.0 * nd / max(1, I(rs)), ", ".join("%s %.0f%%" % (k[6:], 100.0 * v / t) for k, v in sorted(d.items(), key=lambda kv: -kv[1])[:5])))
