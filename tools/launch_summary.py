"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): per-kernel count, mean and total time."""
import csv
import sys
from collections import defaultdict


def main(path):
    rows = []
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            v = float(r["Metric Value"].replace(",", ""))
            unit = r.get("Metric Unit", "ns")
            v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0}.get(unit, 1e-6)
            rows.append((r["Kernel Name"].split("(")[0][:60], v))
    agg = defaultdict(list)
    for k, v in rows:
        agg[k].append(v)
    tot = sum(v for _, v in rows)
    print("kernel,launches,mean_ms,total_ms,share")
    for k, vs in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print("%s,%d,%.4f,%.3f,%.3f" % (k, len(vs), sum(vs) / len(vs), sum(vs), sum(vs) / tot))


if __name__ == "__main__":
    main(sys.argv[1])
