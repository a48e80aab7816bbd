"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, mean, share."""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        v = float(row["Metric Value"].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3}.get(row["Metric Unit"], v)
        agg.setdefault(name, []).append(v)
    tot = sum(sum(v) for v in agg.values())
    print(f"total {tot:.1f} us over {sum(len(v) for v in agg.values())} launches")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k[:60]:60s} n={len(v):3d} mean={sum(v) / len(v):9.1f}us sum={sum(v):10.1f}us share={100 * sum(v) / tot:5.1f}%")


if __name__ == "__main__":
    main(sys.argv[1])
