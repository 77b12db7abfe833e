"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a markdown table
(per-kernel launches, total device time, share).  Usage: summarize_ncu.py launches.csv > profiles/x.md"""
import csv
import re
import sys
from collections import defaultdict


def short(name):
    name = re.sub(r"\(.*$", "", name)
    name = re.sub(r"void |igm::|\(anonymous namespace\)::|unnamed>::|<unnamed>::", "", name)
    return name.strip()[:90]


def main(path):
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.DictReader(lines)
    agg = defaultdict(lambda: [0, 0.0])
    total = 0.0
    n = 0
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = v / 1e3 if unit == "ns" else (v if unit in ("us", "usecond") else v * 1e3)
        k = short(r["Kernel Name"])
        agg[k][0] += 1
        agg[k][1] += us
        total += us
        n += 1
    print(f"launches: {n}, summed device time: {total / 1e3:.3f} ms (ncu-serialised, cold cache: compare SHARES)\n")
    print("| kernel | launches | total us | share |")
    print("|---|---:|---:|---:|")
    for k, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {c} | {us:.1f} | {100 * us / total:.1f}% |")


if __name__ == "__main__":
    main(sys.argv[1])
