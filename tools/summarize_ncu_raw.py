"""Per-kernel table out of an `ncu -i X.ncu-rep --page raw --csv` dump: duration, registers, achieved occupancy,
issue-active %, DRAM throughput % and the top warp-stall reason.  Usage: summarize_ncu_raw.py raw.csv > profiles/x.md"""
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, data = rows[0], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}


def f(r, k):
    try:
        return float(r[ix[k]].replace(",", ""))
    except Exception:
        return float("nan")


stalls = [h for h in hdr if "issue_stalled" in h and h.endswith("per_issue_active.ratio")]
print("| kernel | grid | us | regs | occupancy % | issue active % | DRAM % | top stall (warps per issue) |")
print("|---|---|---:|---:|---:|---:|---:|---|")
seen = set()
for r in data:
    name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).split("::")[-1]
    key = (name, r[ix["Grid Size"]])
    if key in seen:
        continue
    seen.add(key)
    top = max(((f(r, s), s) for s in stalls if "selected" not in s), default=(0, ""))
    reason = re.sub(r"smsp__average_warps_issue_stalled_|_per_issue_active.ratio", "", top[1])
    print(f"| `{name}` | {r[ix['Grid Size']]} | {f(r, 'gpu__time_duration.sum'):.1f} | "
          f"{int(f(r, 'launch__registers_per_thread'))} | {f(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):.0f} | "
          f"{f(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):.0f} | "
          f"{f(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.0f} | {reason} {top[0]:.1f} |")
