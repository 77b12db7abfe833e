"""profiles/r<N>_conv_tc_dram.json (bench.py's roofline.traffic) out of an ncu CSV with dram__bytes_read.sum / dram__bytes_write.sum /
gpu__time_duration.sum for every conv_tc_kernel launch of ONE training step (tools/profile_step.py --what train)."""
import csv
import json
import sys
from collections import defaultdict

rows = [l for l in open(sys.argv[1]) if l.startswith('"')]
agg = defaultdict(lambda: defaultdict(float))
for r in csv.DictReader(rows):
    if "conv_tc_kernel" not in r["Kernel Name"]:
        continue
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3}.get(u, 1)
    agg[r["ID"]][r["Metric Name"]] = v * scale
n = len(agg)
rd = sum(a["dram__bytes_read.sum"] for a in agg.values())
wr = sum(a["dram__bytes_write.sum"] for a in agg.values())
us = sum(a["gpu__time_duration.sum"] for a in agg.values())
out = {"launches": n, "avg_dram_bytes_per_launch": (rd + wr) / n, "dram_read_bytes_total": rd, "dram_write_bytes_total": wr,
       "ncu_time_us_total": us,
       "note": f"ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none over all {n} "
               f"conv_tc_kernel launches of one training step (B=128, cold cache per launch): {rd / 1e6:.0f} MB read + {wr / 1e6:.0f} MB written; "
               "launches include the 8 per-image attention convs of the two 32x32 blocks; ncu --set full captures of individual layers: "
               "profiles/r2_ncu_full.md (8x8 256->256: 17.4 MB read, tensor pipe 56 % of active / 42 % of elapsed cycles)"
               }
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps(out))
