"""Print the headline numbers of bench.py JSON lines found in log files (GPU-call helper)."""
import json
import sys

for f in sys.argv[1:]:
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
    except Exception as e:
        print(f, "no json line", e)
        continue
    cls = {k: round(v["ms_per_step"], 3) for k, v in (d.get("roofline") or {}).get("classes", {}).items()}
    sec = {}
    for k in ("pixelcnn", "vqvae"):
        if k in d:
            sec[k] = d[k].get("value") if "value" in d[k] else {kk: vv.get("value") for kk, vv in d[k].items() if isinstance(vv, dict)}
    print(f.split("/")[-1], "value", round(d["value"], 2), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 2),
          "samples/s", d.get("samples_per_sec_1000step"), cls, sec)
