"""profiles/<round>_traffic.json from the ncu summaries (tools/ncu_summary.py output): per-launch DRAM traffic of
the three hot kernels, read by bench.py for roofline.traffic.  Usage: python tools/traffic_json.py [round]"""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rnd = sys.argv[1] if len(sys.argv) > 1 else "r01"
out = {}
for k in ("decode", "encode", "stats"):
    path = os.path.join(ROOT, "profiles", f"{rnd}_final_{k}_c3.summary.txt")
    if not os.path.exists(path):
        path = os.path.join(ROOT, "profiles", f"{rnd}_{k}_c3.summary.txt")
    txt = open(path).read()

    def val(name):
        return float(re.search(rf"{re.escape(name)}\s+([0-9.]+)", txt).group(1))

    rd, wr = val("dram__bytes_read.sum") * 1e6, val("dram__bytes_write.sum") * 1e6  # summaries print MB
    out[k] = {"dram_bytes_read": rd, "dram_bytes_write": wr, "traffic": rd + wr,
              "duration": val("gpu__time_duration.sum"), "duration_unit": "us", "workload": "c3",
              "capture": f"{os.path.basename(path)} (ncu --set full --clock-control none, tools/gpu_final.sh)"}
json.dump(out, open(os.path.join(ROOT, "profiles", f"{rnd}_traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
