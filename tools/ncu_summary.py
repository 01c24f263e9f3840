"""Summarise an .ncu-rep (here, no GPU): key metrics, pipe utilisation, stall reasons, hottest SASS lines.
usage: python tools/ncu_summary.py file.ncu-rep [n_hot_lines]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
nhot = int(sys.argv[2]) if len(sys.argv) > 2 else 12
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rows[0], rows[-1]
m = dict(zip(hdr, vals))
print("kernel:", m.get("Kernel Name", "")[:100])
keys = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_warps", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"]
for k in keys:
    if k in m:
        print(f"  {k:72s} {m[k]}")
print("pipes (pct of peak, active):")
for h, v in m.items():
    if h.startswith("sm__inst_executed_pipe_") and h.endswith(".avg.pct_of_peak_sustained_active"):
        try:
            if float(v) > 1:
                print(f"  {h[len('sm__inst_executed_pipe_'):-len('.avg.pct_of_peak_sustained_active')]:24s} {float(v):6.1f}")
        except ValueError:
            pass
st = {}
for h, v in m.items():
    if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h:
        try:
            st[h.replace("smsp__pcsamp_warps_issue_stalled_", "")] = float(v)
        except ValueError:
            pass
tot = sum(st.values()) or 1
print("stall samples:")
for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]:
    print(f"  {k:28s} {100 * v / tot:5.1f}%")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]
isrc, isamp, iex = h.index("Source"), h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed")
sc = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
data = []
for r in rows[2:]:
    try:
        data.append((int(r[isamp]), r))
    except ValueError:
        pass
tot = sum(s for s, _ in data) or 1
print("hottest SASS lines:")
for s, r in sorted(data, key=lambda x: -x[0])[:nhot]:
    rs = sorted([(int(r[i] or 0), h[i]) for i in sc if r[i] not in ("", "0")], reverse=True)[:2]
    print(f"  {100 * s / tot:5.1f}% ex={r[iex]:>9s} {r[isrc][:58]:58s} {rs}")
