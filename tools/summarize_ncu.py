"""Turns the ncu outputs brought back in gpurun_out/ into the small text/CSV/JSON summaries committed under profiles/.

  python tools/summarize_ncu.py <launches.csv> <full.ncu-rep> <tag> [workload]
"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
launches, rep, tag = sys.argv[1], sys.argv[2], sys.argv[3]
workload = sys.argv[4] if len(sys.argv) > 4 else "llama2-7b-gptq"
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)

# ---- launch list: one decode step, per-kernel time and share
rows = list(csv.DictReader([l for l in open(launches) if not l.startswith("==")]))
names = [(r["Kernel Name"], float(r["Metric Value"]), r["Grid Size"], r["Block Size"]) for r in rows]
adv = [i for i, n in enumerate(names) if "decode_advance" in n[0]]
step = names[adv[1]:adv[2]] if len(adv) >= 3 else names[adv[-1]:]
tot = sum(v for _, v, _, _ in step)
with open(os.path.join(out_dir, f"{tag}_decode_step_launches.csv"), "w") as f:
    f.write("# one un-graphed decode step (second of three), `ncu --metrics gpu__time_duration.sum --clock-control none`;\n")
    f.write("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes\n")
    f.write("kernel,grid,block,duration_us,share\n")
    for n, v, g, b in step:
        short = n.split("(")[0]
        f.write(f'"{short}","{g}","{b}",{v / 1e3:.2f},{v / tot:.4f}\n')
agg = collections.OrderedDict()
for n, v, _, _ in step:
    k = n.split("(")[0].split("<")[0].replace("void ", "")
    agg.setdefault(k, [0, 0.0])
    agg[k][0] += 1
    agg[k][1] += v
lines = [f"decode step of {workload} (2 layers + head), total {tot / 1e3:.1f} us, {len(step)} launches", ""]
tag_note = f"profiles/{tag}"
for k, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
    lines.append(f"{k:45s} n={c:3d} total={v / 1e3:8.1f} us  avg={v / c / 1e3:7.1f} us  share={v / tot * 100:5.1f}%")

# ---- full report: headline metrics per captured kernel
# `rep` is the .ncu-rep, or the CSV of its raw page produced on the GPU box (`ncu -i x.ncu-rep --page raw --csv > x_raw.csv`: the
# reports themselves are too large to bring back)
raw = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(raw.splitlines()))
hdr, units, data = r[0], r[1], r[2:]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size"]
idx = {w: hdr.index(w) for w in want if w in hdr}
ki = hdr.index("Kernel Name")
lines += ["", "ncu --set full (one row per captured launch):", ""]
traffic = {}
for row in data:
    name = row[ki].split("(")[0].replace("void ", "")
    vals = {w: row[i] for w, i in idx.items()}
    lines.append(name)
    for w, i in idx.items():
        lines.append(f"    {w:70s} {row[i]:>14s} {units[i]}")

    def to_bytes(w):
        v, u = float(row[idx[w]].replace(",", "")), units[idx[w]].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
    if "attn_decode_paged" in name and "dram__bytes_read.sum" in idx:
        traffic = {"attn_decode_traffic_bytes_per_launch": to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum"),
                   "attn_decode_duration_us_under_ncu": float(row[idx["gpu__time_duration.sum"]].replace(",", ""))}
open(os.path.join(out_dir, f"{tag}_summary.txt"), "w").write("\n".join(lines) + "\n")
tpath = os.path.join(out_dir, "roofline_traffic.json")
allt = json.load(open(tpath)) if os.path.exists(tpath) else {}
if traffic:
    traffic["source"] = f"profiles/{tag}_summary.txt (ncu --set full, B=64, context 1536, 2-layer model; per launch)"
    allt[workload] = traffic
    json.dump(allt, open(tpath, "w"), indent=1)
print("\n".join(lines[:14]))
print(traffic)
