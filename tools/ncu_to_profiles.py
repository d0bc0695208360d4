#!/usr/bin/env python
"""Turn the ncu outputs tools/profile.sh left in gpurun_out/ into the tracked summaries under
profiles/ (run here, no GPU needed):

    python tools/ncu_to_profiles.py <workload> <tag> <path-name e.g. tc|simt>

Writes profiles/<tag>_<workload>_launches.csv (the launch list), profiles/<tag>_<workload>_ncu.md
(headline metrics, stall mix, hottest source lines) and updates profiles/traffic.json with the
per-launch DRAM traffic bench.py reports as roofline.traffic."""
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
wl, tag, path = sys.argv[1], sys.argv[2], sys.argv[3]
rep = os.path.join(ROOT, "gpurun_out", "prof_%s_%s.ncu-rep" % (wl, tag))
launches = os.path.join(ROOT, "gpurun_out", "launches_%s_%s.csv" % (wl, tag))
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)

if os.path.exists(launches):
    rows = [r for r in csv.reader(open(launches)) if r and not r[0].startswith("==")]
    keep = [rows[0]] + [r for r in rows[1:] if len(r) > 4]
    with open(os.path.join(out_dir, "%s_%s_launches.csv" % (tag, wl)), "w", newline="") as f:
        w = csv.writer(f)
        hdr = keep[0]
        idx = [i for i, h in enumerate(hdr) if h in ("ID", "Kernel Name", "Block Size", "Grid Size", "Metric Name",
                                                     "Metric Unit", "Metric Value")]
        for r in keep:
            w.writerow([r[i] for i in idx])

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}


def num(name):
    v, u = m[name]
    return float(v.replace(",", "")), u


def to_bytes(name):
    v, u = num(name)
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]


want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max"]
lines = ["# ncu summary: workload %s, path %s, tag %s" % (wl, path, tag), "",
         "Source: `ncu --set full --clock-control none --import-source on` of one launch of the step kernel inside "
         "`bench.py --workload %s --steps 10 --warmup 3 --no-graph` (tools/profile.sh).  Times under the profiler are "
         "cold-cache and serialised; bench.py's CUDA-event number is the one quoted as performance." % wl, "",
         "| metric | value | unit |", "|---|---|---|"]
for k in want:
    if k in m:
        lines.append("| %s | %s | %s |" % (k, m[k][0], m[k][1]))
traffic = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
lines += ["| dram traffic (read+write) | %.0f | byte |" % traffic, "", "## warp stall mix (stalled warps per issue-active cycle)", ""]
st = []
for h, (v, u) in m.items():
    if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
        st.append((float(v), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
for v, h in sorted(st, reverse=True)[:8]:
    lines.append("* %s: %.2f" % (h, v))

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True,
                     text=True).stdout
rows = list(csv.reader(src.splitlines()))
if len(rows) > 3:
    h2 = rows[2]
    ci = {}
    for i, h in enumerate(h2):
        ci.setdefault(h, i)
    S, E = ci["# Samples"], ci["Instructions Executed"]
    agg = []
    for r in rows[3:]:
        if r and r[0].strip().isdigit():
            agg.append((int(r[0]), int(r[S]) if r[S].isdigit() else 0, int(r[E]) if r[E].isdigit() else 0, r[1].strip()[:110]))
    tot = sum(a[1] for a in agg) or 1
    lines += ["", "## hottest source lines (warp-stall samples; file %s)" % rows[0][1], "",
              "| line | samples | % | warp instr | source |", "|---|---|---|---|---|"]
    for ln, s, e, t in sorted(agg, key=lambda a: -a[1])[:14]:
        lines.append("| %d | %d | %.1f | %d | `%s` |" % (ln, s, 100.0 * s / tot, e, t.replace("|", "\\|")))
with open(os.path.join(out_dir, "%s_%s_ncu.md" % (tag, wl)), "w") as f:
    f.write("\n".join(lines) + "\n")

tj = os.path.join(out_dir, "traffic.json")
d = json.load(open(tj)) if os.path.exists(tj) else {}
d["%s:%s" % (wl, path)] = traffic
json.dump(d, open(tj, "w"), indent=1, sort_keys=True)
print("\n".join(lines[:30]))
print("traffic", traffic)
