#!/usr/bin/env python
"""Summarise gpurun_out ncu artefacts into profiles/ (tracked).  Usage:
   python tools/ncu_summary.py <tag> [--launches gpurun_out/launches.csv] [--rep gpurun_out/prof_step.ncu-rep]"""
import argparse, collections, csv, io, json, os, subprocess, sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
           "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__block_size", "launch__grid_size",
           "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg",
           "smsp__inst_executed_pipe_fma.sum", "smsp__inst_executed_pipe_alu.sum", "smsp__inst_executed_pipe_lsu.sum",
           "smsp__inst_executed_pipe_xu.sum", "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]

ap = argparse.ArgumentParser()
ap.add_argument("tag")
ap.add_argument("--launches", default="gpurun_out/launches.csv")
ap.add_argument("--rep", default="gpurun_out/prof_step.ncu-rep")
ap.add_argument("--birds", type=int, default=16_000_000)
a = ap.parse_args()
os.makedirs("profiles", exist_ok=True)
out = {"tag": a.tag}
if os.path.exists(a.launches):
    rows = [r for r in csv.reader(open(a.launches)) if len(r) > 5]
    h = rows[0]; ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    agg = collections.defaultdict(list)
    for r in rows[1:]:
        try: agg[r[ki].split("(")[0]].append(float(r[vi].replace(",", "")))
        except ValueError: pass
    tot = sum(sum(v) for v in agg.values())
    out["launch_list"] = {k: {"launches": len(v), "avg_us": sum(v) / len(v) / 1e3, "share": sum(v) / tot}
                          for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]))}
    out["launch_list_note"] = "ncu --metrics gpu__time_duration.sum --clock-control none: cold-cache, serialised; compare SHARES"
if os.path.exists(a.rep):
    res = subprocess.run(["ncu", "-i", a.rep, "--page", "raw", "--csv"], capture_output=True, text=True)
    rows = list(csv.reader(io.StringIO(res.stdout)))
    h = rows[0]
    ker = {}
    for m in METRICS:
        if m in h:
            i = h.index(m)
            ker[m] = {"unit": rows[1][i], "values": [r[i] for r in rows[2:]]}
    name_i = h.index("Kernel Name")
    out["kernel"] = rows[2][name_i]
    out["full_capture"] = ker
    try:
        rd = [float(x) for x in ker["dram__bytes_read.sum"]["values"]]
        wr = [float(x) for x in ker["dram__bytes_write.sum"]["values"]]
        scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[ker["dram__bytes_read.sum"]["unit"]]
        t = (sum(rd) + sum(wr)) / len(rd) * scale
        out["traffic_bytes_per_launch"] = t
        out["traffic_bytes_per_bird"] = t / a.birds
    except Exception as e:
        out["traffic_error"] = str(e)
json.dump(out, open(f"profiles/{a.tag}.json", "w"), indent=1)
print(json.dumps(out, indent=1)[:3000])
