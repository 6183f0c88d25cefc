#!/usr/bin/env python
"""Summarise gpurun_out ncu artefacts into profiles/ (tracked).
Usage: python tools/ncu_summary.py <tag>   reads gpurun_out/launches_<tag>.csv, gpurun_out/prof_<tag>.ncu-rep,
gpurun_out/bench_<tag>.json and writes profiles/<tag>_launches.txt, profiles/<tag>_kernels.csv, profiles/<tag>_bench.json,
and updates profiles/traffic.json (DRAM bytes per launch of the dominant kernel, read by bench.py)."""
import collections
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__inst_executed.sum", "smsp__inst_executed.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_lsu.sum",
        "smsp__inst_executed_pipe_alu.sum", "smsp__inst_executed_pipe_fma.sum", "smsp__inst_executed_pipe_lsu.sum",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__thread_inst_executed_per_inst_executed.ratio"]


def to_unit(v, u, kind):
    v = float(v.replace(",", ""))
    if kind == "time":
        return v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)          # -> us
    if kind == "bytes":
        return v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
    return v


lf = os.path.join(G, "launches_%s.csv" % tag)
if os.path.exists(lf):
    rows = list(csv.reader(open(lf)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h, data = rows[hi], rows[hi + 1:]
    kn, mv, mu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= mv:
            continue
        a = agg.setdefault(r[kn], [0, 0.0])
        a[0] += 1
        a[1] += to_unit(r[mv], r[mu], "time")
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(P, "%s_launches.txt" % tag), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none  python bench.py --steps 2 --warmup 1 --no-cpu-baseline\n")
        f.write("# cold-cache, serialised launch times: compare SHARES, not absolutes.  %d launches, %.1f us total\n" % (len(data), tot))
        for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write("%-110s n=%3d total_us=%10.1f share=%5.1f%%\n" % (n[:110], c, t, 100 * t / tot))
    print("wrote launches summary")

rep = os.path.join(G, "prof_%s.ncu-rep" % tag)
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(raw)))
    h, units, data = r[0], r[1], r[2:]
    with open(os.path.join(P, "%s_kernels.csv" % tag), "w") as f:
        w = csv.writer(f)
        cols = ["Kernel Name"] + [k for k in KEYS if k in h]
        w.writerow(cols)
        w.writerow([""] + [units[h.index(k)] for k in cols[1:]])
        for row in data:
            w.writerow([row[h.index(k)] for k in cols])
    traffic = {}
    for row in data:
        name = row[h.index("Kernel Name")]
        rd = to_unit(row[h.index("dram__bytes_read.sum")], units[h.index("dram__bytes_read.sum")], "bytes")
        wr = to_unit(row[h.index("dram__bytes_write.sum")], units[h.index("dram__bytes_write.sum")], "bytes")
        traffic.setdefault(name, []).append(rd + wr)
    out = {}
    for name, v in traffic.items():
        key = None
        if "expand_heavy" in name:
            key = "expand_heavy_kernel"
        elif "expand_kernel" in name:
            key = "expand_kernel"
        elif "lookup_kernel" in name:
            key = "lookup_kernel"
        elif "bookkeep_kernel<1>" in name or "bookkeep_kernel<(int)1>" in name:
            key = "bookkeep_kernel_1"
        elif "bookkeep_kernel<8>" in name or "bookkeep_kernel<(int)8>" in name:
            key = "bookkeep_kernel_8"
        elif "sw_fast" in name:
            key = "sw_fast_kernel"
        if key:
            # the second bookkeep<8> launch of a tile (hand-overs of the warp launch) is usually empty: keep the largest
            out[key + "_dram_bytes_per_launch"] = max(v)
    # the roofline entry of bench.py covers the whole prefilter (one batch = one launch of each kernel class)
    out["prefilter_dram_bytes_per_launch"] = sum(out.get(k + "_dram_bytes_per_launch", 0.0)
                                                 for k in ("expand_kernel", "expand_heavy_kernel", "lookup_kernel", "bookkeep_kernel_1", "bookkeep_kernel_8"))
    out["source"] = "profiles/%s_kernels.csv (ncu --set full, cold cache)" % tag
    json.dump(out, open(os.path.join(P, "traffic.json"), "w"), indent=1)
    print("wrote kernel metrics", out)

rep = os.path.join(G, "prof_exh_%s.ncu-rep" % tag)
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(raw)))
    h, units, data = r[0], r[1], r[2:]
    with open(os.path.join(P, "%s_exhaustive_kernels.csv" % tag), "w") as f:
        w = csv.writer(f)
        cols = ["Kernel Name"] + [k for k in KEYS + ["smsp__inst_executed_pipe_alu.sum", "smsp__inst_executed_pipe_fmaheavy.sum",
                                                     "smsp__inst_executed_pipe_fmalite.sum", "sm__inst_executed_pipe_alu.sum"] if k in h]
        cols = list(collections.OrderedDict.fromkeys(cols))
        w.writerow(cols)
        w.writerow([""] + [units[h.index(k)] for k in cols[1:]])
        for row in data:
            w.writerow([row[h.index(k)] for k in cols])
    print("wrote exhaustive kernel metrics")

for name in ("bench_%s.json" % tag, "bench_ref_%s.json" % tag):
    src = os.path.join(G, name)
    if os.path.exists(src) and os.path.getsize(src):
        shutil.copy(src, os.path.join(P, "%s" % name.replace("bench_", tag + "_bench_").replace("_%s.json" % tag, ".json")))
