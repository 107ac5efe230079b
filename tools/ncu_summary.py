#!/usr/bin/env python
"""Summarise ncu output brought back from a gpurun call (run HERE, no GPU needed).

    python tools/ncu_summary.py launches gpurun_out/launches.csv            # per-kernel totals and shares
    python tools/ncu_summary.py full gpurun_out/prof.ncu-rep [--traffic]    # key metrics per captured launch
    python tools/ncu_summary.py sass gpurun_out/prof.ncu-rep k_trace [skip] # instruction mix by SASS region of launch #skip

`full --traffic` also writes profiles/ktrace_dram_traffic.json (mean dram bytes per k_trace launch), which bench.py
reports as roofline.traffic.
"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

METRICS = collections.OrderedDict([
    ("ms", "gpu__time_duration.sum"), ("dram_rd_MB", "dram__bytes_read.sum"), ("dram_wr_MB", "dram__bytes_write.sum"),
    ("dram%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), ("l2hit%", "lts__t_sector_hit_rate.pct"),
    ("l1hit%", "l1tex__t_sector_hit_rate.pct"), ("warps%", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("regs", "launch__registers_per_thread"), ("lanes", "smsp__thread_inst_executed_per_inst_executed.ratio"),
    ("issue%", "smsp__issue_active.avg.pct_of_peak_sustained_active"), ("fma%", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
    ("alu%", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"), ("Minst", "smsp__inst_executed.sum"),
    ("st_long_sb", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
    ("st_wait", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"),
    ("st_branch", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"),
    ("st_not_sel", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"),
    ("st_no_inst", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"),
])
SCALE = {"Gbyte": 1e3, "Mbyte": 1.0, "Kbyte": 1e-3, "byte": 1e-6, "ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}


def short(name):
    return re.sub(r".*::", "", name.split("(")[0]).replace("void ", "")


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        v = float(r[vi].replace(",", "")) * SCALE.get(r[ui], 1.0) * 1e3  # -> us
        agg[short(r[ki])][0] += 1
        agg[short(r[ki])][1] += v
    tot = sum(v[1] for v in agg.values())
    print("| kernel | launches | total us | share |\n|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {k} | {v[0]} | {v[1]:.0f} | {v[1] / tot * 100:.1f} % |")


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def full(path, traffic):
    hdr, units, rows = raw(path)
    print("| kernel | " + " | ".join(METRICS) + " |\n|" + "---|" * (len(METRICS) + 1))
    trace_bytes, trace_pipe = [], []  # per k_trace launch: dram bytes; (ms, issue %, alu pipe %, fma pipe %, lanes per instruction)
    for r in rows:
        name = short(r[hdr.index("Kernel Name")])
        vals = []
        for k, m in METRICS.items():
            if m not in hdr:
                vals.append("n/a")
                continue
            i = hdr.index(m)
            v = float(r[i].replace(",", ""))
            if k in ("ms", "dram_rd_MB", "dram_wr_MB"):
                v *= SCALE.get(units[i], 1.0)
            if k == "Minst":
                v /= 1e6
            vals.append(f"{v:.2f}" if k in ("ms", "st_long_sb", "st_wait", "st_branch", "st_not_sel", "st_no_inst") else f"{v:.1f}")
        print(f"| {name} | " + " | ".join(vals) + " |")
        if name.startswith("k_trace") and not name.startswith("k_trace_array"):  # the two probe launches of the tree choice are not the wavefront
            i, j = hdr.index(METRICS["dram_rd_MB"]), hdr.index(METRICS["dram_wr_MB"])
            trace_bytes.append((float(r[i]) * SCALE[units[i]] + float(r[j]) * SCALE[units[j]]) * 1e6)
            t = hdr.index(METRICS["ms"])
            trace_pipe.append([float(r[t].replace(",", "")) * SCALE.get(units[t], 1.0)] + [float(r[hdr.index(METRICS[k])].replace(",", "")) for k in ("issue%", "alu%", "fma%", "lanes")])
    if traffic and trace_bytes:
        w = sum(p[0] for p in trace_pipe)
        d = {"dram_bytes_per_launch": sum(trace_bytes) / len(trace_bytes), "launches": len(trace_bytes), "source": os.path.basename(path),
             "how": "mean of dram__bytes_read.sum + dram__bytes_write.sum over the captured k_trace launches (ncu --set full)",
             # duration-weighted means over the same launches: what the kernel is actually bound by when the BVH is L2-resident
             "issue_active_pct": sum(p[0] * p[1] for p in trace_pipe) / w, "alu_pipe_pct": sum(p[0] * p[2] for p in trace_pipe) / w,
             "fma_pipe_pct": sum(p[0] * p[3] for p in trace_pipe) / w, "active_lanes_per_instruction": sum(p[0] * p[4] for p in trace_pipe) / w}
        # stamped with the CUDA sources of the tree this runs in: summarise a capture BEFORE changing the kernels again (bench.py prints
        # "traffic_stale": true when the hash differs)
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        from srchash import csrc_sha
        d["csrc_sha"] = csrc_sha()
        json.dump(d, open(os.path.join(ROOT, "profiles", "ktrace_dram_traffic.json"), "w"), indent=1)
        print("\nwrote profiles/ktrace_dram_traffic.json:", d)


def sass(path, kernel, seg=20, skip=0):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", f"regex:{kernel}", "--launch-skip", str(skip), "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    ie, te, sm = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
    data = [(r[1].strip(), int(r[ie]), int(r[te]), int(r[sm])) for r in rows[2:] if len(r) > te and r[ie].isdigit()]
    half = len(data) // 2
    if half and data[0][0] == data[half][0]:
        data = data[:half]
    tot, tott, tots = sum(d[1] for d in data), sum(d[2] for d in data), max(sum(d[3] for d in data), 1)
    print(f"{len(data)} SASS instructions, {tot} warp instructions executed, {tott / tot:.2f} active lanes on average")
    for k in range(0, len(data), seg):
        s = data[k:k + seg]
        e, t, p = sum(d[1] for d in s), sum(d[2] for d in s), sum(d[3] for d in s)
        if e / tot > 0.002:
            print(f"{k:5d} inst={e / tot * 100:5.2f}% lanes={t / max(e, 1):5.1f} samples={p / tots * 100:5.2f}% exec={max(d[1] for d in s):11d}  {s[0][0][:60]}")


if __name__ == "__main__":
    mode = sys.argv[1]
    if mode == "launches":
        launches(sys.argv[2])
    elif mode == "full":
        full(sys.argv[2], "--traffic" in sys.argv)
    elif mode == "sass":
        sass(sys.argv[2], sys.argv[3], skip=int(sys.argv[4]) if len(sys.argv) > 4 else 0)
