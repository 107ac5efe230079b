#!/usr/bin/env python
"""Warp instructions of k_trace per unit of its own scheduling counters (run HERE on files brought back from a gpurun call).

    LMB_STATS_PER_LAUNCH=1 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:'^k_trace$' \
        --csv --log-file gpurun_out/ktrace_inst.csv python tools/ncu_target.py 16 1 2> gpurun_out/ktrace_counters.log
    python tools/calibrate_ktrace.py gpurun_out/ktrace_inst.csv gpurun_out/ktrace_counters.log [more pairs ...]

ncu gives smsp__inst_executed.sum for every k_trace launch; the library prints the launch's own counters (traversal-loop trips of all
warps: lmb_stats.trace_warp_iters, and the rays traced). A non-negative least-squares fit over the launches gives the warp
instructions per loop trip (node step + its share of triangle rounds, pops and stores) and per ray (fetch, preparation, refill);
bench.py multiplies the live counters of its run by them to get the instructions issued -- the numerator of the issue roofline.
(Triangle rounds and refills were counted too at first: across launches they are collinear with trips and rays, their coefficients
are not identifiable, and the two extra integer adds per trip cost 0.9 ms of a 60 ms trace stage -- dropped.) Writes profiles/ktrace_calibration.json stamped with the source hash.
"""
import csv
import json
import os
import re
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from srchash import ROOT, csrc_sha  # noqa: E402


def ncu_inst(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, mi, vi, idi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    per = {}
    for r in rows[1:]:
        if "k_trace" not in r[ki] or "k_trace_array" in r[ki]:
            continue
        per.setdefault(int(r[idi]), {})[r[mi]] = float(r[vi].replace(",", ""))
    return [per[k] for k in sorted(per)]


def counters(path):
    out = []
    for line in open(path, errors="replace"):
        m = re.match(r"k_trace_launch (\{.*\})", line.strip())
        if m:
            out.append(json.loads(m.group(1)))
    return out


def main():
    pairs = list(zip(sys.argv[1::2], sys.argv[2::2]))
    A, y, names = [], [], ("iters", "rays")
    for inst_csv, log in pairs:
        n, c = ncu_inst(inst_csv), counters(log)
        if len(n) != len(c):
            raise SystemExit(f"{inst_csv}: {len(n)} k_trace launches in the ncu list, {len(c)} counter lines in {log}")
        for a, b in zip(n, c):
            if b["rays"] == 0:
                continue
            A.append([b[k] for k in names])
            y.append(a["smsp__inst_executed.sum"])
    A, y = np.array(A, dtype=np.float64), np.array(y, dtype=np.float64)
    # non-negative least squares by active-set elimination (4 unknowns): drop a column while its coefficient is negative
    cols = list(range(len(names)))
    while True:
        coef, *_ = np.linalg.lstsq(A[:, cols], y, rcond=None)
        if (coef >= 0).all():
            break
        cols.pop(int(np.argmin(coef)))
    full = np.zeros(len(names))
    full[cols] = coef
    pred = A @ full
    rel = np.abs(pred - y) / y
    out = {"warp_inst_per": dict(zip(names, [float(v) for v in full])), "launches_fitted": int(len(y)), "max_rel_residual": float(rel.max()),
           "mean_rel_residual": float(rel.mean()), "total_inst_measured": float(y.sum()), "total_inst_predicted": float(pred.sum()),
           "how": "least squares of ncu smsp__inst_executed.sum per k_trace launch on the launch's own counters (tools/calibrate_ktrace.py)",
           "csrc_sha": csrc_sha()}
    json.dump(out, open(os.path.join(ROOT, "profiles", "ktrace_calibration.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
