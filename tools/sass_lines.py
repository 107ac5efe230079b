#!/usr/bin/env python
"""Where a kernel's warp instructions and stall samples go, by SOURCE LINE (run HERE on an ncu report brought back from gpurun).

    python tools/sass_lines.py gpurun_out/prof_shade.ncu-rep k_shade lumen_b200/csrc/wavefront.o [launch_skip] [--outer]

(the kernel regex is ncu's; launch_skip picks among the captured launches it matches; --outer attributes inlined code to the line of the
outermost frame, i.e. of the kernel body or noinline function it was inlined into)

ncu's source page gives executed counts and stall samples per SASS instruction (in CSV form without the line correlation); nvdisasm
--print-line-info-inline on the cubin of the same build gives the source line (innermost inlined frame first) of every SASS instruction.
The two listings are joined by instruction order inside the kernel (checked: same opcode sequence). Output: lines sorted by share of
executed warp instructions, with their share of stall samples.
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile


def ncu_sass(rep, kernel, skip):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", f"regex:{kernel}", "--launch-skip", str(skip),
                          "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    name = rows[0][1]
    hdr = rows[1]
    ie, sm = hdr.index("Instructions Executed"), hdr.index("# Samples")
    data = [(r[1].strip(), int(r[ie]), int(r[sm])) for r in rows[2:] if len(r) > ie and r[ie].isdigit()]
    half = len(data) // 2
    if half and data[0][0] == data[half][0]:
        data = data[:half]
    return name, data


def cubin_lines(obj, mangled_part):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
    cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "--print-line-info-inline", cubin], capture_output=True, text=True).stdout
    out, cur, active, frames = [], None, False, []
    for line in txt.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", line) or re.match(r"\s*\.section\s+\.text\.(\S+),", line)
        if m:
            active = mangled_part in m.group(1)
            continue
        if not active:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', line)
        if m:
            inl = re.findall(r'inlined at "([^"]+)", line (\d+)', m.group(3))
            frames = [(os.path.basename(m.group(1)), int(m.group(2)))] + [(os.path.basename(f), int(n)) for f, n in inl]
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
        if m:
            out.append((m.group(1).strip(), list(frames)))
    return out


def main():
    rep, kernel, obj = sys.argv[1:4]
    skip = int(sys.argv[4]) if len(sys.argv) > 4 and sys.argv[4].isdigit() else 0
    name, prof = ncu_sass(rep, kernel, skip)
    m = re.search(r"(k_\w+)<?\(?(?:unsigned int\))?(\d+)?", name)
    part = m.group(1)
    # mangled name fragment: k_shade<1, 0> -> k_shadeILj1ELb0
    if "k_shade" in name:
        t = re.search(r"k_shade<\(unsigned int\)(\d+), \(bool\)(\d)>", name) or re.search(r"k_shade<(\d+), (\d)>", name)
        part = f"k_shadeILj{t.group(1)}ELb{t.group(2)}"
    lines = cubin_lines(obj, part)
    if len(lines) != len(prof):
        print(f"warning: {len(prof)} SASS instructions in the report, {len(lines)} in the cubin ({part}): joined up to the shorter", file=sys.stderr)
    n = min(len(lines), len(prof))
    mism = sum(1 for i in range(n) if prof[i][0].split()[0].lstrip("@!PU0123456789T ") [:4] != lines[i][0].split()[0].lstrip("@!PU0123456789T ")[:4])
    tot_e, tot_s = sum(p[1] for p in prof) or 1, sum(p[2] for p in prof) or 1
    outer = "--outer" in sys.argv
    agg = collections.defaultdict(lambda: [0, 0])
    for i in range(n):
        fr = lines[i][1]
        if not fr:
            key = "?"
        elif outer:
            key = f"{fr[-1][0]}:{fr[-1][1]}"
        else:
            key = f"{fr[0][0]}:{fr[0][1]}"
        agg[key][0] += prof[i][1]
        agg[key][1] += prof[i][2]
    print(f"{name[:80]}: {len(prof)} SASS, {tot_e} warp instructions, opcode mismatches in the join: {mism}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:45]:
        print(f"{v[0] / tot_e * 100:6.2f}% inst {v[1] / tot_s * 100:6.2f}% samples  {k}")


if __name__ == "__main__":
    main()
