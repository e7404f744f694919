#!/usr/bin/env python3
"""Per-phase (source function) instruction and stall-sample shares of the sweep kernel.

    python tools/ncu_phases.py sass.csv dis.txt n_states [kernel-section-substring]

``sass.csv``: ``ncu -i rep --page source --csv``; ``dis.txt``: ``nvdisasm -g -c`` of the cubin.
Out-of-line device functions are separate .text sections in the disassembly and separate
address ranges in the ncu page; both lists are in the same order, so rows are joined by index
and attributed to the function whose source range contains the line."""
import csv
import re
import sys
from collections import defaultdict


def main():
    sass, dis, states = sys.argv[1], sys.argv[2], float(sys.argv[3])
    want = sys.argv[4] if len(sys.argv) > 4 else "okin_sweep_kernelILb0ELb0"
    core = open(__file__.replace("tools/ncu_phases.py", "open-kinematics_b200/csrc/okin_core.cuh")).read().split("\n")
    starts = []
    for i, line in enumerate(core, 1):
        m = re.match(r"OKIN_(?:FN_HOT|FN|HD) \w[\w\s\*]*?\b(okin_\w+)\(", line)
        if m:
            starts.append((i, m.group(1)))

    def region(loc):
        if loc is None:
            return "?"
        if loc[0] != "okin_core.cuh":
            return loc[0]
        name = "okin_core.cuh(top)"
        for line, fn in starts:
            if line <= loc[1]:
                name = fn
        return name

    locs, cur, active = [], None, False
    for raw in open(dis, errors="replace"):
        m = re.match(r"\s*\.section\s+(\S+)", raw)
        if m:   # one .text section per kernel instantiation: keep the profiled one
            active = m.group(1).startswith(".text.") and want in m.group(1)
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', raw)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        if active and re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", raw):
            locs.append(cur)
    rows = list(csv.reader(open(sass)))
    agg = defaultdict(lambda: [0, 0, 0, 0.0])
    total = [0, 0]
    kernels = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    h = next(i for i, r in enumerate(rows) if "Address" in r and "Source" in r)
    hdr = rows[h]
    ci, si, ti = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
    wi = hdr.index("L1 Wavefronts Shared") if "L1 Wavefronts Shared" in hdr else None
    data = [r for r in rows[h + 1:] if len(r) == len(hdr)]
    if len(locs) != len(data):
        print(f"warning: {len(locs)} disassembled instructions vs {len(data)} profiled rows", file=sys.stderr)
    for r, loc in zip(data, locs):
        n, s, t = int(r[ci]), int(r[si]), int(r[ti])
        a = agg[region(loc)]
        a[0] += n; a[1] += s; a[2] += t
        if wi is not None and r[wi]:
            a[3] += float(r[wi])
        total[0] += n; total[1] += s
    print(f"warp instructions {total[0]:,} = {total[0] / states:,.0f} per state; samples {total[1]:,}")
    for k, (n, s, t, w) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"{k:26s} {100 * n / total[0]:6.2f}% inst {100 * s / max(total[1], 1):6.2f}% smp "
              f"{n / states:8.0f} inst/state  lanes {t / max(n, 1):5.1f}  smem wavefronts/state {w / states:7.0f}")


if __name__ == "__main__":
    main()
