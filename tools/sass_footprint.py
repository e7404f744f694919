#!/usr/bin/env python3
"""Static code footprint of one kernel by source function.

    python tools/sass_footprint.py lib.so [kernel-name-substring]

Extracts the cubin, disassembles it with ``nvdisasm -g -c`` (needs -lineinfo) and counts SASS
instructions per source function of csrc/okin_core.cuh (16 bytes each): the instruction-cache
footprint that the per-step loop has to fit."""
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict


def main():
    lib = os.path.abspath(sys.argv[1])
    want = sys.argv[2] if len(sys.argv) > 2 else "okin_sweep_kernelILb0ELb0"
    here = os.path.dirname(os.path.abspath(__file__))
    core = open(os.path.join(here, "..", "open-kinematics_b200", "csrc", "okin_core.cuh")).read().split("\n")
    starts = []
    for i, line in enumerate(core, 1):
        m = re.match(r"OKIN_(?:FN|HD) \w[\w\s\*]*?\b(okin_\w+)\(", line)
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

    with tempfile.TemporaryDirectory() as tmp:
        subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, check=True, capture_output=True)
        cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
        dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    counts, cur, section, active = defaultdict(int), None, "", False
    for raw in dis.split("\n"):
        m = re.match(r"\s*\.section\s+(\S+)", raw)
        if m:
            section = m.group(1)
            active = want in section or (".text." in section and "okin_sweep_kernel" not in section and "kernel" not in section)
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', raw)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        if active and re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", raw):
            counts[(section.split(".text.")[-1][:40], region(cur))] += 1
    total = sum(counts.values())
    print(f"{total} instructions = {total * 16 / 1024:.0f} KiB")
    for (sec, fn), n in sorted(counts.items(), key=lambda kv: -kv[1])[:40]:
        print(f"{n:7d} {n * 16 / 1024:7.1f} KiB  {fn:28s} in {sec}")


if __name__ == "__main__":
    main()
