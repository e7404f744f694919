#!/usr/bin/env python3
"""Attribute an ncu SASS-level source page to CUDA source lines.

    ncu -i prof.ncu-rep --page source --csv > sass.csv
    cuobjdump -xelf all libokin.so && nvdisasm -g -c okin_abi.sm_100a.cubin > dis.txt
    python tools/ncu_lines.py sass.csv dis.txt [kernel-substring] [top-N]

Joins by instruction order within the kernel (ncu rows and nvdisasm rows are both in
address order) and prints instructions executed / stall samples per source line and per
source *function region* (consecutive line ranges given with --regions)."""
import csv
import re
import sys
from collections import defaultdict


def load_dis(path, kernel):
    lines, cur, active = [], None, False
    for raw in open(path, errors="replace"):
        if raw.startswith("//-----") and ".text." in raw:
            active = kernel in raw
            continue
        if not active:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', raw)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", raw)
        if m:
            lines.append((int(m.group(1), 16), cur, m.group(2).strip()))
    return lines


def main():
    sass_csv, dis = sys.argv[1], sys.argv[2]
    kernel = sys.argv[3] if len(sys.argv) > 3 else "okin_sweep"
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    rows = list(csv.reader(open(sass_csv)))
    h = next(i for i, r in enumerate(rows) if "Address" in r and "Source" in r)
    hdr = rows[h]
    ci, si, ai = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Address")
    ti = hdr.index("Thread Instructions Executed")
    data = [r for r in rows[h + 1:] if len(r) == len(hdr)]
    d = load_dis(dis, kernel)
    if len(d) != len(data):
        print(f"warning: {len(data)} profiled vs {len(d)} disassembled instructions", file=sys.stderr)
    by_line = defaultdict(lambda: [0, 0, 0])
    by_op = defaultdict(lambda: [0, 0])
    tot_i = tot_s = 0
    for r, (_, loc, text) in zip(data, d):
        n, s, t = int(r[ci]), int(r[si]), int(r[ti])
        by_line[loc][0] += n
        by_line[loc][1] += s
        by_line[loc][2] += t
        op = text.split()[0] if not text.startswith("@") else text.split()[1]
        by_op[op.split(".")[0]][0] += n
        by_op[op.split(".")[0]][1] += s
        tot_i += n
        tot_s += s
    print(f"total warp instructions {tot_i:,}  samples {tot_s:,}")
    print("--- by source line")
    for loc, (n, s, t) in sorted(by_line.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{100 * n / tot_i:6.2f}% inst {100 * s / max(tot_s, 1):6.2f}% smp  lanes {t / max(n, 1):5.1f}  {loc}")
    print("--- by opcode")
    for op, (n, s) in sorted(by_op.items(), key=lambda kv: -kv[1][0])[:25]:
        print(f"{100 * n / tot_i:6.2f}% inst {100 * s / max(tot_s, 1):6.2f}% smp  {op}")


if __name__ == "__main__":
    main()
