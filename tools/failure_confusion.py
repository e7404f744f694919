#!/usr/bin/env python3
"""Per-instance (status, failed_step) of the device core against the reference on the failing
batches (tests/golden/fail*.npz, generated from the reference by tests/golden/generate_batches.py):
confusion matrix + the contract check of tests/test_emu_core.py::check_flag_contract.

    python tools/failure_confusion.py [--gpu] [--out profiles/r02_failure_flag_confusion.json]

Default: the lane emulation of the device core (the exact kernel source, CPU); --gpu: the CUDA
library through the C ABI.
"""

from __future__ import annotations

import argparse
import collections
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

FAIL_BATCHES = ["fail256_c3", "fail256_c1", "fail128_c4_roll"]


def run_batch(name: str, solve) -> dict:
    from test_batches import failure_batch_report
    return failure_batch_report(name, solve)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpu", action="store_true")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    if args.gpu:
        from open_kinematics_b200 import _lib

        def solve(prog, hp, values, par):
            topo = _lib.DeviceTopology(prog)
            try:
                return topo.solve_batch(hp, values, params=par, want_design=True)
            finally:
                topo.close()
    else:
        from helpers import emu_solve

        def solve(prog, hp, values, par):
            return emu_solve(prog, hp, values, params=par)
    report = {"engine": "cuda" if args.gpu else "lane emulation of csrc/okin_core.cuh",
              "batches": {name: run_batch(name, solve) for name in FAIL_BATCHES}}
    text = json.dumps(report, indent=1)
    print(text)
    if args.out:
        open(args.out, "w").write(text + "\n")


if __name__ == "__main__":
    main()
