#!/usr/bin/env python3
"""Small batches through every output path of the C ABI, for compute-sanitizer.

    compute-sanitizer --tool memcheck|racecheck|synccheck python tools/sanitizer_batch.py

Three topologies (rocker / U-bar axle, shimmed T-bar axle, MacPherson corner), 40 perturbed instances x 8
sweep steps each: once positions-only (lean kernel family; OKIN_LEAN_REGS=128|168 selects which) and once
with every optional output (full kernel + continuity kernel)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import build_case, load_golden  # noqa: E402
from test_emu_core import _nominal  # noqa: E402
from open_kinematics_b200 import _lib  # noqa: E402
from open_kinematics_b200.core.topology import compile_suspension  # noqa: E402

for case in ("c3_rocker_ubar_coilover_roll", "c4_tbar_heave_shim_bump", "c2_macpherson_bump_steer"):
    meta, arr = load_golden(case)
    sus, sweep = build_case(meta)
    prog = compile_suspension(sus, sweep)
    rng = np.random.default_rng(0)
    hp = np.repeat(_nominal(sus, prog), 40, axis=0)
    if not case.startswith("c2"):       # the MacPherson strut axis must stay consistent: nominal instances
        hp = hp + rng.normal(0, 0.2, size=hp.shape)
    hp[3] += 400.0                       # one instance that fails (residual rejection / invalid geometry paths)
    topo = _lib.DeviceTopology(prog)
    v = arr["sweep_values"][:, :8]
    lean = topo.solve_batch(hp, v, want_worst_row=True)
    full = topo.solve_batch(hp, v, want_metrics=True, want_velocities=True, want_health=True,
                            want_diagnostics=True, want_tangents=True, want_design=True, want_worst_row=True)
    print(case, "ok fraction lean %.3f full %.3f" % ((lean["status"] == 0).mean(), (full["status"] == 0).mean()), flush=True)
    topo.close()
