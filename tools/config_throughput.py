#!/usr/bin/env python3
"""Device-resident throughput of the other BASELINE.json configurations (the bench line is config 2).

    python tools/config_throughput.py [instances]

C1: double-wishbone corner, 36-step bump sweep; C2: MacPherson corner, 41-step bump + steer;
C4: T-bar / torsion-bar / heave-link axle with a camber shim, 101-step bump and 101-step roll, shim
set-up thickness drawn per instance (the Monte-Carlo tolerance configuration); C5: C3 with metrics.
Synthetic inputs as SURVEY.md section 8(d): sigma 0.5 mm (0.25 mm for C4) on the authored hardpoints."""
import ctypes
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from helpers import build_case, load_golden  # noqa: E402
from open_kinematics_b200 import _lib  # noqa: E402
from open_kinematics_b200.core.enums import PointID as P  # noqa: E402
from open_kinematics_b200.core.input import build_sweep  # noqa: E402
from open_kinematics_b200.core.sweep import BatchSolver  # noqa: E402


def axle_sweep(steps: int, amp: float, roll: bool) -> dict:
    z = np.linspace(-amp, amp, steps)
    side = lambda name, v: {"point": "wheel_center", "side": name, "direction": {"axis": "z"}, "mode": "relative",  # noqa: E731
                            "values": [float(x) for x in v]}
    return {"version": 1, "targets": [side("left", z), side("right", -z if roll else z),
                                      {"point": "trackrod_inboard", "side": "left", "direction": {"axis": "y"},
                                       "mode": "relative", "values": [0.0] * steps}]}


def run(label, sus, sweep, n, sigma, metrics=False, shim_mc=False):
    solver = BatchSolver(sus, sweep, tune_layout=True)
    prog = solver.program
    nominal = solver.nominal_hardpoints()
    rng = np.random.default_rng(7)
    if getattr(sus, "is_axle", False):
        hp = bench.make_hardpoints_numpy(nominal, prog, n, 7)
        hp = nominal[None, :] + (hp - nominal[None, :]) * (sigma / bench.SIGMA_MM)
    else:
        pts = nominal.reshape(-1, 3)[None] + rng.normal(0.0, sigma, size=(n, nominal.size // 3, 3))
        if P.STRUT_BOTTOM in prog.in_keys and P.STRUT_TOP in prog.in_keys and "macpherson" in label:
            i_lbj, i_top, i_sb = (prog.in_keys.index(k) for k in (P.LOWER_WISHBONE_OUTBOARD, P.STRUT_TOP, P.STRUT_BOTTOM))
            nom = nominal.reshape(-1, 3)
            ax = nom[i_top] - nom[i_lbj]
            frac = float((nom[i_sb] - nom[i_lbj]) @ ax / (ax @ ax))
            pts[:, i_sb] = pts[:, i_lbj] + frac * (pts[:, i_top] - pts[:, i_lbj])
        hp = pts.reshape(n, -1)
    params = None
    if shim_mc and prog.param_names:
        params = np.repeat(prog.param_default[None, :], n, axis=0)
        col = [i for i, name in enumerate(prog.param_names) if name.endswith("setup_thickness")]
        params[:, col] = rng.uniform(29.5, 30.5, size=(n, len(col)))
    S = solver.values.shape[1]
    d = lambda a: torch.tensor(a, device="cuda")  # noqa: E731
    t_hp, t_tv = d(hp), d(solver.values)
    t_par = d(params) if params is not None else None
    pos = torch.empty((n, S, 3 * prog.n_out), device="cuda", dtype=torch.float64)
    st = torch.empty(n, device="cuda", dtype=torch.int32)
    fs = torch.empty_like(st)
    it = torch.empty((n, S), device="cuda", dtype=torch.int32)
    mr = torch.empty((n, S), device="cuda", dtype=torch.float64)
    met = torch.empty((n, S, len(prog.metric_names)), device="cuda", dtype=torch.float64) if metrics else None
    io = _lib.BatchIO.of(hardpoints=t_hp.data_ptr(), params=t_par.data_ptr() if t_par is not None else None,
                         target_values=t_tv.data_ptr(), positions=pos.data_ptr(), status=st.data_ptr(),
                         failed_step=fs.data_ptr(), iters=it.data_ptr(), max_residual=mr.data_ptr(),
                         metrics=met.data_ptr() if met is not None else None)
    lib, cfg = _lib.load(), _lib.default_cfg()

    def launch():
        _lib.check(lib.okin_solve_batch_device(solver.topology.handle, ctypes.byref(cfg), 0, None, n, S, ctypes.byref(io)), label)

    for _ in range(3):
        launch()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        launch()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    ok = st == 0
    rec = {"config": label, "instances": n, "steps": S, "n_unknowns": prog.n_unknowns, "metrics": bool(metrics),
           "states_per_s": n * S / (ms * 1e-3), "ms_per_launch": ms, "ok_fraction": float(ok.double().mean()),
           "mean_nfev": float(it[ok].double().mean()), "launch": solver.topology.launch_geometry(n),
           "lean_family": solver.topology.lean_calibration(0)}
    print(json.dumps(rec), flush=True)
    solver.close()
    return rec


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    print(json.dumps({"lib": _lib.LIB_PATH}), flush=True)
    meta, _ = load_golden("c3_rocker_ubar_coilover_roll")
    run("c3_axle_roll21_lean", *build_case(meta), 2 * n, 0.5)
    if os.environ.get("OKIN_C3_ONLY"):
        run("c5_c3_axle_roll21_metrics", *build_case(meta), n, 0.5, metrics=True)
        return
    meta, _ = load_golden("c1_dw_corner_bump")
    run("c1_dw_corner_bump36", *build_case(meta), n, 0.5)
    meta, _ = load_golden("c2_macpherson_bump_steer")
    run("c2_macpherson_bump_steer41", *build_case(meta), n, 0.5)
    if os.environ.get("OKIN_CORNERS_ONLY"):
        return
    meta, _ = load_golden("c4_tbar_heave_shim_bump")
    sus, _ = build_case(meta)
    run("c4_tbar_heave_shim_bump101_mc", sus, build_sweep(axle_sweep(101, 50.0, False), sus), n // 2, 0.25, shim_mc=True)
    run("c4_tbar_heave_shim_roll101_mc", sus, build_sweep(axle_sweep(101, 50.0, True), sus), n // 2, 0.25, shim_mc=True)
    meta, _ = load_golden("c3_rocker_ubar_coilover_roll")
    run("c5_c3_axle_roll21_metrics", *build_case(meta), n, 0.5, metrics=True)


if __name__ == "__main__":
    main()
