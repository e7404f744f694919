#!/usr/bin/env python3
"""Large reference batches (round 2): 256 hardpoint-perturbed instances per BASELINE topology,
solved by the *reference itself* (nickmccleery/open-kinematics at /root/reference), one process
per host core.

    PYTHONPATH=/root/reference/src python tests/golden/generate_batches.py [name ...]

Writes ``tests/golden/<name>.npz`` + ``.json``.  Compact format (the 6-instance ``batch_c*`` files of
round 1 keep every perturbed geometry as a dict; at 256 instances that is megabytes of JSON):

  hp_names          json   ["<block>/<point>", ...]  block = "" (corner) | left | right | center
  hardpoints        [N, len(hp_names), 3]  perturbed authored hardpoints
  shim_setup        [N]    per-instance camber-shim setup thickness (configs[3] batches only)
  steps             [K]    sweep steps whose states are stored (first, middle, last)
  positions_tight   [N, K, P, 3]  reference solve with ftol=xtol=gtol=1e-15 (NaN = not solved)
  status            [N]    0 ok | 1 "Solver failed to converge" | 2 residual rejection  (solver.py:726-747),
  failed_step       [N]    for the reference run at DEFAULT tolerances (the flags a user sees)
  status_tight / failed_step_tight   same for the tight run that produced positions_tight
  nfev_default      [N, S] SolverInfo.nfev of the default-tolerance run (0 from the failed step on)
  metrics           [N, K, M]  compute_sweep_metrics rows of the tight states (metric batches only)

The per-step loop below is solve_suspension_sweep's own (solver.py:654-776) driven through the
reference's ResidualComputer / solve_least_squares_problem, so that a failing step index is known
without bisecting.  Nothing here imports the product package or the oracle.
"""

from __future__ import annotations

import copy
import json
import multiprocessing as mp
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import generate_golden as G  # noqa: E402  (also puts /root/reference/src on sys.path)

from kinematics.core.input import build_suspension, build_sweep  # noqa: E402
from kinematics.core.points.derived.manager import DerivedPointsManager  # noqa: E402
from kinematics.core.solver import (  # noqa: E402
    ResidualComputer, SolverConfig, convert_targets_to_absolute, solve_least_squares_problem,
)
from kinematics.core.sweep import compute_sweep_metrics  # noqa: E402


def hp_blocks(geom: dict) -> list:
    hp = geom["hardpoints"]
    if "left" not in hp:
        return [("", hp)]
    return [(side, hp[side]) for side in ("left", "right", "center") if hp.get(side)]


def hp_names(geom: dict) -> list:
    return [f"{block}/{name}" for block, pts in hp_blocks(geom) for name in sorted(pts)]


def hp_array(geom: dict) -> np.ndarray:
    return np.array([[float(pts[name][a]) for a in "xyz"] for _, pts in hp_blocks(geom) for name in sorted(pts)])


def reference_sweep(geom: dict, sweep: dict, config: SolverConfig):
    """(states, status, failed_step, nfev[S]) following solver.py:716-774 step by step."""
    sus = build_suspension(geom)
    cfg = build_sweep(sweep, sus)
    init = sus.initial_state()
    targets = [convert_targets_to_absolute([sw[i] for sw in cfg.target_sweeps], init) for i in range(cfg.n_steps)]
    ws = init.copy()
    dm = DerivedPointsManager(sus.derived_spec())
    rc = ResidualComputer(constraints=sus.constraints(), derived_manager=dm, state_buffer=ws,
                          n_target_variables=len(cfg.target_sweeps))
    x0 = ws.get_free_array()
    states, nfev = [], np.zeros(cfg.n_steps, np.int32)
    for s, step_targets in enumerate(targets):
        res = solve_least_squares_problem(residual_function=rc.compute, x_0=x0, args=(step_targets,),
                                          solver_config=config, n_residuals=rc.n_residuals,
                                          jacobian_function=rc.compute_jacobian)
        if not res.success:
            return sus, cfg, states, 1, s, nfev
        if float(np.max(np.abs(res.fun))) > config.residual_tolerance:
            return sus, cfg, states, 2, s, nfev
        nfev[s] = res.nfev
        ws.update_from_array(res.x)
        dm.update_in_place(ws.positions)
        states.append(ws.copy())
        x0 = res.x
    return sus, cfg, states, 0, -1, nfev


def solve_instance(job):
    geom, sweep, want_metrics, flags_only = job
    out = {}
    try:
        _, _, _, st, fs, nfev = reference_sweep(geom, sweep, SolverConfig())
    except Exception as err:  # noqa: BLE001 - invalid perturbed geometry (build-time validation)
        return {"status": 3, "failed_step": 0, "error": f"{type(err).__name__}: {err}"[:200]}
    out.update(status=st, failed_step=fs, nfev_default=nfev)
    if flags_only:
        return out
    sus, cfg, states, st_t, fs_t, _ = reference_sweep(geom, sweep, G.TIGHT)
    keys = sorted(sus.initial_state().positions)
    S = cfg.n_steps
    steps = sorted({0, S // 2, S - 1})
    pos = np.full((len(steps), len(keys), 3), np.nan)
    for k, s in enumerate(steps):
        if s < len(states):
            pos[k] = G.positions_array([states[s]], keys)[0]
    out.update(status_tight=st_t, failed_step_tight=fs_t, positions=pos, steps=steps,
               point_keys=[G.key_name(k) for k in keys])
    if want_metrics and st_t == 0:
        result = compute_sweep_metrics(sus, cfg, states)
        rows = [row.flat_row() if hasattr(row, "flat_row") else row for row in result.rows]
        names = list(rows[0].keys())
        out["metric_names"] = names
        out["metrics"] = np.array([[np.nan if rows[s][k] is None else float(rows[s][k]) for k in names] for s in steps])
    return out


def run(name: str, geom: dict, sweep: dict, n_inst: int, sigma: float, seed: int, *, shim_sigma: float = 0.0,
        want_metrics: bool = False, flags_only: bool = False) -> None:
    rng = np.random.default_rng(seed)
    geoms, shim = [], []
    for _ in range(n_inst):
        gi = G.perturb_geometry(geom, rng, sigma)
        # centre-line points (T-bar pivot) must keep y = 0 (axle/mechanisms.py:626-643)
        for point, pt in (gi["hardpoints"].get("center") or {}).items() if "left" in gi["hardpoints"] else ():
            if abs(float(geom["hardpoints"]["center"][point]["y"])) <= 1e-9:
                pt["y"] = float(geom["hardpoints"]["center"][point]["y"])
        if shim_sigma:
            t = float(gi["axle_config"]["left_setup"]["camber_shim"]["setup_thickness"]) + float(rng.normal(0, shim_sigma))
            gi["axle_config"]["left_setup"]["camber_shim"]["setup_thickness"] = t
            shim.append(t)
        geoms.append(gi)
    with mp.Pool(os.cpu_count()) as pool:
        recs = pool.map(solve_instance, [(g, sweep, want_metrics, flags_only) for g in geoms], chunksize=1)
    arrays = {
        "hardpoints": np.array([hp_array(g) for g in geoms]),
        "status": np.array([r["status"] for r in recs], np.int32),
        "failed_step": np.array([r["failed_step"] for r in recs], np.int32),
    }
    S = max([len(r["nfev_default"]) for r in recs if "nfev_default" in r] or [0])
    arrays["nfev_default"] = np.array([r.get("nfev_default", np.zeros(S, np.int32)) for r in recs])
    meta = {"geometry": geom, "sweep": sweep, "hp_names": hp_names(geom), "sigma": sigma, "seed": seed,
            "n_instances": n_inst, "errors": {i: r["error"] for i, r in enumerate(recs) if "error" in r}}
    if shim:
        arrays["shim_setup"] = np.array(shim)
    if not flags_only:
        first = next(r for r in recs if "positions" in r)
        meta["point_keys"] = first["point_keys"]
        arrays["steps"] = np.array(first["steps"], np.int32)
        blank = np.full_like(first["positions"], np.nan)
        arrays["positions_tight"] = np.array([r.get("positions", blank) for r in recs])
        arrays["status_tight"] = np.array([r.get("status_tight", 3) for r in recs], np.int32)
        arrays["failed_step_tight"] = np.array([r.get("failed_step_tight", 0) for r in recs], np.int32)
        if want_metrics:
            firstm = next(r for r in recs if "metrics" in r)
            meta["metric_names"] = firstm["metric_names"]
            blankm = np.full_like(firstm["metrics"], np.nan)
            arrays["metrics"] = np.array([r.get("metrics", blankm) for r in recs])
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrays)
    json.dump(meta, open(os.path.join(HERE, name + ".json"), "w"))
    st = arrays["status"]
    print(f"{name}: {n_inst} instances, sigma {sigma}: ok {int((st == 0).sum())}, not converged {int((st == 1).sum())}, "
          f"residual rejected {int((st == 2).sum())}, invalid {int((st == 3).sum())}", flush=True)


JOBS = {
    # sigma = 0.5 mm on every authored hardpoint (the bench's perturbation): parity of positions
    "batch256_c1": lambda: run("batch256_c1", *G.CASES["c1_dw_corner_bump"], 256, 0.5, 11),
    "batch256_c2": lambda: run("batch256_c2", *G.CASES["c2_macpherson_bump_steer"], 256, 0.5, 12),
    # flagship + configs[4]: the 256 instances carry the reference's metric rows too
    "batch256_c3": lambda: run("batch256_c3", *G.CASES["c3_rocker_ubar_coilover_roll"], 256, 0.5, 13, want_metrics=True),
    # configs[3]: per-instance shim thickness (sigma 0.5 mm around the authored setup thickness)
    "batch64_c4_roll": lambda: run("batch64_c4_roll", *G.CASES["c4_tbar_heave_shim_roll"], 64, 0.5, 14, shim_sigma=0.5),
    "batch64_c4_bump": lambda: run("batch64_c4_bump", *G.CASES["c4_tbar_heave_shim_bump"], 64, 0.5, 15, shim_sigma=0.5),
    # heterogeneous batches with failing instances (sigma = 10 mm): per-instance flags
    "fail256_c3": lambda: run("fail256_c3", *G.CASES["c3_rocker_ubar_coilover_roll"], 256, 10.0, 21, flags_only=True),
    "fail256_c1": lambda: run("fail256_c1", G.CASES["c1_dw_corner_bump"][0], G.bump_sweep(41, 0.0, 300.0), 256, 10.0, 22,
                              flags_only=True),
    "fail128_c4_roll": lambda: run("fail128_c4_roll", *G.CASES["c4_tbar_heave_shim_roll"], 128, 10.0, 23, flags_only=True),
}


if __name__ == "__main__":
    only = sys.argv[1:] or list(JOBS)
    for name in only:
        JOBS[name]()
