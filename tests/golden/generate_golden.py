#!/usr/bin/env python3
"""Generate golden vectors by running the *reference itself* (nickmccleery/open-kinematics,
mounted read-only at /root/reference in the build container).

    PYTHONPATH=/root/reference/src python tests/golden/generate_golden.py

Writes ``tests/golden/<case>.json`` (inputs + structure) and ``<case>.npz`` (arrays).  The GPU
box has no /root/reference, so tests only ever read these files.  Nothing here imports the
product package or the oracle.

Per sweep case:
  positions_tight    reference solve with SolverConfig(ftol=xtol=gtol=1e-15)   [S, P, 3]
  positions_default  reference solve with default tolerances                    [S, P, 3]
  nfev_*, max_residual_*                                                        [S]
  tangents           compute_state_tangents at the tight states                 [S, T, n]
  velocities         TangentField.velocities of every point (sorted keys)    [S, T, P, 3]
  tangent_rank / tangent_sigma_min / tangent_cond                               [S]
Per family: residual + jacobian rows at seeded random points (core/constraints.py, core/jacobians.py).
Diagnostics: diagnose_sweep issues, continuity thresholds and U-bar chirality / transmission
quantities on sweeps with an uneven step (diagnostics.json / .npz).
Failure cases: first failed step and failure class for out-of-reach sweeps.
"""

from __future__ import annotations

import copy
import json
import os
import sys

import numpy as np
import yaml

REF = "/root/reference"
sys.path.insert(0, os.path.join(REF, "src"))

from kinematics.core import constraints as C  # noqa: E402
from kinematics.core import jacobians as JAC  # noqa: E402
from kinematics.core.enums import Axis, PointID  # noqa: E402
from kinematics.core.input import build_suspension, build_sweep  # noqa: E402
from kinematics.core.points.derived.manager import DerivedPointsManager  # noqa: E402
from kinematics.core.primitives.geometry import Direction3, Point3  # noqa: E402
from kinematics.core.sensitivity import compute_state_tangents  # noqa: E402
from kinematics.core.sweep import compute_sweep_metrics  # noqa: E402
from kinematics.core.solver import (  # noqa: E402
    ResidualComputer, SolverConfig, convert_targets_to_absolute, solve_suspension_sweep,
)

OUT = os.path.dirname(os.path.abspath(__file__))
TIGHT = SolverConfig(ftol=1e-15, xtol=1e-15, gtol=1e-15)


def key_name(k) -> str:
    return k.name.lower()


def load(path):
    return yaml.safe_load(open(os.path.join(REF, path)))


def add_coilovers(geom: dict) -> dict:
    """BASELINE config 3: the shipped rocker/U-bar axle with coilovers."""
    g = copy.deepcopy(geom)
    g["axle_config"]["spring"] = {"type": "coilover"}
    g["hardpoints"]["left"]["strut_top"] = {"x": 0, "y": 200, "z": 600}
    g["hardpoints"]["left"]["strut_bottom"] = {"x": 0, "y": 330, "z": 520}
    return g


def roll_sweep(steps: int, amp: float) -> dict:
    return {"version": 1, "steps": steps, "targets": [
        {"point": "wheel_center", "side": "left", "direction": {"axis": "z"}, "mode": "relative", "start": -amp, "stop": amp},
        {"point": "wheel_center", "side": "right", "direction": {"axis": "z"}, "mode": "relative", "start": amp, "stop": -amp},
        {"point": "trackrod_inboard", "side": "left", "direction": {"axis": "y"}, "mode": "relative", "start": 0, "stop": 0},
    ]}


def bump_sweep(steps: int, lo: float, hi: float) -> dict:
    return {"version": 1, "steps": steps, "targets": [
        {"point": "trackrod_inboard", "direction": {"axis": "y"}, "mode": "relative", "start": 0, "stop": 0},
        {"point": "wheel_center", "direction": {"axis": "z"}, "mode": "relative", "start": lo, "stop": hi}]}


def with_shim(geom: dict, setup: float) -> dict:
    g = copy.deepcopy(geom)
    g["config"]["camber_shim"]["setup_thickness"] = setup
    return g


def c4_full(geom: dict, setup: float) -> dict:
    """BASELINE config 4: T-bar + torsion bars + rocker-to-rocker heave link + camber shim
    (SURVEY.md section 8d)."""
    g = copy.deepcopy(geom)
    g["axle_config"]["heave_link"] = {"type": "rocker_to_rocker"}
    g["hardpoints"]["left"]["heave_link_rocker"] = {"x": 0, "y": 300, "z": 400}
    g["axle_config"]["left_setup"] = {"camber_shim": {
        "shim_face_point_a": {"x": -25.0, "y": 750.0, "z": 510.0},
        "shim_face_point_b": {"x": -25.0, "y": 750.0, "z": 490.0},
        "shim_face_normal": {"x": 0.0, "y": 1.0, "z": 0.0},
        "design_thickness": 30.0, "setup_thickness": setup}}
    return g


CASES = {
    # BASELINE.json configs[0]
    "c1_dw_corner_bump": (load("tests/data/geometry.yaml"), load("scripts/bump_sweep.yaml")),
    # e2e golden sweep of the reference (tests/data/sweep.yaml: bump + steer)
    "c1_dw_corner_bump_steer": (load("tests/data/geometry.yaml"), load("tests/data/sweep.yaml")),
    # configs[1] topology
    "c2_macpherson_bump_steer": (load("tests/data/macpherson_geometry.yaml"), load("tests/data/sweep.yaml")),
    # configs[2] as shipped and with coilovers (flagship)
    "c3_rocker_ubar_roll_shipped": (load("tests/data/axle_geometry_rocker.yaml"), load("tests/data/axle_rocker_sweep.yaml")),
    "c3_rocker_ubar_coilover_roll": (add_coilovers(load("tests/data/axle_geometry_rocker.yaml")), roll_sweep(21, 20.0)),
    # configs[3] topology family
    "c4_tbar_roll": (load("tests/data/axle_geometry_t_bar.yaml"), load("tests/data/axle_t_bar_roll_sweep.yaml")),
    "c4_tbar_bump": (load("tests/data/axle_geometry_t_bar.yaml"), load("tests/data/axle_t_bar_bump_sweep.yaml")),
    "dw_axle_direct": (load("tests/data/axle_geometry.yaml"), load("tests/data/axle_sweep.yaml")),
    "macpherson_axle": (load("tests/data/macpherson_axle_geometry.yaml"), load("tests/data/axle_sweep.yaml")),
    "dw_corner_coilover_direct": (load("tests/data/corner_strut_geometry.yaml"), load("scripts/bump_sweep.yaml")),
    "dw_corner_rocker": (load("tests/data/corner_rocker_geometry.yaml"), bump_sweep(21, -40.0, 40.0)),
    # camber-shim pre-solve (config/shims.py): 2 mm thicker shim on the shipped corner, upright-mounted
    # pushrod corner (rocker coupling), and BASELINE configs[3] in full
    "c1_shim_plus2mm": (with_shim(load("tests/data/geometry.yaml"), 32.0), load("scripts/bump_sweep.yaml")),
    "c4_tbar_heave_shim_roll": (c4_full(load("tests/data/axle_geometry_t_bar.yaml"), 31.0),
                                load("tests/data/axle_t_bar_roll_sweep.yaml")),
    "c4_tbar_heave_shim_bump": (c4_full(load("tests/data/axle_geometry_t_bar.yaml"), 29.25),
                                load("tests/data/axle_t_bar_bump_sweep.yaml")),
}


def describe_constraints(cons) -> list:
    out = []
    for c in cons:
        rec = {"type": type(c).__name__, "points": [key_name(getattr(c, a)) for a in c._POINT_ATTRS]}
        for attr in ("target_distance", "target_angle", "target_volume", "scale"):
            if hasattr(c, attr):
                rec[attr] = float(getattr(c, attr))
        if isinstance(c, C.PointOnLineConstraint):
            rec["line_point"] = c.line_point.data.tolist()
            rec["line_direction"] = c.line_direction.data.tolist()
        if isinstance(c, C.MidpointOnPlaneConstraint):
            rec["plane_point"] = c.plane_point.data.tolist()
            rec["plane_normal"] = c.plane_normal.data.tolist()
        out.append(rec)
    return out


def positions_array(states, keys) -> np.ndarray:
    return np.array([[st.positions[k].data for k in keys] for st in states])


def run_case(name: str, geom: dict, sweep: dict, with_default=True) -> None:
    sus = build_suspension(geom)
    cfg = build_sweep(sweep, sus)
    init = sus.initial_state()
    cons = sus.constraints()
    keys = sorted(init.positions)
    arrays, meta = {}, {}
    dm = DerivedPointsManager(sus.derived_spec())
    states_t, stats_t = solve_suspension_sweep(init, cons, cfg, dm, TIGHT)
    arrays["positions_tight"] = positions_array(states_t, keys)
    arrays["nfev_tight"] = np.array([s.nfev for s in stats_t])
    arrays["max_residual_tight"] = np.array([s.max_residual for s in stats_t])
    if with_default:
        states_d, stats_d = solve_suspension_sweep(init, cons, cfg, DerivedPointsManager(sus.derived_spec()))
        arrays["positions_default"] = positions_array(states_d, keys)
        arrays["nfev_default"] = np.array([s.nfev for s in stats_d])
        arrays["max_residual_default"] = np.array([s.max_residual for s in stats_d])
    tang, rank, smin, cond, vel = [], [], [], [], []
    for step, st in enumerate(states_t):
        targets = convert_targets_to_absolute([sw[step] for sw in cfg.target_sweeps], init)
        fields, info = compute_state_tangents(st, cons, dm, targets)
        free = st.free_points_order
        tang.append([np.concatenate([f.velocities[k] for k in free]) for f in fields])
        vel.append([[f.velocity(k) for k in keys] for f in fields])
        rank.append(info.rank)
        smin.append(info.smallest_singular_value)
        cond.append(info.condition_number)
    arrays["tangents"] = np.array(tang)
    arrays["velocities"] = np.array(vel)
    arrays["tangent_rank"] = np.array(rank)
    arrays["tangent_sigma_min"] = np.array(smin)
    arrays["tangent_cond"] = np.array(cond)
    # metric rows (state + derivative metrics) evaluated by the reference on its tight states
    result = compute_sweep_metrics(sus, cfg, states_t)
    assert result.derivative_error is None, result.derivative_error
    flat_rows = [row.flat_row() if hasattr(row, "flat_row") else row for row in result.rows]
    names = list(flat_rows[0].keys())
    assert all(list(r.keys()) == names for r in flat_rows)
    arrays["metrics"] = np.array([[np.nan if r[k] is None else float(r[k]) for k in names] for r in flat_rows])
    meta["metric_names"] = names
    arrays["design_positions"] = np.array([init.positions[k].data for k in keys])
    arrays["sweep_values"] = np.array([[t.value for t in sw] for sw in cfg.target_sweeps])
    meta.update(
        geometry=geom, sweep=sweep, point_keys=[key_name(k) for k in keys],
        free_order=[key_name(k) for k in init.free_points_order],
        output_points=[key_name(k) for k in sus.output_points()],
        derived=[key_name(k) for k in dm.update_order],
        constraints=describe_constraints(cons),
        targets=[{"point": key_name(sw[0].point_id), "mode": str(sw[0].mode.value)} for sw in cfg.target_sweeps],
        n_unknowns=3 * len(init.free_points_order), n_residuals=len(cons) + len(cfg.target_sweeps),
    )
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrays)
    json.dump(meta, open(os.path.join(OUT, name + ".json"), "w"), indent=1)
    print(f"{name}: S={cfg.n_steps} P={len(keys)} n={meta['n_unknowns']} m={meta['n_residuals']} "
          f"default-vs-tight {np.abs(arrays.get('positions_default', arrays['positions_tight']) - arrays['positions_tight']).max():.2e}")


# --------------------------------------------------------------------------- perturbed batches
def perturb_geometry(geom: dict, rng: np.random.Generator, sigma: float) -> dict:
    """Gaussian perturbation of every authored hardpoint coordinate (left + centre for mirrored
    axles).  MacPherson: strut_bottom is re-projected on the ball-joint -> strut-top line at its
    nominal axial fraction."""
    g = copy.deepcopy(geom)
    blocks = [g["hardpoints"]] if "left" not in g["hardpoints"] else [
        g["hardpoints"][side] for side in ("left", "right", "center") if g["hardpoints"].get(side)]
    for block in blocks:
        nominal = copy.deepcopy(block)
        for name in sorted(block):
            for ax in "xyz":
                block[name][ax] = float(block[name][ax]) + float(rng.normal(0.0, sigma))
        if g["type"] == "macpherson":
            lbj0 = np.array([nominal["lower_wishbone_outboard"][a] for a in "xyz"], float)
            top0 = np.array([nominal["strut_top"][a] for a in "xyz"], float)
            sb0 = np.array([nominal["strut_bottom"][a] for a in "xyz"], float)
            frac = float((sb0 - lbj0) @ (top0 - lbj0) / ((top0 - lbj0) @ (top0 - lbj0)))
            lbj = np.array([block["lower_wishbone_outboard"][a] for a in "xyz"], float)
            top = np.array([block["strut_top"][a] for a in "xyz"], float)
            sb = lbj + frac * (top - lbj)
            block["strut_bottom"] = {"x": float(sb[0]), "y": float(sb[1]), "z": float(sb[2])}
    return g


def run_batch(name: str, geom: dict, sweep: dict, n_inst: int, sigma: float, seed: int) -> None:
    rng = np.random.default_rng(seed)
    geoms, pos = [], []
    for _ in range(n_inst):
        gi = perturb_geometry(geom, rng, sigma)
        sus = build_suspension(gi)
        cfg = build_sweep(sweep, sus)
        init = sus.initial_state()
        keys = sorted(init.positions)
        states, _ = solve_suspension_sweep(init, sus.constraints(), cfg, DerivedPointsManager(sus.derived_spec()), TIGHT)
        pos.append(positions_array(states, keys))
        geoms.append(gi)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), positions_tight=np.array(pos))
    json.dump({"geometries": geoms, "sweep": sweep, "point_keys": [key_name(k) for k in keys],
               "sigma": sigma, "seed": seed}, open(os.path.join(OUT, name + ".json"), "w"))
    print(f"{name}: {n_inst} perturbed instances")


# --------------------------------------------------------------------------- failure flags
def first_failure(geom: dict, sweep: dict) -> dict:
    """Run the reference; on RuntimeError report the failure class and the first failed step."""
    sus = build_suspension(geom)
    cfg = build_sweep(sweep, sus)
    try:
        solve_suspension_sweep(sus.initial_state(), sus.constraints(), cfg, DerivedPointsManager(sus.derived_spec()))
        return {"status": 0, "failed_step": -1, "message": ""}
    except RuntimeError as err:
        msg = str(err)
    if "acceptable residual" in msg:
        return {"status": 2, "failed_step": int(msg.split("sweep step ")[1].split(" ")[0]), "message": msg[:300]}
    # "Solver failed to converge" carries no step index: bisect over sweep prefixes.
    values = [[t.value for t in sw] for sw in cfg.target_sweeps]
    failed = -1
    for k in range(1, cfg.n_steps + 1):
        sub = copy.deepcopy(sweep)
        sub.pop("steps", None)
        for t, vals in zip(sub["targets"], values):
            t.pop("start", None), t.pop("stop", None)
            t["values"] = [float(v) for v in vals[:k]]
        s2 = build_suspension(geom)
        try:
            solve_suspension_sweep(s2.initial_state(), s2.constraints(), build_sweep(sub, s2),
                                   DerivedPointsManager(s2.derived_spec()))
        except RuntimeError:
            failed = k - 1
            break
    return {"status": 1, "failed_step": failed, "message": msg[:300]}


def run_failures() -> None:
    c1 = load("tests/data/geometry.yaml")
    cases = {f"c1_bump_to_{stop:+.0f}": (c1, bump_sweep(41, 0.0, stop)) for stop in (-400.0, 400.0, 500.0, -600.0)}
    cases["dw_corner_rocker_bump_-60_+80"] = (load("tests/data/corner_rocker_geometry.yaml"), load("scripts/bump_sweep.yaml"))
    cases["c2_macpherson_bump_to_+400"] = (load("tests/data/macpherson_geometry.yaml"), bump_sweep(41, 0.0, 400.0))
    out = {}
    for label, (geom, sweep) in cases.items():
        rec = first_failure(geom, sweep)
        out[label] = {"geometry": geom, "sweep": sweep, **rec}
        print(label, rec["status"], rec["failed_step"])
    json.dump(out, open(os.path.join(OUT, "failures.json"), "w"), indent=1)


# --------------------------------------------------------------------------- family vectors
def run_families() -> None:
    rng = np.random.default_rng(1234)
    K = [PointID(i) for i in range(1, 5)]
    recs = {}
    n = 8
    P = rng.normal(0, 100.0, size=(n, 4, 3))
    L = rng.uniform(50, 200, size=n)
    A = rng.uniform(0.2, 2.8, size=n)
    lp, ld = rng.normal(0, 50, size=(n, 3)), rng.normal(0, 1, size=(n, 3))
    ld /= np.linalg.norm(ld, axis=1)[:, None]

    def pos(i, k):
        return {K[j]: Point3(P[i, j]) for j in range(k)}

    def run(fam, k, make, jac, consts):
        res, grads = [], []
        for i in range(n):
            c = make(i)
            res.append(c.residual(pos(i, k)))
            grads.append(jac(i))
        recs[fam] = {"points": P[:, :k].tolist(), "consts": consts, "residual": res,
                     "jacobian": [np.asarray(g).reshape(k, 3).tolist() for g in grads]}

    run("distance", 2, lambda i: C.DistanceConstraint(K[0], K[1], L[i]), lambda i: JAC.jac_distance(P[i, 0], P[i, 1]), [[v] for v in L])
    run("spherical", 2, lambda i: C.SphericalJointConstraint(K[0], K[1]), lambda i: JAC.jac_distance(P[i, 0], P[i, 1]), [[] for _ in L])
    run("angle", 4, lambda i: C.AngleConstraint(*K, A[i]), lambda i: JAC.jac_angle(*P[i]), [[v] for v in A])
    run("three_point_angle", 3, lambda i: C.ThreePointAngleConstraint(*K[:3], A[i]), lambda i: JAC.jac_three_point_angle(*P[i, :3]), [[v] for v in A])
    run("vectors_parallel", 4, lambda i: C.VectorsParallelConstraint(*K), lambda i: JAC.jac_vectors_parallel(*P[i]), [[] for _ in L])
    run("vectors_perpendicular", 4, lambda i: C.VectorsPerpendicularConstraint(*K), lambda i: JAC.jac_vectors_perpendicular(*P[i]), [[] for _ in L])
    run("equal_distance", 4, lambda i: C.EqualDistanceConstraint(*K), lambda i: JAC.jac_equal_distance(*P[i]), [[] for _ in L])
    run("point_on_line", 1, lambda i: C.PointOnLineConstraint(K[0], Point3(lp[i]), Direction3(ld[i])),
        lambda i: JAC.jac_point_on_line(P[i, 0], lp[i], ld[i]), [[*lp[i], *ld[i]] for i in range(n)])
    run("linear_point", 1, lambda i: C.PointOnPlaneConstraint(K[0], Point3(lp[i]), Direction3(ld[i])),
        lambda i: JAC.jac_point_on_plane(P[i, 0], lp[i], ld[i]), [[*lp[i], *ld[i]] for i in range(n)])
    run("midpoint_on_plane", 2, lambda i: C.MidpointOnPlaneConstraint(K[0], K[1], Point3(lp[i]), Direction3(ld[i])),
        lambda i: np.concatenate([ld[i] / 2, ld[i] / 2]), [[*lp[i], *ld[i]] for i in range(n)])
    run("coplanar", 4, lambda i: C.CoplanarPointsConstraint(*K), lambda i: JAC.jac_coplanar(*P[i]), [[] for _ in L])
    V = rng.normal(0, 1e5, size=n)
    run("scalar_triple", 4, lambda i: C.ScalarTripleProductConstraint(*K, target_volume=V[i], scale=abs(V[i])),
        lambda i: JAC.jac_coplanar(*P[i]) / abs(V[i]), [[V[i], 1.0 / abs(V[i])] for i in range(n)])
    json.dump(recs, open(os.path.join(OUT, "families.json"), "w"))
    print("families:", sorted(recs))


# --------------------------------------------------------------------------- sweep diagnostics
def values_sweep(targets: list) -> dict:
    """Sweep with explicit per-step values: [(point, side, axis, [values...]), ...]."""
    out = []
    for point, side, axis, values in targets:
        t = {"point": point, "direction": {"axis": axis}, "mode": "relative", "values": [float(v) for v in values]}
        if side:
            t["side"] = side
        out.append(t)
    return {"version": 1, "targets": out}


def run_diagnostics() -> None:
    """diagnose_sweep (core/diagnostics.py:118-134) on sweeps with a deliberately uneven step (a
    jump), plus the U-bar chirality / transmission quantities per state."""
    from statistics import median

    from kinematics.core.diagnostics import (CONTINUITY_ABS_FLOOR_MM, CONTINUITY_MEDIAN_FACTOR, diagnose_sweep,
                                             _point_step_displacements)
    from kinematics.core.primitives.point_ref import PointRef, Side
    from kinematics.core.suspensions.axle.mechanisms import (ArbUBar, calculate_arb_branch_volume,
                                                            calculate_arb_chirality_margin,
                                                            calculate_transmission_margin)
    c1 = load("tests/data/geometry.yaml")
    c3 = add_coilovers(load("tests/data/axle_geometry_rocker.yaml"))
    z1 = [0, 2, 4, 6, 8, 10, 12, 14, 16, 50, 52, 54]
    z3 = [-4, -2, 0, 2, 4, 6, 30, 32]
    cases = {
        "c1_uneven_bump": (c1, values_sweep([("trackrod_inboard", None, "y", [0.0] * len(z1)),
                                             ("wheel_center", None, "z", z1)])),
        "c3_uneven_roll": (c3, values_sweep([("wheel_center", "left", "z", z3),
                                             ("wheel_center", "right", "z", [-v for v in z3]),
                                             ("trackrod_inboard", "left", "y", [0.0] * len(z3))])),
        "c3_roll_pm45": (c3, roll_sweep(19, 45.0)),
    }
    out, arrays = {}, {}
    for label, (geom, sweep) in cases.items():
        sus = build_suspension(geom)
        cfg = build_sweep(sweep, sus)
        init = sus.initial_state()
        states, stats = solve_suspension_sweep(init, sus.constraints(), cfg, DerivedPointsManager(sus.derived_spec()), TIGHT)
        report = diagnose_sweep(sus, states, stats)
        issues = [{"step": i.step, "category": str(i.category.value), "severity": str(i.severity.value),
                   "value": i.value, "message": i.message} for i in report.issues]
        thresholds = {}
        for key in sus.free_points():
            disp = _point_step_displacements(states, key)
            nonzero = [d for d in disp if d > 0]
            thresholds[key_name(key)] = max(CONTINUITY_ABS_FLOOR_MM,
                                            CONTINUITY_MEDIAN_FACTOR * (median(nonzero) if nonzero else 0.0))
        rec = {"geometry": geom, "sweep": sweep, "issues": issues, "thresholds": thresholds,
               "free_points": [key_name(k) for k in sus.free_points()]}
        arb = getattr(sus, "anti_roll", None)
        if isinstance(arb, ArbUBar):
            cols, names = [], []
            for side in (Side.LEFT, Side.RIGHT):
                tag = side.name.lower()
                names += [f"arb_branch_volume_{tag}", f"arb_chirality_margin_{tag}",
                          f"transmission_droplink_at_droplink_u_bar_{tag}",
                          f"transmission_pushrod_at_pushrod_inboard_{tag}",
                          f"transmission_droplink_at_droplink_rocker_{tag}"]
            for st in states:
                row = []
                a = st.get(PointRef(Side.CENTER, PointID.ARB_U_BAR_AXIS_A)).data
                b = st.get(PointRef(Side.CENTER, PointID.ARB_U_BAR_AXIS_B)).data
                for side in (Side.LEFT, Side.RIGHT):
                    pt = lambda pid: st.get(PointRef(side, pid)).data   # noqa: E731
                    drop = pt(PointID.DROPLINK_U_BAR) - pt(PointID.DROPLINK_ROCKER)
                    ra = pt(PointID.ROCKER_AXIS_A)
                    rax = pt(PointID.ROCKER_AXIS_B) - ra
                    push = pt(PointID.PUSHROD_OUTBOARD) - pt(PointID.PUSHROD_INBOARD)
                    margins = [calculate_transmission_margin(pt(PointID.DROPLINK_U_BAR), a, b - a, drop),
                               calculate_transmission_margin(pt(PointID.PUSHROD_INBOARD), ra, rax, push),
                               calculate_transmission_margin(pt(PointID.DROPLINK_ROCKER), ra, rax, drop)]
                    row += [calculate_arb_branch_volume(st, side), calculate_arb_chirality_margin(st, side),
                            *[np.nan if m is None else m for m in margins]]
                cols.append(row)
            arrays[label + "_columns"] = np.array(cols)
            rec["column_names"] = names
        keys = sorted(init.positions)
        arrays[label + "_positions"] = positions_array(states, keys)
        rec["point_keys"] = [key_name(k) for k in keys]
        out[label] = rec
        print(label, "issues:", [(i["step"], i["category"]) for i in issues])
    json.dump(out, open(os.path.join(OUT, "diagnostics.json"), "w"), indent=1)
    np.savez_compressed(os.path.join(OUT, "diagnostics.npz"), **arrays)


# --------------------------------------------------------------------------- generic families
def generic_mechanism() -> dict:
    """A synthetic one-degree-of-freedom linkage written with the generic constraint families
    that no shipped topology uses: planar four-bar A-B-C-D in the XZ
    plane, coupler point E on the bisector of BC, out-of-plane point F."""
    A, D = np.array([0.0, 0.0, 0.0]), np.array([300.0, 0.0, 0.0])
    B, C = np.array([30.0, 0.0, 95.0]), np.array([260.0, 0.0, 150.0])
    bc = (C - B) / np.linalg.norm(C - B)
    E = B + (C - B) / 2 + 100.0 * np.array([-bc[2], 0.0, bc[0]])
    F = E + np.array([0.0, 120.0, 0.0])
    pts = {"lower_wishbone_inboard_front": A, "lower_wishbone_inboard_rear": D, "lower_wishbone_outboard": B,
           "upper_wishbone_outboard": C, "axle_inboard": E, "axle_outboard": F}
    a, d, b, c, e, f = pts
    v1, v2 = B - E, C - E
    alpha = float(np.arctan2(np.linalg.norm(np.cross(v1, v2)), v1 @ v2))
    dist = lambda p, q: float(np.linalg.norm(pts[p] - pts[q]))   # noqa: E731
    cons = [
        {"family": "distance", "points": [a, b], "value": dist(a, b)},
        {"family": "point_on_plane", "points": [b], "plane_point": [0, 0, 0], "plane_normal": [0, 1, 0]},
        {"family": "distance", "points": [b, c], "value": dist(b, c)},
        {"family": "distance", "points": [c, d], "value": dist(c, d)},
        {"family": "fixed_axis", "points": [c], "axis": 1, "value": 0.0},
        {"family": "equal_distance", "points": [b, e, c, e]},
        {"family": "three_point_angle", "points": [b, e, c], "value": alpha},
        {"family": "coplanar", "points": [a, d, b, e]},
        {"family": "distance", "points": [e, f], "value": dist(e, f)},
        {"family": "vectors_perpendicular", "points": [e, f, b, c]},
        {"family": "vectors_perpendicular", "points": [e, f, a, d]},
    ]
    return {"points": {k: v.tolist() for k, v in pts.items()}, "free": [b, c, e, f], "constraints": cons,
            "target": {"point": b, "axis": 2, "values": np.linspace(0.0, -45.0, 16).tolist()}}


def run_generic() -> None:
    from kinematics.core.state import SuspensionState
    from kinematics.core.points.derived.manager import DerivedPointsSpec
    from kinematics.core.targeting import PointTarget, PointTargetAxis, SweepConfig
    from kinematics.core.enums import TargetPositionMode
    spec = generic_mechanism()
    key = lambda name: PointID[name.upper()]   # noqa: E731
    state = SuspensionState(positions={key(k): Point3(np.array(v, float)) for k, v in spec["points"].items()},
                            free_points={key(k) for k in spec["free"]})
    cons = []
    for c in spec["constraints"]:
        k = [key(n) for n in c["points"]]
        fam = c["family"]
        if fam == "distance":
            cons.append(C.DistanceConstraint(k[0], k[1], c["value"]))
        elif fam == "point_on_plane":
            cons.append(C.PointOnPlaneConstraint(k[0], Point3(np.array(c["plane_point"], float)),
                                                 Direction3(np.array(c["plane_normal"], float))))
        elif fam == "fixed_axis":
            cons.append(C.FixedAxisConstraint(k[0], Axis(c["axis"]), c["value"]))
        elif fam == "equal_distance":
            cons.append(C.EqualDistanceConstraint(*k))
        elif fam == "three_point_angle":
            cons.append(C.ThreePointAngleConstraint(*k, c["value"]))
        elif fam == "coplanar":
            cons.append(C.CoplanarPointsConstraint(*k))
        elif fam == "vectors_perpendicular":
            cons.append(C.VectorsPerpendicularConstraint(*k))
        else:
            raise KeyError(fam)
    t = spec["target"]
    sweep = SweepConfig([[PointTarget(key(t["point"]), PointTargetAxis(Axis(t["axis"])), v, TargetPositionMode.RELATIVE)
                          for v in t["values"]]])
    dm = DerivedPointsManager(DerivedPointsSpec(functions={}, dependencies={}))
    states, stats = solve_suspension_sweep(state, cons, sweep, dm, TIGHT)
    keys = sorted(state.positions)
    vel = []
    for step, st in enumerate(states):
        targets = convert_targets_to_absolute([sw[step] for sw in sweep.target_sweeps], state)
        fields, info = compute_state_tangents(st, cons, dm, targets)
        assert not info.rank_deficient
        vel.append([[fld.velocity(k) for k in keys] for fld in fields])
    np.savez_compressed(os.path.join(OUT, "generic_mechanism.npz"), positions_tight=positions_array(states, keys),
                        velocities=np.array(vel), max_residual=np.array([s.max_residual for s in stats]))
    spec["point_keys"] = [key_name(k) for k in keys]
    json.dump(spec, open(os.path.join(OUT, "generic_mechanism.json"), "w"), indent=1)
    print("generic mechanism:", len(states), "states, max residual", max(s.max_residual for s in stats))


def run_generic_parallel_spherical() -> None:
    """The two generic families no other golden exercises in a solve: a parallelogram A-B-C-D in the XZ
    plane whose sides AB and DC are kept parallel by VectorsParallelConstraint, and a point E tied to C
    by a SphericalJointConstraint (one scalar row whose gradient vanishes when the joint is closed)."""
    from kinematics.core.state import SuspensionState
    from kinematics.core.points.derived.manager import DerivedPointsSpec
    from kinematics.core.targeting import PointTarget, PointTargetAxis, SweepConfig
    from kinematics.core.enums import TargetPositionMode
    a, d, b, c, e = (PointID.LOWER_WISHBONE_INBOARD_FRONT, PointID.UPPER_WISHBONE_INBOARD_FRONT,
                     PointID.LOWER_WISHBONE_OUTBOARD, PointID.UPPER_WISHBONE_OUTBOARD, PointID.WHEEL_CENTER)
    pts = {a: [0, 0, 0], d: [100, 0, 0], b: [0, 0, 100], c: [100, 0, 100], e: [100, 0, 100]}
    state = SuspensionState(positions={k: Point3(np.array(v, float)) for k, v in pts.items()}, free_points={b, c, e})
    y0 = (Point3(np.zeros(3)), Direction3(np.array([0.0, 1.0, 0.0])))
    cons = [C.DistanceConstraint(a, b, 100.0), C.DistanceConstraint(b, c, 100.0), C.DistanceConstraint(d, c, 100.0),
            C.VectorsParallelConstraint(a, b, d, c), C.PointOnPlaneConstraint(b, *y0), C.PointOnPlaneConstraint(c, *y0),
            C.SphericalJointConstraint(c, e), C.FixedAxisConstraint(e, Axis.Y, 0.0), C.DistanceConstraint(d, e, 100.0)]
    values = np.linspace(0.0, 30.0, 7).tolist()
    sweep = SweepConfig([[PointTarget(b, PointTargetAxis(Axis.X), v, TargetPositionMode.RELATIVE) for v in values]])
    dm = DerivedPointsManager(DerivedPointsSpec(functions={}, dependencies={}))
    states, stats = solve_suspension_sweep(state, cons, sweep, dm, TIGHT)
    keys = sorted(state.positions)
    np.savez_compressed(os.path.join(OUT, "generic_parallel_spherical.npz"), positions_tight=positions_array(states, keys),
                        max_residual=np.array([s.max_residual for s in stats]), values=np.array(values))
    json.dump({"point_keys": [key_name(k) for k in keys], "free": [key_name(k) for k in (b, c, e)]},
              open(os.path.join(OUT, "generic_parallel_spherical.json"), "w"), indent=1)
    print("generic parallel + spherical:", len(states), "states, max residual", max(s.max_residual for s in stats),
          "joint gap", float(np.linalg.norm(states[-1].positions[c].data - states[-1].positions[e].data)))


# --------------------------------------------------------------------------- result files
def run_result_files() -> None:
    """The reference's own sweep file for the C1 case (cli/commands/sweep.py:39-79 -> CSV, format
    version 3) and the unit symbol of every flat metric column of every golden case."""
    import tempfile
    from pathlib import Path

    from kinematics.cli.commands.sweep import run_sweep_files
    from kinematics.core.metrics.registry import flat_specs_for_suspension
    units = {}
    for case, (geom, _sweep) in CASES.items():
        sus = build_suspension(geom)
        units.update({k: spec.unit.symbol for k, spec in flat_specs_for_suspension(sus).items()})
    geom, sweep = CASES["c1_dw_corner_bump_steer"]
    with tempfile.TemporaryDirectory() as tmp:
        gp, sp, op = Path(tmp) / "geometry.yaml", Path(tmp) / "sweep.yaml", Path(tmp) / "out.csv"
        gp.write_text(yaml.safe_dump(geom))
        sp.write_text(yaml.safe_dump(sweep))
        run_sweep_files(gp, sp, op)
        lines = op.read_text().splitlines()
    keep = [ln for ln in lines if not ln.startswith(("# timestamp", "# geometry_", "# sweep_"))]
    open(os.path.join(OUT, "e2e_c1_bump_steer.csv"), "w").write("\n".join(keep) + "\n")
    # setup-reference pose of analyze_sweep (core/analysis.py:182-216) for a corner and an axle
    from kinematics.core.analysis import analyze_sweep
    analysis = {}
    for case in ("c1_dw_corner_bump_steer", "c3_rocker_ubar_coilover_roll", "c4_tbar_roll"):
        g, sw = CASES[case]
        sus = build_suspension(g)
        res = analyze_sweep(sus, build_sweep(sw, sus))
        ref = res.references["setup"]
        analysis[case] = {
            "metric_keys": res.metric_keys, "corner_metric_keys": res.corner_metric_keys,
            "locations": res.locations, "steps": res.steps,
            "sweep_parameters": [[p.point, p.axis, p.side] for p in res.sweep_parameters],
            "setup_metrics": ref.metrics, "setup_corner_metrics": ref.corner_metrics,
            "setup_positions": {k: list(v) for k, v in ref.positions.items()},
            "last_frame_metrics": res.frames[-1].metrics,
            "point_keys": res.point_keys,
            "last_frame_positions": {k: list(v) for k, v in res.frames[-1].positions.items()},
            "diagnostics": [[d.step, str(d.category.value)] for d in res.diagnostics],
        }
    json.dump({"metric_units": units, "case": "c1_dw_corner_bump_steer", "analysis": analysis},
              open(os.path.join(OUT, "result_files.json"), "w"), indent=1)
    print("result files:", len(units), "metric units,", len(keep), "csv lines")


if __name__ == "__main__":
    only = set(sys.argv[1:])
    for case, (g, s) in CASES.items():
        if not only or case in only:
            run_case(case, g, s)
    if not only or "batches" in only:
        run_batch("batch_c1", *CASES["c1_dw_corner_bump"], n_inst=6, sigma=0.5, seed=1)
        run_batch("batch_c2", *CASES["c2_macpherson_bump_steer"], n_inst=6, sigma=0.5, seed=2)
        run_batch("batch_c3", *CASES["c3_rocker_ubar_coilover_roll"], n_inst=4, sigma=0.5, seed=3)
    if not only or "failures" in only:
        run_failures()
    if not only or "families" in only:
        run_families()
    if not only or "diagnostics" in only:
        run_diagnostics()
    if not only or "generic" in only:
        run_generic()
    if not only or "generic2" in only:
        run_generic_parallel_spherical()
    if not only or "results" in only:
        run_result_files()
