"""Post-sweep diagnostics (reference core/diagnostics.py:36-226 and the topology-owned checks of
axle/mechanisms.py:432-549).

The checks run on the device: per-state flag bits and topology quantities inside the sweep
kernel (``okin_diagnostics``), the continuity check as a second pass over the position rows
(``okin_continuity``).  Batches get them as arrays (``BatchSweepResult.diagnostics`` / ``.jumps``);
``diagnose_sweep`` gives one solved sweep the reference's ``SweepDiagnostics`` report, issue for
issue, by re-pinning the states on the device and reading those arrays.
"""

from __future__ import annotations

from dataclasses import dataclass
from enum import StrEnum
from math import acos, degrees

import numpy as np

from .primitives.constants import SOLVE_ACCEPT_RESIDUAL

CONTINUITY_ABS_FLOOR_MM: float = 5.0
CONTINUITY_MEDIAN_FACTOR: float = 4.0


class DiagnosticCategory(StrEnum):
    CONVERGENCE = "convergence"
    RESIDUAL = "residual"
    JUMP = "jump"
    DERIVATIVES = "derivatives"
    DIAGNOSTICS = "diagnostics"
    REFERENCE = "reference"
    CHIRALITY = "chirality"
    TRANSMISSION = "transmission"


class DiagnosticSeverity(StrEnum):
    WARNING = "warning"
    ERROR = "error"


@dataclass(frozen=True)
class DiagnosticIssue:
    step: int | None
    category: DiagnosticCategory
    severity: DiagnosticSeverity
    message: str
    value: float | None


@dataclass
class SweepDiagnostics:
    issues: list

    @property
    def ok(self) -> bool:
        return not self.errors

    @property
    def warnings(self) -> list:
        return [i for i in self.issues if i.severity is DiagnosticSeverity.WARNING]

    @property
    def errors(self) -> list:
        return [i for i in self.issues if i.severity is DiagnosticSeverity.ERROR]


def convergence_issues(stats: list) -> list:
    """Non-convergence and residuals above the acceptance threshold (diagnostics.py:136-173)."""
    issues = []
    for step, info in enumerate(stats):
        if not info.converged:
            issues.append(DiagnosticIssue(step, DiagnosticCategory.CONVERGENCE, DiagnosticSeverity.ERROR,
                                          f"Step {step} did not converge.", None))
        if info.max_residual > SOLVE_ACCEPT_RESIDUAL:
            issues.append(DiagnosticIssue(
                step, DiagnosticCategory.RESIDUAL, DiagnosticSeverity.ERROR,
                f"Step {step} residual {info.max_residual:.6g} exceeds the acceptance tolerance "
                f"{SOLVE_ACCEPT_RESIDUAL:.6g}.", float(info.max_residual)))
    return issues


def continuity_issues(free_points, free_order: list, jumps: np.ndarray) -> list:
    """Jump warnings from one instance's ``jumps[n_steps, n_free]`` array (row 0 = thresholds), in
    the reference's order: point by point, then step by step (diagnostics.py:176-226)."""
    issues = []
    for key in free_points:
        k = free_order.index(key)
        threshold = float(jumps[0, k])
        for step in range(1, jumps.shape[0]):
            displacement = float(jumps[step, k])
            if displacement <= 0.0:
                continue
            name = getattr(key, "name", str(key))
            issues.append(DiagnosticIssue(
                step, DiagnosticCategory.JUMP, DiagnosticSeverity.WARNING,
                f"Point '{name}' jumped {displacement:.3g} mm from step {step - 1} to step {step} "
                f"(threshold {threshold:.3g} mm); possible branch snap.", displacement))
    return issues


def topology_issues(checks: list, diag: np.ndarray) -> list:
    """Topology-owned issues from one instance's ``diag[n_steps, n_diagnostics]`` rows.  ``checks``
    is ``TopologyProgram.diagnostic_checks``; emission order = side, step, check
    (axle/mechanisms.py:432-549)."""
    issues = []
    sides = []
    for _, side, _, _ in checks:
        if side not in sides:
            sides.append(side)
    for side in sides:
        tag = side.name.lower()
        for step in range(diag.shape[0]):
            for kind, check_side, label, col in checks:
                if check_side is not side:
                    continue
                if kind == "chirality":
                    volume, margin, code = (float(diag[step, col + k]) for k in range(3))
                    if code == 1.0:
                        issues.append(DiagnosticIssue(
                            step, DiagnosticCategory.CHIRALITY, DiagnosticSeverity.ERROR,
                            f"{tag} U-bar arm reached its chirality boundary at step {step}.", margin))
                    elif code == 2.0:
                        issues.append(DiagnosticIssue(
                            step, DiagnosticCategory.CHIRALITY, DiagnosticSeverity.ERROR,
                            f"{tag} U-bar arm inverted at step {step}.", volume))
                elif kind == "transmission":
                    margin = float(diag[step, col])
                    if np.isnan(margin) or margin >= 0.15:
                        continue
                    angle_from_toggle = 90.0 - degrees(acos(min(1.0, margin)))
                    issues.append(DiagnosticIssue(
                        step, DiagnosticCategory.TRANSMISSION, DiagnosticSeverity.WARNING,
                        f"{tag} {label} is {angle_from_toggle:.1f} deg from toggle at step {step} "
                        f"(margin {margin:.3g}).", margin))
    return issues


def evaluate_diagnostics_on_device(suspension, states: list):
    """Re-pin the solved states on the device and return ``(diag, jumps, program)`` of the one
    instance."""
    from .sweep import BatchSolver, SweepConfig
    from .sensitivity import STATE_MATCH_TOL_MM, measured_targets
    heads = suspension.default_state_targets()
    pinned = [[measured_targets([h], st)[0] for st in states] for h in heads]
    solver = BatchSolver(suspension, SweepConfig(pinned))
    try:
        res = solver.solve(solver.nominal_hardpoints()[None, :], want_diagnostics=True)
    finally:
        solver.close()
    if int(res.status[0]) != 0:
        raise RuntimeError("Diagnostic evaluation failed: the states are not solutions of this suspension "
                           f"(device status {int(res.status[0])} at step {int(res.failed_step[0])}).")
    for s, st in enumerate(states):
        given = np.array([st.positions[k].data for k in solver.program.out_keys])
        drift = float(np.abs(res.positions[0, s] - given).max())
        if drift > STATE_MATCH_TOL_MM:
            raise RuntimeError(f"Diagnostic evaluation failed: state {s} is {drift:.3g} mm away from the "
                               "nearest solution of this suspension.")
    return res.diagnostics[0], res.jumps[0], solver.program


def diagnose_sweep(suspension, states: list, stats: list) -> SweepDiagnostics:
    """Topology-independent checks, then the topology's own (reference diagnostics.py:118-134)."""
    issues = convergence_issues(stats)
    if states:
        diag, jumps, program = evaluate_diagnostics_on_device(suspension, states)
        if len(states) >= 2:
            issues.extend(continuity_issues(suspension.free_points(), program.free_order, jumps))
        issues.extend(topology_issues(program.diagnostic_checks, diag))
    return SweepDiagnostics(issues=issues)
