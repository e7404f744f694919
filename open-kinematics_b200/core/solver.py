"""Per-state solve and sweep continuation behind the reference's boundary B1.

``solve_suspension_sweep`` has the signature, return types and error behaviour of
reference ``core/solver.py:654-776``; the arithmetic (residuals, Jacobian,
damped least-squares iteration, continuation loop) runs in the CUDA library
through ``okin_solve_batch`` (``include/okin.h``).  ``solve_sweep_batch`` is the
batched entry point the reference does not have.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import NamedTuple

import numpy as np

from .. import _lib
from .constraints import Constraint
from .enums import TargetPositionMode
from .points.derived.manager import DerivedPointsManager
from .primitives.constants import (
    SOLVE_ACCEPT_RESIDUAL, SOLVE_TOLERANCE_GRAD, SOLVE_TOLERANCE_STEP, SOLVE_TOLERANCE_VALUE,
)
from .primitives.geometry import Point3
from .state import SuspensionState
from .targeting import PointTarget, SweepConfig, resolve_target
from .topology import TopologyProgram, compile_topology

SOLVE_METHOD = "lm"

STATUS_OK, STATUS_NOT_CONVERGED, STATUS_RESIDUAL_REJECTED, STATUS_INVALID_GEOMETRY = 0, 1, 2, 3


class SolverConfig(NamedTuple):
    """Reference ``SolverConfig`` (solver.py:65-80).  ``residual_tolerance`` is honoured
    exactly.  ``ftol/xtol/gtol`` are MINPACK stopping rules with no counterpart in the
    device iteration (Gauss-Newton ended by a step below ``fine_tol`` = 1e-4 mm, which leaves an
    error of second order in it, ~1e-10 mm on mm-scale linkages, or by a verification step below
    1e-6 mm; measured <= 3e-8 mm from the reference's tight-tolerance solution on every golden
    sweep); they are kept for call compatibility."""

    ftol: float = SOLVE_TOLERANCE_VALUE
    xtol: float = SOLVE_TOLERANCE_STEP
    gtol: float = SOLVE_TOLERANCE_GRAD
    verbose: int = 0
    residual_tolerance: float = SOLVE_ACCEPT_RESIDUAL


@dataclass
class SolverInfo:
    converged: bool
    nfev: int
    max_residual: float


def validate_least_squares_dimensions(n_vars: int, n_residuals: int, *, method: str = SOLVE_METHOD) -> None:
    if method == "lm" and n_vars > n_residuals:
        raise ValueError(
            f"System is underdetermined (n_vars={n_vars} > m_res={n_residuals}). "
            "The solve method (Levenberg-Marquardt) requires at least as many residuals as variables."
        )


def convert_targets_to_absolute(targets: list, initial_state: SuspensionState) -> list:
    """RELATIVE -> ABSOLUTE: ``dot(p_design, dir) + value`` (solver.py:584-627)."""
    out = []
    for t in targets:
        if t.mode == TargetPositionMode.ABSOLUTE:
            out.append(t)
            continue
        d = resolve_target(t.direction).data
        base = float(np.dot(initial_state.positions[t.point_id].data, d))
        out.append(PointTarget(t.point_id, t.direction, base + t.value, TargetPositionMode.ABSOLUTE))
    return out


def describe_constraint(constraint: Constraint) -> str:
    names = ", ".join(sorted(getattr(p, "name", str(p)) for p in constraint.involved_points))
    return f"{type(constraint).__name__}({names})"


def sweep_target_values(sweep_config: SweepConfig) -> tuple:
    """Split a sweep into per-dimension target declarations and a ``[T, S]`` value table.
    Point, direction and mode must be constant along each dimension."""
    heads, values = [], []
    for dim in sweep_config.target_sweeps:
        head = dim[0]
        d0 = resolve_target(head.direction).data
        for t in dim:
            if t.point_id != head.point_id or t.mode != head.mode or not np.array_equal(
                    resolve_target(t.direction).data, d0):
                raise ValueError("Every step of a sweep dimension must target the same point, "
                                 "direction and mode")
        heads.append(head)
        values.append([float(t.value) for t in dim])
    return heads, np.asarray(values, dtype=np.float64).reshape(len(heads), sweep_config.n_steps)


def _device_cfg(solver_config: SolverConfig):
    return _lib.default_cfg(residual_tol=float(solver_config.residual_tolerance))


def describe_worst_residual(program: TopologyProgram, worst_row: int, constraints: list, step_targets: list) -> str:
    """The constraint or target owning the largest residual row (solver.py:640-651).  The device
    reports its own row index; ``program.row_source`` maps it back to the reference's row (a pin or
    report row of a point-on-line constraint maps to that constraint)."""
    source = program.row_source[worst_row]
    if source[0] == "target":
        target = step_targets[source[1]]
        point_name = getattr(target.point_id, "name", str(target.point_id))
        return f"target on point '{point_name}' (direction {target.direction})"
    return f"constraint {describe_constraint(constraints[source[1]])}"


def _failure(program: TopologyProgram, constraints, heads, values, step, status, max_residual, tol,
             worst_row: int = -1) -> RuntimeError:
    step_targets = [PointTarget(h.point_id, h.direction, float(values[j, step]), h.mode) for j, h in enumerate(heads)]
    if status == STATUS_RESIDUAL_REJECTED:
        worst = describe_worst_residual(program, worst_row, constraints, step_targets) if worst_row >= 0 \
            else "unknown"
        return RuntimeError(
            f"Solve at sweep step {step} did not reach an acceptable residual: worst residual "
            f"{max_residual:.6g} exceeds the acceptance tolerance {tol:.6g}. Worst residual row: "
            f"{worst}. The mechanism likely cannot reach the requested targets "
            "(kinematic lock-out / infeasible target combination)."
        )
    reason = "invalid geometry (NaN in the design pose)" if status == STATUS_INVALID_GEOMETRY else \
        "iteration limit reached"
    return RuntimeError(f"Solver failed to converge for targets: {step_targets}.\nMessage: {reason}")


def solve_suspension_sweep(
    initial_state: SuspensionState,
    constraints: list,
    sweep_config: SweepConfig,
    derived_manager: DerivedPointsManager,
    solver_config: SolverConfig = SolverConfig(),
) -> tuple:
    """Solve a sweep sequentially from the initial state (reference solver.py:654-776).

    Inputs are not mutated; one fresh ``SuspensionState`` per step is returned.
    Raises ``ValueError`` for an underdetermined system and ``RuntimeError`` at the
    first step that does not converge or whose worst residual exceeds
    ``solver_config.residual_tolerance``.
    """
    n_steps = sweep_config.n_steps
    if n_steps == 0:
        validate_least_squares_dimensions(3 * len(initial_state.free_points), len(constraints))
        return [], []
    heads, values = sweep_target_values(sweep_config)
    program = compile_topology(initial_state, constraints, derived_manager.spec, heads, design_rules=False)
    topo = _lib.DeviceTopology(program)
    try:
        hardpoints = np.array([initial_state.positions[k].data for k in program.in_keys]).reshape(1, -1)
        out = topo.solve_batch(hardpoints, values, _device_cfg(solver_config), want_worst_row=True)
    finally:
        topo.close()
    status, failed = int(out["status"][0]), int(out["failed_step"][0])
    if status != STATUS_OK:
        raise _failure(program, constraints, heads, values, failed, status,
                       float(out["max_residual"][0, failed]), solver_config.residual_tolerance,
                       int(out["worst_row"][0]))
    states, stats = [], []
    for s in range(n_steps):
        positions = {k: Point3(out["positions"][0, s, i]) for i, k in enumerate(program.out_keys)}
        states.append(SuspensionState(positions=positions, free_points=set(initial_state.free_points)))
        stats.append(SolverInfo(converged=True, nfev=int(out["iters"][0, s]),
                                max_residual=float(out["max_residual"][0, s])))
    return states, stats
