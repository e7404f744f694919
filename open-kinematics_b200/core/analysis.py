"""Structured sweep analysis (reference core/analysis.py:52-316): frames of named positions and
metric rows, the solved setup-reference pose, sweep parameters and diagnostics.

The numbers come from the device through the sweep facade (``core/sweep.py``).  Frames and the
setup reference carry the reference's named positions: every point of the solved state under its
public name plus the presentation points (rocker-pickup axis projections, T-bar midpoint;
``core/presentation.py``).  The drawing metadata of the reference's analysis object (element
paths, wheel dimensions, metric display labels) is outside the solve path and is not mirrored.
"""

from __future__ import annotations

from dataclasses import dataclass, field

from .diagnostics import DiagnosticCategory, DiagnosticIssue, DiagnosticSeverity
from .enums import TargetPositionMode
from .metrics.main import AxleMetricRows
from .presentation import named_point_keys, resolve_positions
from .primitives.point_ref import PointRef, Side
from .solver import SolverInfo
from .sweep import (EvaluatedSweep, compute_sweep_metrics, evaluate_solved_sweep, solve_evaluated_sweep, solve_sweep)
from .targeting import PointTarget, SweepConfig


def point_key_name(key) -> str:
    return key.name.lower()


@dataclass(frozen=True)
class SuspensionInfo:
    name: str
    type_key: str
    units: str


@dataclass(frozen=True)
class SweepParameter:
    point: str
    axis: str
    side: str | None


@dataclass(frozen=True)
class AnalyzedFrame:
    index: int
    positions: dict
    metrics: dict
    corner_metrics: dict
    solver: SolverInfo


@dataclass(frozen=True)
class ReferenceCondition:
    label: str
    positions: dict
    metrics: dict
    corner_metrics: dict


@dataclass(frozen=True)
class SweepAnalysis:
    suspension: SuspensionInfo
    point_keys: list
    metric_keys: list
    corner_metric_keys: list
    locations: list
    sweep_parameters: list
    references: dict
    diagnostics: list
    frames: list = field(default_factory=list)

    @property
    def steps(self) -> int:
        return len(self.frames)


def named_positions(positions: dict) -> dict:
    return {point_key_name(k): tuple(float(v) for v in p.data) for k, p in positions.items()}


def sweep_parameters(sweep_config: SweepConfig) -> list:
    """Every principal-axis dimension of a sweep (reference analysis.py:133-152)."""
    out = []
    for dimension in sweep_config.target_sweeps:
        if not dimension:
            continue
        target = dimension[0]
        axis = getattr(target.direction, "axis", None)
        if axis is None:
            continue
        key = target.point_id
        side = key.side.name.lower() if isinstance(key, PointRef) and key.side is not Side.CENTER else None
        out.append(SweepParameter(point=point_key_name(key), axis=axis.name.lower(), side=side))
    return out


def _split_metric_rows(rows) -> tuple:
    if isinstance(rows, AxleMetricRows):
        return rows.axle, {side.name.lower(): row for side, row in rows.corners.items()}
    return rows, {}


def setup_reference(suspension, sweep_config: SweepConfig) -> tuple:
    """Solved nominal setup pose: every sweep dimension held at zero relative displacement
    (reference analysis.py:155-216).  Returns ``(ReferenceCondition | None, DiagnosticIssue | None)``."""
    hold = [[PointTarget(dim[0].point_id, dim[0].direction, 0.0, TargetPositionMode.RELATIVE)]
            for dim in sweep_config.target_sweeps if dim]
    if not hold:
        return None, None
    hold_config = SweepConfig(hold)
    try:
        states, _ = solve_sweep(suspension, hold_config)
        if not states:
            return None, None
        row = compute_sweep_metrics(suspension, hold_config, states).rows[0]
    except Exception as error:  # noqa: BLE001 - the reference pose is optional
        return None, DiagnosticIssue(
            None, DiagnosticCategory.REFERENCE, DiagnosticSeverity.WARNING,
            f"Setup reference unavailable: reference solve failed ({type(error).__name__}: {error}).", None)
    metrics, corner_metrics = _split_metric_rows(row)
    return ReferenceCondition("Setup", resolve_positions(states[0].positions, suspension), metrics, corner_metrics), None


def analyze_evaluated_sweep(suspension, sweep_config: SweepConfig, evaluated: EvaluatedSweep) -> SweepAnalysis:
    frames = []
    for index, (state, info, row) in enumerate(zip(evaluated.states, evaluated.solver_stats, evaluated.metrics.rows)):
        metrics, corner_metrics = _split_metric_rows(row)
        frames.append(AnalyzedFrame(index, resolve_positions(state.positions, suspension), metrics, corner_metrics, info))
    metric_keys, corner_metric_keys, locations = [], [], []
    for frame in frames:
        if not frame.metrics and not frame.corner_metrics:
            continue
        metric_keys, locations = list(frame.metrics), list(frame.corner_metrics)
        for row in frame.corner_metrics.values():
            corner_metric_keys += [k for k in row if k not in corner_metric_keys]
        break
    references = {}
    setup, issue = setup_reference(suspension, sweep_config)
    if setup is not None:
        references["setup"] = setup
    diagnostics = list(evaluated.diagnostics) + ([issue] if issue is not None else [])
    units = getattr(getattr(suspension, "units", None), "symbol", "mm")
    return SweepAnalysis(
        suspension=SuspensionInfo(suspension.name, str(suspension.reported_type_key()), units),
        point_keys=named_point_keys(suspension, evaluated.states[0].positions) if frames else [],
        metric_keys=metric_keys, corner_metric_keys=corner_metric_keys, locations=locations,
        sweep_parameters=sweep_parameters(sweep_config), references=references, diagnostics=diagnostics,
        frames=frames)


def analyze_sweep(suspension, sweep_config: SweepConfig) -> SweepAnalysis:
    """Solve a sweep and assemble its structured analysis (reference analysis.py:219-225)."""
    return analyze_evaluated_sweep(suspension, sweep_config, solve_evaluated_sweep(suspension, sweep_config))


def analyze_solved_sweep(suspension, sweep_config: SweepConfig, states: list, solver_stats: list) -> SweepAnalysis:
    return analyze_evaluated_sweep(suspension, sweep_config,
                                   evaluate_solved_sweep(suspension, sweep_config, states, solver_stats))
