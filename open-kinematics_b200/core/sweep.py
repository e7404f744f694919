"""Sweep orchestration (reference core/sweep.py:35-65) and its batched counterpart."""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .. import _lib
from .points.derived.manager import DerivedPointsManager
from .solver import SolverConfig, solve_suspension_sweep, sweep_target_values
from .targeting import SweepConfig, validate_sweep_controls
from .topology import TopologyProgram, compile_topology


def solve_sweep(suspension, sweep_config: SweepConfig) -> tuple:
    """Single-instance sweep: same call and results as reference ``solve_sweep``."""
    validate_sweep_controls(sweep_config, suspension.actuator_dofs())
    manager = DerivedPointsManager(suspension.derived_spec())
    return solve_suspension_sweep(
        initial_state=suspension.initial_state(),
        constraints=suspension.constraints(),
        sweep_config=sweep_config,
        derived_manager=manager,
    )


@dataclass
class BatchSweepResult:
    """Arrays for a batch of hardpoint-perturbed instances of one topology.

    positions    [n_instances, n_steps, n_out_points, 3]  (NaN from the failed step on)
    status       [n_instances]  0 ok | 1 not converged | 2 residual rejected | 3 invalid geometry
    failed_step  [n_instances]  -1 or first failed step
    nfev         [n_instances, n_steps]
    max_residual [n_instances, n_steps]
    tangents     [n_instances, n_steps, n_targets, n_unknowns] or None
    metrics      [n_instances, n_steps, n_metrics] or None; columns = ``metric_names`` (the
                 reference's flat export order), NaN where the reference yields None
    """

    program: TopologyProgram
    positions: np.ndarray | None
    status: np.ndarray
    failed_step: np.ndarray
    nfev: np.ndarray
    max_residual: np.ndarray
    tangents: np.ndarray | None
    metrics: np.ndarray | None = None

    @property
    def point_keys(self) -> list:
        return self.program.out_keys

    @property
    def metric_names(self) -> list:
        return self.program.metric_names


class BatchSolver:
    """A topology compiled once and resident on the device(s); solves any number of
    hardpoint sets of that topology.  Design constants (link lengths, upright angle,
    signed volumes, rack line point, strut clamp offset) are recomputed per instance on
    the device from that instance's hardpoints (reference recomputes them in
    ``Suspension.constraints()``, core/sweep.py:58-63)."""

    def __init__(self, suspension, sweep_config: SweepConfig, output_points=None):
        validate_sweep_controls(sweep_config, suspension.actuator_dofs())
        self.suspension = suspension
        self.heads, self.values = sweep_target_values(sweep_config)
        from .metrics_program import build_metric_program
        from .shim_program import shim_records
        state, constraints = suspension.structure()
        self.program = compile_topology(
            state, constraints, suspension.derived_spec(), self.heads,
            output_points=output_points, design_rules=True,
            metrics=(lambda pidx: build_metric_program(suspension, self.heads, pidx))
            if suspension.config is not None else None,
            shims=shim_records(suspension),
        )
        self.topology = _lib.DeviceTopology(self.program)

    def nominal_hardpoints(self) -> np.ndarray:
        """Authored positions of the input points, shape ``[n_in*3]`` (input slot order)."""
        authored = self.authored_positions()
        return np.array([authored[k].data for k in self.program.in_keys]).reshape(-1)

    def authored_positions(self) -> dict:
        sus = self.suspension
        if getattr(sus, "is_axle", False):
            from .enums import PointID
            from .primitives.point_ref import PointRef, Side
            out = {PointRef(side, k): p for side, corner in sus.corners.items() for k, p in corner.hardpoints.items()}
            for point, p in getattr(sus.anti_roll, "center_points", {}).items():
                out[PointRef(Side.CENTER, point)] = p
            for side, p in getattr(sus.anti_roll, "droplink_points", {}).items():
                arm = PointID.DROPLINK_U_BAR if PointRef(side, PointID.DROPLINK_U_BAR) in self.program.in_keys \
                    else PointID.DROPLINK_T_BAR
                out[PointRef(side, arm)] = p
            return out
        return dict(sus.hardpoints)

    def solve(self, hardpoints: np.ndarray, solver_config: SolverConfig = SolverConfig(), devices=None,
              want_positions: bool = True, want_tangents: bool = False,
              want_metrics: bool = False, params: np.ndarray | None = None) -> BatchSweepResult:
        """``params``: optional ``[n_instances, n_params]`` per-instance scalars in the order of
        ``program.param_names`` (camber-shim datums and thicknesses); default = the model's."""
        cfg = _lib.default_cfg(residual_tol=float(solver_config.residual_tolerance))
        hp = np.asarray(hardpoints, dtype=np.float64)
        hp = hp.reshape(hp.shape[0], -1)
        out = self.topology.solve_batch(hp, self.values, cfg, devices=devices,
                                        want_positions=want_positions, want_tangents=want_tangents,
                                        want_metrics=want_metrics, params=params)
        return BatchSweepResult(self.program, out["positions"], out["status"], out["failed_step"],
                                out["iters"], out["max_residual"], out["tangents"], out["metrics"])

    def close(self) -> None:
        self.topology.close()


def solve_sweep_batch(suspension, sweep_config: SweepConfig, hardpoints: np.ndarray, **kwargs) -> BatchSweepResult:
    """One-shot batched solve (compile + solve + release)."""
    solver = BatchSolver(suspension, sweep_config, output_points=kwargs.pop("output_points", None))
    try:
        return solver.solve(hardpoints, **kwargs)
    finally:
        solver.close()


def compute_sweep_metrics_batch(suspension, sweep_config: SweepConfig, hardpoints: np.ndarray, **kwargs):
    """Batched counterpart of reference ``compute_sweep_metrics`` (core/sweep.py:144-173): solves
    the sweeps and returns ``(metric_names, metrics[n_instances, n_steps, n_metrics], result)``
    with tangents and derivative metrics evaluated on the device."""
    result = solve_sweep_batch(suspension, sweep_config, hardpoints, want_metrics=True, **kwargs)
    return result.metric_names, result.metrics, result
