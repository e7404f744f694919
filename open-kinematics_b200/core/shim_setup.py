"""Setup pose of a shimmed corner, computed by the device (``okin_shim_presolve``)."""

from __future__ import annotations

import numpy as np

from .. import _lib
from .enums import Axis, PointID, TargetPositionMode
from .primitives.geometry import Point3
from .shim_program import corner_shim_record
from .targeting import PointTarget, PointTargetAxis
from .topology import compile_topology


def device_setup_pose(corner) -> dict:
    """Positions (authored + derived) after the camber-shim assembly pre-solve.

    Compiles the corner with its authored pose (only the structure matters: design constants are
    recomputed on the device), runs a zero-step batch of one instance and reads the design pose
    back.  Raises ``RuntimeError`` like ``solve_camber_shim_assembly`` when the assembly cannot
    be closed."""
    state = corner.authored_state()
    targets = [PointTarget(PointID.WHEEL_CENTER, PointTargetAxis(Axis.Z), 0.0, TargetPositionMode.RELATIVE)]
    rack = corner.rack_attachment_point()
    if rack is not None:
        targets.append(PointTarget(rack, PointTargetAxis(Axis.Y), 0.0, TargetPositionMode.RELATIVE))
    program = compile_topology(state, corner.constraints_at(state.positions), corner.derived_spec(), targets,
                               design_rules=True, shims=[corner_shim_record(corner)])
    topo = _lib.DeviceTopology(program)
    try:
        hp = np.array([corner.hardpoints[k].data for k in program.in_keys]).reshape(1, -1)
        out = topo.solve_batch(hp, np.zeros((len(targets), 0)), want_positions=False, want_design=True)
    finally:
        topo.close()
    if int(out["status"][0]) != 0:
        raise RuntimeError("Camber shim assembly solve failed to converge or did not satisfy its constraints.")
    return {k: Point3(out["design"][0, i]) for i, k in enumerate(program.out_keys)}
