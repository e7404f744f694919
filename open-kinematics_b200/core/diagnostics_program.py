"""Compiles a suspension's advisory sweep checks into the flat diagnostic ops interpreted by
``okin_diagnostics`` (csrc/okin_core.cuh).

Topology-independent checks (convergence, residual, continuity; reference
core/diagnostics.py:136-226) need no ops: they are the base columns every diagnostic program
has.  Topology-owned checks are declared here in the order the reference emits their issues:
the U-bar's branch chirality and its three transmission margins per side
(axle/mechanisms.py:432-549).
"""

from __future__ import annotations

from dataclasses import dataclass, field

from .enums import PointID as P
from .primitives.point_ref import PointRef, Side
from .topology import D

BASE_COLUMNS = ["flags", "jump_count", "jump_max_mm", "jump_point_slot", "jump_threshold_mm"]
assert len(BASE_COLUMNS) == D["OKIN_DIAG_BASE"]


@dataclass
class DiagnosticProgram:
    names: list = field(default_factory=lambda: list(BASE_COLUMNS))
    ops: list = field(default_factory=list)
    # (kind, side, joint label, column) per topology check, reference emission order
    checks: list = field(default_factory=list)


def build_diagnostic_program(suspension, pidx: dict, design_pts: list) -> DiagnosticProgram:
    """``design_pts`` is the topology's list of points whose design position is kept on the device
    (shared with the metric program); slots are appended here as needed."""
    prog = DiagnosticProgram()

    def dslot(key) -> int:
        point = pidx[key]
        if point not in design_pts:
            design_pts.append(point)
        return design_pts.index(point)

    def op(kind, points, column, dslots=()):
        rec = [kind, *[pidx[k] for k in points]] + [0] * (5 - len(points)) + [column, *dslots]
        prog.ops.append(rec + [0] * (D["OKIN_DGOP_STRIDE"] - len(rec)))

    from .suspensions.axle import ArbUBar, AxleSuspension
    if isinstance(suspension, AxleSuspension) and isinstance(suspension.anti_roll, ArbUBar):
        axis_a, axis_b = PointRef(Side.CENTER, P.ARB_U_BAR_AXIS_A), PointRef(Side.CENTER, P.ARB_U_BAR_AXIS_B)
        for side in (Side.LEFT, Side.RIGHT):
            tag = side.name.lower()
            key = lambda point, side=side: PointRef(side, point)   # noqa: E731
            branch = [axis_a, axis_b, key(P.DROPLINK_ROCKER), key(P.DROPLINK_U_BAR)]
            col = len(prog.names)
            prog.names += [f"arb_branch_volume_{tag}", f"arb_chirality_margin_{tag}", f"arb_chirality_state_{tag}"]
            op(D["OKIN_DG_CHIRALITY"], branch, col, [dslot(k) for k in branch])
            prog.checks.append(("chirality", side, "U-bar arm", col))
            joints = [("droplink @ DROPLINK_U_BAR",
                       [key(P.DROPLINK_U_BAR), axis_a, axis_b, key(P.DROPLINK_ROCKER), key(P.DROPLINK_U_BAR)])]
            group = [key(p) for p in (P.ROCKER_AXIS_A, P.ROCKER_AXIS_B, P.PUSHROD_INBOARD, P.PUSHROD_OUTBOARD)]
            if all(k in pidx for k in group):
                ra, rb, pin, pout = group
                joints += [
                    ("pushrod @ PUSHROD_INBOARD", [pin, ra, rb, pin, pout]),
                    ("droplink @ DROPLINK_ROCKER",
                     [key(P.DROPLINK_ROCKER), ra, rb, key(P.DROPLINK_ROCKER), key(P.DROPLINK_U_BAR)]),
                ]
            for label, points in joints:
                col = len(prog.names)
                prog.names.append("transmission_" + label.replace(" @ ", "_at_").lower() + f"_{tag}")
                op(D["OKIN_DG_TRANSMISSION"], points, col)
                prog.checks.append(("transmission", side, label, col))
    return prog
