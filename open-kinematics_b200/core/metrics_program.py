"""Host builder of the on-device metric program.

Lowers the reference's metric declarations for a built suspension -- the 19 corner state
metrics (``core/metrics/catalog.py:86-159``), the mechanism state metrics
(``corner/mechanisms.py:379-407``, ``:611-623``; ``axle/mechanisms.py:402-430``, ``:768-808``,
``:931-938``), the axle state metrics (``metrics/axle_metrics.py:18-95``) and the derivative
metrics (``catalog.py:169-308``, ``corner/mechanisms.py:316-377``, ``:512-536``, ``:589-605``,
``corner/macpherson.py:224-245``, ``axle/mechanisms.py:344-396``, ``:718-762``, ``:901-925``) --
to the flat records interpreted by ``csrc/okin_metrics.cuh``.  The column order is the
reference's flat export order (``metrics/main.py:50-60``).
"""

from __future__ import annotations

from dataclasses import dataclass, field

from .enums import AxlePosition, PointID
from .primitives.point_ref import PointRef, Side
from .suspensions.axle import ArbTBar, ArbUBar, AxleSuspension, T_BAR_LEFT_KEY, T_BAR_PIVOT_KEY, T_BAR_RIGHT_KEY
from .suspensions.corner import ActuationPushrodRocker, DoubleWishboneSuspension, MacPhersonSuspension

P = PointID

CORNER_STATE_METRICS = (
    "camber", "caster", "kpi", "scrub_radius", "mechanical_trail", "roadwheel_angle", "svic_x", "svic_z",
    "svsa_length", "fvic_y", "fvic_z", "fvsa_length", "wheel_travel", "half_track", "damper_length",
    "svsa_angle", "anti_dive", "anti_lift", "anti_squat",
)
AXLE_STATE_METRICS = ("heave", "roll", "ride_height_change", "track", "roll_center_y", "roll_center_z",
                      "rack_displacement")

# response / op codes of csrc/okin_metrics.cuh
R_COORD, R_DIST, R_CAMBER, R_TOE, R_CASTER, R_KPI, R_ROTATION, R_ROTATION_DIFF, R_MID_X = range(9)
R_TBAR_TWIST_DEG, R_TBAR_TWIST_DELTA, R_TBAR_HEAVE = 9, 10, 11
MOP_VALUE, MOP_DERIV = 0, 1
MOP_STRIDE, MCORNER_STRIDE, MAXLE_STRIDE = 16, 24, 16
IC_DW, IC_MAC, IC_NONE = 0, 1, 2
MF_FRONT, MF_REAR, MF_HAS_BIAS, MF_DRIVEN_HERE = 1, 2, 4, 8


@dataclass
class MetricProgram:
    names: list = field(default_factory=list)
    locations: list = field(default_factory=list)    # per column: (Side | None, key without side suffix)
    corners: list = field(default_factory=list)
    mops: list = field(default_factory=list)
    axle: list = field(default_factory=list)
    design_pts: list = field(default_factory=list)
    fconst: list = field(default_factory=list)


class _Builder:
    def __init__(self, heads, pidx):
        self.heads, self.pidx = heads, pidx
        self.prog = MetricProgram()
        self._dslot: dict = {}
        self.side, self.suffix = None, ""

    def col(self, name: str) -> int:
        if name in self.prog.names:
            raise ValueError(f"Duplicate metric column: {name}")
        self.prog.names.append(name)
        located = self.side is not None and name.endswith(self.suffix)
        self.prog.locations.append((self.side, name[:-len(self.suffix)]) if located else (None, name))
        return len(self.prog.names) - 1

    def dslot(self, key) -> int:
        if key not in self._dslot:
            self._dslot[key] = len(self.prog.design_pts)
            self.prog.design_pts.append(self.pidx[key])
        return self._dslot[key]

    def fc(self, values) -> int:
        off = len(self.prog.fconst)
        self.prog.fconst.extend(float(v) for v in values)
        return off

    def mop(self, kind, rtype, pts, out_col, d0=-1, d1=-1, consts=(0.0,), driver=None, mask=0):
        idx = [self.pidx[k] if k is not None else -1 for k in pts] + [-1] * (4 - len(pts))
        drv_point, drv_axis = (self.pidx[driver[0]], int(driver[1])) if driver else (-1, 0)
        rec = [kind, rtype, *idx, d0, d1, drv_point, drv_axis, mask, out_col, self.fc(consts)]
        self.prog.mops.append(rec + [0] * (MOP_STRIDE - len(rec)))


def _corner_block(b: _Builder, corner, key, suffix: str, candidates) -> None:
    """State metrics, mechanism metrics and derivative metrics of one corner, in the order of
    ``compute_metrics_for_state`` (metrics/main.py:145-184)."""
    side = corner.side.lateral_sign
    cfg = corner.config
    ai, ao = corner.wheel_axis_points()
    lo, up = corner.steering_axis_points()
    if isinstance(corner, DoubleWishboneSuspension):
        ic = [IC_DW, P.UPPER_WISHBONE_INBOARD_FRONT, P.UPPER_WISHBONE_INBOARD_REAR, P.UPPER_WISHBONE_OUTBOARD,
              P.LOWER_WISHBONE_INBOARD_FRONT, P.LOWER_WISHBONE_INBOARD_REAR, P.LOWER_WISHBONE_OUTBOARD]
    elif isinstance(corner, MacPhersonSuspension):
        ic = [IC_MAC, P.LOWER_WISHBONE_INBOARD_FRONT, P.LOWER_WISHBONE_INBOARD_REAR, P.LOWER_WISHBONE_OUTBOARD,
              P.STRUT_TOP, None, None]
    else:   # user-defined architecture composed through the role hooks: no instant-centre declaration
        ic = [IC_NONE, None, None, None, None, None, None]
    damper = corner.damper_points() or (None, None)
    flags = 0
    if cfg.axle_position is AxlePosition.FRONT:
        flags |= MF_FRONT
    if cfg.axle_position is AxlePosition.REAR:
        flags |= MF_REAR
    if cfg.front_brake_bias is not None:
        flags |= MF_HAS_BIAS
    if cfg.driven_axle is not None and cfg.axle_position is not None and cfg.driven_axle == cfg.axle_position:
        flags |= MF_DRIVEN_HERE

    def ix(pid):
        return -1 if pid is None else b.pidx[key(pid)]

    out_base = len(b.prog.names)
    for name in CORNER_STATE_METRICS:
        b.col(name + suffix)
    caux = b.fc([side, cfg.cg_position.data[2], cfg.wheelbase, cfg.front_brake_bias or 0.0])
    rec = [ix(ai), ix(ao), ix(P.WHEEL_CENTER), ix(P.CONTACT_PATCH_CENTER), ix(lo), ix(up), ic[0],
           *[ix(p) for p in ic[1:]], ix(damper[0]), ix(damper[1]),
           b.dslot(key(P.WHEEL_CENTER)), b.dslot(key(P.CONTACT_PATCH_CENTER)), out_base, caux, flags]
    b.prog.corners.append(rec + [0] * (MCORNER_STRIDE - len(rec)))

    # mechanism state metrics (double wishbone only)
    actuation = getattr(corner, "actuation", None)
    spring_kind = getattr(getattr(corner, "spring", None), "kind", "none")
    rocker = isinstance(actuation, ActuationPushrodRocker)

    def rotation(col: int, kind: int, driver=None, mask=0):
        b.mop(kind, R_ROTATION, [key(P.PUSHROD_INBOARD), key(P.ROCKER_AXIS_A), key(P.ROCKER_AXIS_B)], col,
              d0=b.dslot(key(P.PUSHROD_INBOARD)), consts=[side], driver=driver, mask=mask)

    if rocker:
        rotation(b.col("rocker_angle" + suffix), MOP_VALUE)
    if spring_kind == "torsion_bar":
        rotation(b.col("torsion_bar_twist" + suffix), MOP_VALUE)

    # derivative metrics (catalog.py:169-308 then mechanism declarations)
    hub = (key(P.WHEEL_CENTER), 2)
    hub_mask = candidates(P.WHEEL_CENTER)

    def deriv(name, driver_name, rtype, pts, consts=(0.0,), driver=hub, mask=hub_mask, d0=-1):
        b.mop(MOP_DERIV, rtype, [key(p) for p in pts], b.col(f"deriv_{name}_wrt_{driver_name}{suffix}"),
              d0=d0, consts=consts, driver=driver, mask=mask)

    deriv("camber", "hub_z", R_CAMBER, [ai, ao], [side])
    deriv("roadwheel_angle", "hub_z", R_TOE, [ai, ao], [side])
    deriv("caster", "hub_z", R_CASTER, [lo, up])
    deriv("kpi", "hub_z", R_KPI, [lo, up], [side])
    deriv("half_track", "hub_z", R_COORD, [P.CONTACT_PATCH_CENTER], [0.0, side, 0.0])
    deriv("wheel_center_x", "hub_z", R_COORD, [P.WHEEL_CENTER], [1.0, 0.0, 0.0])
    rack = corner.rack_attachment_point()
    if rack is not None:
        rack_driver, rack_mask = (key(rack), 1), candidates(rack)
        deriv("roadwheel_angle", "rack_displacement", R_TOE, [ai, ao], [side], rack_driver, rack_mask)
        deriv("camber", "rack_displacement", R_CAMBER, [ai, ao], [side], rack_driver, rack_mask)
    if rocker:
        rotation(b.col("deriv_rocker_angle_wrt_hub_z" + suffix), MOP_DERIV, hub, hub_mask)
    if spring_kind == "coilover" or isinstance(corner, MacPhersonSuspension):
        deriv("damper_length", "hub_z", R_DIST, [P.STRUT_TOP, P.STRUT_BOTTOM])
    if spring_kind == "torsion_bar":
        rotation(b.col("deriv_torsion_bar_twist_wrt_hub_z" + suffix), MOP_DERIV, hub, hub_mask)


def build_metric_program(suspension, heads, pidx) -> MetricProgram:
    b = _Builder(heads, pidx)
    if not isinstance(suspension, AxleSuspension):
        def candidates(point):
            return sum(1 << j for j, h in enumerate(heads) if h.point_id == point)
        _corner_block(b, suspension, lambda pid: pid, "", candidates)
        return b.prog

    axle = suspension
    dofs = axle.actuator_dofs()
    for side in (Side.LEFT, Side.RIGHT):
        def local_target(target_key, side=side):
            """Side-local target of an axle tangent, or its shared-actuator equivalent
            (metrics/main.py:102-142)."""
            if isinstance(target_key, PointRef) and target_key.side is side:
                return target_key.point
            for dof in dofs:
                if target_key in dof.point_keys:
                    for k in dof.point_keys:
                        if isinstance(k, PointRef) and k.side is side:
                            return k.point
            return None

        def candidates(point, local_target=local_target):
            return sum(1 << j for j, h in enumerate(heads) if local_target(h.point_id) == point)

        suffix = "_" + side.name.lower()
        b.side, b.suffix = side, suffix
        _corner_block(b, axle.corners[side], lambda pid, side=side: PointRef(side, pid), suffix, candidates)
        if isinstance(axle.anti_roll, ArbUBar):   # per-corner row appended after the derivatives
            arm = PointRef(side, P.DROPLINK_U_BAR)
            b.mop(MOP_VALUE, R_ROTATION,
                  [arm, PointRef(Side.CENTER, P.ARB_U_BAR_AXIS_A), PointRef(Side.CENTER, P.ARB_U_BAR_AXIS_B)],
                  b.col("arb_arm_angle" + suffix), d0=b.dslot(arm), consts=[1.0])

    # axle state metrics
    b.side, b.suffix = None, ""
    L, R = Side.LEFT, Side.RIGHT
    out_base = len(b.prog.names)
    for name in AXLE_STATE_METRICS:
        b.col(name)
    rack = axle.corners[L].rack_attachment_point()
    rack_key = PointRef(L, rack) if rack is not None else None
    wc = {s: PointRef(s, P.WHEEL_CENTER) for s in (L, R)}
    cp = {s: PointRef(s, P.CONTACT_PATCH_CENTER) for s in (L, R)}
    rec = [pidx[wc[L]], pidx[wc[R]], pidx[cp[L]], pidx[cp[R]], b.dslot(wc[L]), b.dslot(wc[R]), b.dslot(cp[L]),
           b.dslot(cp[R]), pidx[rack_key] if rack_key else -1, b.dslot(rack_key) if rack_key else -1, out_base]
    b.prog.axle = rec + [0] * (MAXLE_STRIDE - len(rec))

    def axle_candidates(key):
        return sum(1 << j for j, h in enumerate(heads) if h.point_id == key)

    arb = axle.anti_roll
    if isinstance(arb, ArbUBar):
        arms = {s: PointRef(s, P.DROPLINK_U_BAR) for s in (L, R)}
        pts = [arms[L], PointRef(Side.CENTER, P.ARB_U_BAR_AXIS_A), PointRef(Side.CENTER, P.ARB_U_BAR_AXIS_B), arms[R]]
        d = (b.dslot(arms[L]), b.dslot(arms[R]))
        b.mop(MOP_VALUE, R_ROTATION_DIFF, pts, b.col("arb_twist"), d0=d[0], d1=d[1])
    elif isinstance(arb, ArbTBar):
        pts = [T_BAR_LEFT_KEY, T_BAR_RIGHT_KEY, T_BAR_PIVOT_KEY]
        d = (b.dslot(T_BAR_LEFT_KEY), b.dslot(T_BAR_RIGHT_KEY))
        b.mop(MOP_VALUE, R_TBAR_HEAVE, pts, b.col("t_bar_heave_angle"), d0=d[0], d1=d[1])
        b.mop(MOP_VALUE, R_TBAR_TWIST_DELTA, pts, b.col("arb_twist"), d0=d[0], d1=d[1])
    heave_link = axle.heave_link.kind == "rocker_to_rocker"
    hl = [PointRef(L, P.HEAVE_LINK_ROCKER), PointRef(R, P.HEAVE_LINK_ROCKER)]
    if heave_link:
        b.mop(MOP_VALUE, R_DIST, hl, b.col("heave_link_length"))

    # axle derivative metrics
    if isinstance(arb, ArbUBar):
        for s in (L, R):
            b.mop(MOP_DERIV, R_ROTATION_DIFF, pts, b.col(f"deriv_arb_twist_wrt_hub_z_{s.name.lower()}"),
                  d0=d[0], d1=d[1], driver=(wc[s], 2), mask=axle_candidates(wc[s]))
    elif isinstance(arb, ArbTBar):
        for s in (L, R):
            drv, mask = (wc[s], 2), axle_candidates(wc[s])
            b.mop(MOP_DERIV, R_MID_X, pts[:2], b.col(f"deriv_t_bar_center_x_wrt_hub_z_{s.name.lower()}"),
                  driver=drv, mask=mask)
            b.mop(MOP_DERIV, R_TBAR_TWIST_DEG, pts, b.col(f"deriv_arb_twist_wrt_hub_z_{s.name.lower()}"),
                  driver=drv, mask=mask)
    if heave_link:
        for s in (L, R):
            b.mop(MOP_DERIV, R_DIST, hl, b.col(f"deriv_heave_link_length_wrt_hub_z_{s.name.lower()}"),
                  driver=(wc[s], 2), mask=axle_candidates(wc[s]))
    return b.prog
