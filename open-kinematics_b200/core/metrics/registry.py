"""Metric column metadata: flat keys and units (reference core/metrics/registry.py, metrics/units.py).

Only what the export boundary needs is mirrored: which flat columns a suspension produces, in the
device's column order (= the reference's export order), and the unit symbol of each.  Display
labels and kinds are presentation metadata and stay out of scope.
"""

from __future__ import annotations

from dataclasses import dataclass

LOCATIONS = ("left", "right")
_DEG = {"camber", "caster", "kpi", "roadwheel_angle", "svsa_angle", "roll", "rocker_angle", "torsion_bar_twist",
        "arb_arm_angle", "arb_twist", "t_bar_heave_angle"}
_PERCENT = {"anti_dive", "anti_lift", "anti_squat"}


def flat_key(key: str, location: str | None) -> str:
    return key if location is None else f"{key}_{location}"


def metric_unit(name: str) -> str:
    """Unit symbol of a (flat or structural) metric key: angles in deg, anti-geometry in %, everything
    else in mm; ``deriv_<response>_wrt_<driver>`` is the quotient of its response unit and mm (every
    driver is a displacement)."""
    key = name
    for location in LOCATIONS:
        if key.endswith("_" + location):
            key = key[: -len(location) - 1]
    if key.startswith("deriv_") and "_wrt_" in key:
        return f"{metric_unit(key[len('deriv_'):key.index('_wrt_')])}/mm"
    if key in _DEG:
        return "deg"
    if key in _PERCENT:
        return "%"
    return "mm"


@dataclass(frozen=True)
class MetricSpec:
    key: str
    unit: str
    scope: str          # "corner" | "axle"


def flat_specs_for_suspension(suspension, targets=None) -> dict:
    """``{flat column: MetricSpec}`` in export order.  ``targets`` (one ``PointTarget`` per sweep
    dimension) select the derivative drivers exactly as they do on the device; without them the
    default state targets of the suspension are used."""
    from ..metrics_program import build_metric_program
    from ..topology import structure_point_index
    heads = list(targets) if targets else suspension.default_state_targets()
    prog = build_metric_program(suspension, heads, structure_point_index(suspension))
    return {name: MetricSpec(name, metric_unit(name), "corner" if side is not None or not suspension.is_axle else "axle")
            for name, (side, _) in zip(prog.names, prog.locations)}
