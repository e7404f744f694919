"""Metric row containers (reference core/metrics/main.py:36-60).

The numbers are computed on the device by the metric program compiled in
``core/metrics_program.py`` (one flat column per metric, reference export order); this module
only gives them the reference's row shapes: an ``OrderedDict`` per corner state, and
``AxleMetricRows`` (axle row + one row per side) per axle state.
"""

from __future__ import annotations

from collections import OrderedDict
from dataclasses import dataclass
from typing import Mapping

import numpy as np

from ..primitives.point_ref import Side

MetricRow = OrderedDict


def flat_key(key: str, location: str) -> str:
    """Flat export name of a located metric (reference metrics/registry.py ``flat_key``)."""
    return f"{key}_{location}"


@dataclass(frozen=True)
class AxleMetricRows:
    """Location-independent axle metrics plus one row per corner."""

    axle: MetricRow
    corners: dict

    def flat_row(self) -> MetricRow:
        return flatten_metric_rows(self.axle, self.corners)


def flatten_metric_rows(metrics: MetricRow, corner_metrics: Mapping) -> MetricRow:
    flat: MetricRow = OrderedDict()
    for side, row in corner_metrics.items():
        for key, value in row.items():
            flat[flat_key(key, side.name.lower())] = value
    flat.update(metrics)
    return flat


def rows_from_columns(values: np.ndarray, locations: list, is_axle: bool, with_derivatives: bool = True):
    """One device metric record (``values[n_metrics]``, NaN == None) -> the reference's row shape.
    ``locations[c] = (Side | None, key)`` as recorded by the metric-program builder.  A derivative
    column holding +inf is the device's mark for tied driver tangents: the reference raises there
    (metrics/derivatives.py:299-304), and so does this."""
    axle_row: MetricRow = OrderedDict()
    corner_rows = {Side.LEFT: OrderedDict(), Side.RIGHT: OrderedDict()} if is_axle else {}
    for value, (side, key) in zip(values, locations):
        if not with_derivatives and key.startswith("deriv_"):
            continue
        if key.startswith("deriv_") and np.isposinf(value):
            raise ValueError(f"Ambiguous derivative driver for column '{key}': "
                             "multiple matching tangents have equal strength")
        item = None if np.isnan(value) else float(value)
        (corner_rows[side] if (is_axle and side is not None) else axle_row)[key] = item
    return AxleMetricRows(axle=axle_row, corners=corner_rows) if is_axle else axle_row
