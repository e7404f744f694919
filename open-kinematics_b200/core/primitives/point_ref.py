"""Side-qualified point keys for axle models (reference core/primitives/point_ref.py:24-109).

``PointRef`` sorts as the tuple ``(side, point)`` with LEFT < RIGHT < CENTER; that
ordering is the solver's column order for axles (reference core/state.py:50).
"""

from enum import IntEnum
from typing import NamedTuple, Union

from ..enums import PointID


class Side(IntEnum):
    LEFT = 0
    RIGHT = 1
    CENTER = 2

    @property
    def lateral_sign(self) -> float:
        if self is Side.LEFT:
            return 1.0
        if self is Side.RIGHT:
            return -1.0
        raise ValueError("CENTER does not have a lateral sign")


class PointRef(NamedTuple):
    side: Side
    point: PointID

    @property
    def name(self) -> str:
        return f"{self.side.name}_{self.point.name}"


PointKey = Union[PointID, PointRef]


def point_key_name(key: PointKey) -> str:
    return key.name.lower()


def side_qualified(side: Side, point: PointKey) -> PointRef:
    if not isinstance(point, PointID):
        raise TypeError(f"Cannot side-qualify a non-corner key: {point!r}")
    return PointRef(side, point)
