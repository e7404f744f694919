"""Integer constants of ``csrc/okin_defs.h`` (flag bits, status codes, layout), parsed once."""

from .core.topology import D

__all__ = ["D"]
