"""Topology declarations: which points move, which constraints tie them, which
points are derived.  They run once per topology on the host; per-instance design
constants are recomputed on the device (csrc/okin_core.cuh, setup phase)."""
