"""Derived points and their Jacobian blocks (reference ``core/points/derived/definitions.py:24-180``,
``corner/macpherson.py:307-313``; the reference differentiates them with dual numbers,
``points/derived/manager.py:271-324`` -- here the closed forms: d unit(v) = (I - u u^T)/|v|)."""

from __future__ import annotations

import numpy as np


def _unit(v):
    n = float(np.sqrt(v @ v))
    if n < 1e-6:
        raise ValueError("Cannot normalize a zero-length vector")  # generic.py:158-171
    u = v / n
    return u, (np.eye(3) - np.outer(u, u)) / n


def midpoint(a, b, param=None):
    """a + (b - a)/2  (definitions.py:76-89)."""
    return a + (b - a) / 2.0, [np.eye(3) * 0.5, np.eye(3) * 0.5]


def along_line(a, b, param):
    """a + unit(b - a) * param  (definitions.py:24-33; wheel centre/rim faces :92-155)."""
    u, du = _unit(b - a)
    return a + u * param, [np.eye(3) - du * param, du * param]


def contact_patch(wc, ai, ao, param):
    """wc + unit(down - (down.u)u) * r, u = unit(ao - ai), down = -Z  (definitions.py:36-73, :158-180)."""
    u, du = _unit(ao - ai)
    down = np.array([0.0, 0.0, -1.0])
    w = down - (down @ u) * u
    dw_du = -(np.outer(u, down) + (down @ u) * np.eye(3))
    q, dq = _unit(w)
    j_u = dq @ dw_du @ du * param
    return wc + q * param, [np.eye(3), -j_u, j_u]


OPS = {"midpoint": midpoint, "along_line": along_line, "contact_patch": contact_patch}
