"""Scalar residuals and analytical gradient rows per constraint family.

Restates reference ``core/constraints.py`` (residuals) and ``core/jacobians.py`` (gradients,
SymPy-derived there; written here in closed vector form).  ``softnorm(s) = sqrt(s + EPS_SQ) - EPS``
(``core/primitives/soft_math.py:16-27``); gradients differentiate ``sqrt(s + EPS_SQ)`` only
(``tools/generate_jacobians.py:34-44``).

Every function takes the points as a ``(k, 3)`` array and the constants as a sequence and
returns ``(residual, gradient (k, 3))``.
"""

from __future__ import annotations

import math

import numpy as np

EPS = 1e-6
EPS_SQ = EPS * EPS


def softnorm(s: float) -> float:
    return math.sqrt(s + EPS_SQ) - EPS


def distance(p, c):
    """constraints.py:125-134, jacobians.py:35-51."""
    d = p[1] - p[0]
    q = math.sqrt(float(d @ d) + EPS_SQ)
    return q - EPS - c[0], np.array([-d / q, d / q])


def spherical(p, c):
    """constraints.py:162-170 (gradient = distance)."""
    d = p[1] - p[0]
    q = math.sqrt(float(d @ d) + EPS_SQ)
    return q - EPS, np.array([-d / q, d / q])


def _angle_core(v1, v2):
    """atan2(softnorm(|v1 x v2|^2), v1.v2) and its gradient w.r.t. v1, v2 (of the smooth form)."""
    cr = np.cross(v1, v2)
    s = float(cr @ cr)
    q = math.sqrt(s + EPS_SQ)
    dt = float(v1 @ v2)
    value = math.atan2(q - EPS, dt)
    # d atan2(q, dt) = (dt dq - q d(dt)) / (q^2 + dt^2);  dq = (cr . d cr)/q
    den = q * q + dt * dt
    dq_dv1 = np.cross(v2, cr) / q
    dq_dv2 = np.cross(cr, v1) / q
    g1 = (dt * dq_dv1 - q * v2) / den
    g2 = (dt * dq_dv2 - q * v1) / den
    return value, g1, g2


def angle(p, c):
    """constraints.py:223-243, jacobians.py:55-122.  v1 = p2 - p1, v2 = p4 - p3."""
    value, g1, g2 = _angle_core(p[1] - p[0], p[3] - p[2])
    return value - c[0], np.array([-g1, g1, -g2, g2])


def three_point_angle(p, c):
    """constraints.py:287-308, jacobians.py:127-188.  Vertex p2."""
    value, g1, g2 = _angle_core(p[0] - p[1], p[2] - p[1])
    return value - c[0], np.array([g1, -g1 - g2, g2])


def vectors_parallel(p, c):
    """constraints.py:351-371, jacobians.py:192-262 (gradient of the sqrt-only form)."""
    v1, v2 = p[1] - p[0], p[3] - p[2]
    cr = np.cross(v1, v2)
    qc = math.sqrt(float(cr @ cr) + EPS_SQ)
    q1 = math.sqrt(float(v1 @ v1) + EPS_SQ)
    q2 = math.sqrt(float(v2 @ v2) + EPS_SQ)
    value = (qc - EPS) / ((q1 - EPS) * (q2 - EPS))
    f = qc / (q1 * q2)
    g1 = np.cross(v2, cr) / (qc * q1 * q2) - f * v1 / (q1 * q1)
    g2 = np.cross(cr, v1) / (qc * q1 * q2) - f * v2 / (q2 * q2)
    return value, np.array([-g1, g1, -g2, g2])


def vectors_perpendicular(p, c):
    """constraints.py:414-429, jacobians.py:266-318."""
    v1, v2 = p[1] - p[0], p[3] - p[2]
    q1 = math.sqrt(float(v1 @ v1) + EPS_SQ)
    q2 = math.sqrt(float(v2 @ v2) + EPS_SQ)
    dt = float(v1 @ v2)
    value = dt / ((q1 - EPS) * (q2 - EPS))
    f = dt / (q1 * q2)
    g1 = v2 / (q1 * q2) - f * v1 / (q1 * q1)
    g2 = v1 / (q1 * q2) - f * v2 / (q2 * q2)
    return value, np.array([-g1, g1, -g2, g2])


def equal_distance(p, c):
    """constraints.py:466-477, jacobians.py:322-367."""
    d1, d2 = p[1] - p[0], p[3] - p[2]
    q1 = math.sqrt(float(d1 @ d1) + EPS_SQ)
    q2 = math.sqrt(float(d2 @ d2) + EPS_SQ)
    return q1 - q2, np.array([-d1 / q1, d1 / q1, d2 / q2, -d2 / q2])


def point_on_line(p, c):
    """constraints.py:560-576, jacobians.py:372-403.  c = [p0(3), d(3)]."""
    p0, d = np.asarray(c[0:3]), np.asarray(c[3:6])
    cr = np.cross(p[0] - p0, d)
    q = math.sqrt(float(cr @ cr) + EPS_SQ)
    return q - EPS, np.array([np.cross(d, cr) / q])


def linear_point(p, c):
    """n.(p - p0): PointOnPlane (constraints.py:616-627), FixedAxis (:508-516), the tangent
    pins of sensitivity.py:146-174.  c = [p0(3), n(3)]."""
    p0, n = np.asarray(c[0:3]), np.asarray(c[3:6])
    return float((p[0] - p0) @ n), np.array([n])


def midpoint_on_plane(p, c):
    """constraints.py:657-666, solver.py:439-448."""
    p0, n = np.asarray(c[0:3]), np.asarray(c[3:6])
    mid = p[0] + (p[1] - p[0]) / 2.0
    return float((mid - p0) @ n), np.array([n / 2.0, n / 2.0])


def coplanar(p, c):
    """constraints.py:698-709, jacobians.py:426-483."""
    a, b, d = p[1] - p[0], p[2] - p[0], p[3] - p[0]
    g2, g3, g4 = np.cross(b, d), np.cross(d, a), np.cross(a, b)
    return float(a @ g2), np.array([-(g2 + g3 + g4), g2, g3, g4])


def scalar_triple(p, c):
    """constraints.py:731-733, solver.py:463-472.  c = [V, 1/scale]."""
    value, g = coplanar(p, c)
    return (value - c[0]) * c[1], g * c[1]


FAMILIES = {
    "distance": distance, "spherical": spherical, "angle": angle, "three_point_angle": three_point_angle,
    "vectors_parallel": vectors_parallel, "vectors_perpendicular": vectors_perpendicular,
    "equal_distance": equal_distance, "point_on_line": point_on_line, "linear_point": linear_point,
    "midpoint_on_plane": midpoint_on_plane, "coplanar": coplanar, "scalar_triple": scalar_triple,
}
