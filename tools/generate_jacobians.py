#!/usr/bin/env python3
"""Emit fp64 CUDA device functions (residual + analytical gradient) for every
constraint family of the solve path.

Retargets the idea of the reference generator (``tools/generate_jacobians.py:47-220``,
which prints SymPy-CSE'd *Python* snippets to paste into ``core/jacobians.py``) to
CUDA: for each family one ``__host__ __device__`` function computing the scalar
residual and its gradient with a *shared* CSE, plus a residual-only variant.  The
output file is ``open-kinematics_b200/csrc/okin_gen_constraints.cuh`` (committed;
regenerate with ``python tools/generate_jacobians.py``).

Formulas (reference ``core/constraints.py``, ``core/primitives/soft_math.py:16-27``):
the residual *value* uses ``softnorm(s) = sqrt(s + EPS_SQ) - EPS``; the gradient is
that of the smooth part ``sqrt(s + EPS_SQ)`` exactly as the reference differentiates
it (``generate_jacobians.py:34-44``: "the bias correction is constant and vanishes
under differentiation" -- also inside ``atan2`` and in the normalised families,
where the reference gradient therefore differs from the true one at O(1e-6)
relative; we reproduce the reference's choice).

Calling convention of every generated function::

    double okin_<family>_res(const double* p, const double* c)
    double okin_<family>_resgrad(const double* p, const double* c, double* g)

``p`` holds the family's points packed ``[x1,y1,z1,x2,...]``, ``c`` its constants,
``g`` receives ``dR/dp`` in the same packing.
"""

from __future__ import annotations

import os
import sys

import sympy as sp
from sympy.printing.c import C99CodePrinter

EPS = sp.Symbol("OKIN_EPS", positive=True)
EPS_SQ = sp.Symbol("OKIN_EPS_SQ", positive=True)


def pts(n):
    """Symbols for n points and the list in packing order."""
    syms = []
    for i in range(1, n + 1):
        syms += list(sp.symbols(f"x{i} y{i} z{i}", real=True))
    return syms


def consts(n):
    return [sp.Symbol(f"c{i}", real=True) for i in range(n)]


def smooth(s):
    return sp.sqrt(s + EPS_SQ)


def cross(a, b):
    return (a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0])


def dot(a, b):
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]


def sq(a):
    return dot(a, a)


def sub(a, b):
    return tuple(x - y for x, y in zip(a, b))


class Family:
    """value: residual as evaluated; smooth_value: expression that is differentiated."""

    def __init__(self, name, n_points, n_consts, doc, build):
        self.name, self.n_points, self.n_consts, self.doc = name, n_points, n_consts, doc
        self.p = pts(n_points)
        self.c = consts(n_consts)
        P = [tuple(self.p[3 * i: 3 * i + 3]) for i in range(n_points)]
        self.value, self.smooth_value = build(P, self.c)


def fam_distance(P, c):
    s = sq(sub(P[1], P[0]))
    return smooth(s) - EPS - c[0], smooth(s)


def fam_spherical(P, c):
    s = sq(sub(P[1], P[0]))
    return smooth(s) - EPS, smooth(s)


def _angle(v1, v2, alpha):
    cr = cross(v1, v2)
    s = sq(cr)
    return sp.atan2(smooth(s) - EPS, dot(v1, v2)) - alpha, sp.atan2(smooth(s), dot(v1, v2))


def fam_angle(P, c):
    return _angle(sub(P[1], P[0]), sub(P[3], P[2]), c[0])


def fam_three_point_angle(P, c):
    return _angle(sub(P[0], P[1]), sub(P[2], P[1]), c[0])


def fam_vectors_parallel(P, c):
    v1, v2 = sub(P[1], P[0]), sub(P[3], P[2])
    s = sq(cross(v1, v2))
    value = (smooth(s) - EPS) / ((smooth(sq(v1)) - EPS) * (smooth(sq(v2)) - EPS))
    return value, smooth(s) / (smooth(sq(v1)) * smooth(sq(v2)))


def fam_vectors_perpendicular(P, c):
    v1, v2 = sub(P[1], P[0]), sub(P[3], P[2])
    value = dot(v1, v2) / ((smooth(sq(v1)) - EPS) * (smooth(sq(v2)) - EPS))
    return value, dot(v1, v2) / (smooth(sq(v1)) * smooth(sq(v2)))


def fam_equal_distance(P, c):
    d1, d2 = smooth(sq(sub(P[1], P[0]))), smooth(sq(sub(P[3], P[2])))
    return d1 - d2, d1 - d2


def fam_point_on_line(P, c):
    w = sub(P[0], (c[0], c[1], c[2]))
    s = sq(cross(w, (c[3], c[4], c[5])))
    return smooth(s) - EPS, smooth(s)


def fam_linear_point(P, c):
    v = dot(sub(P[0], (c[0], c[1], c[2])), (c[3], c[4], c[5]))
    return v, v


def fam_midpoint_on_plane(P, c):
    a, b = P
    mid = tuple(a[i] + (b[i] - a[i]) / 2 for i in range(3))
    v = dot(sub(mid, (c[0], c[1], c[2])), (c[3], c[4], c[5]))
    return v, v


def _triple(P):
    return dot(sub(P[1], P[0]), cross(sub(P[2], P[0]), sub(P[3], P[0])))


def fam_coplanar(P, c):
    v = _triple(P)
    return v, v


def fam_scalar_triple(P, c):
    v = (_triple(P) - c[0]) * c[1]
    return v, v


FAMILIES = [
    Family("distance", 2, 1, "softnorm(|p2-p1|^2) - L; c = [L]  (constraints.py:125-134, jacobians.py:35-51)", fam_distance),
    Family("spherical", 2, 0, "softnorm(|p2-p1|^2)  (constraints.py:162-170)", fam_spherical),
    Family("angle", 4, 1, "atan2(softnorm(|v1 x v2|^2), v1.v2) - alpha, v1=p2-p1, v2=p4-p3; c = [alpha]  (constraints.py:223-243, jacobians.py:55-122)", fam_angle),
    Family("three_point_angle", 3, 1, "vertex p2, v1=p1-p2, v2=p3-p2; c = [alpha]  (constraints.py:287-308, jacobians.py:127-188)", fam_three_point_angle),
    Family("vectors_parallel", 4, 0, "softnorm(|v1 x v2|^2)/(softnorm(|v1|^2) softnorm(|v2|^2))  (constraints.py:351-371, jacobians.py:192-262)", fam_vectors_parallel),
    Family("vectors_perpendicular", 4, 0, "(v1.v2)/(softnorm(|v1|^2) softnorm(|v2|^2))  (constraints.py:414-429, jacobians.py:266-318)", fam_vectors_perpendicular),
    Family("equal_distance", 4, 0, "softnorm(|p2-p1|^2) - softnorm(|p4-p3|^2)  (constraints.py:466-477, jacobians.py:322-367)", fam_equal_distance),
    Family("point_on_line", 1, 6, "softnorm(|(p-p0) x d|^2); c = [p0, d]  (constraints.py:560-576, jacobians.py:372-403)", fam_point_on_line),
    Family("linear_point", 1, 6, "n.(p - p0); c = [p0, n] -- point-on-plane, fixed-axis and the point-on-line pin rows  (constraints.py:508-516, :616-627; sensitivity.py:146-174)", fam_linear_point),
    Family("midpoint_on_plane", 2, 6, "n.(a + (b-a)/2 - p0); c = [p0, n]  (constraints.py:657-666, solver.py:439-448)", fam_midpoint_on_plane),
    Family("coplanar", 4, 0, "(p2-p1).((p3-p1) x (p4-p1))  (constraints.py:698-709, jacobians.py:426-483)", fam_coplanar),
    Family("scalar_triple", 4, 2, "(triple - V) * inv_scale; c = [V, 1/scale]  (constraints.py:731-733, solver.py:463-472)", fam_scalar_triple),
]


class CudaPrinter(C99CodePrinter):
    """Integer powers as products, x**-1/2 as OKIN_RSQRT, x**(3/2) as x*sqrt(x)."""

    def _print_Pow(self, expr):
        base, exp = expr.base, expr.exp
        b = self.parenthesize(base, sp.printing.precedence.PRECEDENCE["Mul"] + 1)
        if exp == sp.Rational(-1, 2):
            return f"OKIN_RSQRT({self._print(base)})"
        if exp == sp.Rational(1, 2):
            return f"sqrt({self._print(base)})"
        if exp == sp.Rational(3, 2):
            return f"({b}*sqrt({self._print(base)}))"
        if exp == sp.Rational(-3, 2):
            return f"(OKIN_RSQRT({self._print(base)})/{b})"
        if exp.is_Integer and 1 < int(exp) <= 4:
            return "(" + "*".join([b] * int(exp)) + ")"
        if exp == -1:
            return f"(1.0/{b})"
        if exp.is_Integer and -4 <= int(exp) < -1:
            return "(1.0/(" + "*".join([b] * (-int(exp))) + "))"
        return super()._print_Pow(expr)

    def _print_Rational(self, expr):
        return f"({int(expr.p)}.0/{int(expr.q)}.0)"


PRINTER = CudaPrinter()


def emit(exprs, outputs, symbols_in):
    """CSE the expressions; return C statements assigning ``outputs``."""
    tsyms = sp.numbered_symbols("t")
    repl, red = sp.cse(exprs, symbols=tsyms, optimizations="basic")
    lines = [f"  const double {PRINTER.doprint(s)} = {PRINTER.doprint(e)};" for s, e in repl]
    for out, e in zip(outputs, red):
        lines.append(f"  {out} = {PRINTER.doprint(e)};")
    return lines


def load_lines(f: Family):
    lines = []
    for k, s in enumerate(f.p):
        lines.append(f"  const double {s} = p[{k}];")
    for k, s in enumerate(f.c):
        lines.append(f"  const double {s} = c[{k}];")
    return lines


def generate() -> str:
    out = []
    out.append("// GENERATED by tools/generate_jacobians.py -- do not edit by hand.")
    out.append("// fp64 residual + analytical-gradient device functions, one pair per constraint family.")
    out.append("// Reference formulas: src/kinematics/core/constraints.py, core/jacobians.py (cited per family).")
    out.append("#pragma once")
    out.append('#include "okin_defs.h"')
    out.append("")
    for f in FAMILIES:
        grads = [sp.diff(f.smooth_value, v) for v in f.p]
        out.append(f"// {f.name}: {f.doc}")
        out.append(f"OKIN_HD double okin_{f.name}_res(const double* __restrict__ p, const double* __restrict__ c) {{")
        out += load_lines(f)
        out.append("  double r;")
        out += emit([f.value], ["r"], f.p + f.c)
        if not f.c:
            out.append("  (void)c;")
        out.append("  return r;")
        out.append("}")
        out.append(f"OKIN_HD double okin_{f.name}_resgrad(const double* __restrict__ p, const double* __restrict__ c, double* __restrict__ g) {{")
        out += load_lines(f)
        out.append("  double r;")
        out += emit([f.value] + grads, ["r"] + [f"g[{k}]" for k in range(len(f.p))], f.p + f.c)
        if not f.c:
            out.append("  (void)c;")
        out.append("  return r;")
        out.append("}")
        out.append("")
    # dispatch tables
    out.append("// Family codes (must match OKIN_FAM_* in okin_defs.h and core/topology.py).")
    out.append("OKIN_HD double okin_family_res(int fam, const double* p, const double* c) {")
    out.append("  switch (fam) {")
    for f in FAMILIES:
        out.append(f"    case OKIN_FAM_{f.name.upper()}: return okin_{f.name}_res(p, c);")
    out.append("    default: return 0.0;")
    out.append("  }")
    out.append("}")
    out.append("OKIN_HD double okin_family_resgrad(int fam, const double* p, const double* c, double* g) {")
    out.append("  switch (fam) {")
    for f in FAMILIES:
        out.append(f"    case OKIN_FAM_{f.name.upper()}: return okin_{f.name}_resgrad(p, c, g);")
    out.append("    default: return 0.0;")
    out.append("  }")
    out.append("}")
    out.append("")
    out.append("// points / constants per family, indexed by family code")
    out.append("#define OKIN_FAMILY_TABLE(X) \\")
    for i, f in enumerate(FAMILIES):
        tail = " \\" if i + 1 < len(FAMILIES) else ""
        out.append(f"  X({f.name.upper()}, {i}, {f.n_points}, {f.n_consts}){tail}")
    out.append("")
    return "\n".join(out)


def main() -> None:
    here = os.path.dirname(os.path.abspath(__file__))
    target = os.path.join(here, "..", "open-kinematics_b200", "csrc", "okin_gen_constraints.cuh")
    text = generate()
    if len(sys.argv) > 1 and sys.argv[1] == "--stdout":
        print(text)
        return
    with open(target, "w", encoding="utf-8") as fh:
        fh.write(text)
    print(f"wrote {os.path.normpath(target)} ({len(text.splitlines())} lines)")


if __name__ == "__main__":
    main()
