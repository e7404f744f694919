"""Reference algorithm for the hot path, restated on a neutral problem description.

Follows reference ``core/solver.py``: ``ResidualComputer.compute`` (:226-275),
``compute_jacobian`` (:502-581), ``convert_targets_to_absolute`` (:584-627),
``solve_suspension_sweep`` (:654-776); ``core/sensitivity.py:57-174`` for tangents; design
constants as the shipped topologies compute them (``suspensions/corner/double_wishbone.py:259-308``,
``corner/attachments.py:23-120``, ``corner/track_rod.py:60-97``, ``corner/macpherson.py:199-204``,
``axle/suspension.py:196-209``, ``axle/mechanisms.py:307-342``, ``:669-716``).
"""

from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
from scipy.optimize import least_squares

from . import derived as D
from . import families as F

STATUS_OK, STATUS_NOT_CONVERGED, STATUS_RESIDUAL_REJECTED, STATUS_INVALID = 0, 1, 2, 3


@dataclass
class OracleConstraint:
    family: str
    keys: tuple
    consts: list
    from_design: bool = True   # recompute the constant(s) from the instance's design pose


@dataclass
class OracleDerived:
    op: str
    out: object
    inputs: tuple
    param: float
    design_projection: object = None   # authored point projected on the op's line -> param


@dataclass
class OracleTarget:
    key: object
    direction: np.ndarray
    relative: bool = True


@dataclass
class OracleProblem:
    point_keys: list                 # every point, any order
    free_order: list                 # sorted free keys = column order (state.py:50)
    constraints: list
    derived: list                    # topological order
    targets: list
    col: dict = field(init=False)

    def __post_init__(self):
        self.col = {k: 3 * i for i, k in enumerate(self.free_order)}
        self.n = 3 * len(self.free_order)
        self.derived_by_key = {d.out: d for d in self.derived}


# ---------------------------------------------------------------------------
def update_derived(problem: OracleProblem, pos: dict, jac: bool = False) -> dict:
    """Evaluate derived points in place; with ``jac`` also return, per derived point, the
    3x3 blocks w.r.t. each *free* base point (chain rule through derived inputs)."""
    blocks: dict = {}
    for d in problem.derived:
        ins = [pos[k] for k in d.inputs]
        value, jacs = D.OPS[d.op](*ins, d.param)
        pos[d.out] = value
        if jac:
            mine: dict = {}
            for k, j in zip(d.inputs, jacs):
                if k in problem.col:
                    mine[k] = mine.get(k, 0) + j
                elif k in blocks:
                    for base, b in blocks[k].items():
                        mine[base] = mine.get(base, 0) + j @ b
            blocks[d.out] = mine
    return blocks


def design_setup(problem: OracleProblem, authored: dict) -> tuple:
    """Design pose and per-instance constants from authored hardpoints.

    Returns ``(positions, consts)`` with ``consts[i]`` the constant list of constraint i.
    True norms / angles / signed volumes, no softnorm (``geometric.py:17-28``, ``:71-104``,
    ``:197-214``).  Also fixes derived-op parameters declared as design projections.
    """
    pos = {k: np.array(v, dtype=np.float64) for k, v in authored.items()}
    for d in problem.derived:
        if d.design_projection is not None:
            a, b = pos[d.inputs[0]], pos[d.inputs[1]]
            axis = (b - a) / np.linalg.norm(b - a)
            d.param = float((pos[d.design_projection] - a) @ axis)
    update_derived(problem, pos)
    consts = []
    for c in problem.constraints:
        p = [pos[k] for k in c.keys]
        if not c.from_design:
            consts.append(list(c.consts))
        elif c.family == "distance":
            consts.append([float(np.linalg.norm(p[1] - p[0]))])
        elif c.family == "angle":
            u1 = (p[1] - p[0]) / np.linalg.norm(p[1] - p[0])
            u2 = (p[3] - p[2]) / np.linalg.norm(p[3] - p[2])
            consts.append([float(np.arctan2(np.linalg.norm(np.cross(u1, u2)), u1 @ u2))])
        elif c.family == "scalar_triple":
            v = float((p[1] - p[0]) @ np.cross(p[2] - p[0], p[3] - p[0]))
            consts.append([v, 1.0 / abs(v)])
        elif c.family == "point_on_line":
            consts.append([*p[0], *c.consts[3:6]])
        else:
            consts.append(list(c.consts))
    return pos, consts


def target_bases(problem: OracleProblem, pos: dict) -> np.ndarray:
    """``dot(p_design, dir)`` for relative targets, 0 for absolute (solver.py:612-615)."""
    return np.array([float(pos[t.key] @ t.direction) if t.relative else 0.0 for t in problem.targets])


class ResidualComputer:
    """``fun(x)`` / ``jac(x)`` pair on a working copy of the positions (solver.py:172-581)."""

    def __init__(self, problem: OracleProblem, pos: dict, consts: list):
        self.pb, self.pos, self.consts = problem, {k: v.copy() for k, v in pos.items()}, consts
        self.m = len(problem.constraints) + len(problem.targets)

    def _refresh(self, x, jac=False):
        for k, c in self.pb.col.items():
            self.pos[k] = x[c: c + 3]
        return update_derived(self.pb, self.pos, jac)

    def compute(self, x, tabs):
        self._refresh(x)
        r = np.empty(self.m)
        for i, (c, cst) in enumerate(zip(self.pb.constraints, self.consts)):
            r[i] = F.FAMILIES[c.family](np.array([self.pos[k] for k in c.keys]), cst)[0]
        off = len(self.pb.constraints)
        for j, t in enumerate(self.pb.targets):
            r[off + j] = float(self.pos[t.key] @ t.direction) - tabs[j]
        return r

    def _scatter(self, row, key, g, blocks):
        if key in self.pb.col:
            c = self.pb.col[key]
            row[c: c + 3] += g
        elif key in blocks:
            for base, b in blocks[key].items():
                c = self.pb.col[base]
                row[c: c + 3] += g @ b

    def compute_jacobian(self, x, tabs):
        blocks = self._refresh(x, jac=True)
        J = np.zeros((self.m, self.pb.n))
        for i, (c, cst) in enumerate(zip(self.pb.constraints, self.consts)):
            try:
                g = F.FAMILIES[c.family](np.array([self.pos[k] for k in c.keys]), cst)[1]
            except ZeroDivisionError:     # solver.py:539-545
                continue
            for key, gk in zip(c.keys, g):
                self._scatter(J[i], key, gk, blocks)
        off = len(self.pb.constraints)
        for j, t in enumerate(self.pb.targets):
            self._scatter(J[off + j], t.key, t.direction, blocks)
        return J


def pin_rows(problem: OracleProblem, consts: list) -> list:
    """``(column, normal, p0)`` per pin: two per point-on-line constraint (sensitivity.py:146-174)."""
    rows = []
    for c, cst in zip(problem.constraints, consts):
        if c.family != "point_on_line" or c.keys[0] not in problem.col:
            continue
        d = np.asarray(cst[3:6], dtype=np.float64)
        d = d / np.linalg.norm(d)
        least = np.zeros(3)
        least[int(np.argmin(np.abs(d)))] = 1.0
        n1 = np.cross(d, least)
        n1 /= np.linalg.norm(n1)
        n2 = np.cross(d, n1)
        for n in (n1, n2):
            rows.append((problem.col[c.keys[0]], n, np.asarray(cst[0:3])))
    return rows


def solve_sweep(problem: OracleProblem, authored: dict, values: np.ndarray, ftol=1e-5, xtol=1e-9, gtol=1e-9,
                residual_tol=1e-3, lm="scipy") -> dict:
    """The reference's sweep (solver.py:654-776) for one instance.

    ``values``: ``[T, S]`` sweep values.  Instead of raising at the first failed step the
    result carries ``status`` / ``failed_step`` (0 ok, 1 = "Solver failed to converge",
    2 = residual rejection), which is what a batch reports per instance.
    """
    pos0, consts = design_setup(problem, authored)
    bases = target_bases(problem, pos0)
    rc = ResidualComputer(problem, pos0, consts)
    x = np.concatenate([pos0[k] for k in problem.free_order])
    n_steps = values.shape[1]
    keys = list(problem.point_keys)
    out = {
        "positions": np.full((n_steps, len(keys), 3), np.nan), "nfev": np.zeros(n_steps, int),
        "max_residual": np.full(n_steps, np.nan), "status": STATUS_OK, "failed_step": -1, "keys": keys,
        "design": pos0, "consts": consts, "x": [],
    }
    for s in range(n_steps):
        tabs = bases + values[:, s]
        if lm == "scipy":
            res = least_squares(rc.compute, x, jac=rc.compute_jacobian, method="lm", ftol=ftol, xtol=xtol,
                                gtol=gtol, args=(tabs,))
            xs, fun, nfev, success = res.x, res.fun, res.nfev, res.success
        else:
            from .minpack_lm import lmder
            xs, fun, info, nfev, _ = lmder(lambda v: rc.compute(v, tabs), lambda v: rc.compute_jacobian(v, tabs),
                                           x, ftol=ftol, xtol=xtol, gtol=gtol)
            success = info in (1, 2, 3, 4)
        out["nfev"][s] = nfev
        if not success:
            out["status"], out["failed_step"] = STATUS_NOT_CONVERGED, s
            break
        rmax = float(np.max(np.abs(fun)))
        out["max_residual"][s] = rmax
        if rmax > residual_tol:
            out["status"], out["failed_step"] = STATUS_RESIDUAL_REJECTED, s
            break
        rc._refresh(xs)
        out["positions"][s] = np.array([rc.pos[k] for k in keys])
        out["x"].append(xs.copy())
        x = xs
    return out


def pinned_root(problem: OracleProblem, rc: ResidualComputer, x0: np.ndarray, tabs: np.ndarray, iters: int = 12):
    """Gauss-Newton on ``[J without point-on-line rows; pins]`` to round-off."""
    keep = [i for i, c in enumerate(problem.constraints) if c.family != "point_on_line"]
    keep += list(range(len(problem.constraints), rc.m))
    pins = pin_rows(problem, rc.consts)
    x = x0.copy()
    for _ in range(iters):
        r = rc.compute(x, tabs)[keep]
        J = rc.compute_jacobian(x, tabs)[keep]
        pr, pj = [], []
        for col, n, p0 in pins:
            pr.append(float((x[col: col + 3] - p0) @ n))
            row = np.zeros(problem.n)
            row[col: col + 3] = n
            pj.append(row)
        if pins:
            r, J = np.concatenate([r, pr]), np.vstack([J, pj])
        dx = np.linalg.lstsq(J, -r, rcond=None)[0]
        x = x + dx
        if np.max(np.abs(dx)) < 1e-13:
            break
    return x


def state_tangents(problem: OracleProblem, rc: ResidualComputer, x: np.ndarray, tabs: np.ndarray) -> dict:
    """``J dq/dt_j = e_j`` by SVD least squares on ``[J; pins]`` (sensitivity.py:57-143)."""
    J = rc.compute_jacobian(x, tabs)
    pins = pin_rows(problem, rc.consts)
    for col, n, _ in pins:
        row = np.zeros(problem.n)
        row[col: col + 3] = n
        J = np.vstack([J, row])
    nt = len(problem.targets)
    rhs = np.zeros((J.shape[0], nt))
    for j in range(nt):
        rhs[len(problem.constraints) + j, j] = 1.0
    v, _, rank, sv = np.linalg.lstsq(J, rhs, rcond=None)
    return {"tangents": v.T.copy(), "rank": int(rank), "smallest_singular_value": float(sv[-1]),
            "condition_number": float(sv[0] / sv[-1]) if sv[-1] > 0 else math.inf}
