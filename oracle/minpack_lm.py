"""NumPy restatement of MINPACK ``lmder`` (Levenberg-Marquardt with analytic Jacobian).

The reference's solve iteration is third-party code absent from ``/root/reference``:
``scipy.optimize.least_squares(method="lm")`` -> ``scipy.optimize._minpack._lmder`` (reference call
sites ``core/solver.py:13``, ``:169``, ``:717-724``; scipy pinned 1.14.1 in the reference's
``uv.lock``, 1.18.1 installed here).  This file restates the published algorithm -- J. J. More,
B. S. Garbow, K. E. Hillstrom, "User Guide for MINPACK-1", ANL-80-74, routines ``lmder``,
``lmpar``, ``qrfac``, ``qrsolv``, ``enorm`` -- with SciPy's wrapper conventions
(``least_squares.py:46-99``): ``factor = 100``, ``maxfev = 100 n``, ``diag=None`` i.e. MINPACK
``mode = 1`` internal column scaling (SciPy >= 1.16 default ``x_scale='jac'``) or a fixed
``diag`` (``mode = 2``; SciPy 1.14.1's ``x_scale = 1``).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  ``tests/test_oracle_minpack.py`` pins it
against SciPy's own ``_lmder`` on the reference's problems (same ``info``, ``nfev``, iterate).
"""

from __future__ import annotations

import numpy as np

EPSMCH = np.finfo(np.float64).eps
DWARF = np.finfo(np.float64).tiny


def enorm(v) -> float:
    return float(np.sqrt(np.dot(v, v)))


def qrfac(a: np.ndarray):
    """Householder QR with column pivoting, in place (MINPACK qrfac, pivot = true).
    Returns (ipvt, rdiag, acnorm); ``a`` holds R's strict upper triangle and the
    Householder vectors in its lower trapezoid."""
    m, n = a.shape
    acnorm = np.array([enorm(a[:, j]) for j in range(n)])
    rdiag = acnorm.copy()
    wa = rdiag.copy()
    ipvt = np.arange(n)
    for j in range(min(m, n)):
        kmax = j + int(np.argmax(rdiag[j:]))
        if kmax != j:
            a[:, [j, kmax]] = a[:, [kmax, j]]
            rdiag[kmax] = rdiag[j]
            wa[kmax] = wa[j]
            ipvt[[j, kmax]] = ipvt[[kmax, j]]
        ajnorm = enorm(a[j:, j])
        if ajnorm != 0.0:
            if a[j, j] < 0.0:
                ajnorm = -ajnorm
            a[j:, j] /= ajnorm
            a[j, j] += 1.0
            for k in range(j + 1, n):
                temp = float(np.dot(a[j:, j], a[j:, k])) / a[j, j]
                a[j:, k] -= temp * a[j:, j]
                if rdiag[k] != 0.0:
                    temp = a[j, k] / rdiag[k]
                    rdiag[k] *= np.sqrt(max(0.0, 1.0 - temp * temp))
                    if 0.05 * (rdiag[k] / wa[k]) ** 2 <= EPSMCH:
                        rdiag[k] = enorm(a[j + 1:, k])
                        wa[k] = rdiag[k]
        rdiag[j] = -ajnorm
    return ipvt, rdiag, acnorm


def qrsolv(r: np.ndarray, ipvt, diag, qtb):
    """Solve ``[R P^T; D P^T] x ~ [Q^T b; 0]`` by Givens rotations (MINPACK qrsolv).
    ``r``: n x n with R in the upper triangle (modified: strict lower triangle receives S).
    Returns (x, sdiag)."""
    n = r.shape[0]
    for j in range(n):
        r[j:, j] = r[j, j:]
    x = np.diag(r).copy()
    wa = np.array(qtb[:n], dtype=np.float64)
    sdiag = np.zeros(n)
    for j in range(n):
        l = ipvt[j]
        if diag[l] != 0.0:
            sdiag[j:] = 0.0
            sdiag[j] = diag[l]
            qtbpj = 0.0
            for k in range(j, n):
                if sdiag[k] == 0.0:
                    continue
                if abs(r[k, k]) < abs(sdiag[k]):
                    cotan = r[k, k] / sdiag[k]
                    sin = 0.5 / np.sqrt(0.25 + 0.25 * cotan * cotan)
                    cos = sin * cotan
                else:
                    tan = sdiag[k] / r[k, k]
                    cos = 0.5 / np.sqrt(0.25 + 0.25 * tan * tan)
                    sin = cos * tan
                r[k, k] = cos * r[k, k] + sin * sdiag[k]
                temp = cos * wa[k] + sin * qtbpj
                qtbpj = -sin * wa[k] + cos * qtbpj
                wa[k] = temp
                if k + 1 < n:
                    col = r[k + 1:, k].copy()
                    r[k + 1:, k] = cos * col + sin * sdiag[k + 1:]
                    sdiag[k + 1:] = -sin * col + cos * sdiag[k + 1:]
        sdiag[j] = r[j, j]
        r[j, j] = x[j]
    nsing = n
    for j in range(n):
        if sdiag[j] == 0.0 and nsing == n:
            nsing = j
        if nsing < n:
            wa[j] = 0.0
    for k in range(nsing):
        j = nsing - 1 - k
        s = float(np.dot(r[j + 1:nsing, j], wa[j + 1:nsing]))
        wa[j] = (wa[j] - s) / sdiag[j]
    out = np.zeros(n)
    out[ipvt] = wa
    return out, sdiag


def lmpar(r: np.ndarray, ipvt, diag, qtb, delta: float, par: float):
    """Levenberg-Marquardt parameter (MINPACK lmpar).  Returns (par, x, sdiag)."""
    n = r.shape[0]
    nsing = n
    wa1 = np.array(qtb[:n], dtype=np.float64)
    for j in range(n):
        if r[j, j] == 0.0 and nsing == n:
            nsing = j
        if nsing < n:
            wa1[j] = 0.0
    for k in range(nsing):
        j = nsing - 1 - k
        wa1[j] /= r[j, j]
        wa1[:j] -= r[:j, j] * wa1[j]
    x = np.zeros(n)
    x[ipvt] = wa1
    it = 0
    wa2 = diag * x
    dxnorm = enorm(wa2)
    fp = dxnorm - delta
    sdiag = np.zeros(n)
    if fp <= 0.1 * delta:
        return 0.0, x, sdiag
    parl = 0.0
    if nsing >= n:
        wa1 = diag[ipvt] * (wa2[ipvt] / dxnorm)
        for j in range(n):
            s = float(np.dot(r[:j, j], wa1[:j]))
            wa1[j] = (wa1[j] - s) / r[j, j]
        temp = enorm(wa1)
        parl = ((fp / delta) / temp) / temp
    wa1 = np.array([float(np.dot(r[: j + 1, j], qtb[: j + 1])) / diag[ipvt[j]] for j in range(n)])
    gnorm = enorm(wa1)
    paru = gnorm / delta
    if paru == 0.0:
        paru = DWARF / min(delta, 0.1)
    par = min(max(par, parl), paru)
    if par == 0.0:
        par = gnorm / dxnorm
    while True:
        it += 1
        if par == 0.0:
            par = max(DWARF, 0.001 * paru)
        wa1 = np.sqrt(par) * diag
        x, sdiag = qrsolv(r, ipvt, wa1, qtb)
        wa2 = diag * x
        dxnorm = enorm(wa2)
        temp = fp
        fp = dxnorm - delta
        if abs(fp) <= 0.1 * delta or (parl == 0.0 and fp <= temp and temp < 0.0) or it == 10:
            break
        wa1 = diag[ipvt] * (wa2[ipvt] / dxnorm)
        for j in range(n):
            wa1[j] /= sdiag[j]
            wa1[j + 1:] -= r[j + 1:, j] * wa1[j]
        temp = enorm(wa1)
        parc = ((fp / delta) / temp) / temp
        if fp > 0.0:
            parl = max(parl, par)
        if fp < 0.0:
            paru = min(paru, par)
        par = max(parl, par + parc)
    return par, x, sdiag


def lmder(fun, jac, x0, ftol=1e-5, xtol=1e-9, gtol=1e-9, maxfev=None, factor=100.0, diag=None):
    """MINPACK lmder.  Returns ``(x, fvec, info, nfev, njev)`` with MINPACK's ``info`` codes
    (1 ftol, 2 xtol, 3 both, 4 gtol, 5 maxfev, 6-8 tolerances too small)."""
    x = np.array(x0, dtype=np.float64)
    n = x.size
    mode = 1 if diag is None else 2
    diag = np.ones(n) if diag is None else np.array(diag, dtype=np.float64)
    maxfev = 100 * n if maxfev is None else maxfev
    fvec = np.array(fun(x), dtype=np.float64)
    m = fvec.size
    nfev, njev = 1, 0
    fnorm = enorm(fvec)
    par, it, info = 0.0, 1, 0
    xnorm = delta = 0.0
    while True:
        fjac = np.array(jac(x), dtype=np.float64)
        njev += 1
        ipvt, rdiag, acnorm = qrfac(fjac)
        if it == 1:
            if mode != 2:
                diag = np.where(acnorm == 0.0, 1.0, acnorm)
            xnorm = enorm(diag * x)
            delta = factor * xnorm
            if delta == 0.0:
                delta = factor
        # (Q^T fvec)[:n] and R with its diagonal restored
        wa4 = fvec.copy()
        qtf = np.zeros(n)
        for j in range(n):
            if fjac[j, j] != 0.0:
                temp = -float(np.dot(fjac[j:, j], wa4[j:])) / fjac[j, j]
                wa4[j:] += fjac[j:, j] * temp
            fjac[j, j] = rdiag[j]
            qtf[j] = wa4[j]
        r = np.triu(fjac[:n, :n]).copy()
        gnorm = 0.0
        if fnorm != 0.0:
            for j in range(n):
                l = ipvt[j]
                if acnorm[l] != 0.0:
                    s = float(np.dot(r[: j + 1, j], qtf[: j + 1] / fnorm))
                    gnorm = max(gnorm, abs(s / acnorm[l]))
        if gnorm <= gtol:
            info = 4
            break
        if mode != 2:
            diag = np.maximum(diag, acnorm)
        while True:
            par, p, _ = lmpar(r.copy(), ipvt, diag, qtf, delta, par)
            wa1 = -p
            wa2 = x + wa1
            wa3 = diag * wa1
            pnorm = enorm(wa3)
            if it == 1:
                delta = min(delta, pnorm)
            wa4 = np.array(fun(wa2), dtype=np.float64)
            nfev += 1
            fnorm1 = enorm(wa4)
            actred = -1.0
            if 0.1 * fnorm1 < fnorm:
                actred = 1.0 - (fnorm1 / fnorm) ** 2
            wa3 = r @ wa1[ipvt]
            temp1 = enorm(wa3) / fnorm
            temp2 = (np.sqrt(par) * pnorm) / fnorm
            prered = temp1 * temp1 + temp2 * temp2 / 0.5
            dirder = -(temp1 * temp1 + temp2 * temp2)
            ratio = actred / prered if prered != 0.0 else 0.0
            if ratio <= 0.25:
                temp = 0.5 if actred >= 0.0 else 0.5 * dirder / (dirder + 0.5 * actred)
                if 0.1 * fnorm1 >= fnorm or temp < 0.1:
                    temp = 0.1
                delta = temp * min(delta, pnorm / 0.1)
                par /= temp
            elif par == 0.0 or ratio >= 0.75:
                delta = pnorm / 0.5
                par *= 0.5
            if ratio >= 1e-4:
                x = wa2
                wa2 = diag * x
                fvec = wa4
                xnorm = enorm(wa2)
                fnorm = fnorm1
                it += 1
            small = abs(actred) <= ftol and prered <= ftol and 0.5 * ratio <= 1.0
            if small:
                info = 1
            if delta <= xtol * xnorm:
                info = 2
            if small and info == 2:
                info = 3
            if info != 0:
                break
            if nfev >= maxfev:
                info = 5
            if abs(actred) <= EPSMCH and prered <= EPSMCH and 0.5 * ratio <= 1.0:
                info = 6
            if delta <= EPSMCH * xnorm:
                info = 7
            if gnorm <= EPSMCH:
                info = 8
            if info != 0 or ratio >= 1e-4:
                break
        if info != 0:
            break
    return x, fvec, info, nfev, njev
