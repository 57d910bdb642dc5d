#!/usr/bin/env python
"""Generates tests/golden/golden_small.json — an INDEPENDENT pin for the C++ oracle.

kryst is pure Rust and cannot be built in this image (no cargo/rustc), so the reference itself cannot
produce vectors here.  This script is a second, independent line-by-line transliteration of the reference's
algorithms in dense numpy (left-to-right Python sums instead of the oracle's canonical tree), following:
  PCG        src/solver/pcg.rs:114-222          GMRES (None/Left/Right, MGS+2nd pass) src/solver/gmres.rs:216-402
  BiCGStab   src/solver/bicgstab.rs:69-293      Jacobi src/preconditioner/jacobi.rs:69-95
  Ilu0 (dense, literal) src/preconditioner/ilu.rs:59-122       Convergence::check src/utils/convergence.rs:18-34
  FGMRES (classical GS, flexible right pc) src/solver/fgmres.rs:114-340
plus the Chronopoulos-Gear single-reduction PCG (the reference's flag of that name is a no-op),
plus a dense textbook ILU(0) (Saad Alg. 10.4, IKJ) used to pin the Tier-T factorisation.
The JSON holds iteration counts, final residuals and solutions for small problems; tests/test_oracle_golden.py
requires the C++ oracle to reproduce them (counts exactly, floats to 1e-10: the reduction order differs).
Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import math
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def dot(a, b):
    s = 0.0
    for x, y in zip(a, b):
        s = s + x * y
    return s


def norm(a):
    return math.sqrt(dot(a, a))


def check(res, res0, i, tol, max_iters):
    rel = res / res0
    return (rel <= tol) or (i >= max_iters)


def jacobi_inv(a):
    return np.array([1.0 / a[i, i] if a[i, i] != 0.0 else 0.0 for i in range(a.shape[0])])


def pcg(a, inv, b, x0, tol, max_iters):
    x = np.array(x0, dtype=float)
    r = b - a @ x
    z = inv * r if inv is not None else r.copy()
    p = z.copy()
    rz = dot(r, z)
    res0 = math.sqrt(abs(rz))
    hist = [norm(r)]
    it, res, conv = 0, res0, False
    for i in range(max_iters):
        ap = a @ p
        pap = dot(p, ap)
        if pap <= 0.0:
            return dict(err="IndefiniteMatrix", iterations=i + 1)
        alpha = rz / pap
        x = x + alpha * p
        r = r - alpha * ap
        z = inv * r if inv is not None else r.copy()
        rz_new = dot(r, z)
        res = norm(r)
        hist.append(res)
        it = i + 1
        conv = check(res, res0, it, tol, max_iters)
        if conv:
            break
        beta = rz_new / rz
        if beta < 0.0:
            return dict(err="IndefinitePreconditioner", iterations=i + 1)
        p = z + beta * p
        rz = rz_new
    return dict(iterations=it, final_residual=res, converged=conv, x=x.tolist(), history=hist)


def pcg_single_reduction(a, inv, b, x0, tol, max_iters):
    """Chronopoulos-Gear PCG with the literal PCG's conventions (res0 = sqrt|r.u|, history, Convergence::check)."""
    x = np.array(x0, dtype=float)
    r = b - a @ x
    u = inv * r if inv is not None else r.copy()
    w = a @ u
    gamma = dot(r, u)
    delta = dot(u, w)
    res0 = math.sqrt(abs(gamma))
    hist = [norm(r)]
    if max_iters == 0:
        return dict(iterations=0, final_residual=res0, converged=False, x=x.tolist(), history=hist)
    if delta <= 0.0:
        return dict(err="IndefiniteMatrix", iterations=1)
    alpha, beta = gamma / delta, 0.0
    p = np.zeros_like(x)
    s = np.zeros_like(x)
    it, res, conv = 0, res0, False
    for i in range(max_iters):
        p = u + beta * p
        s = w + beta * s
        x = x + alpha * p
        r = r - alpha * s
        u = inv * r if inv is not None else r.copy()
        gamma_new = dot(r, u)
        res = norm(r)
        w = a @ u
        delta = dot(u, w)
        hist.append(res)
        it = i + 1
        conv = check(res, res0, it, tol, max_iters)
        if conv:
            break
        beta = gamma_new / gamma
        if beta < 0.0:
            return dict(err="IndefinitePreconditioner", iterations=i + 1)
        pap = delta - beta * gamma_new / alpha
        if pap <= 0.0:
            return dict(err="IndefiniteMatrix", iterations=i + 2)
        alpha = gamma_new / pap
        gamma = gamma_new
    return dict(iterations=it, final_residual=res, converged=conv, x=x.tolist(), history=hist)


def fgmres_literal(a, pc_apply, b, x0, tol, max_iters, restart, haptol=1e-12):
    """fgmres.rs:114-340 with the defaults of FgmresSolver::new (Orthog::Classical, preallocate = false)."""
    n = len(b)
    x = np.array(x0, dtype=float)
    r = b - a @ x
    beta = norm(r)
    if beta == 0.0:
        return dict(iterations=0, final_residual=0.0, converged=True, x=x.tolist())
    v = [np.zeros(n) for _ in range(restart + 1)]
    z = [np.zeros(n) for _ in range(restart)]
    h = [[0.0] * restart for _ in range(restart + 1)]
    cs, sn, s = [0.0] * restart, [0.0] * restart, [0.0] * (restart + 1)
    s[0] = beta
    v[0] = r / beta
    total = 0
    res_norm_outer = beta                      # fgmres.rs:171: the value the epilogue reports (:334)
    stats = dict(iterations=0, final_residual=res_norm_outer, converged=False)
    while total < max_iters:
        m = min(restart, max_iters - total)
        converged = False
        steps = m
        for j in range(m):
            z[j] = pc_apply(v[j]) if pc_apply is not None else v[j].copy()
            w = a @ z[j]
            hcol = [dot(w, v[i]) for i in range(j + 1)]
            for i in range(j + 1):
                w = w - hcol[i] * v[i]
            h[j + 1][j] = norm(w)
            for i in range(j + 1):
                h[i][j] = hcol[i]
            if not (abs(h[j + 1][j]) < haptol * abs(s[j])):
                v[j + 1] = w / h[j + 1][j]
            else:
                v[j + 1] = np.zeros(n)
            for i in range(j):
                t = cs[i] * h[i][j] + sn[i] * h[i + 1][j]
                h[i + 1][j] = -sn[i] * h[i][j] + cs[i] * h[i + 1][j]
                h[i][j] = t
            h1, h2 = h[j][j], h[j + 1][j]
            den = math.sqrt(h1 * h1 + h2 * h2)
            c, s_ = (1.0, 0.0) if den == 0.0 else (h1 / den, h2 / den)
            cs[j], sn[j] = c, s_
            t = c * s[j] + s_ * s[j + 1]
            s[j + 1] = -s_ * s[j] + c * s[j + 1]
            s[j] = t
            h[j][j] = c * h[j][j] + s_ * h[j + 1][j]
            h[j + 1][j] = 0.0
            res = abs(s[j + 1])
            total += 1
            stop = check(res, s[0], total, tol, max_iters)      # NB: s[0] was just rotated when j == 0 (fgmres.rs:290)
            stats = dict(iterations=total, final_residual=res, converged=stop)
            if stop:
                steps = j + 1
                converged = True
                break
        k = steps
        y = [0.0] * k
        for i in reversed(range(k)):
            acc = s[i]
            for l in range(i + 1, k):
                acc = acc - h[i][l] * y[l]
            y[i] = acc / h[i][i]
        for i in range(k):
            x = x + y[i] * z[i]
        r_new = b - a @ x
        rn = norm(r_new)
        if rn < tol or converged:
            stats = dict(iterations=total, final_residual=rn, converged=True)
            break
        beta = rn
        v[0] = r_new / beta
        s = [0.0] * (restart + 1)
        s[0] = beta
    stats["final_residual"] = res_norm_outer   # fgmres.rs:334-335 (shadowed variable: always ||r0||)
    stats["iterations"] = total
    stats["x"] = x.tolist()
    return stats


def gmres_literal(a, pc_apply, b, x0, restart, tol, max_iters, mode):
    n = len(b)
    xk = np.array(x0, dtype=float)
    r0 = b - a @ xk
    beta = norm(r0)
    res0 = beta
    stats = dict(iterations=0, final_residual=beta, converged=False)
    eps = 1e-14
    n_outer = -(-max_iters // restart)
    iteration = 0
    if pc_apply is None:
        mode = "none"
    for _ in range(n_outer):
        V, Z = [], []
        r0_norm = beta
        if mode == "left":
            v0 = r0 / r0_norm
            V.append(v0)
            Z.append(pc_apply(v0))
        elif mode == "right":
            z0 = pc_apply(r0)
            r0_norm = norm(z0)
            v0 = z0 / r0_norm
            V.append(v0)
            Z.append(pc_apply(v0))
            beta = r0_norm
        else:
            V.append(r0 / r0_norm)
        h = np.zeros((restart + 1, restart))
        g = np.zeros(restart + 1)
        g[0] = r0_norm
        cs, sn = np.zeros(restart), np.zeros(restart)
        m, happy = 0, False
        for j in range(restart):
            iteration += 1
            if mode == "left":
                w = pc_apply(a @ V[j])
                basis = Z
            elif mode == "right":
                w = a @ pc_apply(V[j])
                basis = V
            else:
                w = a @ V[j]
                basis = V
            for i in range(j + 1):
                h[i, j] = dot(w, basis[i])
                w = w - h[i, j] * basis[i]
            for i in range(j + 1):
                t = dot(w, basis[i])
                h[i, j] += t
                w = w - t * basis[i]
            h[j + 1, j] = norm(w)
            if abs(h[j + 1, j]) < eps:
                happy = True
                if mode != "none":
                    break
            else:
                vj1 = w / h[j + 1, j]
                V.append(vj1)
                if mode == "left":
                    Z.append(vj1)
                if mode == "right":
                    Z.append(pc_apply(vj1))
            for i in range(j):
                temp = cs[i] * h[i, j] + sn[i] * h[i + 1, j]
                h[i + 1, j] = -sn[i] * h[i, j] + cs[i] * h[i + 1, j]
                h[i, j] = temp
            hkk, hk1k = h[j, j], h[j + 1, j]
            r = math.sqrt(hkk * hkk + hk1k * hk1k)
            if abs(r) < eps:
                cs[j], sn[j] = 1.0, 0.0
            else:
                cs[j], sn[j] = hkk / r, hk1k / r
            h[j, j] = cs[j] * hkk + sn[j] * hk1k
            h[j + 1, j] = 0.0
            temp = cs[j] * g[j] + sn[j] * g[j + 1]
            g[j + 1] = -sn[j] * g[j] + cs[j] * g[j + 1]
            g[j] = temp
            res_norm = abs(g[j + 1])
            stop = check(res_norm, res0, iteration, tol, max_iters)
            stats = dict(iterations=iteration, final_residual=res_norm, converged=stop)
            m = j + 1
            if stop or happy:
                break
        y = np.zeros(m)
        for i in reversed(range(m)):
            y[i] = g[i]
            for jj in range(i + 1, m):
                y[i] -= h[i, jj] * y[jj]
            y[i] = y[i] / h[i, i] if abs(h[i, i]) > eps else 0.0
        upd = Z if mode == "right" else V
        for j in range(m):
            xk = xk + y[j] * upd[j]
        r0 = b - a @ xk
        beta = norm(r0)
        stats["final_residual"] = beta
        stats["converged"] = bool(beta < tol * res0)
        if stats["converged"] or iteration >= max_iters:
            break
    stats["x"] = xk.tolist()
    return stats


def bicgstab(a, b, x0, tol, max_iters):
    eps = np.finfo(float).eps
    xk = np.array(x0, dtype=float)
    r = b - a @ xk
    rhat = r.copy()
    rho_prev = alpha = omega_prev = 1.0
    v = np.zeros_like(r)
    p = r.copy()
    res0 = norm(r)
    stats = dict(iterations=0, final_residual=res0, converged=False)
    if res0 <= tol:
        stats["converged"] = True
        stats["x"] = xk.tolist()
        return stats
    for i in range(1, max_iters + 1):
        rho = dot(rhat, r)
        if abs(rho) < eps:
            break
        beta = 0.0 if i == 1 else (rho / rho_prev) * (alpha / omega_prev)
        p = r + beta * (p - omega_prev * v)
        v = a @ p
        den = dot(rhat, v)
        if abs(den) < eps:
            break
        alpha = rho / den
        s = r - alpha * v
        sn = norm(s)
        if sn <= tol:
            xk = xk + alpha * p
            stats = dict(iterations=i, final_residual=sn, converged=True)
            break
        t = a @ s
        on, od = dot(t, s), dot(t, t)
        if abs(od) < eps:
            break
        omega = on / od
        xk = xk + alpha * p + omega * s
        r = s - omega * t
        rn = norm(r)
        stats = dict(iterations=i, final_residual=rn, converged=bool(rn <= tol))
        if rn <= tol:
            break
        if abs(omega) < eps:
            break
        rho_prev, omega_prev = rho, omega
    stats["x"] = xk.tolist()
    return stats


def ilu_literal(a):
    n = a.shape[0]
    l, u = np.zeros((n, n)), np.zeros((n, n))
    for i in range(n):
        u[i, i] = a[i, i]
        for j in range(i + 1, n):
            if a[i, j] != 0.0:
                u[i, j] = a[i, j]
        l[i, i] = 1.0
        for j in range(i + 1, n):
            if a[j, i] != 0.0:
                l[j, i] = a[j, i] / u[i, i]
        for j in range(i + 1, n):
            for k in range(i + 1, n):
                if a[j, k] != 0.0:
                    v = a[j, k] - l[j, i] * u[i, k]
                    if v != 0.0:
                        if k >= j:
                            u[j, k] = v
                        else:
                            l[j, k] = v
    return l, u


def ilu_literal_apply(l, u, x):
    n = len(x)
    y = np.array(x, dtype=float)
    for i in range(n):
        for j in range(i):
            y[i] -= l[i, j] * y[j]
    for i in reversed(range(n)):
        for j in range(i + 1, n):
            y[i] -= u[i, j] * y[j]
    return y


def ilu0_textbook_dense(a):
    """Saad Alg. 10.4 (IKJ) restricted to the pattern of a; returns the combined LU matrix."""
    n = a.shape[0]
    lu = a.copy()
    pat = a != 0.0
    for i in range(1, n):
        for k in range(i):
            if pat[i, k]:
                lu[i, k] = lu[i, k] / lu[k, k]
                for j in range(k + 1, n):
                    if pat[i, j] and pat[k, j]:
                        lu[i, j] = lu[i, j] - lu[i, k] * lu[k, j]
    return lu


def tridiag(n, lo, d, up):
    a = np.zeros((n, n))
    for i in range(n):
        a[i, i] = d
        if i > 0:
            a[i, i - 1] = lo
        if i + 1 < n:
            a[i, i + 1] = up
    return a


def poisson2d(N):
    n = N * N
    a = np.zeros((n, n))
    for r in range(n):
        i, j = r % N, r // N
        a[r, r] = 4.0
        if i > 0: a[r, r - 1] = -1.0
        if i < N - 1: a[r, r + 1] = -1.0
        if j > 0: a[r, r - N] = -1.0
        if j < N - 1: a[r, r + N] = -1.0
    return a


def convdiff2d(N, px=0.4, py=0.2):
    n = N * N
    a = np.zeros((n, n))
    for r in range(n):
        i, j = r % N, r // N
        a[r, r] = (4.0 + px) + py
        if i > 0: a[r, r - 1] = -(1.0 + px)
        if i < N - 1: a[r, r + 1] = -1.0
        if j > 0: a[r, r - N] = -(1.0 + py)
        if j < N - 1: a[r, r + N] = -1.0
    return a


def rowscaled(a):
    """Row scaling with a non-constant factor: makes Jacobi a genuinely different preconditioner from the identity."""
    return np.diag(1.0 + 0.05 * np.arange(a.shape[0])) @ a


def main():
    out = {}
    # PCG + Jacobi / no pc on small Poisson and the reference's tridiagonal
    for name, a in (("poisson2d_8", poisson2d(8)), ("poisson2d_12", poisson2d(12)), ("tridiag_spd_10", tridiag(10, -1, 2, -1))):
        b = a @ np.ones(a.shape[0])
        out["pcg_jacobi/" + name] = pcg(a, jacobi_inv(a), b, np.zeros(len(b)), 1e-10, 500)
        out["pcg_none/" + name] = pcg(a, None, b, np.zeros(len(b)), 1e-10, 500)
    # GMRES literal, three modes with Jacobi, restart shorter than convergence
    for name, a in (("convdiff2d_8", convdiff2d(8)), ("tridiag_nonsym_10", tridiag(10, -1, 2, 0.5))):
        b = a @ np.ones(a.shape[0])
        inv = jacobi_inv(a)
        for mode in ("none", "left", "right"):
            out["gmres_literal_%s_jacobi/%s" % (mode, name)] = gmres_literal(
                a, (lambda v, inv=inv: inv * v) if mode != "none" else None, b, np.zeros(len(b)), 5, 1e-9, 400, mode)
    # GMRES literal left + literal dense Ilu0 (tests/preconditioner_integration.rs:169-179)
    a = tridiag(10, -1, 2, 0.5)
    l, u = ilu_literal(a)
    b = a @ np.ones(10)
    out["gmres_literal_left_iluliteral/tridiag_nonsym_10"] = gmres_literal(a, lambda v: ilu_literal_apply(l, u, v), b, np.zeros(10), 10, 1e-12, 100, "left")
    out["ilu_literal_apply/tridiag_nonsym_6"] = dict(z=ilu_literal_apply(*ilu_literal(tridiag(6, -1, 2, 0.5)), np.arange(1.0, 7.0)).tolist())
    # single-reduction PCG and literal FGMRES
    for name, a in (("poisson2d_8", poisson2d(8)), ("tridiag_spd_10", tridiag(10, -1, 2, -1))):
        b = a @ np.ones(a.shape[0])
        out["pcg_sr_jacobi/" + name] = pcg_single_reduction(a, jacobi_inv(a), b, np.zeros(len(b)), 1e-10, 500)
        out["pcg_sr_none/" + name] = pcg_single_reduction(a, None, b, np.zeros(len(b)), 1e-10, 500)
    for name, a in (("convdiff2d_8", convdiff2d(8)), ("tridiag_nonsym_10", tridiag(10, -1, 2, 0.5)),
                    ("rowscaled_convdiff2d_8", rowscaled(convdiff2d(8)))):
        b = a @ np.ones(a.shape[0])
        inv = jacobi_inv(a)
        out["fgmres_literal_jacobi/" + name] = fgmres_literal(a, lambda v, inv=inv: inv * v, b, np.zeros(len(b)), 1e-9, 400, 5)
        out["fgmres_literal_none/" + name] = fgmres_literal(a, None, b, np.zeros(len(b)), 1e-9, 400, 5)
    # BiCGStab literal
    for name, a in (("convdiff2d_8", convdiff2d(8)),):
        b = a @ np.ones(a.shape[0])
        out["bicgstab_literal/" + name] = bicgstab(a, b, np.zeros(len(b)), 1e-9, 400)
    # textbook ILU(0) factors (dense IKJ on the pattern)
    for name, a in (("convdiff2d_6", convdiff2d(6)), ("poisson2d_5", poisson2d(5)), ("tridiag_nonsym_8", tridiag(8, -1, 2, 0.5))):
        out["ilu0_textbook/" + name] = dict(lu=ilu0_textbook_dense(a).tolist())
    with open(os.path.join(HERE, "golden_small.json"), "w") as f:
        json.dump(out, f)
    print("wrote", len(out), "cases")


if __name__ == "__main__":
    main()
