"""NumPy/SciPy oracle for the problems and solvers on top of the operators.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  Restates

* convex_envelope.operator / solve     /root/reference/convex_envelope.py:8-123
* pucci.operator / solve               /root/reference/pucci.py:7-160
* solvers.euler / NewtonEulerLS        /root/reference/solvers.py:13-233
* scipy.sparse.linalg.bicgstab         SciPy 1.18.1 _isolve/iterative.py (third
  party, not under /root/reference; the reference pins no version.  The
  restatement below is checked against the installed SciPy by
  ``oracle/make_golden.py``; the reference itself has no solver test, so solver
  parity is pinned only by outputs of the reference run in the build container.)

``dd`` is the operator module to use (``oracle.ddg_oracle`` by default).
"""
import itertools
import time
import warnings

import numpy as np
import scipy.sparse as sp

from . import ddg_oracle


# ---------------------------------------------------------------------------
# SciPy's BiCGStab, restated (iterative.py, `bicgstab`)
# ---------------------------------------------------------------------------
def bicgstab(A, b, rtol=1e-5, atol=0.0, maxiter=None, M=None):
    """x0 = 0.  Returns (x, info, iterations)."""
    n = len(b)
    matvec = A.dot
    psolve = (lambda v: v) if M is None else M.dot
    bnrm2 = np.linalg.norm(b)
    atol = max(float(atol), float(rtol) * float(bnrm2))
    if bnrm2 == 0:
        return b, 0, 0
    if maxiter is None:
        maxiter = n * 10
    rhotol = np.finfo(np.float64).eps ** 2
    omegatol = rhotol
    x = np.zeros(n)
    r = b.copy()
    rtilde = r.copy()
    rho_prev = omega = alpha = p = v = None
    for iteration in range(maxiter):
        if np.linalg.norm(r) < atol:
            return x, 0, iteration
        rho = np.dot(rtilde, r)
        if np.abs(rho) < rhotol:
            return x, -10, iteration
        if iteration > 0:
            if np.abs(omega) < omegatol:
                return x, -11, iteration
            beta = (rho / rho_prev) * (alpha / omega)
            p -= omega * v
            p *= beta
            p += r
        else:
            s = np.empty_like(r)
            p = r.copy()
        phat = psolve(p)
        v = matvec(phat)
        rv = np.dot(rtilde, v)
        if rv == 0:
            return x, -11, iteration
        alpha = rho / rv
        r -= alpha * v
        s[:] = r[:]
        if np.linalg.norm(s) < atol:
            x += alpha * phat
            return x, 0, iteration
        shat = psolve(s)
        t = matvec(shat)
        omega = np.dot(t, s) / np.dot(t, t)
        x += alpha * phat
        x += omega * shat
        r -= omega * t
        rho_prev = rho
    return x, maxiter, maxiter


# ---------------------------------------------------------------------------
# convex envelope
# ---------------------------------------------------------------------------
def ce_operator(G, W, g, jacobian=True, dd=ddg_oracle):
    """convex_envelope.py:42-56."""
    lam, M, _ = dd.d2min(G, W, jacobian=jacobian, control=False)
    N = G.num_interior
    FW = W - g
    b = -lam > FW[:N]
    FW[np.flatnonzero(b)] = -lam[b]
    Jac = None
    if jacobian:
        n = G.num_points
        keep = np.ones(n)
        keep[np.flatnonzero(b)] = 0.0
        sel = sp.diags(b.astype(float))
        rows = sp.vstack([-(sel.dot(M)), sp.csr_matrix((n - N, n))]).tocsr()
        Jac = (sp.diags(keep) + rows).tocsr()
    return FW, Jac


def ce_dt(G):
    return 1 / 2 * np.max([G.min_interior_nb_dist, G.min_radius]) ** 2


def ce_solve(G, g, U0=None, solver="newton", dd=ddg_oracle, **kw):
    """convex_envelope.py:100-123."""
    if callable(g):
        g = g(G.points)
    if U0 is None:
        U0 = np.copy(g)
    dt = ce_dt(G)
    if solver == "euler":
        return euler(U0, lambda W: (ce_operator(G, W, g, False, dd)[0], dt), **kw)

    def op(W, jacobian=True):
        F, J = ce_operator(G, W, g, jacobian, dd)
        return F, J, dt
    return newton(U0, op, **kw)


# ---------------------------------------------------------------------------
# Pucci
# ---------------------------------------------------------------------------
def pucci_residual(G, W, Fvec, A, jacobian=True, dd=ddg_oracle):
    """GW - F of pucci.py:123-141 (with operator pucci.py:38-58 inlined)."""
    (lmin, Mmin, _), (lmax, Mmax, _) = dd.d2eigs(G, W, jacobian=jacobian, control=False)
    N, n = G.num_interior, G.num_points
    val = A[0] * lmax + A[1] * lmin
    GW = np.zeros(n)
    GW[:N] = val
    GW[N:] = W[N:]
    Jac = None
    if jacobian:
        if np.isscalar(A[0]) and np.isscalar(A[1]):
            M = A[0] * Mmax + A[1] * Mmin
        else:
            M = sp.diags(A[0]).dot(Mmax) + sp.diags(A[1]).dot(Mmin)
        ident = np.zeros(n)
        ident[N:] = 1.0
        Jac = (sp.vstack([M, sp.csr_matrix((n - N, n))]) + sp.diags(ident)).tocsr()
    return GW - Fvec, Jac


def pucci_solve(G, f, g, A=(2, 1), U0=None, dd=ddg_oracle, **kw):
    """pucci.py:105-160, Newton path."""
    N, n = G.num_interior, G.num_points
    if callable(f):
        f = f(G.points[:N])
    if callable(g):
        g = g(G.points[N:])
    Fv = np.zeros(n)
    Fv[N:] = g
    Fv[:N] = f
    dt = 1 / 2 * np.max([G.min_radius, G.min_interior_nb_dist]) ** 2
    if U0 is None:
        U0 = np.ones(n)
        U0[N:] = g

    def op(W, jacobian=True):
        R, J = pucci_residual(G, W, Fv, A, jacobian, dd)
        return R, J, dt
    return newton(U0, op, **kw)


# ---------------------------------------------------------------------------
# solvers
# ---------------------------------------------------------------------------
def euler(U, operator, solution_tol=1e-6, operator_tol=1e-6, max_iters=None, timeout=None):
    """solvers.py:56-91 (without the plot / zero-mean options)."""
    if max_iters is None:
        max_iters = U.size
    t0 = time.time()
    t_end = None if timeout is None else t0 + timeout
    for i in itertools.count(1):
        FU, dt = operator(U)
        U_new = U - dt * FU
        diff = np.amax(np.absolute(U - U_new))
        U = U_new
        if diff < solution_tol and np.abs(FU).max() < operator_tol:
            break
        if i >= max_iters:
            break
        if t_end is not None and time.time() > t_end:
            break
    return U, diff, i, time.time() - t0


def newton(U, operator, solution_tol=1e-6, operator_tol=1e-6, max_iters=1e2, euler_ratio=1,
           stats=None, use_scipy=False):
    """solvers.py:147-233.  Returns (U, diff, i+1, seconds)."""
    t0 = time.time()
    rho, p = 0.5, 2
    timeout = 1.0 * euler_ratio
    max_euler_iters = 10 * U.size
    krylov = 0
    fallbacks = 0
    for i in itertools.count(0):
        tstart = time.time()
        Gu, Jac, _ = operator(U, jacobian=True)
        P = sp.diags(1 / Jac.diagonal(), format="csr")
        tol = np.min([operator_tol, solution_tol])
        if use_scipy:
            from scipy.sparse.linalg import bicgstab as sb
            d, info = sb(Jac, -Gu, M=P, rtol=tol)
            kit = -1
        else:
            d, info, kit = bicgstab(Jac, -Gu, rtol=tol, M=P)
        krylov += max(kit, 0)
        timeout = euler_ratio * (time.time() - tstart)
        ok = info <= 0
        if ok:
            U_new = U + d
            gradTheta = Jac.transpose().dot(Gu)
            normp = np.linalg.norm(d) ** p
            if gradTheta.dot(d) > -rho * normp:
                ok = False
        if ok:
            diff = np.amax(np.absolute(U - U_new))
            U = U_new
        else:
            fallbacks += 1
            U, diff, _, _ = euler(U, lambda W: (lambda t: (t[0], t[2]))(operator(W, jacobian=False)),
                                  timeout=timeout, solution_tol=solution_tol,
                                  operator_tol=operator_tol, max_iters=max_euler_iters)
        if diff < solution_tol and np.abs(Gu).max() < operator_tol:
            break
        if i >= max_iters:
            break
    if stats is not None:
        stats.update(newton_iters=i + 1, krylov_iters=krylov, euler_fallbacks=fallbacks)
    return U, diff, i + 1, time.time() - t0
