"""Pucci equation  A[0]*lambda_max[u] + A[1]*lambda_min[u] = f  with Dirichlet data
(mirror of /root/reference/pucci.py).

``operator`` and ``solve`` keep the reference signatures and conventions
(pucci.py:7-58, :60-160).  Documented deviations from reference defects
(SURVEY.md section 7.3-G): ``operator(..., jacobian=False)`` returns
``(-val, None)`` instead of raising TypeError on ``-None`` (pucci.py:58, Q2), and
``solver='euler'`` integrates the stable sign of the residual (Q3) -- the
reference's Euler path crashes before it does anything.
"""
import time
import warnings

import numpy as np
import torch

from . import ddg, ddi, ddf, solvers
from ._context import GridContext, like_input, to_device
from .linop import PolicyJacobian
from .convex_envelope import _is_pairs_grid


def _coef(ctx, a):
    """Scalar, or per-interior-point array -> state-layout tensor."""
    if np.isscalar(a):
        return float(a)
    a = to_device(a, ctx.device)
    full = torch.cat([a, torch.zeros(ctx.nb, dtype=a.dtype, device=a.device)])
    return ctx.pack(full)


def operator(Grid, W, A, jacobian=True, fdmethod='interpolate'):
    """Returns ``(-val, -M)`` with val = A[0]*lambda_max + A[1]*lambda_min over the
    interior and M = A[0]*M_max + A[1]*M_min (pucci.py:38-58)."""
    if Grid.dim == 3:
        raise NotImplementedError("Pucci operator not yet implemented in 3d.")
    if fdmethod == 'grid' or (fdmethod == 'froese' and _is_pairs_grid(Grid)):
        ctx = GridContext.get(Grid)
        Ws = ctx.pack(W)
        zero = ctx.zeros()
        F, pmin, pmax, _ = ctx.pucci_residual(Ws, zero, _coef(ctx, A[0]), _coef(ctx, A[1]),
                                              policy=jacobian)
        val = ctx.unpad(F)
        M = None
        if jacobian:
            M = _InteriorRows(PolicyJacobian(ctx, pmin, pmax, cmin=_neg(_coef(ctx, A[1])),
                                             cmax=_neg(_coef(ctx, A[0]))))
        return like_input(W, -val), M
    mod = {'interpolate': ddi, 'froese': ddf}[fdmethod]
    return mod._pucci_operator(Grid, W, A, jacobian)


def _neg(c):
    return -c


class _InteriorRows(object):
    """(num_interior, num_points) view of a square PolicyJacobian: what
    pucci.operator returns as ``-M`` (pucci.py:58)."""

    def __init__(self, J):
        self.J = J

    @property
    def shape(self):
        return (self.J.ctx.n_interior_local, self.J.ctx.n_flat())

    def dot(self, x):
        y = self.J.dot(x)
        return y[:self.J.ctx.n_interior_local]

    def tocsr(self):
        return self.J.tocsr()[:self.J.ctx.n_interior_local]


def solve(Grid, f, dirichlet, A=(2, 1), U0=None, fdmethod='interpolate', solver="newton",
          **kwargs):
    """Solve the Pucci Dirichlet problem (pucci.py:60-160).  ``f`` over interior
    points and ``dirichlet`` over boundary points may be arrays or callables.
    Returns ``(U, diff, iterations, seconds)``."""
    g = dirichlet
    if callable(f):
        f = f(Grid.interior_points)
    if callable(g):
        g = g(Grid.bdry_points)
    if Grid.dim == 3:
        raise NotImplementedError("Pucci operator not yet implemented in 3d.")
    if fdmethod == 'grid' or (fdmethod == 'froese' and _is_pairs_grid(Grid)):
        return _solve_grid(Grid, f, g, A, U0, solver, **kwargs)
    mod = {'interpolate': ddi, 'froese': ddf}[fdmethod]
    return mod._pucci_solve(Grid, f, g, A, U0, solver, **kwargs)


def _solve_grid(Grid, f, g, A, U0, solver, stats=None, **kwargs):
    ctx = GridContext.get(Grid)
    t0 = time.time()
    dev = ctx.device
    N, nb = ctx.n_interior_local, ctx.nb
    proto = f if not np.isscalar(f) else (g if g is not None and not np.isscalar(g) else np.zeros(1))
    # forcing over the whole domain (pucci.py:115-117)
    Fv = torch.zeros(N + nb, dtype=torch.float64, device=dev)
    Fv[:N] = to_device(f, dev) if not np.isscalar(f) else float(f)
    if g is not None:
        Fv[N:] = to_device(g, dev) if not np.isscalar(g) else float(g)
    Fs = ctx.pack(Fv)
    dx = np.max([Grid.min_radius, Grid.min_interior_nb_dist])
    dt = 1 / 2 * dx ** 2                                              # pucci.py:119-120
    if U0 is None:                                                    # pucci.py:145-149
        U0v = torch.ones(N + nb, dtype=torch.float64, device=dev)
        if g is not None:
            U0v[N:] = Fv[N:]
        Us = ctx.pack(U0v)
    else:
        Us = ctx.pack(U0)
    A0, A1 = _coef(ctx, A[0]), _coef(ctx, A[1])
    sol = kwargs.get("solution_tol", 1e-6)
    opt = kwargs.get("operator_tol", 1e-6)

    def residual(U, jacobian):
        # GW - F: interior A0*lmax + A1*lmin - f, wall W - g (pucci.py:127-141)
        R, pmin, pmax, red = ctx.pucci_residual(U, Fs, A0, A1, policy=jacobian)
        Jac = PolicyJacobian(ctx, pmin, pmax, cmin=A1, cmax=A0) if jacobian else None
        return R, Jac, red[0]

    def euler_residual(U):
        R, _, _, red = ctx.pucci_residual(U, Fs, A0, A1, policy=False)
        R[:ctx.n_lattice].neg_()           # stable sign for explicit Euler (quirk Q3)
        return R, red[0]

    if solver == "euler":
        Us, diff, iters = solvers.euler_state(ctx, Us, euler_residual, dt, sol, opt,
                                              kwargs.get("max_iters"), kwargs.get("timeout"))
        return like_input(proto, ctx.unpack(Us)), diff, iters, time.time() - t0
    elif solver == "newton":
        solvers.drop_unsupported(kwargs)

        def euler_fallback(U, timeout):
            U, diff, _ = solvers.euler_state(ctx, U, euler_residual, dt, sol, opt,
                                             10 * ctx.n_flat(), timeout)
            return U, diff

        with warnings.catch_warnings():
            Us, diff, iters = solvers.newton_state(ctx, Us, residual, euler_fallback, stats=stats,
                                                   **kwargs)
        return like_input(proto, ctx.unpack(Us)), diff, iters, time.time() - t0
    raise ValueError("solver must be 'euler' or 'newton'")
