"""Convex envelope of an obstacle via  max(u - g, -lambda_min[u]) = 0
(mirror of /root/reference/convex_envelope.py).

``operator(Grid, W, g, jacobian=True, fdmethod=...)`` and ``solve(Grid, g, U0=None,
fdmethod=..., solver="newton", **kwargs)`` keep the reference's signatures and
return conventions (convex_envelope.py:8-56, :58-123).  For ``fdmethod='grid'``
the residual, the active set and the policy come out of ONE fused kernel
(``pde_ce_residual``) and the Jacobian is a matrix-free
:class:`~pyellipticfd_b200.linop.PolicyJacobian`; ``solver='euler'`` runs the
device-resident loop ``pde_ce_euler``.
"""
import time
import warnings

import numpy as np
import torch

from . import ddg, ddi, ddf, solvers
from ._context import GridContext, like_input, to_device
from .linop import PolicyJacobian


def _is_pairs_grid(Grid):
    return hasattr(Grid, "stencil_pairs") or getattr(Grid, "pairs", None) is not None


def operator(Grid, W, g, jacobian=True, fdmethod='interpolate'):
    """F(W) = max(W - g, -lambda_min[W]) on the interior, W - g on the boundary,
    and its Jacobian (identity rows, active rows replaced by -M_min)
    (convex_envelope.py:44-54).  Returns ``(FW, Jac)``; ``Jac`` is None when
    ``jacobian=False``."""
    if fdmethod == 'grid' or (fdmethod == 'froese' and _is_pairs_grid(Grid)):
        # ddf delegates to ddg when pairs exist (ddf.py:36-37, :165-166)
        ctx = GridContext.get(Grid)
        Ws, gs = ctx.pack(W), ctx.pack(g)
        F, P, _ = ctx.ce_residual(Ws, gs, policy=jacobian)
        FW = like_input(W, ctx.unpack(F))
        Jac = PolicyJacobian(ctx, P, None, cmin=-1.0) if jacobian else None
        return FW, Jac
    mod = {'interpolate': ddi, 'froese': ddf}[fdmethod]
    return mod._ce_operator(Grid, W, g, jacobian)


def solve(Grid, g, U0=None, fdmethod='interpolate', solver="newton", **kwargs):
    """Convex envelope of ``g`` (array over all points, or a callable on
    ``Grid.points``) -- convex_envelope.py:58-123.  Returns
    ``(U, diff, iterations, seconds)``."""
    if callable(g):
        g = g(Grid.points)
    if fdmethod == 'grid' or (fdmethod == 'froese' and _is_pairs_grid(Grid)):
        return _solve_grid(Grid, g, U0, solver, **kwargs)
    mod = {'interpolate': ddi, 'froese': ddf}[fdmethod]
    return mod._ce_solve(Grid, g, U0, solver, **kwargs)


def cfl_dt(Grid):
    """dt = 1/2 * max(min_interior_nb_dist, min_radius)^2 (convex_envelope.py:110)."""
    return 1 / 2 * np.max([Grid.min_interior_nb_dist, Grid.min_radius]) ** 2


def _solve_grid(Grid, g, U0, solver, stats=None, **kwargs):
    ctx = GridContext.get(Grid)
    t0 = time.time()
    gs = ctx.pack(g)
    Us = gs.clone() if U0 is None else ctx.pack(U0)
    dt = cfl_dt(Grid)
    proto = g if U0 is None else U0

    if solver == "euler":
        plain = set(kwargs) <= {"solution_tol", "operator_tol", "max_iters"}
        if plain:
            sol = kwargs.get("solution_tol", 1e-6)
            opt = kwargs.get("operator_tol", 1e-6)
            mi = kwargs.get("max_iters") or ctx.n_flat()          # solvers.py:56-57
            Us, diff, iters, fmax, reason = ctx.ce_euler(Us, gs, dt, sol, opt, mi)
            if reason == 2:
                warnings.warn("Maximum iterations reached")
            if stats is not None:
                stats.update(residual=fmax, reason=reason)
            return like_input(proto, ctx.unpack(Us)), diff, iters, time.time() - t0

        def G(W):
            return operator(Grid, W, g, jacobian=False, fdmethod='grid')[0], dt
        U0f = like_input(proto, ctx.unpack(Us))
        return solvers.euler(U0f, G, **kwargs)

    elif solver == "newton":
        solvers.drop_unsupported(kwargs)

        def residual(U, jacobian):
            F, P, red = ctx.ce_residual(U, gs, policy=jacobian)
            Jac = PolicyJacobian(ctx, P, None, cmin=-1.0) if jacobian else None
            return F, Jac, red[0]

        sol = kwargs.get("solution_tol", 1e-6)
        opt = kwargs.get("operator_tol", 1e-6)

        def euler_fallback(U, timeout):
            # solvers.py:209-212: Euler steps bounded by wall-clock `timeout`
            # (and 10*U.size iterations, solvers.py:166)
            t_end = time.time() + timeout
            budget = 10 * ctx.n_flat()
            diff = float("inf")
            while budget > 0:
                n = min(budget, 256)
                U, diff, it, fmax, reason = ctx.ce_euler(U, gs, dt, sol, opt, n, check_every=n)
                budget -= it
                if reason == 1 or time.time() > t_end:
                    break
            return U, diff

        Us, diff, iters = solvers.newton_state(ctx, Us, residual, euler_fallback, stats=stats,
                                               **kwargs)
        return like_input(proto, ctx.unpack(Us)), diff, iters, time.time() - t0
    elif solver == "sweep":
        # extension (csrc/ce_sweep.cu): directional convex-hull sweeps; stops on the criterion
        # of solvers.euler (solvers.py:84) -- largest change < solution_tol and max|F| < operator_tol
        sol = kwargs.get("solution_tol", 1e-6)
        opt = kwargs.get("operator_tol", 1e-6)
        mi = int(kwargs.get("max_iters") or 1000)
        chunk = int(kwargs.get("check_every", 8))
        timeout = kwargs.get("timeout")            # seconds, like solvers.euler's (solvers.py:58-61)
        # band relaxation after every 2nd direction pass: fewer sweeps than after every 4th
        # (4097^2: 48 vs 53; host build at 1025^2: 33 with 1, 39 with 4) for 8 more tiny launches
        relax_every = int(kwargs.get("relax_every", 2))
        if U0 is not None:
            # sweeps only lower values, so they must start on or above the solution: from g
            warnings.warn("solver='sweep' starts from the obstacle; U0 is ignored")
        iters, diff, fmax, why = 0, float("inf"), float("inf"), "Maximum iterations reached"
        Us = gs.clone()
        while iters < mi:
            # the device ends a batch early after a sweep whose largest change is below what an
            # Euler step of the reference at max|F| = operator_tol / 2 would make (dt * |F|)
            Us, diff, n = ctx.ce_sweep(gs, min(sol, 0.5 * opt * dt), min(chunk, mi - iters), U=Us,
                                       relax_every=relax_every)
            iters += n
            fmax = float(ctx.ce_residual(Us, gs, policy=False)[2][0])
            if diff < sol and fmax < opt:
                break
            if diff == 0.0:
                why = "The sweeps stalled above operator_tol"
                break
            if timeout is not None and time.time() - t0 > timeout:
                why = "Maximum computation time reached"      # solvers.py:89
                break
        if not (diff < sol and fmax < opt):
            warnings.warn(why)
        if stats is not None:
            stats.update(residual=fmax)
        return like_input(proto, ctx.unpack(Us)), diff, iters, time.time() - t0
    raise ValueError("solver must be 'euler', 'newton' or 'sweep'")
