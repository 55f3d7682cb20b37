"""Solvers for the finite difference schemes (mirror of /root/reference/solvers.py).

``euler`` (solvers.py:13-93) and ``NewtonEulerLS`` (solvers.py:98-233) keep the
reference signatures, stopping rules, iteration counting and return tuple
``(U, diff, iterations, seconds)``.  They accept any operator closure whose
values are NumPy arrays or torch tensors; the Newton linear solve, however,
runs only on the device: the Jacobian returned by the operator must be a
:class:`~pyellipticfd_b200.linop.PolicyJacobian` (there is no SciPy fallback).

``newton_state`` / ``euler_state`` are the device-resident loops on
state-layout tensors that ``convex_envelope.solve`` and ``pucci.solve`` use.
"""
import itertools
import time
import warnings
from warnings import warn

import numpy as np
import torch
from scipy.sparse.linalg import MatrixRankWarning

from .linop import PolicyJacobian
from ._context import like_input


class NewtonDecreaseWarning(UserWarning):
    """Raised (as a warning) when the Newton step fails the sufficient-decrease
    test (solvers.py:95-96, :192-197)."""
    pass


def _amax(x):
    return float(abs(x).max())


def euler(U, operator, solution_tol=1e-6, operator_tol=1e-6, max_iters=None,
          timeout=None, zeromean=False, zeromax=False, plotter=None):
    """Explicit Euler to steady state: iterate U <- U - dt*F(U) until
    max|U_new-U| < solution_tol and max|F(U)| < operator_tol (solvers.py:63-91).

    ``operator(U)`` returns ``(F(U), dt)``.  Returns ``(U, diff, i, seconds)``
    with ``i`` counted from 1, exactly as the reference.
    """
    if max_iters is None:
        max_iters = U.numel() if isinstance(U, torch.Tensor) else U.size
    t0 = time.time()
    if timeout is not None:
        timeout = time.time() + timeout
    diff = float("inf")
    i = 0
    for i in itertools.count(1):
        try:
            FU, dt = operator(U)
            U_new = U - dt * FU
            if plotter:
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    plotter(U_new)
            diff = _amax(U - U_new)
            if zeromean:
                U = U_new - U_new.mean()
            elif zeromax:
                U = U_new - U_new.max()
            else:
                U = U_new
            if diff < solution_tol and _amax(FU) < operator_tol:
                return U, diff, i, time.time() - t0
            elif i >= max_iters:
                warn("Maximum iterations reached")
                return U, diff, i, time.time() - t0
            elif timeout is not None and time.time() > timeout:
                warn("Maximum computation time reached")
                return U, diff, i, time.time() - t0
        except KeyboardInterrupt:
            return U, diff, i, time.time() - t0


def NewtonEulerLS(U, operator, solution_tol=1e-6, operator_tol=1e-6, max_iters=1e2,
                  euler_ratio=1, plotter=None, max_euler_iters=None,
                  verbose=False, force_euler=False):
    """Semismooth Newton (policy iteration) with an explicit-Euler fallback
    (solvers.py:98-233).

    ``operator(U, jacobian=bool)`` returns ``(F(U), Jac, dt)``; ``Jac`` must be a
    PolicyJacobian.  Per iteration: Jacobi-preconditioned BiCGStab on the device
    with rtol = min(operator_tol, solution_tol) (solvers.py:179-181), update,
    sufficient-decrease test (Jac^T F).d > -rho*||d||^p with rho=.5, p=2
    (solvers.py:158,192-197); a failed solve or failed test switches to Euler
    steps for euler_ratio * (time of the Newton step) seconds.  Returns
    ``(U, diff, i+1, seconds)``.
    """
    t0 = time.time()
    if euler_ratio <= 0:
        raise ValueError('euler_ratio must be positive')
    newtontime = 1.0
    timeout = euler_ratio * newtontime      # reference leaves this undefined until the first
    rho, p = 0.5, 2                         # successful solve (quirk Q10); initialised here

    def G(W):
        tup = operator(W, jacobian=False)
        return (tup[0], tup[2])

    n = U.numel() if isinstance(U, torch.Tensor) else U.size
    max_euler_iters = 10 * n                # solvers.py:166 overrides the argument (quirk Q8)
    diff = float("inf")
    for i in itertools.count(0):
        euler_flag = False
        try:
            tstart = time.time()
            Gu, Jac, _ = operator(U, jacobian=True)
            # any device Jacobian of this package: regular grids (linop.PolicyJacobian, state
            # layout) or point clouds (_ell.EllJacobian, the reference's flat vectors)
            if not (hasattr(Jac, "bicgstab_state") and hasattr(Jac, "matvec_state")
                    and hasattr(getattr(Jac, "ctx", None), "newton_update")):
                raise TypeError("NewtonEulerLS needs a device Jacobian (PolicyJacobian / EllJacobian) "
                                "from the operator; pyellipticfd_b200 has no CPU linear solver")
            ctx = Jac.ctx
            B = ctx.pack(Gu)
            B.neg_()
            D, info, _ = Jac.bicgstab_state(B, rtol=min(operator_tol, solution_tol))
            newtontime = time.time() - tstart
            timeout = euler_ratio * newtontime
            if info > 0:
                raise _Fallback(MatrixRankWarning('Failed to invert Jacobian'))
            Us = ctx.pack(U)
            JD = Jac.matvec_state(D)
            Un, dmax, gtd, dd = ctx.newton_update(Us, D, ctx.pack(Gu), JD)
            # (Jac^T Gu).d == Gu.(Jac d)
            if gtd > -rho * np.sqrt(dd) ** p:
                raise _Fallback(NewtonDecreaseWarning('Insufficient decrease in Newton step'))
            diff = dmax
            U = like_input(U, ctx.unpack(Un))
        except _Fallback as e:
            if verbose:
                print('At iteratate', i, ':', e.warning, '; switching to Euler')
            euler_flag = True
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                U, diff, _, _ = euler(U, G, timeout=timeout, solution_tol=solution_tol,
                                      operator_tol=operator_tol, max_iters=max_euler_iters)
        if force_euler and not euler_flag:
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                U, diff, _, _ = euler(U, G, timeout=timeout, solution_tol=solution_tol,
                                      operator_tol=operator_tol, max_iters=max_euler_iters)
            Gu, _, _ = operator(U, jacobian=False)
        if plotter:
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                plotter(U)
        if diff < solution_tol and _amax(Gu) < operator_tol:
            return U, diff, i + 1, time.time() - t0
        elif i >= max_iters:
            warn("Maximum iterations reached")
            return U, diff, i + 1, time.time() - t0


def drop_unsupported(kwargs, names=("plotter", "max_euler_iters", "zeromean", "zeromax")):
    """Solver keywords of the reference that the device-resident loops of ``*.solve`` do not
    implement are dropped LOUDLY (``plotter`` needs the host copy of every iterate,
    ``max_euler_iters`` is overridden by the reference itself, solvers.py:166; ``zeromean`` /
    ``zeromax`` exist in solvers.euler only)."""
    for k in names:
        if kwargs.pop(k, None):
            warn("solver keyword %r is not supported by the device-resident loop and is ignored; "
                 "use solvers.NewtonEulerLS / solvers.euler with an operator closure for it" % k)


class _Fallback(Exception):
    def __init__(self, warning):
        self.warning = warning


# ---------------------------------------------------------------------------------------
# device-resident loops on state-layout tensors
# ---------------------------------------------------------------------------------------

def euler_state(ctx, U, residual, dt, solution_tol=1e-6, operator_tol=1e-6, max_iters=None,
                timeout=None):
    """Generic explicit Euler on a state tensor; ``residual(U) -> (F_state, max|F|)``.
    Used where no fused step kernel exists (Pucci)."""
    if max_iters is None:
        max_iters = ctx.n_flat()
    t_end = None if timeout is None else time.time() + timeout
    diff = float("inf")
    i = 0
    for i in itertools.count(1):
        F, fmax = residual(U)
        U_new = U - dt * F
        diff = float((U - U_new).abs().max())
        U = U_new
        if diff < solution_tol and float(fmax) < operator_tol:
            break
        if i >= max_iters:
            warn("Maximum iterations reached")
            break
        if t_end is not None and time.time() > t_end:
            warn("Maximum computation time reached")
            break
    return U, diff, i


def newton_state(ctx, U, residual, euler_fallback, solution_tol=1e-6, operator_tol=1e-6,
                 max_iters=1e2, euler_ratio=1, verbose=False, force_euler=False, stats=None,
                 bicgstab_check_every=32, krylov_rtol=None, krylov_restarts=0, trace=None):
    """NewtonEulerLS (solvers.py:168-233) on state tensors.

    ``residual(U, jacobian) -> (F_state, PolicyJacobian|None, max|F|)``;
    ``euler_fallback(U, timeout) -> (U, diff)`` runs Euler steps for at most
    ``timeout`` seconds.  ``stats`` (dict) collects iteration counts.

    Extensions (defaults = the reference's behaviour): ``krylov_rtol`` replaces the reference's
    ``min(operator_tol, solution_tol)`` as the relative tolerance of the linear solves;
    ``krylov_restarts = k`` restarts BiCGStab up to k times on the current residual when it
    leaves on a breakdown (SciPy's info -10 / -11, which the reference accepts as a solve,
    solvers.py:186-187: it only tests ``info > 0``) before its tolerance is met."""
    if euler_ratio <= 0:
        raise ValueError('euler_ratio must be positive')
    rho, p = 0.5, 2
    newtontime = 1.0
    timeout = euler_ratio * newtontime
    diff = float("inf")
    krylov_total = 0
    fallbacks = 0
    i = 0
    for i in itertools.count(0):
        euler_flag = False
        tstart = time.time()
        F, Jac, fmax = residual(U, True)
        B = -F
        # reference: tol = min(operator_tol, solution_tol) (solvers.py:181); `krylov_rtol`
        # is an extension for fine grids where a 1e-6 linear solve lets the active set chatter
        rtol = min(operator_tol, solution_tol) if krylov_rtol is None else krylov_rtol
        D, info, kit = Jac.bicgstab_state(B, rtol=rtol, check_every=bicgstab_check_every)
        krylov_total += kit
        if krylov_restarts and info < 0:
            target = rtol * float(torch.linalg.vector_norm(B))
            for _ in range(int(krylov_restarts)):
                R = B - Jac.matvec_state(D)
                if float(torch.linalg.vector_norm(R)) <= target:
                    info = 0
                    break
                dD, info, k2 = Jac.bicgstab_state(R, rtol=0.0, atol=target,
                                                  check_every=bicgstab_check_every)
                D += dD
                krylov_total += k2
                kit += k2
                if info == 0:
                    break
        if callable(trace):
            trace(i, float(fmax), kit, info)
        elif trace is not None:
            trace.append((i, float(fmax), kit, info))
        newtontime = time.time() - tstart
        timeout = euler_ratio * newtontime
        ok = info <= 0
        reason = 'Failed to invert Jacobian'
        if ok:
            JD = Jac.matvec_state(D)
            Un, dmax, gtd, dd = ctx.newton_update(U, D, F, JD)
            if gtd > -rho * np.sqrt(dd) ** p:
                ok = False
                reason = 'Insufficient decrease in Newton step'
        if ok:
            diff = dmax
            U = Un
        else:
            if verbose:
                print('At iteratate', i, ':', reason, '; switching to Euler')
            euler_flag = True
            fallbacks += 1
            U, diff = euler_fallback(U, timeout)
        if force_euler and not euler_flag:
            U, diff = euler_fallback(U, timeout)
            F, _, fmax = residual(U, False)
        if diff < solution_tol and float(fmax) < operator_tol:
            break
        elif i >= max_iters:
            warn("Maximum iterations reached")
            break
    if stats is not None:
        stats.update(newton_iters=i + 1, krylov_iters=krylov_total, euler_fallbacks=fallbacks,
                     residual=float(fmax))
    return U, diff, i + 1
