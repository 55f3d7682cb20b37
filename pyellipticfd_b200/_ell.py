"""Shared implementation of the point-cloud operator modules ``ddi`` and ``ddf``.

Both reference modules have the same shape (ddi.py:304-414, ddf.py:137-252): a
set of ``Nds`` fixed unit directions, one 5-point finite-difference row per
(interior point, direction), cached on the grid after the first call, then per
call ``Nds`` mat-vecs and a min/max over directions.  Here the rows live in one
direction-major ELL table built on the device (``pde_ell_build``) and every
call is one kernel (``pde_ell_d2eigs``).
"""
import ctypes as C

import numpy as np
import scipy.sparse as sp
import torch

from . import _lib, _ddutils
from ._lib import check, vp
from ._context import require_cuda, to_device, like_input, _ptr, _stream

EPS = 1e-15          # ddi.py:10


def _grouped_simplices(G, lo=0, hi=None):
    """Simplices whose centre lies in [lo, hi), grouped by centre point, original row order
    kept inside a group (the reference takes the FIRST matching row per point, ddi.py:270-271)."""
    S = np.asarray(G.simplices, dtype=np.int64)
    hi = int(G.num_interior) if hi is None else hi
    S = S[(S[:, 0] >= lo) & (S[:, 0] < hi)]
    order = np.argsort(S[:, 0], kind="stable")
    S = np.ascontiguousarray(S[order])
    counts = np.bincount(S[:, 0] - lo, minlength=hi - lo)
    ptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    return S, ptr


class EllTable(object):
    """Device ELL table of one (grid, scheme).  ``scheme``: 0 ddi second derivative,
    1 Froese, 2 ddi first derivative (interior rows), 3 ddi first derivative on the
    boundary rows (``pde_ell_build``)."""

    def __init__(self, G, V, froese, device=None, per_point=False, scheme=None, rows=None):
        self.lib = _lib.lib()
        self.device = require_cuda(device)
        self.G = G
        self.scheme = int(froese) if scheme is None else int(scheme)
        self.n_points = int(np.asarray(G.points).shape[0])
        n_int = int(G.num_interior)
        self.row_offset = n_int if self.scheme == 3 else 0
        self.N = self.n_points - n_int if self.scheme == 3 else n_int
        if rows is not None:                       # point-range shard [lo, hi) of the interior
            self.row_offset, self.N = int(rows[0]), int(rows[1] - rows[0])
        self.V = np.ascontiguousarray(V, dtype=np.float64)
        self.nds = 1 if per_point else self.V.shape[0]
        dev = self.device
        P = torch.from_numpy(np.ascontiguousarray(G.points, dtype=np.float64)).to(dev)
        lo, hi = self.row_offset, self.row_offset + self.N
        if hasattr(G, "device_simplices") and getattr(G, "_simplices", None) is None:
            dS, dptr = G.device_simplices(dev, lo, hi)   # FDTriMesh: stencils already on the device
        else:
            S, ptr = _grouped_simplices(G, lo, hi)
            dS, dptr = torch.from_numpy(S).to(dev), torch.from_numpy(ptr).to(dev)
        dV = torch.from_numpy(self.V).to(dev)
        self.idx = torch.empty((self.nds, self.N, 4), dtype=torch.int32, device=dev)
        self.w = torch.empty((self.nds, self.N, 4), dtype=torch.float64, device=dev)
        fail = torch.zeros(1, dtype=torch.int32, device=dev)
        eps_h = 0.0 if self.scheme == 1 else EPS / G.spatial_resolution
        check(self.lib.pde_ell_build(dev.index, self.N, _ptr(P), _ptr(dS), _ptr(dptr), self.nds,
                                     _ptr(dV), int(per_point), float(eps_h), self.scheme,
                                     self.row_offset, n_int,
                                     _ptr(self.idx), _ptr(self.w), _ptr(fail), _stream()))
        nfail = int(fail.item())
        if nfail:
            raise ValueError("%d (point, direction) pairs are not enclosed by any simplex cone; "
                             "the stencil does not resolve the requested angular resolution"
                             % nfail)
        self.h = None
        if self.scheme != 3:
            h = vp()
            check(self.lib.pde_ell_create(C.byref(h), dev.index, self.N, self.n_points, self.nds,
                                          _ptr(self.idx), _ptr(self.w)))
            if self.row_offset:
                check(self.lib.pde_ell_set_row_offset(h, self.row_offset))
            self.h = h

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.pde_ell_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # ---- the pieces of GridContext the solvers use (solvers.NewtonEulerLS, newton_state,
    # euler_state): point clouds work on the reference's flat vectors, so pack / unpack only
    # move the vector to the device
    def n_flat(self):
        return self.n_points

    def pack(self, u):
        return to_device(u, self.device).to(torch.float64).clone()

    def unpack(self, S):
        return S

    def newton_update(self, U, D, F, JD):
        Unew = torch.empty_like(U)
        out = (C.c_double * 3)()
        check(self.lib.pde_ell_newton_update(self.h, _ptr(U), _ptr(D), _ptr(F), _ptr(JD),
                                             _ptr(Unew), out, _stream()))
        return Unew, float(out[0]), float(out[1]), float(out[2])

    def d2eigs(self, u, want_k=False):
        N, dev = self.N, self.device
        lmin = torch.empty(N, dtype=torch.float64, device=dev)
        lmax = torch.empty(N, dtype=torch.float64, device=dev)
        kmin = torch.empty(N, dtype=torch.int32, device=dev) if want_k else None
        kmax = torch.empty(N, dtype=torch.int32, device=dev) if want_k else None
        check(self.lib.pde_ell_d2eigs(self.h, _ptr(u), _ptr(lmin), _ptr(lmax), _ptr(kmin),
                                      _ptr(kmax), _stream()))
        return lmin, lmax, kmin, kmax

    def apply(self, k, x):
        y = torch.empty(self.N, dtype=torch.float64, device=self.device)
        check(self.lib.pde_ell_apply_policy(self.h, _ptr(k), _ptr(x), _ptr(y), _stream()))
        return y


class EllMatrix(object):
    """Selected-direction rows as a (num_interior, num_points) operator: stands in for
    the CSR matrices assembled row by row in ddi.py:369-404."""

    def __init__(self, table, k, coef=1.0):
        self.table, self.k, self.coef = table, k, coef

    @property
    def shape(self):
        return (self.table.N, self.table.n_points)

    def __neg__(self):
        return EllMatrix(self.table, self.k, -self.coef)

    def dot(self, x):
        X = to_device(x, self.table.device)
        y = self.table.apply(self.k, X)
        if self.coef != 1.0:
            y = y * self.coef
        return like_input(x, y)

    __matmul__ = dot

    def tocsr(self):
        t = self.table
        k = self.k.long()
        i = torch.arange(t.N, device=t.device)
        idx = t.idx[k, i].cpu().numpy().astype(np.int64)
        w = t.w[k, i].cpu().numpy() * self.coef
        I = np.arange(t.N)
        rows = np.concatenate([np.repeat(I, 4), I])
        cols = np.concatenate([idx.ravel(), I])
        vals = np.concatenate([w.ravel(), -w.sum(1)])
        return sp.coo_matrix((vals, (rows, cols)), shape=self.shape).tocsr()


def directions(nds):
    th = np.arange(0, 2 * np.pi, 2 * np.pi / nds)       # ddi.py:342-343
    return np.stack([np.cos(th), np.sin(th)], axis=1)


def cached_table(G, module_name, nds, froese):
    """The table lives where the reference keeps its matrices: G._d2cache =
    (module name, V, payload) (ddi.py:355-356).  A cache written by the other module
    is never reused (the reference's intent; its `cache_exist` typo does reuse it, Q5)."""
    cache = getattr(G, "_d2cache", None)
    if cache is not None and cache[0] == module_name and isinstance(cache[2], EllTable):
        return cache[2]
    V = directions(nds)
    table = EllTable(G, V, froese)
    G._d2cache = (module_name, V, table)
    return table


def d2eigs(G, u, module_name, nds, froese, jacobian=False, control=False,
           cache_fdmatrices=True):
    if G.dim == 3:
        raise NotImplementedError("Eigenvalues not yet implemented in 3d.")
    if cache_fdmatrices:
        table = cached_table(G, module_name, nds, froese)
    else:
        table = EllTable(G, directions(nds), froese)
    U = to_device(u, table.device)
    lmin, lmax, kmin, kmax = table.d2eigs(U, want_k=jacobian or control)
    M_min = M_max = v_min = v_max = None
    if jacobian:
        M_min, M_max = EllMatrix(table, kmin), EllMatrix(table, kmax)
    if control:
        V = torch.from_numpy(table.V).to(table.device)
        v_min, v_max = like_input(u, V[kmin.long()]), like_input(u, V[kmax.long()])
    return (like_input(u, lmin), M_min, v_min), (like_input(u, lmax), M_max, v_max)


def d2(G, v, u, jacobian, froese):
    if u is None:
        jacobian = True
    if froese and G.dim != 2:
        raise TypeError('Dimensions other than two not supported.')
    V = np.ascontiguousarray(_ddutils.process_v(G, v), dtype=np.float64)
    table = EllTable(G, V, froese, per_point=True)
    k = torch.zeros(table.N, dtype=torch.int32, device=table.device)
    M = EllMatrix(table, k)
    d2u = None
    if u is not None:
        d2u = like_input(u, table.apply(k, to_device(u, table.device)))
    return d2u, (M if jacobian else None)


def d1(G, v, u, jacobian, domain="interior"):
    """ddi.d1 (ddi.py:12-92): first derivative along ``v`` by linear interpolation inside
    the first simplex (row order) whose cone contains ``v``; on the boundary domain only
    simplices made of interior points qualify (ddi.py:66-68).  Returns ``(d1u, M)``."""
    from ._nbops import RowsMatrix
    if u is None:
        jacobian = True
    if G.dim != 2:
        raise NotImplementedError("only 2-D point clouds are implemented")
    V = np.ascontiguousarray(_ddutils.process_v(G, v, domain=domain), dtype=np.float64)
    table = EllTable(G, V, False, per_point=True, scheme=2 if domain == "interior" else 3)
    M = RowsMatrix(table.idx[0], table.w[0], table.row_offset, table.n_points, table.device)
    d1u = M.dot(u) if u is not None else None
    return d1u, (M if jacobian else None)


def d1grad(G, u, module_name, nds, jacobian=False, control=False, cache_fdmatrices=True):
    """ddi.d1grad (ddi.py:124-221): maximum over ``nds`` fixed directions of the
    interpolated first derivative; ``np.argmax`` keeps the FIRST maximal direction, which
    the table reproduces by storing the directions in reverse order (the min/max kernel
    keeps the highest index on ties).  Returns ``(d1_max, M, v)``."""
    if G.dim == 3:
        raise NotImplementedError("Gradient not yet implemented in 3d.")        # ddi.py:157-158
    V = directions(nds)
    cache = getattr(G, "_d1cache", None)
    if cache is not None and cache[0] == module_name and isinstance(cache[2], EllTable):
        table = cache[2]
    else:
        table = EllTable(G, V[::-1], False, scheme=2)
        if cache_fdmatrices:
            G._d1cache = (module_name, V, table)
    U = to_device(u, table.device)
    _, d1max, _, krev = table.d2eigs(U, want_k=jacobian or control)
    M = EllMatrix(table, krev) if jacobian else None
    vout = None
    if control:
        Vd = torch.from_numpy(V).to(table.device)
        vout = like_input(u, Vd[(nds - 1 - krev).long()])
    return like_input(u, d1max), M, vout


# ---------------------------------------------------------------------------------------------
# problems on point clouds: convex_envelope / pucci with fdmethod 'interpolate' | 'froese'
# ---------------------------------------------------------------------------------------------

class EllJacobian(object):
    """Square Jacobian over all points of a direction policy (convex_envelope.py:48-54,
    pucci.py:136-139): interior row i = cmin*row(kmin[i]) + cmax*row(kmax[i]), the identity
    where kmin[i] < 0; boundary rows identity."""

    def __init__(self, table, kmin, kmax=None, cmin=1.0, cmax=0.0):
        self.table, self.kmin, self.kmax, self.cmin, self.cmax = table, kmin, kmax, cmin, cmax
        self.ctx = table
        J = _lib.pde_ell_jac()
        J.d_kmin = kmin.data_ptr()
        J.d_kmax = kmax.data_ptr() if kmax is not None else None
        J.cmin, J.cmax = float(cmin), float(cmax)
        self._J = J

    @property
    def shape(self):
        return (self.table.n_points, self.table.n_points)

    def matvec_state(self, X, out=None):
        t = self.table
        Y = out if out is not None else torch.empty(t.n_points, dtype=torch.float64, device=t.device)
        check(t.lib.pde_ell_jac_matvec(t.h, C.byref(self._J), _ptr(X), _ptr(Y), _stream()))
        return Y

    def dot(self, x):
        return like_input(x, self.matvec_state(to_device(x, self.table.device)))

    __matmul__ = dot

    def diagonal(self):
        t = self.table
        D = torch.empty(t.n_points, dtype=torch.float64, device=t.device)
        check(t.lib.pde_ell_jac_diagonal(t.h, C.byref(self._J), _ptr(D), _stream()))
        return D

    def bicgstab_state(self, B, rtol, atol=0.0, maxiter=0, check_every=32):
        t = self.table
        x = torch.zeros(t.n_points, dtype=torch.float64, device=t.device)
        work = torch.empty(8 * t.n_points, dtype=torch.float64, device=t.device)
        info = (C.c_int64 * 2)()
        check(t.lib.pde_ell_bicgstab(t.h, C.byref(self._J), _ptr(B), _ptr(x), float(rtol),
                                     float(atol), int(maxiter), int(check_every), _ptr(work), info,
                                     _stream()))
        return x, int(info[0]), int(info[1])

    def tocsr(self):
        t = self.table
        N, n = t.N, t.n_points
        km = self.kmin.cpu().numpy()
        active = km >= 0

        def rows(k, c):
            k = np.where(active, k, 0)
            M = EllMatrix(t, torch.from_numpy(k.astype(np.int32)).to(t.device), c).tocsr()
            sel = sp.diags(active.astype(float))
            return sp.vstack([sel.dot(M), sp.csr_matrix((n - N, n))])

        M = rows(km, self.cmin)
        if self.kmax is not None:
            M = M + rows(self.kmax.cpu().numpy(), self.cmax)
        ident = np.ones(n)
        ident[:N][active] = 0.0
        return (M + sp.diags(ident)).tocsr()


def _problem_table(G, module_name, nds, froese):
    return cached_table(G, module_name, nds, froese)


def residual(table, W, g, mode, A=(0.0, 0.0), policy=True):
    n, dev = table.n_points, table.device
    F = torch.empty(n, dtype=torch.float64, device=dev)
    kmin = torch.empty(table.N, dtype=torch.int32, device=dev) if policy else None
    kmax = torch.empty(table.N, dtype=torch.int32, device=dev) if (policy and mode == 1) else None
    red = torch.zeros(1, dtype=torch.float64, device=dev)
    check(table.lib.pde_ell_residual(table.h, _ptr(W), _ptr(g), int(mode), float(A[0]), float(A[1]),
                                     _ptr(F), _ptr(kmin), _ptr(kmax), _ptr(red), _stream()))
    return F, kmin, kmax, red


def ce_operator(G, W, g, jacobian, module_name, nds, froese):
    """convex_envelope.operator for a point cloud (convex_envelope.py:37-56)."""
    table = _problem_table(G, module_name, nds, froese)
    F, kmin, _, _ = residual(table, to_device(W, table.device), to_device(g, table.device), 0,
                             policy=jacobian)
    Jac = EllJacobian(table, kmin, None, cmin=-1.0) if jacobian else None
    return like_input(W, F), Jac


def ce_solve(G, g, U0, solver, module_name, nds, froese, stats=None, **kwargs):
    """convex_envelope.solve for a point cloud (convex_envelope.py:100-123)."""
    import time
    from . import solvers
    table = _problem_table(G, module_name, nds, froese)
    t0 = time.time()
    dev = table.device
    gs = to_device(g, dev)
    Us = gs.clone() if U0 is None else to_device(U0, dev).clone()
    proto = g if U0 is None else U0
    dt = 1 / 2 * np.max([G.min_interior_nb_dist, G.min_radius]) ** 2      # convex_envelope.py:110
    sol, opt = kwargs.get("solution_tol", 1e-6), kwargs.get("operator_tol", 1e-6)

    def euler_res(U):
        F, _, _, red = residual(table, U, gs, 0, policy=False)
        return F, red[0]

    if solver == "euler":
        Us, diff, iters = solvers.euler_state(table, Us, euler_res, dt, sol, opt,
                                              kwargs.get("max_iters"), kwargs.get("timeout"))
        return like_input(proto, Us), diff, iters, time.time() - t0
    solvers.drop_unsupported(kwargs)

    def res(U, jacobian):
        F, kmin, _, red = residual(table, U, gs, 0, policy=jacobian)
        return F, (EllJacobian(table, kmin, None, cmin=-1.0) if jacobian else None), red[0]

    def fallback(U, timeout):
        U, diff, _ = solvers.euler_state(table, U, euler_res, dt, sol, opt, 10 * table.n_points,
                                         timeout)
        return U, diff

    Us, diff, iters = solvers.newton_state(table, Us, res, fallback, stats=stats, **kwargs)
    return like_input(proto, Us), diff, iters, time.time() - t0


def pucci_operator(G, W, A, jacobian, module_name, nds, froese):
    """pucci.operator for a point cloud (pucci.py:40-58): returns (-val, -M)."""
    if not (np.isscalar(A[0]) and np.isscalar(A[1])):
        raise NotImplementedError("per-point Pucci coefficients on point clouds")
    table = _problem_table(G, module_name, nds, froese)
    Wd = to_device(W, table.device)
    zero = torch.zeros(table.n_points, dtype=torch.float64, device=table.device)
    F, kmin, kmax, _ = residual(table, Wd, zero, 1, A, policy=jacobian)
    val = F[:table.N]
    M = None
    if jacobian:
        J = EllJacobian(table, kmin, kmax, cmin=-float(A[1]), cmax=-float(A[0]))
        M = _InteriorView(J)
    return like_input(W, -val), M


class _InteriorView(object):
    def __init__(self, J):
        self.J = J

    @property
    def shape(self):
        return (self.J.table.N, self.J.table.n_points)

    def dot(self, x):
        return self.J.dot(x)[:self.J.table.N]

    def tocsr(self):
        return self.J.tocsr()[:self.J.table.N]


def pucci_solve(G, f, g, A, U0, solver, module_name, nds, froese, stats=None, **kwargs):
    """pucci.solve for a point cloud (pucci.py:105-160)."""
    import time
    from . import solvers
    if not (np.isscalar(A[0]) and np.isscalar(A[1])):
        raise NotImplementedError("per-point Pucci coefficients on point clouds")
    table = _problem_table(G, module_name, nds, froese)
    t0 = time.time()
    dev, N, n = table.device, table.N, table.n_points
    proto = f if not np.isscalar(f) else (g if not np.isscalar(g) else np.zeros(1))
    Fv = torch.zeros(n, dtype=torch.float64, device=dev)
    Fv[:N] = to_device(f, dev) if not np.isscalar(f) else float(f)
    if g is not None:
        Fv[N:] = to_device(g, dev) if not np.isscalar(g) else float(g)
    dt = 1 / 2 * np.max([G.min_radius, G.min_interior_nb_dist]) ** 2
    if U0 is None:
        Us = torch.ones(n, dtype=torch.float64, device=dev)
        if g is not None:
            Us[N:] = Fv[N:]
    else:
        Us = to_device(U0, dev).clone()
    sol, opt = kwargs.get("solution_tol", 1e-6), kwargs.get("operator_tol", 1e-6)

    def res(U, jacobian):
        R, kmin, kmax, red = residual(table, U, Fv, 1, A, policy=jacobian)
        Jac = EllJacobian(table, kmin, kmax, cmin=float(A[1]), cmax=float(A[0])) if jacobian else None
        return R, Jac, red[0]

    def euler_res(U):
        R, _, _, red = residual(table, U, Fv, 1, A, policy=False)
        R[:N].neg_()                      # stable sign for explicit Euler
        return R, red[0]

    if solver == "euler":
        Us, diff, iters = solvers.euler_state(table, Us, euler_res, dt, sol, opt,
                                              kwargs.get("max_iters"), kwargs.get("timeout"))
        return like_input(proto, Us), diff, iters, time.time() - t0
    solvers.drop_unsupported(kwargs)

    def fallback(U, timeout):
        U, diff, _ = solvers.euler_state(table, U, euler_res, dt, sol, opt, 10 * n, timeout)
        return U, diff

    Us, diff, iters = solvers.newton_state(table, Us, res, fallback, stats=stats, **kwargs)
    return like_input(proto, Us), diff, iters, time.time() - t0
