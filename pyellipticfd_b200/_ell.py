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


def _grouped_simplices(G):
    """Interior simplices grouped by centre point, original row order kept inside a
    group (the reference takes the FIRST matching row per point, ddi.py:270-271)."""
    S = np.asarray(G.simplices, dtype=np.int64)
    N = int(G.num_interior)
    S = S[S[:, 0] < N]
    order = np.argsort(S[:, 0], kind="stable")
    S = np.ascontiguousarray(S[order])
    counts = np.bincount(S[:, 0], minlength=N)
    ptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    return S, ptr


class EllTable(object):
    """Device ELL table of one (grid, scheme)."""

    def __init__(self, G, V, froese, device=None, per_point=False):
        self.lib = _lib.lib()
        self.device = require_cuda(device)
        self.G = G
        self.N = int(G.num_interior)
        self.n_points = int(np.asarray(G.points).shape[0])
        self.V = np.ascontiguousarray(V, dtype=np.float64)
        self.nds = 1 if per_point else self.V.shape[0]
        dev = self.device
        S, ptr = _grouped_simplices(G)
        P = torch.from_numpy(np.ascontiguousarray(G.points, dtype=np.float64)).to(dev)
        dS, dptr = torch.from_numpy(S).to(dev), torch.from_numpy(ptr).to(dev)
        dV = torch.from_numpy(self.V).to(dev)
        self.idx = torch.empty((self.nds, self.N, 4), dtype=torch.int32, device=dev)
        self.w = torch.empty((self.nds, self.N, 4), dtype=torch.float64, device=dev)
        fail = torch.zeros(1, dtype=torch.int32, device=dev)
        eps_h = 0.0 if froese else EPS / G.spatial_resolution
        check(self.lib.pde_ell_build(dev.index, self.N, _ptr(P), _ptr(dS), _ptr(dptr), self.nds,
                                     _ptr(dV), int(per_point), float(eps_h), int(froese),
                                     _ptr(self.idx), _ptr(self.w), _ptr(fail), _stream()))
        nfail = int(fail.item())
        if nfail:
            raise ValueError("%d (point, direction) pairs are not enclosed by any simplex cone; "
                             "the stencil does not resolve the requested angular resolution"
                             % nfail)
        h = vp()
        check(self.lib.pde_ell_create(C.byref(h), dev.index, self.N, self.n_points, self.nds,
                                      _ptr(self.idx), _ptr(self.w)))
        self.h = h

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.pde_ell_destroy(self.h)
                self.h = None
        except Exception:
            pass

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
