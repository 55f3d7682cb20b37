"""First derivatives on neighbour lists (``ddg.d1`` / ``d1n`` / ``d1grad``) and the
row-major finite-difference matrices they return.

Mirrors ddg.py:8-175.  The neighbour list of the grid (``G.neighbours``, rows ``(I, J)``)
is kept on the device as a CSR structure per domain (interior / boundary points); one
kernel (``pde_nb_d1``) selects, per centre point, the neighbour best aligned with the
direction (``d1``) or with the steepest difference quotient (``d1grad``).  The Jacobians
are :class:`RowsMatrix` objects (``.dot``, ``.tocsr()``, ``.shape``) backed by the selected
rows, applied by ``pde_rows_apply``.
"""
import ctypes as C

import numpy as np
import scipy.sparse as sp
import torch

from . import _lib, _ddutils
from ._lib import check
from ._context import require_cuda, to_device, like_input, _ptr, _stream


class RowsMatrix(object):
    """(n_rows, n_points) finite-difference matrix with at most 4 off-centre entries per
    row: row r = sum_j w[r,j] * (e[idx[r,j]] - e[r + centre_offset]).  Stands in for the
    ``csr_matrix((val,(i,j)))[nonzero rows]`` of ddg.py:73-79 and ddi.py:84-90."""

    def __init__(self, idx, w, centre_offset, n_points, device):
        self.idx, self.w = idx.contiguous(), w.contiguous()
        self.centre_offset, self.n_points, self.device = int(centre_offset), int(n_points), device

    @property
    def shape(self):
        return (self.idx.shape[0], self.n_points)

    def dot(self, x):
        X = to_device(x, self.device)
        y = torch.empty(self.idx.shape[0], dtype=torch.float64, device=self.device)
        check(_lib.lib().pde_rows_apply(self.device.index, self.idx.shape[0], self.centre_offset,
                                        _ptr(self.idx), _ptr(self.w), _ptr(X), _ptr(y), _stream()))
        return like_input(x, y)

    __matmul__ = dot

    def tocsr(self):
        n = self.idx.shape[0]
        idx = self.idx.cpu().numpy().astype(np.int64)
        w = self.w.cpu().numpy()
        R = np.arange(n)
        rows = np.concatenate([np.repeat(R, 4), R])
        cols = np.concatenate([idx.ravel(), R + self.centre_offset])
        vals = np.concatenate([w.ravel(), -w.sum(1)])
        M = sp.coo_matrix((vals, (rows, cols)), shape=self.shape).tocsr()
        M.eliminate_zeros()
        return M


def neighbour_csr(G, domain, device):
    """(ptr, J, centre_offset, n_rows) device CSR of the neighbour rows whose centre lies
    in ``domain``; cached on the grid."""
    cache = getattr(G, "_b200nb", None)
    if cache is None:
        cache = {}
        try:
            G._b200nb = cache
        except AttributeError:
            pass
    key = (domain, device.index)
    if key in cache:
        return cache[key]
    N, n = int(G.num_interior), int(G.num_interior + G.num_bdry)
    lo, hi = (0, N) if domain == "interior" else (N, n)
    if hasattr(G, "_count") and getattr(G, "_neighbours", None) is None:   # FDTriMesh, device copy
        count = G._count[lo:hi].to(device)
        mask = torch.arange(G._kmax, device=count.device)[None, :] < count[:, None]
        J = G._nb[lo:hi].to(device)[mask].contiguous()
        ptr = torch.zeros(hi - lo + 1, dtype=torch.int64, device=device)
        ptr[1:] = torch.cumsum(count.long(), 0)
    else:
        nb = getattr(G, "neighbours", None)
        if nb is None:
            raise TypeError("The grid has no neighbour list")
        nb = np.asarray(nb, dtype=np.int64)
        nb = nb[(nb[:, 0] >= lo) & (nb[:, 0] < hi)]
        nb = nb[np.argsort(nb[:, 0], kind="stable")]           # row order inside a centre kept
        counts = np.bincount(nb[:, 0] - lo, minlength=hi - lo)
        if (counts == 0).any():
            raise TypeError("Point {0} has no neighbours".format(int(np.flatnonzero(counts == 0)[0]) + lo))
        ptr = torch.from_numpy(np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)).to(device)
        J = torch.from_numpy(nb[:, 1].astype(np.int32)).to(device)
    out = (ptr, J, lo, hi - lo)
    cache[key] = out
    return out


def _points(G, device):
    P = getattr(G, "_b200pts", None)
    if P is None or P.device != device:
        P = torch.from_numpy(np.ascontiguousarray(G.points, dtype=np.float64)).to(device)
        try:
            G._b200pts = P
        except AttributeError:
            pass
    return P


def _run(G, domain, u, v, want_vout):
    device = require_cuda()
    ptr, J, off, nrows = neighbour_csr(G, domain, device)
    P = _points(G, device)
    U = to_device(u, device) if u is not None else None
    V = torch.from_numpy(np.ascontiguousarray(v, dtype=np.float64)).to(device) if v is not None else None
    d1 = torch.empty(nrows, dtype=torch.float64, device=device) if U is not None else None
    sel = torch.empty(nrows, dtype=torch.int32, device=device)
    w = torch.empty(nrows, dtype=torch.float64, device=device)
    vout = torch.empty((nrows, 2), dtype=torch.float64, device=device) if want_vout else None
    check(_lib.lib().pde_nb_d1(device.index, nrows, off, _ptr(ptr), _ptr(J), _ptr(P), _ptr(U),
                               _ptr(V), _ptr(d1), _ptr(sel), _ptr(w), _ptr(vout), _stream()))
    return device, off, nrows, d1, sel, w, vout


def _matrix(G, device, off, nrows, sel, w):
    centre = torch.arange(off, off + nrows, dtype=torch.int32, device=device)
    idx = torch.stack([sel, centre, centre, centre], 1)
    W = torch.zeros((nrows, 4), dtype=torch.float64, device=device)
    W[:, 0] = w
    return RowsMatrix(idx, W, off, int(G.num_interior + G.num_bdry), device)


def d1(G, v, u=None, jacobian=False, domain="interior"):
    """ddg.d1 (ddg.py:8-81): upwind difference along the neighbour closest in angle to
    ``v``.  Returns ``(d1u, M)``."""
    if u is None:
        jacobian = True
    if G.dim != 2:
        raise NotImplementedError("only 2-D grids are implemented")
    vv = np.ascontiguousarray(_ddutils.process_v(G, v, domain=domain), dtype=np.float64)
    device, off, nrows, d1u, sel, w, _ = _run(G, domain, u, vv, False)
    M = _matrix(G, device, off, nrows, sel, w) if jacobian else None
    return (like_input(u, d1u) if u is not None else None), M


def d1n(G, u=None, **kwargs):
    """ddg.d1n (ddg.py:86-108): derivative along the inward normal on the boundary points."""
    if getattr(G, "bdry_normals", None) is None:
        raise TypeError("The grid has no boundary normals")
    return d1(G, -np.asarray(G.bdry_normals), u=u, domain="boundary", **kwargs)


def d1grad(G, u, jacobian=False, control=False):
    """ddg.d1grad (ddg.py:110-175): largest difference quotient over the neighbours.
    Returns ``(d1_max, M, v)``."""
    if G.dim != 2:
        raise NotImplementedError("only 2-D grids are implemented")
    device, off, nrows, d1u, sel, w, vout = _run(G, "interior", u, None, control)
    M = _matrix(G, device, off, nrows, sel, w) if jacobian else None
    return like_input(u, d1u), M, (like_input(u, vout) if control else None)
