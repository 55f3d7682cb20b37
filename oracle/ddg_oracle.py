"""NumPy oracle for the regular-grid ("antipodal pairs") operators.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  Restates
/root/reference/ddg.py:177-339.  The arithmetic of every floating-point
quantity follows the reference operation by operation so that results are
bit-identical (checked by ``oracle/make_golden.py`` against the real thing):

    X  = points[J] - points[I]                 ddg.py:276
    h  = sqrt(X_x*X_x + X_y*X_y)               ddg.py:277  (np.linalg.norm, axis=2)
    w  = 1/h ;  w0 = 2/(h_0 + h_1)             ddg.py:279-280
    d  = w0 * ((u[J0]-u[I])*w_0 + (u[J1]-u[I])*w_1)      ddg.py:281

Selection rule (ddg.py:283-292, stable ``lexsort((-d, I))``): per centre
point the minimum is the LAST row of ``pairs`` (in given order) attaining the
smallest ``d`` and the maximum is the FIRST row attaining the largest ``d``.
"""
import numpy as np
import scipy.sparse as sp


def pair_geometry(points, pairs):
    """Stencil vectors and lengths of every pair row.  ddg.py:274-280."""
    I = pairs[:, 0]
    J = pairs[:, 1:]
    X = points[J] - points[I, None]
    h = np.sqrt(X[..., 0] * X[..., 0] + X[..., 1] * X[..., 1]) if X.shape[-1] == 2 \
        else np.linalg.norm(X, axis=2)
    w = 1 / h
    w0 = 2 / (h[:, 0] + h[:, 1])
    return I, J, X, h, w, w0


def pair_values(points, pairs, u):
    """Second difference of every pair row, reference operation order."""
    I, J, X, h, w, w0 = pair_geometry(points, pairs)
    du = u[J] - u[I, None]
    t = du * w
    return w0 * (t[:, 0] + t[:, 1])


def _segments(I):
    """Stable grouping of rows by centre index.

    Returns (order, starts, seg_of_sorted_row, centres)."""
    order = np.argsort(I, kind="stable")
    Is = I[order]
    new = np.empty(Is.size, dtype=bool)
    new[0] = True
    new[1:] = Is[1:] != Is[:-1]
    starts = np.flatnonzero(new)
    seg = np.cumsum(new) - 1
    return order, starts, seg, Is[starts]


def select_extremes(I, d):
    """Row numbers (into the original pair list) of the selected min / max pair.

    min -> last row among exact ties, max -> first row (ddg.py:283-292)."""
    order, starts, seg, centres = _segments(I)
    ds = d[order]
    pos = np.arange(ds.size)
    smin = np.minimum.reduceat(ds, starts)
    smax = np.maximum.reduceat(ds, starts)
    at_min = ds == smin[seg]
    at_max = ds == smax[seg]
    last_min = np.maximum.reduceat(np.where(at_min, pos, -1), starts)
    first_max = np.minimum.reduceat(np.where(at_max, pos, ds.size), starts)
    return centres, order[last_min], order[first_max]


def _three_point_matrix(G, rows, I, J, w, w0):
    """CSR (num_interior, num_points) with J0, J1, I entries.  ddg.py:313-323."""
    Ir = I[rows]
    Jr = J[rows]
    wr = w[rows]
    w0r = w0[rows]
    centre = (-wr[:, 0]) + (-wr[:, 1])          # duplicate (I,I) entries summed by tocsr
    r = np.concatenate([Ir, Ir, Ir])
    c = np.concatenate([Jr[:, 0], Jr[:, 1], Ir])
    v = np.concatenate([w0r * wr[:, 0], w0r * wr[:, 1], w0r * centre])
    return sp.coo_matrix((v, (r, c)), shape=(G.num_interior, G.num_points)).tocsr()


def d2eigs(G, u, jacobian=False, control=False):
    """ddg.d2eigs restated (ddg.py:251-327).  Same return convention."""
    pairs = np.asarray(G.pairs)
    points = np.asarray(G.points)
    u = np.asarray(u, dtype=np.float64)
    I, J, X, h, w, w0 = pair_geometry(points, pairs)
    du = u[J] - u[I, None]
    t = du * w
    d = w0 * (t[:, 0] + t[:, 1])
    centres, rmin, rmax = select_extremes(I, d)
    lam_min, lam_max = d[rmin], d[rmax]

    v_min = v_max = None
    if control:
        what = w[rmin, 0]                       # quirk Q1: BOTH scaled by the min pair's 1/h
        v_min = X[rmin, 0] * what[:, None]      # ddg.py:299-300
        v_max = X[rmax, 0] * what[:, None]
    M_min = M_max = None
    if jacobian:
        M_min = _three_point_matrix(G, rmin, I, J, w, w0)
        M_max = _three_point_matrix(G, rmax, I, J, w, w0)
    return (lam_min, M_min, v_min), (lam_max, M_max, v_max)


def d2eigs_policy(G, u):
    """Like d2eigs but also returns the selected pair-row numbers (the policy)."""
    pairs = np.asarray(G.pairs)
    d = pair_values(np.asarray(G.points), pairs, np.asarray(u, dtype=np.float64))
    centres, rmin, rmax = select_extremes(pairs[:, 0], d)
    return d[rmin], d[rmax], rmin, rmax


def d2min(G, u, **kw):
    return d2eigs(G, u, **kw)[0]


def d2max(G, u, **kw):
    return d2eigs(G, u, **kw)[1]


def process_v(G, v, domain="interior"):
    """_ddutils.process_v restated (_ddutils.py:6-47), 2-D/3-D cases used by d2."""
    N = G.num_interior if domain == "interior" else G.num_bdry
    dim = G.dim
    v = np.array(v)
    if v.size == 1 and dim == 2:
        v = np.broadcast_to([np.cos(v), np.sin(v)], (N, dim))
    elif v.size == 2 and dim == 3:
        v = np.broadcast_to([np.sin(v[0]) * np.cos(v[1]), np.sin(v[0]) * np.sin(v[1]),
                             np.cos(v[1])], (N, dim))
    elif v.size == dim:
        v = np.broadcast_to(v / np.linalg.norm(v), (N, dim))
    elif v.size == N and dim == 2:
        v = np.array([np.cos(v), np.sin(v)]).T
    elif v.shape == (N, 2) and dim == 3:
        v = np.array([np.sin(v[:, 0]) * np.cos(v[:, 1]), np.sin(v[:, 0]) * np.sin(v[:, 1]),
                      np.cos(v[:, 1])]).T
    elif v.shape == (N, dim):
        v = v / np.linalg.norm(v, axis=1)[:, None]
    return v


def d2(G, v, u=None, jacobian=False):
    """ddg.d2 restated (ddg.py:177-249): per point, the pair best aligned with v
    (largest cos^2 sum, first row among ties), same 3-point formula."""
    if G.pairs is None:
        raise TypeError("The grid has no finite difference pairs")
    if u is None:
        jacobian = True
    v = process_v(G, v)
    pairs = np.asarray(G.pairs)
    points = np.asarray(G.points)
    I, J, X, h, w, w0 = pair_geometry(points, pairs)
    Xs = X / h[:, :, None]
    # interior index == graph index (interior-first numbering, pointclasses.py:200-201)
    vv = v[I]
    cos_sq = np.einsum("lji,li->lj", Xs, vv) ** 2
    score = np.sum(cos_sq, axis=1)
    order, starts, seg, centres = _segments(I)
    ss = score[order]
    smax = np.maximum.reduceat(ss, starts)
    pos = np.arange(ss.size)
    first = np.minimum.reduceat(np.where(ss == smax[seg], pos, ss.size), starts)
    rows = order[first]
    d2u = None
    if u is not None:
        u = np.asarray(u, dtype=np.float64)
        # ddg.py:236 uses np.sum over the 2-vector: same as t0 + t1
        du = u[J[rows]] - u[I[rows], None]
        t = du * w[rows]
        d2u = w0[rows] * (t[:, 0] + t[:, 1])
    M = _three_point_matrix(G, rows, I, J, w, w0) if jacobian else None
    return d2u, M
