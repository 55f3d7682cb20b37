"""NumPy oracle for the point-cloud operators (ddi: linear interpolation, ddf: Froese).

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  Restates
/root/reference/ddi.py:223-414 and /root/reference/ddf.py:8-252 (the paths taken
when the grid has simplices and no pairs).  ``G`` needs ``points``, ``simplices``,
``num_interior``, ``spatial_resolution``, ``angular_resolution``.
"""
import numpy as np
import scipy.sparse as sp

from .ddg_oracle import process_v

EPS = 1e-15      # ddi.py:10


def _solve(X, v):
    return np.linalg.solve(X, v[..., None])[..., 0]


def _first_per_point(I, mask):
    """Row numbers (into the unmasked arrays) of the first masked row of every point."""
    rows = np.flatnonzero(mask)
    _, first = np.unique(I[rows], return_index=True)
    return rows[first]


def num_directions(G, froese=False):
    if froese:
        return int(np.max([int(np.ceil(2 * np.pi / G.angular_resolution)), 4]))      # ddf.py:133-135
    return int(np.max([int(2 * np.ceil(np.pi ** 2 / G.angular_resolution ** 2)), 4]))  # ddi.py:120-122


def rows_for_direction(G, v, froese=False):
    """(S (N,4) neighbour ids, W (N,4) weights) of the 5-point row of every interior
    point for per-point unit directions v (N,2)."""
    N = G.num_interior
    P = np.asarray(G.points)
    simp = np.asarray(G.simplices)
    simp = simp[simp[:, 0] < N]
    I, S = simp[:, 0], simp[:, 1:]
    X = np.swapaxes(P[S] - P[I, None], 1, 2)
    Xi = _solve(X, v[I])
    if froese:
        mf, mb = (Xi >= 0).all(1), (Xi <= 0).all(1)                   # ddf.py:62-63
    else:
        tol = EPS / G.spatial_resolution
        mf, mb = (Xi >= -tol).all(1), (Xi <= tol).all(1)              # ddi.py:267-268
    rf, rb = _first_per_point(I, mf), _first_per_point(I, mb)
    assert rf.size == N and rb.size == N, "a direction is not enclosed by any simplex"
    Sf, Sb = S[rf], S[rb]
    if not froese:
        Xif, Xib = Xi[rf], -Xi[rb]
        hf, hb = 1 / Xif.sum(1), 1 / Xib.sum(1)                      # ddi.py:279-280
        sc = 2 / (hf + hb)
        return np.concatenate([Sf, Sb], 1), np.concatenate([Xif, Xib], 1) * sc[:, None]
    # Froese, ddf.py:70-116 (with the backward simplex taken from the backward match)
    Xc = np.concatenate([X[rf], X[rb]], axis=2)
    h = np.linalg.norm(Xc, axis=1)
    Xth = np.arctan2(Xc[:, 1, :], Xc[:, 0, :])
    vth = np.arctan2(v[:, 1], v[:, 0])
    Sc = np.concatenate([Sf, Sb], axis=1)
    dth = (Xth - vth[:, None]) % (2 * np.pi)
    js = np.argsort(dth, kind="stable")
    ii = np.arange(N)[:, None]
    dth, Sc, h = dth[ii, js], Sc[ii, js], h[ii, js]
    c, s = h * np.cos(dth), h * np.sin(dth)
    m = (c[:, 0] >= 0) & (c[:, 1] >= 0) & (s[:, 0] >= 0) & (s[:, 1] >= 0)
    c[m], s[m], Sc[m] = np.roll(c[m], -1, 1), np.roll(s[m], -1, 1), np.roll(Sc[m], -1, 1)
    A = c[:, 2] * s[:, 1] - c[:, 1] * s[:, 2]
    B = c[:, 0] * s[:, 3] - c[:, 3] * s[:, 0]
    den = A * (c[:, 0] ** 2 * s[:, 3] - c[:, 3] ** 2 * s[:, 0]) - \
        B * (c[:, 2] ** 2 * s[:, 1] - c[:, 1] ** 2 * s[:, 2])
    a = np.stack([2 * s[:, 3] * A / den, 2 * s[:, 2] * B / den,
                  -2 * s[:, 1] * B / den, -2 * s[:, 0] * A / den], 1)
    return Sc, a


def matrix(G, S, W):
    N = G.num_interior
    I = np.arange(N)
    rows = np.concatenate([np.repeat(I, 4), np.repeat(I, 4)])
    cols = np.concatenate([S.ravel(), np.repeat(I, 4)])
    vals = np.concatenate([W.ravel(), -W.ravel()])
    return sp.coo_matrix((vals, (rows, cols)), shape=(N, len(G.points))).tocsr()


def d2(G, v, u=None, jacobian=False, froese=False):
    v = np.ascontiguousarray(process_v(G, v), dtype=np.float64)
    S, W = rows_for_direction(G, v, froese)
    d2u = None
    if u is not None:
        u = np.asarray(u)
        d2u = np.sum(W * (u[S] - u[:G.num_interior, None]), axis=1)
    return d2u, (matrix(G, S, W) if jacobian or u is None else None)


def d2eigs_table(G, froese=False):
    nds = num_directions(G, froese)
    th = np.arange(0, 2 * np.pi, 2 * np.pi / nds)
    V = np.stack([np.cos(th), np.sin(th)], axis=1)
    rows = [rows_for_direction(G, np.broadcast_to(vk, (G.num_interior, 2)), froese) for vk in V]
    return V, np.stack([r[0] for r in rows]), np.stack([r[1] for r in rows])   # (nds, N, 4)


def d2eigs(G, u, froese=False, table=None):
    """Values only: (lmin, lmax, d2u (N, nds))."""
    V, S, W = table if table is not None else d2eigs_table(G, froese)
    u = np.asarray(u)
    d2u = np.sum(W * (u[S] - u[None, :G.num_interior, None]), axis=2).T
    return d2u.min(1), d2u.max(1), d2u
