"""NumPy restatement of the first-derivative operators -- TEST INFRASTRUCTURE.

ddg.d1 (ddg.py:8-81), ddg.d1grad (ddg.py:110-175) on a neighbour list and ddi.d1
(ddi.py:12-92), ddi.d1grad (ddi.py:124-221) on simplices, written per centre point with
the reference's expressions (norm, normalisation, dot product, 1/h; batched 2x2 LAPACK
solve for the cone coordinates).  Pinned against the unmodified reference by
oracle/make_golden.py::d1_case.  Only tests/, smoke() and bench's cpu_baseline may use it.
"""
import numpy as np

EPS = 1e-15                                       # ddi.py:10


def _rows(G, domain):
    N = G.num_interior
    lo, hi = (0, N) if domain == "interior" else (N, G.num_points)
    nb = np.asarray(G.neighbours)
    nb = nb[(nb[:, 0] >= lo) & (nb[:, 0] < hi)]
    return nb[np.argsort(nb[:, 0], kind="stable")], lo, hi


def ddg_d1(G, v, u, domain="interior"):
    """(d1u, J_selected, w) per domain point; v is the (n_domain, 2) array of process_v."""
    nb, lo, hi = _rows(G, domain)
    I, J = nb[:, 0], nb[:, 1]
    X = G.points[J] - G.points[I]
    h = np.linalg.norm(X, axis=1)
    Xs = X / h[:, None]
    C = np.einsum('ij,ij->i', Xs, v[I - lo])                      # ddg.py:59
    ind = np.lexsort((-C, I))                                     # ddg.py:60
    I, J, h = I[ind], J[ind], h[ind]
    m = np.concatenate([[True], I[1:] > I[:-1]])
    I, J, h = I[m], J[m], h[m]
    w = 1 / h
    return w * (u[J] - u[I]), J, w


def ddg_d1grad(G, u):
    """(d1_max, J_selected, w, v) per interior point (ddg.py:146-160)."""
    nb, lo, hi = _rows(G, "interior")
    I, J = nb[:, 0], nb[:, 1]
    X = G.points[J] - G.points[I]
    h = np.linalg.norm(X, axis=1)
    w = 1 / h
    d = w * (u[J] - u[I])
    ind = np.lexsort((-d, I))
    Is = I[ind]
    m = np.concatenate([[True], Is[1:] > Is[:-1]])
    return d[ind][m], J[ind][m], w[ind][m], w[ind][m][:, None] * X[ind][m]


def ddi_d1(G, v, u, domain="interior"):
    """(d1u, S (n,2), xi (n,2)) per domain point (ddi.py:49-80)."""
    N = G.num_interior
    lo, hi = (0, N) if domain == "interior" else (N, G.num_points)
    S = np.asarray(G.simplices)
    S = S[(S[:, 0] >= lo) & (S[:, 0] < hi)]
    S = S[np.argsort(S[:, 0], kind="stable")]
    I, T = S[:, 0], S[:, 1:]
    X = np.swapaxes(G.points[T] - G.points[I, None], 1, 2)
    Xi = np.linalg.solve(X, v[I - lo][..., None])[..., 0]
    ok = (Xi >= -EPS / G.spatial_resolution)
    if domain == "boundary":
        ok &= T < N
    ok = ok.all(axis=1)
    _, first = np.unique(I[ok], return_index=True)
    Sf, Xif = T[ok][first], Xi[ok][first]
    c = np.arange(lo, hi)
    return np.einsum('ij,ij->i', u[Sf] - u[c, None], Xif), Sf, Xif


def ddi_d1grad(G, u, nds):
    th = np.arange(0, 2 * np.pi, 2 * np.pi / nds)
    V = np.stack([np.cos(th), np.sin(th)], axis=1)
    N = G.num_interior
    d = np.stack([ddi_d1(G, np.broadcast_to(vk, (N, 2)), u)[0] for vk in V], axis=1)
    ix = d.argmax(axis=1)
    return d[np.arange(N), ix], ix, V[ix]
