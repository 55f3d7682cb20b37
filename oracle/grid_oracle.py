"""Brute-force restatement of FDRegularGrid(interpolation=False) -- TEST INFRASTRUCTURE.

Any dimension, O(N * (N + Nb)) per grid: usable up to a few hundred interior points.  Follows
/root/reference/pointclasses.py:
  :503-540  lattice points, wall points (verbatim arithmetic, so that exact comparisons on the
            unscaled coordinates give the reference's answers);
  :421-447  neighbours = l-inf distance in (0, r] on the UNSCALED coordinates (the joggled
            Delaunay graph + adjacency powers only pre-select candidates: `depth = dim * r`
            edges always cover that box -- pinned against the reference in make_golden.py);
  :37-103   colinear pruning, tol 1e-3, prefer="min" (SciPy pdist 'cosine'; of a double the
            farther point goes, the first on ties; wall points that only lost to wall points stay);
  :12-35    antipodal pairs: cos < -1 + 1e-3 between kept unit vectors, upper triangular.
Pinned: oracle/make_golden.py asserts pair sets, neighbour sets and points equal to the
unmodified reference constructor on the grids it writes fixtures for.
"""
import numpy as np
from scipy.spatial.distance import pdist

TOL = 1e-3


def create_stencil(r, dim):
    """stencils.py:25-72 (`create`): Stern-Brocot directions in 2-D, gcd == 1 offsets in n-D."""
    if dim == 2:
        sb = np.array([[1, 0], [1, 1], [0, 1]])
        for _ in range(1, r):
            new = []
            for v, w in zip(sb[:-1], sb[1:]):
                new += [v, v + w]
            new.append(sb[-1])
            sb = np.array(new)
        V = sb[:-1]
        V = V[(V <= r).all(1)].T
        R = np.array([[0, -1], [1, 0]])
        return np.concatenate([np.linalg.matrix_power(R, i).dot(V).T for i in range(4)])
    st = (np.indices((2 * r + 1,) * dim) - r).reshape(dim, -1).T
    st = st[(st != 0).any(axis=1)]
    return st[np.gcd.reduce(st, axis=1) == 1]


def unscaled_points(shape, bounds, r):
    """(points on the integer lattice, num_interior), pointclasses.py:513-540."""
    shape = np.array(shape)
    dim = len(shape)
    bounds = np.array(bounds, dtype=float)
    Xint = np.reshape(np.meshgrid(*[np.arange(1, n + 1) for n in shape], indexing="ij"),
                      (dim, -1)).T
    stcl = create_stencil(r, dim)
    sl1 = stcl / np.max(np.abs(stcl), axis=1)[:, None]
    m = ((Xint == 1) | (Xint == shape)).any(axis=1)
    X = Xint[m, :, None] + sl1.T[None, :, :]
    X = np.moveaxis(X, 2, 0).reshape(-1, dim)
    X = X[((X == 0) | (X == shape + 1)).any(axis=1)]
    h = (bounds[1] - bounds[0]) / (shape + 1)
    Xb = (np.unique(h * X + bounds[0], axis=0) - bounds[0]) / h
    return np.concatenate([Xint, Xb]), Xint.shape[0]


def build(shape, bounds, r):
    """-> dict(points, num_interior, neighbours (interior centres), pairs)."""
    bounds = np.array(bounds, dtype=float)
    shape = np.array(shape)
    P, N = unscaled_points(shape, bounds, r)
    scaling = (bounds[1] - bounds[0]) / (shape + 1)
    points = P * scaling + bounds[0]                      # :551
    isb = np.arange(len(P)) >= N
    nb_rows, pair_rows = [], []
    for i in range(N):
        d = np.max(np.abs(P - P[i]), axis=1)
        cand = np.flatnonzero((d <= r) & (d > 0))
        V = points[cand] - points[i]
        D = np.linalg.norm(V, axis=1)
        K = len(cand)
        disc = np.zeros(K, bool)
        disc_mixed = np.zeros(K, bool)
        iu = np.triu_indices(K, 1)
        chk = pdist(V, "cosine") < TOL
        for k, l in zip(iu[0][chk], iu[1][chk]):
            far = k if D[k] >= D[l] else l
            disc[far] = True
            if isb[cand[k]] != isb[cand[l]]:
                disc_mixed[far] = True
        keep = ~disc | (isb[cand] & ~disc_mixed)
        kc = cand[keep]
        nb_rows += [(i, int(j)) for j in kc]
        Vs = (V / D[:, None])[keep]
        a, b = np.nonzero(np.triu(Vs.dot(Vs.T) < -1 + TOL))
        if len(a) == 0:
            raise TypeError("Point {0} has no pairs".format(i))
        pair_rows += [(i, int(kc[x]), int(kc[y])) for x, y in zip(a, b)]
    return dict(points=points, num_interior=N, neighbours=np.array(nb_rows, dtype=np.int64),
                pairs=np.array(pair_rows, dtype=np.int64))


def stencil_table(r):
    """(D, 4) integer offsets (a0, b0, a1, b1) of the antipodal pairs of a deep-interior point of a
    2-D lattice with stencil radius r: the pairs `build` gives the centre point of a small square
    grid (so the colinear-tolerance pruning of pointclasses.py:37-103 is included: radius 6 yields
    the 36 pairs of radius 5).  Used by the CPU baseline so that it needs no product code."""
    n = 4 * r + 3
    bounds = np.array([[0.0, 1.0], [0.0, 1.0]]).T
    b = build([n, n], bounds, r)
    P, _ = unscaled_points([n, n], bounds, r)
    c = (n // 2) * n + n // 2
    rows = b["pairs"][b["pairs"][:, 0] == c]
    return np.concatenate([np.rint(P[rows[:, 1]] - P[c]), np.rint(P[rows[:, 2]] - P[c])],
                          axis=1).astype(np.int64)
