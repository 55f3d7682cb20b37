"""NumPy/SciPy restatement of the stencil search of FDTriMesh -- TEST INFRASTRUCTURE.

Follows pointclasses.py:367-398 (edges, graph depth, adjacency powers, radius mask),
:37-103 (remove_colinear_neighbours, prefer="max") and :105-121 (compute_simplices; in 2-D
the hull of unit vectors = consecutive neighbours in angular order).  Brute force, small
clouds only.  Pinned against the unmodified reference by oracle/make_golden.py::trimesh_case.
"""
import itertools

import numpy as np
from scipy.sparse import coo_matrix
from scipy.spatial.distance import pdist


def stencils(p, t, num_interior, angular_resolution, spatial_resolution):
    """Returns (neighbours (T,2) sorted by (I,J), simplices (S,3) with sorted vertex pairs,
    sorted by row)."""
    n = p.shape[0]
    N = num_interior
    Cd = 2
    th, h = angular_resolution, spatial_resolution
    max_radius = h * (1 + Cd / np.sin(th / 2))                       # :259-264
    min_radius = h * (-1 + Cd / np.sin(th / 2))                      # :266-271
    edges = np.concatenate([np.array([v1, v2]).T
                            for v1, v2 in itertools.permutations(t.T, 2)])      # :368-369
    L = np.linalg.norm(p[edges[:, 1]] - p[edges[:, 0]], axis=1)
    min_nb = L[edges[:, 0] < N].min()                                # :283-285 on the edges
    depth = int(np.ceil(max_radius / min_nb))                        # :373
    A = coo_matrix((np.ones(len(edges), dtype=np.intp), (edges[:, 0], edges[:, 1])),
                   shape=(n, n), dtype=np.intp).tocsr()
    nb = edges
    if depth > 1:
        pows = [A]
        for k in range(1, depth):
            pows.append(pows[k - 1].dot(A))
        S = sum(pows).tocoo()
        nb = np.array([S.row, S.col]).T
    V = p[nb[:, 1]] - p[nb[:, 0]]
    D = np.linalg.norm(V, axis=1)
    mask = (D <= max_radius) & ((D >= min_radius) | (nb[:, 1] >= N)) & (D > 0)   # :387-390
    nb = nb[mask]
    nb = nb[np.lexsort((nb[:, 1], nb[:, 0]))]
    tol = 1 - np.cos(angular_resolution / 8)                        # :395
    keep_rows, simp = [], []
    for i in range(n):
        J = nb[nb[:, 0] == i, 1]
        v = p[J] - p[i]
        d = np.linalg.norm(v, axis=1)
        nv = len(J)
        keep = np.arange(nv)
        if nv > 1:
            check = pdist(v, 'cosine') < tol                        # :51-52
            if check.any():
                ij = np.array(list(itertools.combinations(range(nv), 2)))
                doubles = ij[check]
                arg = np.argmin(d[doubles], axis=1)                 # prefer == "max": :63-64
                ix = np.arange(len(doubles))
                discard = doubles[ix, arg]
                keep = np.arange(nv)[~np.isin(np.arange(nv), discard)]
                isb = J >= N
                mixed = isb[doubles[:, 0]] ^ isb[doubles[:, 1]]     # :80-83
                bdry_discard = doubles[ix[mixed], arg[mixed]]
                boundary = np.arange(nv)[isb]
                boundary_keep = boundary[~np.isin(boundary, bdry_discard)]
                keep = np.unique(np.concatenate([boundary_keep, keep]))
        Jk = J[keep]
        keep_rows.append(np.stack([np.full(len(Jk), i), Jk], 1))
        vk = v[keep]
        order = np.argsort(np.arctan2(vk[:, 1], vk[:, 0]), kind="stable")
        ring = Jk[order]
        nxt = np.roll(ring, -1)
        simp.append(np.stack([np.full(len(ring), i), np.minimum(ring, nxt), np.maximum(ring, nxt)], 1))
    nbk = np.concatenate(keep_rows)
    S = np.concatenate(simp)
    return nbk, S[np.lexsort((S[:, 2], S[:, 1], S[:, 0]))]
