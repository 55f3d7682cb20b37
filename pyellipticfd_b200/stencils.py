"""Lattice direction sets (mirror of /root/reference/stencils.py:6-72).

``create(r, d)`` returns every primitive integer vector (gcd of the components
is 1) with l-inf norm at most ``r``.  In 2-D this is the same *set* the
reference builds from a Stern-Brocot tree rotated four times
(stencils.py:25-43): 8, 16, 32, 48, 80, 96 vectors for r = 1..6; in higher
dimension it is the reference's own gcd construction (stencils.py:46-61).
Only the set matters downstream (wall-point generation and the angular
resolution), so rows are returned in plain lexicographic order.
"""
import math

import numpy as np


def stern_brocot(N):
    """Stern-Brocot tree of depth N (stencils.py:6-23): rows (p, q) from (1, 0) to (0, 1), every
    level inserting the mediant between neighbours."""
    sb = np.array([[1, 0], [1, 1], [0, 1]], dtype=np.intp)
    for _ in range(1, N):
        new = np.empty((2 * sb.shape[0] - 1, 2), dtype=np.intp)
        new[0::2] = sb
        new[1::2] = sb[:-1] + sb[1:]
        sb = new
    return sb


def create_nd(r, n):
    if r < 1:
        raise ValueError('r must be positive')
    axes = [np.arange(-r, r + 1)] * n
    V = np.stack(np.meshgrid(*axes, indexing='ij'), axis=-1).reshape(-1, n)
    g = np.array([math.gcd(*map(int, v)) for v in V])
    return V[g == 1].astype(np.intp)


def create_2d(r):
    return create_nd(r, 2)


def create(r, d):
    if d < 2:
        raise ValueError('dimension must be two or higher')
    return create_nd(r, d)
