"""CPU: the native (C++/OpenMP) band builder of FDRegularGrid against its NumPy specification,
and structural invariants of the fast grid, on randomly drawn shapes (hypothesis)."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from pyellipticfd_b200.pointclasses import FDRegularGrid


def _canon(rows):
    rows = np.asarray(rows)
    return rows[np.lexsort(tuple(rows[:, k] for k in range(rows.shape[1] - 1, -1, -1)))]


@settings(max_examples=12, deadline=None)
@given(nx=st.integers(7, 23), ny=st.integers(7, 23), r=st.integers(1, 3),
       sx=st.sampled_from([1.0, 0.5, 0.37, 1.7]), sy=st.sampled_from([1.0, 0.25, 0.61]),
       x0=st.sampled_from([0.0, -1.0, 0.3]))
def test_native_band_builder_equals_numpy_specification(nx, ny, r, sx, sy, x0):
    if min(nx, ny) < 2 * r + 1:
        return
    bounds = np.array([[x0, x0 + sx * (nx + 1) / 8.0], [0.0, sy * (ny + 1) / 8.0]]).T
    A = FDRegularGrid([nx, ny], bounds, r, interpolation=False, native=True)
    B = FDRegularGrid([nx, ny], bounds, r, interpolation=False, native=False)
    assert np.array_equal(A.band_index, B.band_index)
    assert np.array_equal(A.band_ptr, B.band_ptr)
    assert np.array_equal(A.band_pairs, B.band_pairs)             # same rows, same order
    assert np.array_equal(A.band_nb_ptr, B.band_nb_ptr)
    assert np.array_equal(_canon(A.neighbours), _canon(B.neighbours))
    assert A.min_interior_nb_dist == B.min_interior_nb_dist
    assert np.array_equal(A.stencil_pairs, B.stencil_pairs)


@settings(max_examples=8, deadline=None)
@given(n=st.integers(9, 21), r=st.integers(1, 4))
def test_grid_invariants(n, r):
    if n < 2 * r + 3:
        return
    G = FDRegularGrid([n, n], np.array([[0.0, 1.0], [0.0, 1.0]]).T, r, interpolation=False)
    P = G.pairs
    N = G.num_interior
    # every interior point has pairs; rows grouped by ascending centre; ids in range
    assert np.array_equal(np.unique(P[:, 0]), np.arange(N))
    assert (np.diff(P[:, 0]) >= 0).all() and P.min() >= 0 and P.max() < G.num_points
    # a pair is (almost) antipodal: cos < -1 + 1e-3 (pointclasses.py:24)
    X0 = G.points[P[:, 1]] - G.points[P[:, 0]]
    X1 = G.points[P[:, 2]] - G.points[P[:, 0]]
    cos = (X0 * X1).sum(1) / np.sqrt((X0 ** 2).sum(1) * (X1 ** 2).sum(1))
    assert (cos < -1 + 1e-3).all()
    # neighbours lie in the l-inf box of radius r (unscaled coordinates) and are not the centre
    nb = G.neighbours
    d = np.abs(G.points[nb[:, 1]] - G.points[nb[:, 0]]) / G.scaling
    assert (d.max(1) <= r + 1e-9).all() and (d.max(1) > 0).all()
    # every pair member is a neighbour of its centre
    key = set(map(tuple, nb.tolist()))
    sel = np.random.default_rng(0).integers(0, len(P), 200)
    assert all((int(P[k, 0]), int(P[k, 1])) in key and (int(P[k, 0]), int(P[k, 2])) in key for k in sel)
    # deep-interior table: antipodal offsets, each direction once
    sp = G.stencil_pairs
    assert np.array_equal(sp[:, :2], -sp[:, 2:]) and len({tuple(x) for x in sp[:, :2]}) == len(sp)


import pytest


@pytest.mark.parametrize("nx,ny,r,bounds", [
    (12, 13, 1, [[0, 0], [0.37 * 13 / 8, 0.61 * 14 / 8]]),     # wall coordinate 13.999999999999998
    (7, 13, 1, [[0, 0], [1.0, 0.61 * 14 / 8]]),
    (40, 40, 2, [[0, 0], [1, 1]]),
    (23, 31, 3, [[0.1, -0.3], [1.3, 0.77]])])
def test_rounded_wall_coordinates_against_the_brute_force_restatement(nx, ny, r, bounds):
    """Lattices whose unscaled wall coordinates carry rounding (pointclasses.py:538 gives
    4000.9999999999995 for 4001 on [0,1]^2 with n = 4000): the wall lists of the fast constructor
    must still find those points -- pair sets equal to oracle/grid_oracle.py (pinned against the
    reference constructor), native builder and NumPy specification alike."""
    from oracle import grid_oracle
    b = np.array(bounds, dtype=float)
    ref = grid_oracle.build([nx, ny], b, r)

    def canon(P):
        return sorted((int(a),) + tuple(sorted((int(x), int(y)))) for a, x, y in P)
    for native in (True, False):
        G = FDRegularGrid([nx, ny], b, r, interpolation=False, native=native)
        assert canon(G.pairs) == canon(ref["pairs"])


def test_large_inexact_lattice_builds():
    G = FDRegularGrid([4000, 4000], np.array([[0.0, 1.0], [0.0, 1.0]]).T, 4, interpolation=False)
    assert not G.uniform_weights() and G._pairs is None and G.band_pairs.shape[0] > 2_000_000
