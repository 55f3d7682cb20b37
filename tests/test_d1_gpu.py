"""GPU parity of the first-derivative operators (ddg.d1 / d1n / d1grad, ddi.d1 / d1n /
d1grad) against the reference's outputs (tests/golden/d1_ops.npz) and the NumPy oracle."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLD, "d1_ops.npz"))


@pytest.fixture(scope="module")
def reg_grid():
    from pyellipticfd_b200.pointclasses import FDRegularGrid
    return FDRegularGrid([9, 9], np.array([[0, 1], [0, 1]]).T, 2, interpolation=False)


def _disc(h0):
    from scipy.spatial import Delaunay
    from oracle import disc
    from pyellipticfd_b200.pointclasses import FDTriMesh
    theta = np.pi * h0 ** (1 / 3)
    p, nint = disc.disc_points(h0, theta)
    return FDTriMesh(p, Delaunay(p).simplices, num_interior=nint, angular_resolution=theta,
                     bdry_normals=p[nint:])


def test_ddg_d1_and_grad_match_reference(gold, reg_grid):
    """Bit-exact against the reference on tests/test_ddg.py's grid (9^2, r = 2), plus the
    reference's own assertions (tests/test_ddg.py:31-49)."""
    import pyellipticfd_b200.ddg as ddg
    G = reg_grid
    X, Y = G.points.T
    flds = {"lin": X - Y, "random": np.random.default_rng(0).standard_normal(len(X))}
    for fname, u in flds.items():
        for k, v in enumerate(([1, 0], [0, 1], 0.3)):
            d, M = ddg.d1(G, v, u, jacobian=True)
            assert np.array_equal(d, gold["ddg_d1_%s_%d" % (fname, k)])
            assert np.abs(M.dot(u) - d).max() < 1e-12
            assert np.abs(M.tocsr().dot(u) - d).max() < 1e-12 * max(1.0, np.abs(d).max())
            assert M.shape == (G.num_interior, G.num_interior + G.num_bdry)
        d, M, vv = ddg.d1grad(G, u, jacobian=True, control=True)
        assert np.array_equal(d, gold["ddg_d1grad_%s" % fname])
        assert np.abs(M.dot(u) - d).max() < 1e-12
        if fname == "random":
            assert np.array_equal(vv, gold["ddg_d1grad_v_%s" % fname])
    U1 = flds["lin"]
    d1x, _ = ddg.d1(G, [1, 0], U1)
    d1y, _ = ddg.d1(G, [0, 1], U1)
    assert np.abs(d1x - 1).max() < 1e-8 and np.abs(d1y + 1).max() < 1e-8
    dgrad, _, vgrad = ddg.d1grad(G, U1, control=True)
    assert np.abs(dgrad - np.sqrt(2)).max() < 1e-8
    assert np.abs(vgrad.dot(1 / np.sqrt(2) * np.array([1, -1])) - 1).max() < 1e-8
    _, M = ddg.d1(G, [1, 0])                       # u=None -> Jacobian only (ddg.py:35-36)
    assert M is not None


@pytest.mark.parametrize("name,h0", [("h0.1", 0.1), ("h0.05", 0.05)])
def test_disc_first_derivatives_match_reference(gold, name, h0):
    import pyellipticfd_b200.ddg as ddg
    import pyellipticfd_b200.ddi as ddi
    from oracle import d1_oracle
    from pyellipticfd_b200._ddutils import process_v
    G = _disc(h0)
    p = G.points
    X, Y = p.T
    Un = np.linalg.norm(p, axis=1)
    rnd = np.random.default_rng(1).standard_normal(len(p))
    lin = X - Y
    # neighbour-list operators: bit-exact
    d, M = ddg.d1n(G, u=Un, jacobian=True)
    assert np.array_equal(d, gold["%s_ddg_d1n" % name])
    assert np.abs(M.dot(Un) - d).max() < 1e-12 and M.shape == (G.num_bdry, G.num_points)
    assert np.array_equal(ddg.d1grad(G, rnd)[0], gold["%s_ddg_d1grad_random" % name])
    # interpolated operators: 1e-9 relative (2x2 solves, summation order)
    for k, v in enumerate(([1, 0], 0.3)):
        for fname, u in (("lin", lin), ("random", rnd)):
            d, M = ddi.d1(G, v, u, jacobian=True)
            ref = gold["%s_ddi_d1_%s_%d" % (name, fname, k)]
            assert np.abs(d - ref).max() < 1e-9 * np.abs(ref).max()
            assert np.abs(M.dot(u) - d).max() < 1e-12 * max(1.0, np.abs(d).max())
            od, _, _ = d1_oracle.ddi_d1(G, process_v(G, v), u)
            assert np.abs(d - od).max() < 1e-9 * np.abs(od).max()
    d, M = ddi.d1n(G, u=Un, jacobian=True)
    assert np.abs(d - gold["%s_ddi_d1n" % name]).max() < 1e-9
    assert np.abs(M.dot(Un) - d).max() < 1e-12
    for fname, u in (("lin", lin), ("random", rnd)):
        for call in range(2):                      # second call: cached table
            d, M, vv = ddi.d1grad(G, u, jacobian=True, control=True)
            ref = gold["%s_ddi_d1grad_%s" % (name, fname)]
            assert np.abs(d - ref).max() < 1e-9 * np.abs(ref).max()
            assert np.abs(M.dot(u) - d).max() < 1e-12 * max(1.0, np.abs(d).max())
            if fname == "random":
                assert np.array_equal(vv, gold["%s_ddi_d1grad_v_%s" % (name, fname)])
    # the reference's own assertions (tests/test_ddi.py:18-43) at this coarser resolution
    d1x, _ = ddi.d1(G, [1, 0], lin)
    assert np.abs(d1x - 1).max() < 1e-4
    dgrad, _, vgrad = ddi.d1grad(G, lin, control=True)
    assert np.abs(dgrad - np.sqrt(2)).max() < 5e-2
