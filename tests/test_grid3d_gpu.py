"""GPU: 3-D pairs grids from the fast constructor through the CUDA operators
and solvers, against the outputs of the unmodified reference on the same boxes
(tests/golden/grid3d_*.npz, written by oracle/make_golden.py::grid3d_case) and against the
NumPy restatement at a size the reference constructor cannot build."""
import glob
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import ddg_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GRID3D = sorted(glob.glob(os.path.join(GOLD, "grid3d_*.npz")))


def grid_of(z):
    from pyellipticfd_b200.pointclasses import FDRegularGrid
    return FDRegularGrid([int(s) for s in z["shape"]], z["bounds"], int(z["radius"]),
                         interpolation=False)


@pytest.mark.parametrize("path", GRID3D, ids=[os.path.basename(p) for p in GRID3D])
def test_operators_and_solves_match_reference(path):
    from pyellipticfd_b200 import ddg, convex_envelope
    z = np.load(path)
    G = grid_of(z)
    rng = np.random.default_rng(0)
    X, Y, Z = G.points.T
    for fname, u in (("random", rng.standard_normal(G.num_points)),
                     ("quad", X ** 2 - Y ** 2 + 2 * Z ** 2)):
        (lmin, Mmin, vmin), (lmax, Mmax, vmax) = ddg.d2eigs(G, u, jacobian=True, control=True)
        assert np.array_equal(lmin, z["lmin_" + fname]) and np.array_equal(lmax, z["lmax_" + fname])
        assert np.array_equal(ddg.d2min(G, u)[0], lmin) and np.array_equal(ddg.d2max(G, u)[0], lmax)
        assert np.abs(Mmin.dot(u) - lmin).max() <= 1e-12 * np.abs(lmin).max()
        assert np.abs(Mmax.dot(u) - lmax).max() <= 1e-12 * np.abs(lmax).max()
        if fname == "random":                   # no ties: same pair, possibly stored reversed
            for v, ref in ((vmin, z["vmin_random"]), (vmax, z["vmax_random"])):
                assert v.shape == (G.num_interior, 3)
                c = (v * ref).sum(1) / np.sqrt((v ** 2).sum(1) * (ref ** 2).sum(1))
                assert (np.abs(np.abs(c) - 1) < 1e-3).all()
    if "euler_U" in z.files:
        g = z["g"]
        U, diff, it, _ = convex_envelope.solve(G, g, fdmethod="grid", solver="euler")
        assert it == int(z["euler_iters"]) and np.array_equal(U, z["euler_U"])       # bit-exact
        assert diff == float(z["euler_diff"])
        U, diff, it, _ = convex_envelope.solve(G, g, fdmethod="grid", solver="newton")
        assert abs(it - int(z["newton_iters"])) <= 4
        assert np.abs(U - z["newton_U"]).max() < 2e-6          # reference solver tolerance 1e-6


def test_33cubed_against_the_numpy_restatement():
    """35 937 interior points, r = 2 (49 pairs per deep point, 1.75 M pair rows): beyond the
    reference constructor (O(N^2)); values bit-exact against the NumPy restatement of
    ddg.d2eigs on the same pair rows, known answer for a quadratic."""
    from pyellipticfd_b200 import ddg
    from pyellipticfd_b200.pointclasses import FDRegularGrid
    G = FDRegularGrid([33, 33, 33], np.array([[-1.0, 1.0]] * 3).T, 2, interpolation=False)
    assert G.num_interior == 33 ** 3 and G.dim == 3
    u = np.random.default_rng(1).standard_normal(G.num_points)
    (lmin, _, _), (lmax, _, _) = ddg.d2eigs(G, u)
    (omin, _, _), (omax, _, _) = O.d2eigs(G, u)
    assert np.array_equal(lmin, omin) and np.array_equal(lmax, omax)
    X, Y, Z = G.points.T
    lam, _, v = ddg.d2eigs(G, X ** 2 - Y ** 2 + 2 * Z ** 2, control=True)[0]
    assert np.abs(lam + 2).max() < 1e-9                        # d2/dy2 of -y^2, exact for quadratics
    lam = ddg.d2max(G, X ** 2 - Y ** 2 + 2 * Z ** 2)[0]
    assert np.abs(lam - 4).max() < 1e-9
