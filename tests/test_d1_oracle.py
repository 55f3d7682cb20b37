"""First-derivative operators: the NumPy restatement (oracle/d1_oracle.py) on grids built
by the fast constructors reproduces the reference's outputs on ITS grids (CPU only)."""
import os

import numpy as np
import pytest

from oracle import d1_oracle, disc
from pyellipticfd_b200._ddutils import process_v
from pyellipticfd_b200.pointclasses import FDRegularGrid, FDTriMesh

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_ddg_d1_on_fast_grid():
    z = np.load(os.path.join(GOLD, "d1_ops.npz"))
    G = FDRegularGrid([9, 9], np.array([[0, 1], [0, 1]]).T, 2, interpolation=False)
    nb = np.asarray(G.neighbours)
    assert np.array_equal(nb[np.lexsort((nb[:, 1], nb[:, 0]))], z["grid_neighbours"])
    X, Y = G.points.T
    flds = {"lin": X - Y, "random": np.random.default_rng(0).standard_normal(len(X))}
    for fname, u in flds.items():
        for k, v in enumerate(([1, 0], [0, 1], 0.3)):
            d, _, _ = d1_oracle.ddg_d1(G, process_v(G, v), u)
            assert np.array_equal(d, z["ddg_d1_%s_%d" % (fname, k)])
        d, _, _, vv = d1_oracle.ddg_d1grad(G, u)
        assert np.array_equal(d, z["ddg_d1grad_%s" % fname])
        if fname == "random":                       # the linear field ties between neighbours
            assert np.array_equal(vv, z["ddg_d1grad_v_%s" % fname])
    # tests/test_ddg.py:31-49
    d, _, _ = d1_oracle.ddg_d1(G, process_v(G, [1, 0]), flds["lin"])
    assert np.abs(d - 1).max() < 1e-8


def test_ddi_d1_on_fast_trimesh():
    from scipy.spatial import Delaunay
    z = np.load(os.path.join(GOLD, "d1_ops.npz"))
    h0 = 0.1
    theta = np.pi * h0 ** (1 / 3)
    p, nint = disc.disc_points(h0, theta)
    G = FDTriMesh(p, Delaunay(p).simplices, num_interior=nint, angular_resolution=theta,
                  bdry_normals=p[nint:], device="cpu")
    X, Y = p.T
    Un = np.linalg.norm(p, axis=1)
    rnd = np.random.default_rng(1).standard_normal(len(p))
    d, _, _ = d1_oracle.ddg_d1(G, process_v(G, -G.bdry_normals, "boundary"), Un, "boundary")
    assert np.array_equal(d, z["h0.1_ddg_d1n"])
    assert np.array_equal(d1_oracle.ddg_d1grad(G, rnd)[0], z["h0.1_ddg_d1grad_random"])
    for k, v in enumerate(([1, 0], 0.3)):
        d, _, _ = d1_oracle.ddi_d1(G, process_v(G, v), rnd)
        ref = z["h0.1_ddi_d1_random_%d" % k]
        assert np.abs(d - ref).max() < 1e-9 * np.abs(ref).max()
    d, _, _ = d1_oracle.ddi_d1(G, process_v(G, -G.bdry_normals, "boundary"), Un, "boundary")
    assert np.abs(d - z["h0.1_ddi_d1n"]).max() < 1e-12
    assert np.abs(d + 1).max() < 5e-2               # tests/test_ddi.py:38-43 (coarser disc)
