"""Fast FDTriMesh constructor against the reference's stencils (host build; no GPU needed).

Fixtures: tests/golden/trimesh_*.npz, written by oracle/make_golden.py from the unmodified
reference constructor (pointclasses.py:289-398).  The row order inside one centre point
is arbitrary in the reference, so neighbours and simplices are compared as sets.
"""
import glob
import os

import numpy as np
import pytest

from pyellipticfd_b200.pointclasses import FDTriMesh

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MESH = sorted(glob.glob(os.path.join(GOLD, "trimesh_*.npz")))


def build(z, device="cpu", **kw):
    n = int(z["num_interior"])
    return FDTriMesh(z["points"], z["tri"].astype(np.int64), num_interior=n,
                     angular_resolution=float(z["angular_resolution"]),
                     bdry_normals=z["points"][n:], device=device, **kw)


def canon_simplices(S):
    S = np.asarray(S)
    S = np.stack([S[:, 0], S[:, 1:].min(1), S[:, 1:].max(1)], 1)
    return S[np.lexsort((S[:, 2], S[:, 1], S[:, 0]))]


@pytest.mark.parametrize("path", MESH, ids=[os.path.basename(p) for p in MESH])
def test_host_build_matches_reference(path):
    z = np.load(path)
    G = build(z)
    for a in ("spatial_resolution", "dist_to_bdry", "bdry_resolution", "min_interior_nb_dist",
              "max_radius", "min_radius"):
        assert getattr(G, a) == pytest.approx(float(z[a]), rel=1e-14), a
    nb = np.asarray(G.neighbours)
    assert nb.shape == z["neighbours"].shape
    assert np.array_equal(nb[np.lexsort((nb[:, 1], nb[:, 0]))], z["neighbours"])
    assert np.array_equal(canon_simplices(G.simplices), z["simplices"])
    # every centre's simplices close the ring once: each neighbour appears in exactly two
    S = np.asarray(G.simplices)
    deg = np.bincount(np.concatenate([S[:, 0] * G.num_points + S[:, 1],
                                      S[:, 0] * G.num_points + S[:, 2]]))
    assert set(np.unique(deg)) <= {0, 2}
    assert G.triangulation.shape == z["tri"].shape and G.edges.shape == (6 * len(z["tri"]), 2)
    assert G.adjacency.shape == (G.num_points, G.num_points)
    assert "FDTriMesh in 2D" in repr(G)


def test_constructor_errors():
    z = np.load(MESH[0])
    n = int(z["num_interior"])
    with pytest.raises(TypeError):                       # pointclasses.py:158-159
        FDTriMesh(z["points"], z["tri"], angular_resolution=0.5)
    with pytest.raises(TypeError):                       # pointclasses.py:358-365
        FDTriMesh(z["points"], z["tri"].astype(np.int64), num_interior=n,
                  angular_resolution=float(z["angular_resolution"]) / 8, device="cpu")
    with pytest.raises(NotImplementedError):
        FDTriMesh(np.zeros((5, 3)), np.zeros((1, 4), dtype=np.int64), num_interior=1, device="cpu")


def test_given_resolutions_are_kept():
    z = np.load(MESH[0])
    G = build(z, spatial_resolution=float(z["spatial_resolution"]),
              dist_to_bdry=float(z["dist_to_bdry"]), bdry_resolution=float(z["bdry_resolution"]))
    assert G.spatial_resolution == float(z["spatial_resolution"])
    assert len(G.neighbours) == len(z["neighbours"])


@pytest.mark.parametrize("path", MESH, ids=[os.path.basename(p) for p in MESH])
def test_oracle_restatement_matches_reference(path):
    """oracle/trimesh_oracle.py (brute-force NumPy/SciPy restatement of pointclasses.py:367-398,
    :37-121) reproduces the reference constructor's neighbour and simplex sets."""
    from oracle import trimesh_oracle
    z = np.load(path)
    nb, S = trimesh_oracle.stencils(z["points"], z["tri"].astype(np.int64), int(z["num_interior"]),
                                    float(z["angular_resolution"]), float(z["spatial_resolution"]))
    assert np.array_equal(nb, z["neighbours"]) and np.array_equal(S, z["simplices"])


def test_native_build_equals_oracle_on_random_clouds():
    """Randomly jittered disc clouds (different seeds, spacings, jitter): the native stencil
    search equals the brute-force oracle -- graph-depth rule, radius mask, colinear pruning with
    the boundary re-admission, hull ring."""
    from scipy.spatial import Delaunay
    from oracle import disc, trimesh_oracle
    for seed, h0, jitter in ((1, 0.12, 0.1), (2, 0.09, 0.25), (3, 0.08, 0.05), (4, 0.15, 0.3)):
        theta = np.pi * h0 ** (1 / 3)
        p, nint = disc.disc_points(h0, theta, seed=seed, jitter=jitter)
        tri = Delaunay(p).simplices
        try:
            G = FDTriMesh(p, tri, num_interior=nint, angular_resolution=theta, device="cpu")
        except TypeError:          # interior points too close to the boundary for this jitter
            continue
        nb, S = trimesh_oracle.stencils(p, tri.astype(np.int64), nint, theta, G.spatial_resolution)
        mine = np.asarray(G.neighbours)
        assert np.array_equal(mine[np.lexsort((mine[:, 1], mine[:, 0]))], nb), (seed, h0)
        assert np.array_equal(canon_simplices(G.simplices), S), (seed, h0)
