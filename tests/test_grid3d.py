"""CPU: the fast 3-D FDRegularGrid constructor against fixtures written from
the unmodified reference constructor (oracle/make_golden.py::grid3d_case), against the
brute-force restatement oracle/grid_oracle.py on randomly drawn boxes, and the NumPy operator /
solver restatements evaluated on the fast grid against the reference's outputs."""
import glob
import os

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import ddg_oracle as O, grid_oracle, solvers_oracle as SO
from oracle.make_golden import canon_pairs
from pyellipticfd_b200.pointclasses import FDRegularGrid

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GRID3D = sorted(glob.glob(os.path.join(GOLD, "grid3d_*.npz")))


def grid_of(z):
    return FDRegularGrid([int(s) for s in z["shape"]], z["bounds"], int(z["radius"]),
                         interpolation=False)


def _sorted_rows(rows):
    rows = np.asarray(rows)
    return rows[np.lexsort(tuple(rows[:, k] for k in range(rows.shape[1] - 1, -1, -1)))]


def test_fixtures_present():
    assert len(GRID3D) >= 2


@pytest.mark.parametrize("path", GRID3D, ids=[os.path.basename(p) for p in GRID3D])
def test_constructor_matches_reference(path):
    z = np.load(path)
    G = grid_of(z)
    assert G.dim == 3 and G.num_points == G.num_interior + G.num_bdry
    assert G.num_bdry == int(z["num_bdry"])
    assert np.array_equal(G.points[G.num_interior:], z["bdry_points"])        # bit-equal wall points
    assert np.array_equal(G.bdry_points, z["bdry_points"])
    assert np.array_equal(G.normals, z["normals"])
    assert np.array_equal(canon_pairs(G.pairs), z["pairs"])                   # per-point pair SETS
    assert np.array_equal(_sorted_rows(G.neighbours), z["neighbours"])        # neighbour SETS
    meta = np.array([G.spatial_resolution, G.angular_resolution, G.dist_to_bdry,
                     G.bdry_resolution, G.max_radius, G.min_radius, G.min_interior_nb_dist])
    assert np.array_equal(meta, z["meta"])
    assert G.simplices is None and (np.diff(G.pairs[:, 0]) >= 0).all()


@pytest.mark.parametrize("path", GRID3D, ids=[os.path.basename(p) for p in GRID3D])
def test_oracle_on_fast_grid_matches_reference(path):
    """Extremes do not depend on the order of a point's pairs: the NumPy restatement of
    ddg.d2eigs on the fast grid reproduces the reference's values bit for bit, and the selected
    directions where no exact tie exists (random field)."""
    z = np.load(path)
    G = grid_of(z)
    rng = np.random.default_rng(0)
    X, Y, Z = G.points.T
    for fname, u in (("random", rng.standard_normal(G.num_points)),
                     ("quad", X ** 2 - Y ** 2 + 2 * Z ** 2)):
        (lmin, _, vmin), (lmax, _, vmax) = O.d2eigs(G, u, True, True)
        assert np.array_equal(lmin, z["lmin_" + fname]) and np.array_equal(lmax, z["lmax_" + fname])
        if fname == "random":
            # a pair may be stored as (J0, J1) or (J1, J0): the control vector then points the
            # other way (and carries the other 1/h): same LINE at every point
            for v, ref in ((vmin, z["vmin_random"]), (vmax, z["vmax_random"])):
                c = (v * ref).sum(1) / np.sqrt((v ** 2).sum(1) * (ref ** 2).sum(1))
                assert (np.abs(np.abs(c) - 1) < 1e-3).all()
    if "euler_U" in z.files:
        U, diff, it, _ = SO.ce_solve(G, z["g"], solver="euler")
        assert it == int(z["euler_iters"]) and np.array_equal(U, z["euler_U"])
        U, diff, it, _ = SO.ce_solve(G, z["g"], solver="newton")
        assert it == int(z["newton_iters"]) and np.abs(U - z["newton_U"]).max() < 1e-9


@settings(max_examples=10, deadline=None)
@given(nx=st.integers(2, 6), ny=st.integers(2, 6), nz=st.integers(2, 6), r=st.integers(1, 2),
       sx=st.sampled_from([1.0, 0.5, 0.37]), sz=st.sampled_from([1.0, 0.25, 1.61]),
       x0=st.sampled_from([0.0, -1.0, 0.3]))
def test_native_builder_equals_brute_force_oracle(nx, ny, nz, r, sx, sz, x0):
    shape = [nx, ny, nz]
    bounds = np.array([[x0, x0 + sx * (nx + 1) / 8.0], [0.0, (ny + 1) / 8.0],
                       [-0.2, -0.2 + sz * (nz + 1) / 8.0]]).T
    G = FDRegularGrid(shape, bounds, r, interpolation=False)
    og = grid_oracle.build(shape, bounds, r)
    assert np.array_equal(G.points, og["points"]) and G.num_interior == og["num_interior"]
    assert np.array_equal(canon_pairs(G.pairs), canon_pairs(og["pairs"]))
    assert np.array_equal(_sorted_rows(G.neighbours), _sorted_rows(og["neighbours"]))


def test_2d_grid_through_the_nd_builder():
    """pde_band_build_nd with dim = 2 lists the same rows as the 2-D builder."""
    import ctypes as C
    from pyellipticfd_b200 import _lib
    L = _lib.lib()
    G = FDRegularGrid([9, 12], np.array([[0.0, 1.0], [-1.0, 2.0]]).T, 3, interpolation=False)
    shape = np.array(G.interior_shape, dtype=np.int64)
    W = np.ascontiguousarray(G._wall_unscaled)
    sc, b0 = np.ascontiguousarray(G.scaling), np.ascontiguousarray(G.bounds[0], dtype=np.float64)
    ids = np.ascontiguousarray(G.band_index)
    ptr, pairs, mn = C.c_void_p(), C.c_void_p(), C.c_double()
    _lib.check(L.pde_band_build_nd(2, shape.ctypes.data, 3, sc.ctypes.data, b0.ctypes.data,
                                   W.shape[0], W.ctypes.data, ids.size, ids.ctypes.data, 1e-3,
                                   C.byref(ptr), C.byref(pairs), C.byref(mn), None, None))
    try:
        p = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_int64)), shape=(ids.size + 1,)).copy()
        rows = np.ctypeslib.as_array(C.cast(pairs, C.POINTER(C.c_int64)),
                                     shape=(int(p[-1]), 3)).copy()
    finally:
        L.pde_free(ptr)
        L.pde_free(pairs)
    assert np.array_equal(p, G.band_ptr)
    assert np.array_equal(canon_pairs(rows), canon_pairs(G.band_pairs))


def test_grid_invariants_3d():
    G = FDRegularGrid([9, 8, 7], np.array([[0.0, 1.0], [0.0, 1.0], [0.0, 1.0]]).T, 2,
                      interpolation=False)
    P, N = G.pairs, G.num_interior
    assert np.array_equal(np.unique(P[:, 0]), np.arange(N)) and P.max() < G.num_points
    X0 = G.points[P[:, 1]] - G.points[P[:, 0]]
    X1 = G.points[P[:, 2]] - G.points[P[:, 0]]
    cos = (X0 * X1).sum(1) / np.sqrt((X0 ** 2).sum(1) * (X1 ** 2).sum(1))
    assert (cos < -1 + 1e-3).all()
    # deep points carry the 49 antipodal pairs of the primitive offsets in the 5^3 box
    deep = np.ones(N, bool)
    deep[G.band_index] = False
    counts = np.bincount(P[:, 0], minlength=N)
    assert (counts[deep] == 49).all() and G.stencil_pairs_nd.shape == (49, 6)
    assert np.array_equal(G.stencil_pairs_nd[:, :3], -G.stencil_pairs_nd[:, 3:])
    with pytest.raises(NotImplementedError):
        FDRegularGrid([5, 5, 5], np.array([[0.0, 1.0]] * 3).T, 1, interpolation=True)


def test_nd_helpers():
    G = FDRegularGrid([5, 6, 7], np.array([[0.0, 1.0]] * 3).T, 2, interpolation=False)
    ids = np.array([0, 17, G.num_interior - 1, G.num_interior, G.num_points - 1])
    assert np.array_equal(G.coords(ids), G.points[ids])
    assert G.num_pairs == len(G.pairs) and G.num_deep == 1 * 2 * 3
    assert "3D" in repr(G)
    from pyellipticfd_b200.multigpu import ShardedGrid
    with pytest.raises(NotImplementedError):
        ShardedGrid(None, None, 2, 0, 1, None, grid=G)


def test_4d_box_equals_brute_force_oracle():
    """dim = 4 (3x4x3x3, r = 1; the reference constructor builds the same points and pair sets,
    checked once in this container): native builder == brute-force restatement."""
    shape, b = [3, 4, 3, 3], np.array([[0.0, 1.0]] * 4).T
    G = FDRegularGrid(shape, b, 1, interpolation=False)
    og = grid_oracle.build(shape, b, 1)
    assert G.dim == 4 and np.array_equal(G.points, og["points"])
    assert np.array_equal(canon_pairs(G.pairs), canon_pairs(og["pairs"]))
    assert G.stencil_pairs_nd.shape == (40, 8)


def test_3d_error_behaviour_follows_the_reference():
    """pucci refuses 3-D grids (pucci.py:33-34) before anything touches the device; ddf on a
    pairs grid delegates to ddg (ddf.py:36-37,165-166), so it needs the device like ddg."""
    from pyellipticfd_b200 import pucci, ddi
    G = FDRegularGrid([5, 5, 5], np.array([[0.0, 1.0]] * 3).T, 1, interpolation=False)
    u = np.zeros(G.num_points)
    with pytest.raises(NotImplementedError):
        pucci.operator(G, u, (2, 1), fdmethod="grid")
    with pytest.raises(NotImplementedError):
        pucci.solve(G, np.zeros(G.num_interior), np.zeros(G.num_bdry), fdmethod="grid")
    with pytest.raises(NotImplementedError):
        ddi.d2eigs(G, u)                                   # ddi.py:330-331
    with pytest.raises(NotImplementedError):
        ddi.d1grad(G, u)                                   # ddi.py:157-158


@pytest.mark.parametrize("shape,cell", [((9, 8, 11), 1 / 8.0), ((12, 21, 10), 1 / 32.0)])
def test_band_translation_classes_decode_to_the_explicit_rows(shape, cell):
    """Host logic of pde_grid_set_band_classes (CPU): the class tables built by
    GridContext._band_classes reproduce every explicit band row -- lattice neighbours as
    state-offset deltas from the centre, wall neighbours as positions in the band point's wall
    list -- and the weights, for every band point."""
    from pyellipticfd_b200._context import GridContext, pair_weights
    G = FDRegularGrid(list(shape), np.array([[0.0, (n + 1) * cell] for n in shape]).T, 2,
                      interpolation=False)
    nx, ny, nz = shape
    pitch, rows, N = (nz + 15) // 16 * 16, nx * ny, G.num_interior
    bi, bp, ptr = np.asarray(G.band_index), np.asarray(G.band_pairs), np.asarray(G.band_ptr)
    bw, _ = pair_weights(G.coords(bp[:, 0]), G.coords(bp[:, 1]), G.coords(bp[:, 2]))
    t = GridContext._band_classes(G, bi, bp, bw, ptr, pitch, rows)
    assert t is not None and 100 < t["n_cls"] < 65535
    assert t["cls_id"].shape == (len(bi),) and t["wl_ptr"][-1] == t["wl"].shape[0]

    def soff(ids):
        return (ids // nz) * pitch + ids % nz

    cls = t["cls_id"].astype(np.int64)
    counts = np.diff(ptr)
    assert np.array_equal(counts, np.diff(t["cls_ptr"])[cls])
    pt = np.repeat(np.arange(len(bi)), counts)
    krow = t["cls_ptr"][:-1][cls[pt]] + (np.arange(len(bp)) - ptr[:-1][pt])
    assert np.array_equal(t["cls_w"][krow], bw)
    ctr = soff(bp[:, 0])
    for q in (0, 1):
        ref = t["cls_ref"][krow, q].astype(np.int64)
        J = bp[:, 1 + q]
        wall = (ref & 1) == 1
        assert np.array_equal(wall, J >= N)
        assert np.array_equal(ctr[~wall] + (ref[~wall] >> 1), soff(J[~wall]))
        got = t["wl"][t["wl_ptr"][:-1][pt[wall]] + (ref[wall] >> 1)]
        assert np.array_equal(got, rows * pitch + (J[wall] - N))
    # on a lattice with inexact coordinates the weights are not per-class constants: refused
    Gi = FDRegularGrid([9, 8, 11], np.array([[-1.0, 1.0]] * 3).T, 2, interpolation=False)
    assert not Gi.uniform_weights()
    bpi = np.asarray(Gi.band_pairs)
    bwi, _ = pair_weights(Gi.coords(bpi[:, 0]), Gi.coords(bpi[:, 1]), Gi.coords(bpi[:, 2]))
    assert GridContext._band_classes(Gi, np.asarray(Gi.band_index), bpi, bwi,
                                     np.asarray(Gi.band_ptr), 16, 72) is None
