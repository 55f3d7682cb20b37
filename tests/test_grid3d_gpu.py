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


# ---- the structured 3-D kernel (stencil_deep3.cuh): bricks staged by 3-D TMA loads -------------
def _dyadic_box(shape, cell=(1 / 8.0, 1 / 8.0, 1 / 8.0), origin=(0.0, -1.0, 0.5)):
    return np.array([[o, o + c * (n + 1)] for n, c, o in zip(shape, cell, origin)]).T


def _fields(G, seed=0):
    rng = np.random.default_rng(seed)
    X, Y, Z = G.points.T
    return (("random", rng.standard_normal(G.num_points)),
            ("saddle", X * Y - Z * X),                          # exact ties among the pairs
            ("ints", rng.integers(-3, 4, G.num_points).astype(float)),   # heavy exact ties
            ("zero", np.zeros(G.num_points)))


@pytest.mark.parametrize("shape,cell,r", [
    ((9, 8, 11), (1 / 8.0,) * 3, 1),            # compile-time table, r = 1
    ((9, 8, 11), (1 / 8.0,) * 3, 2),            # compile-time table, r = 2 (49 pairs)
    ((13, 7, 70), (1 / 16.0,) * 3, 2),          # two bricks along z, partial bricks everywhere
    ((21, 14, 9), (1 / 8.0, 1 / 4.0, 1 / 16.0), 2),   # anisotropic cells: run-time table
    ((6, 9, 7), (1 / 2.0, 1 / 8.0, 1 / 8.0), 1),
])
def test_structured_3d_kernel_bitexact(shape, cell, r):
    """Values, selected pairs (tie rule included: explicit matrices and control vectors) of the
    3-D table kernel against the NumPy restatement of ddg.d2eigs on the grid's pair rows, and
    against the explicit-row kernel every 3-D grid ran through before."""
    from pyellipticfd_b200 import ddg, _context
    from pyellipticfd_b200._context import GridContext
    from pyellipticfd_b200.pointclasses import FDRegularGrid
    G = FDRegularGrid(list(shape), _dyadic_box(shape, cell), r, interpolation=False)
    assert G.uniform_weights()
    ctx = GridContext.get(G)
    assert ctx.nd and ctx.structured and ctx.weights == "exact"
    G2 = FDRegularGrid(list(shape), _dyadic_box(shape, cell), r, interpolation=False)
    _context.ND_STRUCTURED = False
    try:
        assert not GridContext.get(G2).structured
    finally:
        _context.ND_STRUCTURED = True
    for name, u in _fields(G):
        (lmin, Mmin, vmin), (lmax, Mmax, vmax) = ddg.d2eigs(G, u, jacobian=True, control=True)
        (omin, OMmin, ovmin), (omax, OMmax, ovmax) = O.d2eigs(G, u, True, True)
        assert np.array_equal(lmin, omin) and np.array_equal(lmax, omax), name
        assert np.array_equal(vmin, ovmin) and np.array_equal(vmax, ovmax), name
        assert (Mmin.tocsr() != OMmin).nnz == 0 and (Mmax.tocsr() != OMmax).nnz == 0, name
        (gmin, _, gvmin), (gmax, _, gvmax) = ddg.d2eigs(G2, u, control=True)
        assert np.array_equal(lmin, gmin) and np.array_equal(lmax, gmax)
        assert np.array_equal(vmin, gvmin) and np.array_equal(vmax, gvmax)
        assert np.array_equal(ddg.d2min(G, u)[0], omin) and np.array_equal(ddg.d2max(G, u)[0], omax)
        assert np.array_equal(Mmin.dot(u), lmin) and np.array_equal(Mmax.dot(u), lmax)
        w = np.random.default_rng(5).standard_normal(G.num_interior)
        assert np.abs(Mmin.transpose().dot(w) - OMmin.T.dot(w)).max() < 1e-10 * np.abs(w).max() / G.scaling.min() ** 2


def test_structured_3d_solves_equal_explicit_rows():
    """convex_envelope on a 3-D lattice: the Euler iteration through the table kernel is
    bit-identical to the explicit-row path (same values -> same iterates), Newton (policy
    Jacobian rows of deep points from the table) agrees to solver tolerance."""
    from pyellipticfd_b200 import convex_envelope, _context
    from pyellipticfd_b200.pointclasses import FDRegularGrid
    shape = (15, 15, 15)
    mk = lambda: FDRegularGrid(list(shape), np.array([[-1.0, 1.0]] * 3).T, 2, interpolation=False)
    G, G2 = mk(), mk()
    X = G.points
    g = np.minimum(np.linalg.norm(X - np.array([0.4, 0.0, 0.1]), axis=1),
                   np.linalg.norm(X + np.array([0.4, 0.1, 0.0]), axis=1)) - 0.3
    U, diff, it, _ = convex_envelope.solve(G, g, fdmethod="grid", solver="euler")
    Un, diffn, itn, _ = convex_envelope.solve(G, g, fdmethod="grid", solver="newton")
    _context.ND_STRUCTURED = False
    try:
        U2, diff2, it2, _ = convex_envelope.solve(G2, g, fdmethod="grid", solver="euler")
        Un2, _, itn2, _ = convex_envelope.solve(G2, g, fdmethod="grid", solver="newton")
    finally:
        _context.ND_STRUCTURED = True
    assert _context.GridContext.get(G).nd and not _context.GridContext.get(G2).structured
    assert it == it2 and diff == diff2 and np.array_equal(U, U2)
    assert abs(itn - itn2) <= 2 and np.abs(Un - Un2).max() < 2e-6
    assert np.abs(Un - U).max() < 1e-3


def test_structured_3d_129cubed_properties():
    """2.1 M interior points, r = 2: no pair rows are materialised for the deep points; the
    kernel's values against the C oracle formula on a sample of deep points, exact answers for
    a quadratic, and translation invariance."""
    from pyellipticfd_b200 import ddg
    from pyellipticfd_b200._context import GridContext
    from pyellipticfd_b200.pointclasses import FDRegularGrid
    n = 129
    G = FDRegularGrid([n, n, n], np.array([[0.0, (n + 1) / 128.0]] * 3).T, 2, interpolation=False)
    ctx = GridContext.get(G)
    assert ctx.nd and G._pairs is None
    rng = np.random.default_rng(3)
    u = rng.standard_normal(G.num_points)
    lmin, lmax = ddg.d2min(G, u)[0], ddg.d2max(G, u)[0]
    assert G._pairs is None
    # sample of deep points against the reference formula (ddg.py:276-292) row by row
    sp = G.stencil_pairs_nd.astype(np.int64)
    stride = np.array([n * n, n, 1])
    w = ctx.deep_pair_weights()
    ids = rng.choice(G._deep_ids(), 4000, replace=False)
    J0 = ids[:, None] + (sp[:, :3] * stride).sum(1)[None]
    J1 = ids[:, None] + (sp[:, 3:] * stride).sum(1)[None]
    d = w[None, :, 2] * ((u[J0] - u[ids, None]) * w[None, :, 0] + (u[J1] - u[ids, None]) * w[None, :, 1])
    assert np.array_equal(d.min(1), lmin[ids]) and np.array_equal(d.max(1), lmax[ids])
    X, Y, Z = G.points.T
    q = X ** 2 - Y ** 2 + 2 * Z ** 2
    assert np.abs(ddg.d2min(G, q)[0] + 2).max() < 1e-8 and np.abs(ddg.d2max(G, q)[0] - 4).max() < 1e-8


def test_band_points_inside_the_brick_loop_equal_the_band_kernel(monkeypatch):
    """Opt-in (PDE_K3_FUSE_BAND=1): lattices with at least two bricks per CTA can evaluate their
    band points inside the brick loop of the 3-D kernel (stencil_deep3.cuh::band_point_warp)
    instead of a separate band launch: values, policies and the fused residual / Euler epilogues
    are bit-identical either way (the separate launch is faster and stays the default)."""
    import torch
    from pyellipticfd_b200._context import GridContext
    from pyellipticfd_b200.pointclasses import FDRegularGrid
    n = 97
    G = FDRegularGrid([n, n - 8, n + 5], _dyadic_box((n, n - 8, n + 5), (1 / 128.0,) * 3), 2,
                      interpolation=False)
    ctx = GridContext.get(G)
    assert ctx.nd
    rng = np.random.default_rng(11)
    S = ctx.pack(torch.from_numpy(rng.integers(-4, 5, G.num_points).astype(float)).cuda())   # ties
    g = ctx.pack(torch.from_numpy(rng.standard_normal(G.num_points)).cuda())
    res = {}
    for fuse in ("1", "0"):
        monkeypatch.setenv("PDE_K3_FUSE_BAND", fuse)
        lmin, lmax, pmin, pmax = ctx.d2eigs(S, policy=True)
        F, P, red = ctx.ce_residual(S, g, policy=True)
        Un, red2 = ctx.ce_euler_step(S, g, 1e-5)
        torch.cuda.synchronize()
        res[fuse] = [t.clone() for t in (lmin, lmax, ctx.unpad_policy(pmin), ctx.unpad_policy(pmax),
                                         F, ctx.unpad_policy(P), red, Un, red2)]
    for a, b in zip(res["1"], res["0"]):
        assert torch.equal(a, b)
    # and against the reference formula on the band points themselves (explicit rows, host)
    u = ctx.unpack(S).cpu().numpy()
    bp = np.asarray(G.band_pairs)
    P0 = G.coords(bp[:, 0]); h0 = np.linalg.norm(G.coords(bp[:, 1]) - P0, axis=1)
    h1 = np.linalg.norm(G.coords(bp[:, 2]) - P0, axis=1)
    d = 2 / (h0 + h1) * ((u[bp[:, 1]] - u[bp[:, 0]]) * (1 / h0) + (u[bp[:, 2]] - u[bp[:, 0]]) * (1 / h1))
    want = np.minimum.reduceat(d, G.band_ptr[:-1])
    got = ctx.unpad(res["1"][0]).cpu().numpy()[G.band_index]
    assert np.array_equal(got, want)


@pytest.mark.parametrize("shape,r", [((9, 8, 7), 3), ((5, 4, 6, 5), 1)])
def test_lattices_outside_the_structured_3d_kernel_use_explicit_rows(shape, r):
    """3-D lattices with r >= 3 (145 pairs: more than the direction table holds) and 4-D lattices
    keep the explicit-row path (rows materialised lazily from the offset table): bit-exact against
    the NumPy restatement."""
    from pyellipticfd_b200 import ddg
    from pyellipticfd_b200._context import GridContext
    from pyellipticfd_b200.pointclasses import FDRegularGrid
    dim = len(shape)
    G = FDRegularGrid(list(shape), np.array([[0.0, (n + 1) / 16.0] for n in shape]).T, r,
                      interpolation=False)
    ctx = GridContext.get(G)
    assert not ctx.nd and not ctx.structured
    u = np.random.default_rng(2).standard_normal(G.num_points)
    (lmin, M, _), (lmax, _, _) = ddg.d2eigs(G, u, jacobian=True)
    (omin, OM, _), (omax, _, _) = O.d2eigs(G, u, True, False)
    assert np.array_equal(lmin, omin) and np.array_equal(lmax, omax)
    assert (M.tocsr() != OM).nnz == 0
    assert G.points.shape[1] == dim


def test_inexact_3d_lattice_nominal_weights_within_tolerance():
    """[-1,1]^3 with 31 interior points per axis (h = 1/16 exactly would be dyadic: use 29 ->
    h = 2/30, not a power of two): with weights='nominal' the 3-D kernel runs with one weight
    triple per direction; values agree with the NumPy restatement (per-pair weights) within
    c * max(|d_ref|, w0 (|du0| w_0 + |du1| w_1)), c = max(1e-12, 8 eps max|x| / h) -- the 2-D
    statement of test_ddg_gpu.py; band points stay per-pair exact."""
    from pyellipticfd_b200 import ddg
    from pyellipticfd_b200._context import GridContext
    from pyellipticfd_b200.pointclasses import FDRegularGrid
    n = 29
    G = FDRegularGrid([n, n - 3, n + 2], np.array([[-1.0, 1.0]] * 3).T, 2, interpolation=False,
                      weights="nominal")
    assert not G.uniform_weights()
    ctx = GridContext.get(G)
    assert ctx.nd and ctx.weights == "nominal"
    rng = np.random.default_rng(4)
    X, Y, Z = G.points.T
    u = np.sin(2 * X) * np.cos(3 * Y) + Z * X + 0.1 * rng.standard_normal(G.num_points)
    (lmin, _, _), (lmax, _, _) = ddg.d2eigs(G, u)
    (omin, _, _), (omax, _, _) = O.d2eigs(G, u)
    P = G.pairs
    I, J = P[:, 0], P[:, 1:]
    hh = np.linalg.norm(G.points[J] - G.points[I, None], axis=2)
    term_rows = 2 / hh.sum(1) * (np.abs(u[J] - u[I, None]) / hh).sum(1)
    term = np.maximum.reduceat(term_rows, np.flatnonzero(np.r_[True, np.diff(I) != 0]))
    c = max(1e-12, 8 * np.finfo(float).eps * 1.0 / G.scaling.min())
    for mine, ref in ((lmin, omin), (lmax, omax)):
        assert (np.abs(mine - ref) <= c * np.maximum(np.abs(ref), term)).all()
    band = np.zeros(G.num_interior, bool)
    band[G.band_index] = True
    assert np.array_equal(lmin[band], omin[band]) and np.array_equal(lmax[band], omax[band])


def test_band_translation_classes_equal_the_explicit_rows(monkeypatch):
    """3-D lattices with exact weights evaluate their band points from translation-class tables
    (pde_grid_set_band_classes, k_band_cls) instead of 40-byte explicit rows: values, policies,
    residual and Euler epilogues bit-identical to the explicit band kernel."""
    import torch
    from pyellipticfd_b200._context import GridContext
    from pyellipticfd_b200.pointclasses import FDRegularGrid
    shape = (23, 17, 70)
    G = FDRegularGrid(list(shape), _dyadic_box(shape, (1 / 64.0,) * 3), 2, interpolation=False)
    ctx = GridContext.get(G)
    assert ctx.nd and ctx.band_classes > 100
    rng = np.random.default_rng(12)
    S = ctx.pack(torch.from_numpy(rng.integers(-4, 5, G.num_points).astype(float)).cuda())   # ties
    g = ctx.pack(torch.from_numpy(rng.standard_normal(G.num_points)).cuda())
    res = {}
    for on in ("1", "0"):
        monkeypatch.setenv("PDE_BAND_CLASSES", on)
        lmin, lmax, pmin, pmax = ctx.d2eigs(S, policy=True)
        F, P, red = ctx.ce_residual(S, g, policy=True)
        Un, red2 = ctx.ce_euler_step(S, g, 1e-5)
        torch.cuda.synchronize()
        res[on] = [t.clone() for t in (lmin, lmax, ctx.unpad_policy(pmin), ctx.unpad_policy(pmax),
                                       F, ctx.unpad_policy(P), red, Un, red2)]
    for a, b in zip(res["1"], res["0"]):
        assert torch.equal(a, b)
