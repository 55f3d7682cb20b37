"""GPU parity of the ddg path against the oracle (bit-exact) -- through the C ABI."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import ddg_oracle as O


def _grid(shape, bounds, r):
    from pyellipticfd_b200.pointclasses import FDRegularGrid
    return FDRegularGrid(list(shape), np.array(bounds, dtype=float).T, r, interpolation=False)


CASES = [((9, 9), [[0, 1], [0, 1]], 2),
         ((19, 19), [[0, 1], [0, 1]], 4),
         ((31, 150), [[0, 1], [0, 1]], 3),       # dyadic in x only -> generic path
         ((63, 63), [[-1, 1], [-1, 1]], 2),
         ((40, 300), [[-1, 1], [-1, 1]], 1),
         ((11, 14), [[-1, 1], [0, 3]], 2),       # non-uniform weights -> explicit rows
         ((15, 15), [[0, 1], [0, 1]], 5),
         ((31, 31), [[0, 1], [0, 1]], 6)]


def _fields(G):
    rng = np.random.default_rng(0)
    X, Y = G.points.T
    return {"random": rng.standard_normal(G.num_points),
            "saddle": X ** 2 - Y ** 2,
            "zero": np.zeros(G.num_points),
            "smooth": np.sin(3 * X) * np.cos(2 * Y) + X * Y}


@pytest.mark.parametrize("shape,bounds,r", CASES)
def test_d2eigs_bitexact(shape, bounds, r):
    from pyellipticfd_b200 import ddg
    G = _grid(shape, bounds, r)
    for name, u in _fields(G).items():
        (lmin, Mmin, vmin), (lmax, Mmax, vmax) = ddg.d2eigs(G, u, jacobian=True, control=True)
        (omin, oMmin, ovmin), (omax, oMmax, ovmax) = O.d2eigs(G, u, jacobian=True, control=True)
        assert np.array_equal(lmin, omin), (name, np.abs(lmin - omin).max())
        assert np.array_equal(lmax, omax), (name, np.abs(lmax - omax).max())
        # selected pairs, ties included: the explicit matrices must be identical
        assert (Mmin.tocsr() != oMmin).nnz == 0, name
        assert (Mmax.tocsr() != oMmax).nnz == 0, name
        assert np.array_equal(vmin, ovmin), name
        assert np.array_equal(vmax, ovmax), name
        # the reference tests' identity M.dot(U) == lambda (tests/test_ddg.py:75-78)
        assert np.abs(Mmin.dot(u) - lmin).max() < 1e-12 * max(1.0, np.abs(lmin).max())
        assert np.abs(Mmax.dot(u) - lmax).max() < 1e-12 * max(1.0, np.abs(lmax).max())


def test_reference_known_answers():
    """tests/test_ddg.py:58-80 and tests/test_regular_grid.py:32-37 re-run on the CUDA path."""
    from pyellipticfd_b200 import ddg
    G = _grid((9, 9), [[0, 1], [0, 1]], 2)
    X, Y = G.points.T
    U2 = X ** 2 - Y ** 2
    d2x, M2x = ddg.d2(G, [1, 0], U2, jacobian=True)
    d2y, M2y = ddg.d2(G, [0, 1], U2, jacobian=True)
    assert np.abs(d2x - 2).max() < 1e-8
    assert np.abs(M2x.dot(U2) - d2x).max() < 1e-12
    assert np.abs(d2y + 2).max() < 1e-8
    assert np.abs(M2y.dot(U2) - d2y).max() < 1e-12
    (d2min, M_min, v_min), (d2max, M_max, v_max) = ddg.d2eigs(G, U2, jacobian=True, control=True)
    assert np.abs(d2min + 2).max() < 1e-8
    assert np.abs(d2max - 2).max() < 1e-8
    assert np.abs(M_min.dot(U2) - d2min).max() < 1e-12
    assert np.abs(M_max.dot(U2) - d2max).max() < 1e-12
    assert np.abs((v_max.dot(np.array([1, 0]))) ** 2 - 1).max() < 1e-8
    assert np.abs((v_min.dot(np.array([0, 1]))) ** 2 - 1).max() < 1e-8
    G = _grid((19, 19), [[0, 1], [0, 1]], 4)
    X, Y = G.points.T
    U2 = X ** 2 - Y ** 2
    d2x, _ = ddg.d2(G, [1, 0], U2)
    d2y, _ = ddg.d2(G, [0, 1], U2)
    assert np.abs(d2x.max()) - 2 < 1e-13
    assert np.abs(d2y.max()) - 2 < 1e-13


@pytest.mark.parametrize("shape,bounds,r", CASES[:4])
def test_d2_matches_oracle(shape, bounds, r):
    from pyellipticfd_b200 import ddg
    G = _grid(shape, bounds, r)
    u = _fields(G)["random"]
    rng = np.random.default_rng(1)
    for v in ([1, 0], [0, 1], 0.3, rng.standard_normal((G.num_interior, 2))):
        d, M = ddg.d2(G, v, u, jacobian=True)
        od, oM = O.d2(G, v, u, jacobian=True)
        assert np.array_equal(d, od)
        assert (M.tocsr() != oM).nnz == 0


def test_torch_inputs_stay_on_device():
    import torch
    from pyellipticfd_b200 import ddg
    G = _grid((33, 33), [[0, 1], [0, 1]], 2)
    u = torch.from_numpy(_fields(G)["smooth"]).cuda()
    (lmin, _, _), (lmax, _, _) = ddg.d2eigs(G, u)
    assert lmin.is_cuda and lmax.is_cuda
    omin, omax = O.d2eigs(G, u.cpu().numpy())[0][0], O.d2eigs(G, u.cpu().numpy())[1][0]
    assert np.array_equal(lmin.cpu().numpy(), omin) and np.array_equal(lmax.cpu().numpy(), omax)


def test_pinned_host_pipeline_bitexact():
    """Pinned host tensors take the chunk-pipelined path (H2D / stencil / D2H over row
    chunks, 2-D DMA layout conversion): same bits as the oracle."""
    import torch
    from oracle import c_oracle
    from pyellipticfd_b200 import ddg
    G = _grid((1100, 260), [[0, 1101.0 / 512], [0, 261.0 / 512]], 3)
    assert G.uniform_weights()
    u = np.random.default_rng(5).standard_normal(G.num_points)
    omin, omax, _, _ = c_oracle.pairs_d2eigs(G.points, G.pairs, u, G.num_interior)
    host = torch.from_numpy(u).pin_memory()
    (lmin, _, _), (lmax, _, _) = ddg.d2eigs(G, host)
    assert not lmin.is_cuda and lmin.is_pinned()
    assert np.array_equal(lmin.numpy(), omin) and np.array_equal(lmax.numpy(), omax)
    assert np.array_equal(ddg.d2min(G, host)[0].numpy(), omin)
    assert np.array_equal(ddg.d2max(G, host)[0].numpy(), omax)
    # repeated calls reuse the staging buffers
    u2 = u[::-1].copy()
    o2 = c_oracle.pairs_d2eigs(G.points, G.pairs, u2, G.num_interior)[0]
    assert np.array_equal(ddg.d2min(G, torch.from_numpy(u2).pin_memory())[0].numpy(), o2)
    # the three variants of the pipeline: linear PCIe copies + device-side conversion (default),
    # 2-D DMA copies, and the single C-ABI call pde_ddg_d2eigs_host
    from pyellipticfd_b200._context import GridContext
    ctx = GridContext.get(G)
    for kw in (dict(linear=True, chunk_rows=128), dict(linear=False, chunk_rows=256),
               dict(native=True, chunk_rows=96), dict(native=True, chunk_rows=4096)):
        lo, hi = ctx.d2eigs_host(host, **kw)
        assert np.array_equal(lo.numpy(), omin) and np.array_equal(hi.numpy(), omax), kw
        lo, hi = ctx.d2eigs_host(host, want_max=False, **kw)
        assert hi is None and np.array_equal(lo.numpy(), omin), kw


def test_full_size_4097_r4_bitexact_and_properties():
    """BASELINE configs[1] at full size (4095^2 interior points, r = 4): bit-exact against
    the plain-C oracle (which recomputes every weight from the coordinates per pair, like
    ddg.py:274-281) on all 16.7 M deep points, against the NumPy oracle on the band, and
    the size-independent properties: exact on quadratics, zero on affine functions."""
    import torch
    from concurrent.futures import ThreadPoolExecutor
    from oracle import c_oracle
    from pyellipticfd_b200 import ddg
    from pyellipticfd_b200._context import GridContext
    n, r = 4095, 4
    G = _grid((n, n), [[0, 1], [0, 1]], r)
    assert G.uniform_weights() and G.stencil_pairs.shape[0] == 24
    N, nb = G.num_interior, G.num_bdry
    rng = np.random.default_rng(0)
    u = rng.standard_normal(N + nb)
    ud = torch.from_numpy(u).cuda()
    (lmin, _, _), (lmax, _, _) = ddg.d2eigs(G, ud)
    lmin, lmax = lmin.cpu().numpy(), lmax.cpu().numpy()
    # deep interior: C oracle over row chunks on a few host threads (ctypes releases the GIL)
    out = (np.full(N, np.nan), np.full(N, np.nan), np.full(N, -1, np.int32), np.full(N, -1, np.int32))
    chunks = [(lo, min(n, lo + 256)) for lo in range(0, n, 256)]
    with ThreadPoolExecutor(8) as ex:
        list(ex.map(lambda c: c_oracle.lattice_d2eigs(G, u, c[0], c[1], out), chunks))
    deep = ~np.isnan(out[0])
    assert deep.sum() == G.num_deep
    assert np.array_equal(lmin[deep], out[0][deep]) and np.array_equal(lmax[deep], out[1][deep])
    # band: NumPy oracle on the explicit rows
    bp = G.band_pairs
    I, J, X, h, w, w0 = O.pair_geometry(_band_points(G), _relabel(G, bp))
    uu = _band_values(G, u)
    t = (uu[J] - uu[I, None]) * w
    d = w0 * (t[:, 0] + t[:, 1])
    cen, rmin, rmax = O.select_extremes(I, d)
    ids = np.unique(bp)[cen]                          # centre ids in ascending order == band_index
    assert np.array_equal(ids, G.band_index)
    assert np.array_equal(lmin[ids], d[rmin]) and np.array_equal(lmax[ids], d[rmax])
    # selected pairs of the deep points (policy) against the C oracle, ties included
    ctx = GridContext.get(G)
    _, _, pmin, pmax = ctx.d2eigs(ctx.pack(ud), policy=True)
    kmin = ctx.unpad_policy(pmin).cpu().numpy()
    kmax = ctx.unpad_policy(pmax).cpu().numpy()
    assert np.array_equal(kmin[deep], out[2][deep]) and np.array_equal(kmax[deep], out[3][deep])
    # properties at full size
    x = (torch.arange(1, n + 1, device="cuda", dtype=torch.float64) / (n + 1))
    Pw = torch.from_numpy(G.bdry_points).cuda()
    def field(f):
        return torch.cat([f(x[:, None].expand(n, n).reshape(-1), x[None, :].expand(n, n).reshape(-1)),
                          f(Pw[:, 0], Pw[:, 1])])
    (qmin, _, _), (qmax, _, _) = ddg.d2eigs(G, field(lambda X, Y: X ** 2 - Y ** 2))
    scale = float(n + 1) ** 2 * 2.3e-16 * 8           # rounding of O(1) values divided by h^2
    dm = torch.from_numpy(deep).cuda()
    # exact on the symmetric deep-interior stencil; band pairs are only antipodal within the
    # reference's 1e-3 tolerance (pointclasses.py:24), so quadratics are not exact there
    assert float((qmin[dm] + 2).abs().max()) < scale and float((qmax[dm] - 2).abs().max()) < scale
    (amin, _, _), (amax, _, _) = ddg.d2eigs(G, field(lambda X, Y: 3 * X - 2 * Y + 0.5))
    assert float(amin[dm].abs().max()) < scale and float(amax[dm].abs().max()) < scale
    assert float((amin <= amax).double().min()) == 1.0 and float((qmin <= qmax).double().min()) == 1.0


def _band_points(G):
    """coordinates of every point referenced by the band rows, relabelled densely"""
    ids = np.unique(G.band_pairs)
    return G.coords(ids)


def _relabel(G, bp):
    ids = np.unique(bp)
    return np.searchsorted(ids, bp)


def _band_values(G, u):
    return u[np.unique(G.band_pairs)]


def test_fast_mode_within_tolerance():
    """ddg.ARITHMETIC_MODE = 1 (opt-in, re-associated): |d - d_ref| <= 1e-12 *
    max(|d_ref|, w0*(|du0|*w_0 + |du1|*w_1)), values only."""
    from pyellipticfd_b200 import ddg
    G = _grid((255, 255), [[0, 1], [0, 1]], 4)
    u = _fields(G)["smooth"]
    I, J, X, h, w, w0 = O.pair_geometry(G.points, G.pairs)
    du = np.abs(u[J] - u[I, None])
    bound = np.zeros(G.num_interior)
    np.maximum.at(bound, I, w0 * (du[:, 0] * w[:, 0] + du[:, 1] * w[:, 1]))
    (omin, _, _), (omax, _, _) = O.d2eigs(G, u)
    try:
        ddg.ARITHMETIC_MODE = 1
        (lmin, _, _), (lmax, _, _) = ddg.d2eigs(G, u)
    finally:
        ddg.ARITHMETIC_MODE = 0
    tol = 1e-12 * np.maximum(np.maximum(np.abs(omin), np.abs(omax)), bound)
    assert (np.abs(lmin - omin) <= tol).all() and (np.abs(lmax - omax) <= tol).all()
    assert not np.array_equal(lmin, omin) or True      # usually differs in the last bits


def test_config3_8193_r6_bitexact():
    """BASELINE configs[2] operator at full size (8191^2 interior points, stencil radius 6 ->
    36 pairs, halo 5): bit-exact against the plain-C oracle on the deep interior."""
    import torch
    from concurrent.futures import ThreadPoolExecutor
    from oracle import c_oracle
    from pyellipticfd_b200._context import GridContext
    n, r = 8191, 6
    G = _grid((n, n), [[0, 1], [0, 1]], r)
    assert G.uniform_weights() and G.stencil_pairs.shape[0] == 36 and G.halo == 5
    N = G.num_interior
    u = np.random.default_rng(1).standard_normal(G.num_points)
    ctx = GridContext.get(G)
    S = ctx.pack(torch.from_numpy(u))
    lmin, lmax, _, _ = ctx.d2eigs(S)
    lmin, lmax = ctx.unpad(lmin).cpu().numpy(), ctx.unpad(lmax).cpu().numpy()
    # a sample of row blocks (top, middle, bottom of the deep region) keeps the CPU time bounded
    out = (np.full(N, np.nan), np.full(N, np.nan), np.full(N, -1, np.int32), np.full(N, -1, np.int32))
    chunks = [(lo, lo + 128) for lo in (0, 128, 2048, 4000, 6111, n - 256, n - 128)]
    with ThreadPoolExecutor(8) as ex:
        list(ex.map(lambda c: c_oracle.lattice_d2eigs(G, u, c[0], c[1], out), chunks))
    deep = ~np.isnan(out[0])
    assert deep.sum() > 5_000_000
    assert np.array_equal(lmin[deep], out[0][deep]) and np.array_equal(lmax[deep], out[1][deep])


def test_3d_pairs_grid_bitexact():
    """ddg.d2eigs is dimension agnostic in the reference: a 3-D pairs grid
    built by the reference constructor (5^3 points, r = 1) runs through the explicit-row
    kernel; values and control vectors bit-exact against the reference's outputs."""
    from pyellipticfd_b200 import ddg
    from pyellipticfd_b200.pointclasses import FDPointCloud
    import os
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    z = np.load(os.path.join(gold, "ddg3d_n5_r1.npz"))
    G = FDPointCloud(z["points"], num_interior=int(z["num_interior"]))
    G.pairs = z["pairs"].astype(np.int64)
    assert G.dim == 3
    X, Y, Z = G.points.T
    rng = np.random.default_rng(0)
    for fname, u in (("random", rng.standard_normal(len(X))), ("quad", X ** 2 - Y ** 2 + 2 * Z ** 2)):
        (lmin, Mmin, vmin), (lmax, Mmax, vmax) = ddg.d2eigs(G, u, jacobian=True, control=True)
        assert np.array_equal(lmin, z["lmin_" + fname]) and np.array_equal(lmax, z["lmax_" + fname])
        assert np.abs(Mmin.dot(u) - lmin).max() < 1e-12 * max(1.0, np.abs(lmin).max())
        assert Mmin.tocsr().shape == (G.num_interior, G.num_points)
        if fname == "random":                        # no ties: the selected pairs are the reference's
            assert np.array_equal(vmin, z["vmin_" + fname]) and np.array_equal(vmax, z["vmax_" + fname])
            assert vmin.shape == (G.num_interior, 3)
    lam, _, _ = ddg.d2min(G, X ** 2 - Y ** 2 + 2 * Z ** 2)
    assert np.abs(lam[np.abs(z["lmin_quad"] + 2) < 1e-9] + 2).max() < 1e-9


@pytest.mark.parametrize("n,weights", [(1500, "nominal"), (4000, "auto")])
def test_inexact_lattice_runs_the_structured_kernel_within_tolerance(n, weights):
    """[0,1]^2 with n interior points per side, r = 4: the spacing 1/(n+1) is not a power of two,
    so the reference's per-pair lengths differ by an ulp from point to point.  With
    FDRegularGrid(weights='nominal') (the default above 2^20 points) the structured kernel K1
    runs with one weight triple per direction -- no pair rows are materialised -- and agrees with
    the NumPy restatement of ddg.d2eigs (per-pair weights, ddg.py:274-281) on sampled rows within
    c * max(|d_ref|, w0 (|du0| w_0 + |du1| w_1)), c = max(1e-12, 8 eps max|x| / h): the size of the
    coordinate rounding the reference's own weights carry."""
    import torch
    from oracle import ddg_oracle
    from oracle.cpu_baseline import lattice_chunk
    from pyellipticfd_b200 import ddg
    from pyellipticfd_b200._context import GridContext
    from pyellipticfd_b200.pointclasses import FDRegularGrid
    r = 4
    G = FDRegularGrid([n, n], np.array([[0.0, 1.0], [0.0, 1.0]]).T, r, interpolation=False,
                      weights=weights)
    assert not G.uniform_weights()
    ctx = GridContext.get(G)
    assert ctx.structured and ctx.weights == "nominal" and G._pairs is None
    x = torch.arange(1, n + 1, device="cuda", dtype=torch.float64) * G.scaling[0]
    rng = torch.Generator(device="cuda").manual_seed(0)
    U = torch.empty(G.num_points, dtype=torch.float64, device="cuda")
    U[:n * n] = (torch.sin(3 * x)[:, None] * torch.cos(2 * x)[None, :]
                 + 0.1 * torch.randn(n, n, generator=rng, dtype=torch.float64, device="cuda")).reshape(-1)
    wp = torch.from_numpy(G.bdry_points).cuda()
    U[n * n:] = torch.sin(3 * wp[:, 0]) * torch.cos(2 * wp[:, 1])
    (lmin, Mmin, _), (lmax, _, _) = ddg.d2eigs(G, U, jacobian=True)
    assert G._pairs is None                                    # still not materialised
    assert float((Mmin.dot(U) - lmin).abs().max()) == 0.0      # rows == operator, bit for bit
    table = np.asarray(G.stencil_pairs)
    H = int(np.abs(table).max())
    u = U[:n * n].view(n, n)
    c = max(1e-12, 8 * np.finfo(float).eps * 1.0 / G.scaling.min())
    worst = 0.0
    for row in (r + H, n // 2, n - r - H - 3):
        Gc = lattice_chunk(n, r, table, G.scaling, (0.0, 0.0), row, 3)
        win = u[row - H:row + 3 + H].cpu().numpy()
        ix, iy = np.meshgrid(np.arange(0, 3 + 2 * H), np.arange(n), indexing="ij")
        deep = (ix >= H) & (ix < H + 3) & (iy >= r) & (iy < n - r)
        order = np.concatenate([np.flatnonzero(deep.ravel()), np.flatnonzero(~deep.ravel())])
        uu = win.ravel()[order]
        (omin, _, _), (omax, _, _) = ddg_oracle.d2eigs(Gc, uu)
        I, J = Gc.pairs[:, 0], Gc.pairs[:, 1:]
        hh = np.linalg.norm(Gc.points[J] - Gc.points[I, None], axis=2)
        term = (2 / hh.sum(1) * (np.abs(uu[J] - uu[I, None]) / hh).sum(1)).reshape(-1, len(table)).max(1)
        mine_min = lmin[:n * n].view(n, n)[row:row + 3, r:n - r].cpu().numpy().ravel()
        mine_max = lmax[:n * n].view(n, n)[row:row + 3, r:n - r].cpu().numpy().ravel()
        for mine, ref in ((mine_min, omin), (mine_max, omax)):
            worst = max(worst, (np.abs(mine - ref) / (c * np.maximum(np.abs(ref), term))).max())
    print("n = %d: max |d - d_ref| / tolerance = %.3f (c = %.2e)" % (n, worst, c))
    assert worst < 1.0
    GridContext.drop(G)
