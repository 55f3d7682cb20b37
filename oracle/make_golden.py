"""Generate the committed golden fixtures from the UNMODIFIED reference.

TEST INFRASTRUCTURE.  Runs only in the build container (needs /root/reference):

    python -m oracle.make_golden            # writes tests/golden/*.npz

For every case it (1) builds the grid with the reference constructor, (2) runs
the reference operators / solvers, (3) checks that the restatements under
``oracle/`` reproduce those outputs (bit-for-bit where the arithmetic is the
same), and (4) stores inputs + reference outputs as small ``.npz`` files.  The
tests then pin the oracle, the fast grid constructor and the CUDA path against
these files on any machine.
"""
import os
import sys
import time
import warnings

import numpy as np

from . import refshim, ddg_oracle, solvers_oracle, grid_oracle

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def canon_pairs(pairs):
    p = np.asarray(pairs).copy()
    lo, hi = np.minimum(p[:, 1], p[:, 2]), np.maximum(p[:, 1], p[:, 2])
    p[:, 1], p[:, 2] = lo, hi
    return p[np.lexsort((p[:, 2], p[:, 1], p[:, 0]))].astype(np.int32)


def fields(G, seed=0):
    rng = np.random.default_rng(seed)
    X, Y = G.points.T
    return {"random": rng.standard_normal(G.num_points),
            "saddle": X ** 2 - Y ** 2,
            "zero": np.zeros(G.num_points),
            "smooth": np.sin(3 * X) * np.cos(2 * Y) + X * Y}


# two-cone obstacle of examples/convex_envelope.py:8-18
def two_cones(X):
    r, n, rot = 3 / 7, 2, np.pi / 5
    C = r * np.array([[np.cos(2 * i * np.pi / n + rot), np.sin(2 * i * np.pi / n + rot)]
                      for i in range(n)])
    x, y = X.T
    v = np.array([np.sqrt((x + c[0]) ** 2 + (y + c[1]) ** 2) for c in C])
    return v.min(axis=0)


def bitwise(a, b):
    return a.shape == b.shape and np.array_equal(a, b)


def ddg_case(R, name, shape, bounds, r):
    t = time.time()
    b = np.array(bounds, dtype=float).T
    G = R.pointclasses.FDRegularGrid(list(shape), b, r, interpolation=False)
    out = dict(shape=np.array(shape), bounds=b, radius=r, num_bdry=G.num_bdry,
               bdry_points=G.points[G.num_interior:], pairs=canon_pairs(G.pairs),
               meta=np.array([G.spatial_resolution, G.angular_resolution, G.dist_to_bdry,
                              G.bdry_resolution, G.max_radius, G.min_radius,
                              G.min_interior_nb_dist]))
    if G.num_interior <= 200:                  # brute-force constructor restatement (2-D too)
        og = grid_oracle.build(shape, b, r)
        assert bitwise(og["points"], G.points) and bitwise(canon_pairs(og["pairs"]), out["pairs"]), name
    for fname, u in fields(G).items():
        (lmin, Mmin, vmin), (lmax, Mmax, vmax) = R.ddg.d2eigs(G, u, jacobian=True, control=True)
        (omin, oMmin, ovmin), (omax, oMmax, ovmax) = ddg_oracle.d2eigs(G, u, True, True)
        assert bitwise(lmin, omin) and bitwise(lmax, omax), (name, fname)
        assert (Mmin != oMmin).nnz == 0 and (Mmax != oMmax).nnz == 0, (name, fname)
        assert bitwise(vmin, ovmin) and bitwise(vmax, ovmax), (name, fname)
        out["lmin_" + fname], out["lmax_" + fname] = lmin, lmax
    u = fields(G)["random"]
    for k, v in enumerate(([1, 0], [0, 1], 0.3)):
        d, M = R.ddg.d2(G, v, u, jacobian=True)
        od, oM = ddg_oracle.d2(G, v, u, jacobian=True)
        assert bitwise(d, od) and (M != oM).nnz == 0
        out["d2_%d" % k] = d
    # convex-envelope and Pucci residuals at a generic iterate
    W, g = fields(G)["smooth"], fields(G)["saddle"]
    F, J = R.convex_envelope.operator(G, W.copy(), g, jacobian=True, fdmethod="grid")
    oF, oJ = solvers_oracle.ce_operator(G, W.copy(), g, True)
    assert bitwise(F, oF) and abs(J - oJ).max() == 0, name
    out["ce_F"] = F
    x = fields(G, 3)["random"]
    out["ce_Jx"] = J.dot(x)
    out["ce_diag"] = J.diagonal()
    val, M = R.pucci.operator(G, W, (2, 1), jacobian=True, fdmethod="grid")
    out["pucci_val"] = val
    out["pucci_Mx"] = M.dot(x)
    np.savez_compressed(os.path.join(OUT, "ddg_%s.npz" % name), **out)
    print("ddg_%s: %d pairs, %.1fs" % (name, len(G.pairs), time.time() - t), flush=True)
    return G


def solve_case(R, name, shape, bounds, r, newton=True, euler=True, pucci=False):
    t = time.time()
    b = np.array(bounds, dtype=float).T
    G = R.pointclasses.FDRegularGrid(list(shape), b, r, interpolation=False)
    g = two_cones(G.points)
    out = dict(shape=np.array(shape), bounds=b, radius=r, g=g)
    if euler:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            U, diff, it, _ = R.convex_envelope.solve(G, g, fdmethod="grid", solver="euler")
        oU, odiff, oit, _ = solvers_oracle.ce_solve(G, g, solver="euler")
        assert bitwise(U, oU) and it == oit and diff == odiff, (name, it, oit)
        out.update(euler_U=U, euler_diff=diff, euler_iters=it)
        print("  euler: %d its" % it, flush=True)
    if newton:
        U, diff, it, _ = R.convex_envelope.solve(G, g, fdmethod="grid", solver="newton")
        st = {}
        oU, odiff, oit, _ = solvers_oracle.ce_solve(G, g, solver="newton", stats=st)
        sU, sdiff, sit, _ = solvers_oracle.ce_solve(G, g, solver="newton", use_scipy=True)
        # the restated BiCGStab against the installed SciPy, and both against the reference
        assert it == sit and np.abs(U - sU).max() < 1e-12, (name, it, sit)
        assert it == oit and np.abs(U - oU).max() < 1e-9, (name, it, oit, np.abs(U - oU).max())
        out.update(newton_U=U, newton_diff=diff, newton_iters=it,
                   newton_krylov=st["krylov_iters"])
        print("  newton: %d its, %d krylov, |ref-oracle| %.2e" %
              (it, st["krylov_iters"], np.abs(U - oU).max()), flush=True)
    if pucci:
        X, Y = G.points.T
        Uex = X ** 2 + 2 * Y ** 2
        N = G.num_interior
        f = np.full(N, 2 * 4.0 + 1 * 2.0)       # A=(2,1): 2*lambda_max(=4) + 1*lambda_min(=2)
        U, diff, it, _ = R.pucci.solve(G, f, Uex[N:], A=(2, 1), fdmethod="grid", solver="newton")
        oU, odiff, oit, _ = solvers_oracle.pucci_solve(G, f, Uex[N:], A=(2, 1))
        assert it == oit and np.abs(U - oU).max() < 1e-9, (it, oit)
        out.update(pucci_U=U, pucci_iters=it, pucci_f=f, pucci_g=Uex[N:])
        print("  pucci: %d its, err %.2e" % (it, np.abs(U - Uex).max()), flush=True)
    np.savez_compressed(os.path.join(OUT, "solve_%s.npz" % name), **out)
    print("solve_%s: %.1fs" % (name, time.time() - t), flush=True)


def cloud_case(R, name, h0):
    """Point-cloud operators on a disc built by the reference FDTriMesh constructor."""
    from scipy.spatial import Delaunay
    from . import ddi_oracle, disc
    import types
    t = time.time()
    theta = np.pi * h0 ** (1 / 3)                       # tests/setup_discs.py:45
    p, nint = disc.disc_points(h0, theta)
    tri = Delaunay(p).simplices
    G = R.pointclasses.FDTriMesh(p, tri, num_interior=nint, angular_resolution=theta,
                                 bdry_normals=p[nint:])
    out = dict(points=G.points, simplices=G.simplices.astype(np.int32), num_interior=nint,
               spatial_resolution=G.spatial_resolution, angular_resolution=G.angular_resolution)
    X, Y = G.points.T
    flds = {"saddle": X ** 2 - Y ** 2, "random": np.random.default_rng(0).standard_normal(len(X)),
            "smooth": np.sin(3 * X) * np.cos(2 * Y) + X * Y}
    for mod, froese, tag in ((R.ddi, False, "ddi"), (R.ddf, True, "ddf")):
        G._d2cache = None
        for fname, u in flds.items():
            (lmin, Mmin, vmin), (lmax, Mmax, vmax) = mod.d2eigs(G, u, jacobian=True, control=True)
            out["%s_lmin_%s" % (tag, fname)], out["%s_lmax_%s" % (tag, fname)] = lmin, lmax
            omin, omax, _ = ddi_oracle.d2eigs(G, u, froese)
            scale = 1e-12 * np.abs(lmin).max() + 1e-9
            assert np.abs(omin - lmin).max() < scale and np.abs(omax - lmax).max() < scale, \
                (tag, fname, np.abs(omin - lmin).max())
        assert np.abs(Mmin.dot(u) - lmin).max() < 1e-9
        V, fdms = G._d2cache[1], G._d2cache[2]
        oV, oS, oW = ddi_oracle.d2eigs_table(G, froese)
        assert np.array_equal(V, oV)
        for k in (0, len(fdms) // 3, len(fdms) - 1):
            assert abs(fdms[k] - ddi_oracle.matrix(G, oS[k], oW[k])).max() < 1e-9 * abs(fdms[k]).max()
        u = flds["random"]
        for k, v in enumerate(([1, 0], 0.3)):
            d, M = mod.d2(G, v, u, jacobian=True)
            od, oM = ddi_oracle.d2(G, v, u, jacobian=True, froese=froese)
            assert np.abs(d - od).max() < 1e-9 * np.abs(d).max(), (tag, np.abs(d - od).max())
            out["%s_d2_%d" % (tag, k)] = d
        out["%s_nds" % tag] = len(fdms)
    # the example's solve (examples/convex_envelope.py:38-42): two-cone obstacle, Newton,
    # fdmethod='interpolate', solution_tol=1e-10 -- and the same with Froese's scheme
    out["min_interior_nb_dist"] = G.min_interior_nb_dist
    out["min_radius"] = G.min_radius
    g = two_cones(G.points)
    for tag, fd in (("ddi", "interpolate"), ("ddf", "froese")):
        G._d2cache = None
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            U, diff, it, _ = R.convex_envelope.solve(G, g, solver="newton", fdmethod=fd,
                                                     solution_tol=1e-10)
            F, J = R.convex_envelope.operator(G, g - 0.01 * (1 - X ** 2 - Y ** 2), g,
                                              jacobian=True, fdmethod=fd)
        out["%s_ce_U" % tag], out["%s_ce_iters" % tag] = U, it
        out["%s_ce_F" % tag] = F
        out["%s_ce_Jx" % tag] = J.dot(flds["random"])
        print("  %s convex envelope: %d Newton its, diff %.2e" % (tag, it, diff), flush=True)
    np.savez_compressed(os.path.join(OUT, "cloud_%s.npz" % name), **out)
    print("cloud_%s: %d points (%d interior), %d simplices, %.1fs" %
          (name, len(p), nint, len(G.simplices), time.time() - t), flush=True)


def _rows5(M, N):
    """CSR matrix with <= 5 entries per row -> (cols (N,4) int32, vals (N,4), diag (N,)): the
    off-diagonal entries in CSR order, padded with (row, 0)."""
    M = M.tocsr()
    M.sum_duplicates()
    cols = np.repeat(np.arange(N, dtype=np.int32)[:, None], 4, 1)
    vals = np.zeros((N, 4))
    diag = np.zeros(N)
    for i in range(N):
        k = 0
        for j, v in zip(M.indices[M.indptr[i]:M.indptr[i + 1]], M.data[M.indptr[i]:M.indptr[i + 1]]):
            if j == i:
                diag[i] = v
            else:
                cols[i, k], vals[i, k] = j, v
                k += 1
    return cols, vals, diag


def ellpin_case(R, name, h0, via):
    """Weight-level pin of the ELL rows (SURVEY 8d parity gate): the unmodified reference's
    first-call path ``ddi.d2`` / ``ddf.d2`` (ddi.py:223-301, ddf.py:8-131) for a few of the
    directions of ``d2eigs`` on a disc -- the finite-difference matrix entries and the values.
    via='reference': the stencils come from the reference constructor (the reference tests' own
    h = 0.03 disc, tests/test_ddi.py:11, setup_discs.py:41-46); via='fast': from the product's
    FDTriMesh host build (a larger cloud the reference constructor cannot do),
    which the test rebuilds from the stored points and triangulation."""
    from scipy.spatial import Delaunay
    from . import disc
    t = time.time()
    theta = np.pi * h0 ** (1 / 3)
    p, nint = disc.disc_points(h0, theta)
    tri = Delaunay(p).simplices
    if via == "reference":
        G = R.pointclasses.FDTriMesh(p, tri, num_interior=nint, angular_resolution=theta,
                                     bdry_normals=p[nint:])
        out = dict(simplices=G.simplices.astype(np.int32))
    else:
        from pyellipticfd_b200.pointclasses import FDTriMesh
        G = FDTriMesh(p, tri, num_interior=nint, angular_resolution=theta, device="cpu")
        out = dict(triangulation=tri.astype(np.int32))
    out.update(points=G.points, num_interior=nint, via=via,
               spatial_resolution=G.spatial_resolution, angular_resolution=G.angular_resolution)
    u = np.random.default_rng(0).standard_normal(len(p))
    # the LAPACK solve of this NumPy build, restated operation by operation (csrc/ell.cu::solve2):
    # checked bit for bit against np.linalg.solve on this cloud's simplices
    S = np.asarray(G.simplices)
    S = S[S[:, 0] < nint]
    a, c = (G.points[S[:, 1]] - G.points[S[:, 0]]).T
    b, d = (G.points[S[:, 2]] - G.points[S[:, 0]]).T
    X = np.stack([np.stack([a, b], 1), np.stack([c, d], 1)], 1)
    for tag, mod in (("ddi", R.ddi), ("ddf", R.ddf)):
        nds = int(mod._num_directions(G))
        # every entry is (N, 4) indices + values: one direction on the large cloud, four on the small
        ks = sorted({0, nds // 3, (2 * nds) // 3 + 1, nds - 1}) if via == "reference" else [nds // 3]
        out[tag + "_nds"], out[tag + "_ks"] = nds, np.array(ks)
        th = np.arange(0, 2 * np.pi, 2 * np.pi / nds)           # the directions of d2eigs, ddi.py:342-343
        for k in ks:
            v = np.array([np.cos(th[k]), np.sin(th[k])])
            dd, M = mod.d2(G, v, u, jacobian=True)
            cols, vals, diag = _rows5(M, nint)
            out["%s_d2_%d" % (tag, k)] = dd
            out["%s_cols_%d" % (tag, k)], out["%s_vals_%d" % (tag, k)] = cols, vals
            out["%s_diag_%d" % (tag, k)] = diag
            if tag == "ddi":
                ref = np.linalg.solve(X, np.broadcast_to(v, (len(a), 2))[..., None])[..., 0]
                mine = solve2_lapack_order(a, b, c, d, v[0], v[1])
                assert np.array_equal(ref, mine), "np.linalg.solve is not the pinned operation order"
    np.savez_compressed(os.path.join(OUT, "ellpin_%s.npz" % name), **out)
    print("ellpin_%s: %d points (%d interior), %d simplices, via %s, %.1fs" %
          (name, len(p), nint, len(G.simplices), via, time.time() - t), flush=True)


def solve2_lapack_order(a, b, c, d, v0, v1):
    """csrc/ell.cu::solve2 in NumPy (exact FMA through integer-exact fractions is too slow for
    arrays: the fused operations are done in extended precision and rounded once, which agrees
    with a true FMA unless the 64-bit-mantissa intermediate rounds a tie -- checked by the caller)."""
    from fractions import Fraction
    v0 = np.broadcast_to(v0, a.shape).astype(float)
    v1 = np.broadcast_to(v1, a.shape).astype(float)
    sw = np.abs(c) > np.abs(a)
    a, c = np.where(sw, c, a), np.where(sw, a, c)
    b, d = np.where(sw, d, b), np.where(sw, b, d)
    v0, v1 = np.where(sw, v1, v0), np.where(sw, v0, v1)
    l = c * (1.0 / a)
    u22 = d - l * b

    def fma(x, y, z):
        return np.array([float(Fraction(p) * Fraction(q) + Fraction(r)) for p, q, r in zip(x, y, z)])
    y1 = fma(-l, v0, v1)
    x1 = y1 / u22
    x0 = fma(-b, x1, v0) / a
    return np.stack([x0, x1], 1)


def helpers_case(R):
    """The public module-level helpers of pointclasses.py (:12-121) and stencils.stern_brocot
    (stencils.py:6-23): inputs (a neighbour list) and the reference's outputs."""
    from scipy.spatial import Delaunay
    from . import disc
    h0 = 0.12
    th = np.pi * h0 ** (1 / 3)
    p, nint = disc.disc_points(h0, th)
    Gr = R.pointclasses.FDTriMesh(p, Delaunay(p).simplices, num_interior=nint, angular_resolution=th)
    D = np.linalg.norm(p[:, None] - p[None], axis=2)
    i, j = np.nonzero((D <= 0.8 * Gr.max_radius) & (D > 0))
    allnb = np.stack([i, j], 1)
    out = dict(points=p, num_interior=nint, angular_resolution=th, neighbours_in=allnb.astype(np.int32),
               spatial_resolution=Gr.spatial_resolution, bdry_resolution=Gr.bdry_resolution,
               dist_to_bdry=Gr.dist_to_bdry, min_dist_to_bdry=Gr.min_dist_to_bdry,
               maximum_bdry_res=Gr.maximum_bdry_res)
    O = R.pointclasses.FDPointCloud(p, angular_resolution=th, num_interior=nint,
                                    spatial_resolution=Gr.spatial_resolution)
    for prefer in ("min", "max"):
        O.neighbours = allnb
        R.pointclasses.remove_colinear_neighbours(O, 1 - np.cos(th / 8), prefer=prefer)
        out["neighbours_" + prefer] = O.neighbours.astype(np.int32)
    R.pointclasses.compute_simplices(O)                       # on the prefer="max" list
    out["simplices_max"] = O.simplices.astype(np.int32)
    out["adjacency_nnz"] = O.adjacency.nnz
    from pyellipticfd_b200.pointclasses import FDRegularGrid
    G = FDRegularGrid([7, 8], np.array([[0, 0], [1.0, 1.3]]), 2, interpolation=False)
    nb = np.asarray(G.neighbours)
    nb = nb[np.lexsort((nb[:, 1], nb[:, 0]))]
    Q = R.pointclasses.FDPointCloud(G.points, num_interior=G.num_interior)
    Q.neighbours = nb
    R.pointclasses.compute_pairs(Q, 1e-3)
    out.update(grid_points=G.points, grid_num_interior=G.num_interior, grid_neighbours=nb.astype(np.int32),
               grid_pairs=Q.pairs.astype(np.int32))
    for N in (1, 2, 4):
        out["stern_brocot_%d" % N] = R.stencils.stern_brocot(N)
    np.savez_compressed(os.path.join(OUT, "helpers.npz"), **out)
    print("helpers: %d points, %d candidate neighbours" % (len(p), len(allnb)), flush=True)


def trimesh_case(R, name, h0):
    """Stencils of the reference FDTriMesh constructor (pointclasses.py:289-398) on a disc:
    the inputs (points, triangulation) and every derived attribute the fast constructor
    must reproduce -- neighbour and simplex SETS (row order inside a point is arbitrary in
    the reference: sparse products / joggled Qhull)."""
    from scipy.spatial import Delaunay
    from . import disc
    t = time.time()
    theta = np.pi * h0 ** (1 / 3)
    p, nint = disc.disc_points(h0, theta)
    tri = Delaunay(p).simplices
    G = R.pointclasses.FDTriMesh(p, tri, num_interior=nint, angular_resolution=theta,
                                 bdry_normals=p[nint:])
    nb = np.asarray(G.neighbours)
    nb = nb[np.lexsort((nb[:, 1], nb[:, 0]))]
    S = np.asarray(G.simplices)
    S = np.stack([S[:, 0], S[:, 1:].min(1), S[:, 1:].max(1)], 1)
    S = S[np.lexsort((S[:, 2], S[:, 1], S[:, 0]))]
    from . import trimesh_oracle
    onb, oS = trimesh_oracle.stencils(p, tri, nint, theta, G.spatial_resolution)
    assert np.array_equal(onb, nb) and np.array_equal(oS, S), "trimesh oracle differs from the reference"
    X, Y = p.T
    d2x, _ = R.ddi.d2(G, [1, 0], X ** 2 + Y ** 2)          # tests/test_trimesh.py:19-25
    np.savez_compressed(
        os.path.join(OUT, "trimesh_%s.npz" % name), points=p, tri=tri.astype(np.int32),
        num_interior=nint, angular_resolution=theta, neighbours=nb.astype(np.int32),
        simplices=S.astype(np.int32), spatial_resolution=G.spatial_resolution,
        dist_to_bdry=G.dist_to_bdry, bdry_resolution=G.bdry_resolution,
        min_interior_nb_dist=G.min_interior_nb_dist, max_radius=G.max_radius,
        min_radius=G.min_radius, d2x_of_r2=d2x)
    print("trimesh_%s: %d points, %d neighbours, %d simplices, %.1fs"
          % (name, len(p), len(nb), len(S), time.time() - t), flush=True)


def d1_case(R):
    """First-derivative operators of the reference (ddg.py:8-175, ddi.py:12-221): outputs on
    the reference tests' own grids (tests/test_ddg.py:17-56, tests/test_ddi.py:10-43), with
    the NumPy restatement oracle/d1_oracle.py pinned against them."""
    from scipy.spatial import Delaunay
    from . import disc, d1_oracle
    from pyellipticfd_b200._ddutils import process_v
    t = time.time()
    out = {}
    G = R.pointclasses.FDRegularGrid([9, 9], np.array([[0, 1], [0, 1]]).T, 2, interpolation=False)
    nb = np.asarray(G.neighbours)
    nb = nb[nb[:, 0] < G.num_interior]
    out["grid_neighbours"] = nb[np.lexsort((nb[:, 1], nb[:, 0]))].astype(np.int32)
    X, Y = G.points.T
    flds = {"lin": X - Y, "random": np.random.default_rng(0).standard_normal(len(X))}
    for fname, u in flds.items():
        for k, v in enumerate(([1, 0], [0, 1], 0.3)):
            d, M = R.ddg.d1(G, v, u, jacobian=True)
            od, oJ, ow = d1_oracle.ddg_d1(G, process_v(G, v), u)
            assert np.array_equal(d, od) and np.abs(M.dot(u) - d).max() < 1e-12
            out["ddg_d1_%s_%d" % (fname, k)] = d
        d, M, vv = R.ddg.d1grad(G, u, jacobian=True, control=True)
        od, oJ, ow, ov = d1_oracle.ddg_d1grad(G, u)
        assert np.array_equal(d, od) and np.array_equal(vv, ov)
        out["ddg_d1grad_%s" % fname], out["ddg_d1grad_v_%s" % fname] = d, vv
    # disc (the trimesh fixture's points): ddg.d1n, ddi.d1 / d1n / d1grad
    for name, h0 in (("h0.1", 0.1), ("h0.05", 0.05)):
        theta = np.pi * h0 ** (1 / 3)
        p, nint = disc.disc_points(h0, theta)
        tri = Delaunay(p).simplices
        G = R.pointclasses.FDTriMesh(p, tri, num_interior=nint, angular_resolution=theta,
                                     bdry_normals=p[nint:])
        X, Y = p.T
        Un = np.linalg.norm(p, axis=1)
        rnd = np.random.default_rng(1).standard_normal(len(p))
        lin = X - Y
        d, M = R.ddg.d1n(G, u=Un, jacobian=True)
        od, _, _ = d1_oracle.ddg_d1(G, process_v(G, -G.bdry_normals, "boundary"), Un, "boundary")
        assert np.array_equal(d, od)
        out["%s_ddg_d1n" % name] = d
        out["%s_ddg_d1grad_random" % name] = R.ddg.d1grad(G, rnd)[0]
        for k, v in enumerate(([1, 0], 0.3)):
            for fname, u in (("lin", lin), ("random", rnd)):
                d, M = R.ddi.d1(G, v, u, jacobian=True)
                od, _, _ = d1_oracle.ddi_d1(G, process_v(G, v), u)
                assert np.abs(d - od).max() <= 1e-12 * np.abs(d).max(), np.abs(d - od).max()
                assert np.abs(M.dot(u) - d).max() < 1e-9 * np.abs(d).max()
                out["%s_ddi_d1_%s_%d" % (name, fname, k)] = d
        d, M = R.ddi.d1n(G, u=Un, jacobian=True)
        od, _, _ = d1_oracle.ddi_d1(G, process_v(G, -G.bdry_normals, "boundary"), Un, "boundary")
        assert np.abs(d - od).max() <= 1e-12
        out["%s_ddi_d1n" % name] = d
        for fname, u in (("lin", lin), ("random", rnd)):
            G._d1cache = None
            d, M, vv = R.ddi.d1grad(G, u, jacobian=True, control=True)
            od, oix, ov = d1_oracle.ddi_d1grad(G, u, int(R.ddi._num_directions(G)))
            assert np.abs(d - od).max() <= 1e-12 * np.abs(d).max() and np.array_equal(vv, ov)
            assert np.abs(M.dot(u) - d).max() < 1e-9 * np.abs(d).max()
            out["%s_ddi_d1grad_%s" % (name, fname)], out["%s_ddi_d1grad_v_%s" % (name, fname)] = d, vv
    np.savez_compressed(os.path.join(OUT, "d1_ops.npz"), **out)
    print("d1_ops: %d arrays, %.1fs" % (len(out), time.time() - t), flush=True)


def ddg3d_case(R):
    """ddg.d2eigs is dimension agnostic: a 3-D pairs grid built by the
    reference constructor (5^3 interior points, r = 1, 13 pairs per deep point)."""
    b = np.array([[0, 1], [0, 1], [0, 1]], dtype=float).T
    G = R.pointclasses.FDRegularGrid([5, 5, 5], b, 1, interpolation=False)
    out = dict(points=G.points, pairs=G.pairs.astype(np.int32), num_interior=G.num_interior)
    rng = np.random.default_rng(0)
    X, Y, Z = G.points.T
    for fname, u in (("random", rng.standard_normal(G.num_points)), ("quad", X ** 2 - Y ** 2 + 2 * Z ** 2)):
        (lmin, Mmin, vmin), (lmax, Mmax, vmax) = R.ddg.d2eigs(G, u, jacobian=True, control=True)
        (omin, _, ovmin), (omax, _, ovmax) = ddg_oracle.d2eigs(G, u, True, True)
        assert bitwise(lmin, omin) and bitwise(lmax, omax) and bitwise(vmin, ovmin) and bitwise(vmax, ovmax)
        assert np.abs(Mmin.dot(u) - lmin).max() < 1e-10
        out["lmin_" + fname], out["lmax_" + fname] = lmin, lmax
        out["vmin_" + fname], out["vmax_" + fname] = vmin, vmax
    np.savez_compressed(os.path.join(OUT, "ddg3d_n5_r1.npz"), **out)
    print("ddg3d_n5_r1: %d pairs" % len(G.pairs), flush=True)


def cones3d(X):
    """3-D analogue of the example's obstacle: distance to the nearer of two centres."""
    C = 3 / 7 * np.array([[np.cos(np.pi / 5), np.sin(np.pi / 5), 0.3],
                          [-np.cos(np.pi / 5), -np.sin(np.pi / 5), -0.3]])
    return np.min([np.sqrt(((X + c) ** 2).sum(1)) for c in C], axis=0)


def grid3d_case(R, name, shape, bounds, r, solve=True):
    """3-D pairs grid of the REFERENCE constructor: points, metadata, pair and
    neighbour sets (what the fast constructor is compared with), ddg.d2eigs outputs and the
    convex-envelope Euler / Newton solves of a two-centre distance obstacle."""
    t = time.time()
    b = np.array(bounds, dtype=float).T
    G = R.pointclasses.FDRegularGrid(list(shape), b, r, interpolation=False)
    N = G.num_interior
    nbI = G.neighbours[G.neighbours[:, 0] < N]
    out = dict(shape=np.array(shape), bounds=b, radius=r, num_bdry=G.num_bdry,
               bdry_points=G.points[N:], pairs=canon_pairs(G.pairs), normals=G.normals,
               neighbours=nbI[np.lexsort((nbI[:, 1], nbI[:, 0]))].astype(np.int32),
               meta=np.array([G.spatial_resolution, G.angular_resolution, G.dist_to_bdry,
                              G.bdry_resolution, G.max_radius, G.min_radius,
                              G.min_interior_nb_dist]))
    # the brute-force restatement of the constructor against the reference
    og = grid_oracle.build(shape, b, r)
    assert bitwise(og["points"], G.points) and og["num_interior"] == N
    assert bitwise(canon_pairs(og["pairs"]), out["pairs"]), name
    assert bitwise(og["neighbours"][np.lexsort((og["neighbours"][:, 1], og["neighbours"][:, 0]))],
                   out["neighbours"].astype(np.int64)), name
    rng = np.random.default_rng(0)
    X, Y, Z = G.points.T
    for fname, u in (("random", rng.standard_normal(G.num_points)),
                     ("quad", X ** 2 - Y ** 2 + 2 * Z ** 2)):
        (lmin, Mmin, vmin), (lmax, Mmax, vmax) = R.ddg.d2eigs(G, u, jacobian=True, control=True)
        (omin, _, ovmin), (omax, _, ovmax) = ddg_oracle.d2eigs(G, u, True, True)
        assert bitwise(lmin, omin) and bitwise(lmax, omax) and bitwise(vmin, ovmin) and bitwise(vmax, ovmax)
        out["lmin_" + fname], out["lmax_" + fname] = lmin, lmax
        out["vmin_" + fname], out["vmax_" + fname] = vmin, vmax
    if solve:
        g = cones3d(G.points)
        out["g"] = g
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            U, diff, it, _ = R.convex_envelope.solve(G, g, fdmethod="grid", solver="euler")
        oU, odiff, oit, _ = solvers_oracle.ce_solve(G, g, solver="euler")
        assert bitwise(U, oU) and it == oit and diff == odiff, (name, it, oit)
        out.update(euler_U=U, euler_diff=diff, euler_iters=it)
        U, diff, it, _ = R.convex_envelope.solve(G, g, fdmethod="grid", solver="newton")
        oU, odiff, oit, _ = solvers_oracle.ce_solve(G, g, solver="newton")
        assert it == oit and np.abs(U - oU).max() < 1e-9, (name, it, oit, np.abs(U - oU).max())
        out.update(newton_U=U, newton_diff=diff, newton_iters=it)
        print("  3-D solves: euler %d its, newton %d its" % (out["euler_iters"], it), flush=True)
    np.savez_compressed(os.path.join(OUT, "grid3d_%s.npz" % name), **out)
    print("grid3d_%s: %d pairs, %.1fs" % (name, len(G.pairs), time.time() - t), flush=True)


def main():
    os.makedirs(OUT, exist_ok=True)
    R = refshim.modules()
    if "--3d-only" in sys.argv:
        ddg3d_case(R)
        grid3d_case(R, "n4x5x6_r2", (4, 5, 6), [[0, 1], [0, 1], [0, 1]], 2, solve=False)
        grid3d_case(R, "n6_r2", (6, 6, 6), [[-1, 1], [-1, 1], [-1, 1]], 2)
        return
    if "--d1-only" in sys.argv:
        d1_case(R)
        return
    if "--trimesh-only" in sys.argv:
        trimesh_case(R, "disc_h0.1", 0.1)
        trimesh_case(R, "disc_h0.05", 0.05)
        return
    if "--helpers-only" in sys.argv:
        helpers_case(R)
        return
    if "--ellpin-only" in sys.argv:
        ellpin_case(R, "disc_h0.03", 0.03, "reference")
        ellpin_case(R, "disc_h0.014", 0.014, "fast")
        return
    if "--clouds-only" in sys.argv:
        cloud_case(R, "disc_h0.1", 0.1)
        cloud_case(R, "disc_h0.06", 0.06)
        return
    ddg_case(R, "n9_r2", (9, 9), [[0, 1], [0, 1]], 2)            # tests/test_ddg.py fixture
    ddg_case(R, "n19_r4", (19, 19), [[0, 1], [0, 1]], 4)         # tests/test_regular_grid.py fixture
    ddg_case(R, "n11x14_r2", (11, 14), [[-1, 1], [0, 3]], 2)     # anisotropic, non-dyadic
    ddg_case(R, "n15_r5", (15, 15), [[-1, 1], [-1, 1]], 5)
    ddg_case(R, "n13_r3", (13, 13), [[0, 1], [0, 1]], 3)
    solve_case(R, "n31_r2", (31, 31), [[-1, 1], [-1, 1]], 2, pucci=True)
    solve_case(R, "n63_r2", (63, 63), [[-1, 1], [-1, 1]], 2)     # BASELINE config 1a (65^2 grid)
    solve_case(R, "n15_r4", (15, 15), [[-1, 1], [-1, 1]], 4, pucci=True)
    ellpin_case(R, "disc_h0.03", 0.03, "reference")
    ellpin_case(R, "disc_h0.014", 0.014, "fast")
    cloud_case(R, "disc_h0.1", 0.1)
    cloud_case(R, "disc_h0.06", 0.06)
    trimesh_case(R, "disc_h0.1", 0.1)
    trimesh_case(R, "disc_h0.05", 0.05)
    d1_case(R)
    helpers_case(R)
    ddg3d_case(R)
    grid3d_case(R, "n4x5x6_r2", (4, 5, 6), [[0, 1], [0, 1], [0, 1]], 2, solve=False)
    grid3d_case(R, "n6_r2", (6, 6, 6), [[-1, 1], [-1, 1], [-1, 1]], 2)


if __name__ == "__main__":
    main()
