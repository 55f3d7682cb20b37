"""Host-side tables of the convex-hull sweep solver (``pde_ce_sweep``, an extension beyond the
reference: see ``csrc/ce_sweep.cu``) for a structured 2-D pairs grid, in state-layout offsets."""
import numpy as np


def band_slots(nx, ny, r):
    """Number of compact flag slots (csrc/ce_sweep.cu::sw_band_slot)."""
    return nx * ny if nx <= 2 * r else 2 * r * ny + (nx - 2 * r) * 2 * r


def band_slot(nx, ny, r, i, j):
    """Compact index of the band point (i, j), 1-based interior coordinates: the 2r full rows next
    to the i-walls first, then 2r columns of every other row (csrc/ce_sweep.cu::sw_band_slot)."""
    i, j = np.asarray(i, dtype=np.int64), np.asarray(j, dtype=np.int64)
    if nx <= 2 * r:
        return (i - 1) * ny + (j - 1)
    top = (i - 1) * ny + (j - 1)
    bot = (i - (nx - r) - 1 + r) * ny + (j - 1)
    mid = 2 * r * ny + (i - r - 1) * 2 * r + np.where(j <= r, j - 1, j - (ny - r) - 1 + r)
    return np.where(i <= r, top, np.where(i > nx - r, bot, mid))


def tables(G, pitch, rows):
    """-> dict(dirs (D,2) int32, ring int64, flags uint64 (one per band point, ``band_slot``), band_ptr (B+1,) int64,
    band_center (B,) int64, rest_j (R,2) int64, theta (R,) f64): the band pairs that are not
    lattice pairs, CSR over the B band points that own any (state offsets).  ``rows`` x ``pitch``
    is the lattice part of the state layout; wall points follow at offset rows*pitch."""
    nx, ny = G.interior_shape
    N, r = G.num_interior, G.stencil_radius
    dirs = np.ascontiguousarray(G.stencil_pairs[:, :2], dtype=np.int32)
    D = dirs.shape[0]
    if D > 64:
        raise ValueError("the sweep solver handles at most 64 stencil directions")
    W = G._wall_unscaled
    # the integer ring of wall points (every lattice line ends on it)
    # (unscaled wall coordinates carry the rounding of pointclasses.py:538: 17.999999999999996)
    integer = (np.abs(W - np.round(W)) < 1e-9).all(axis=1)
    wi, wid = np.round(W[integer]).astype(np.int64), np.flatnonzero(integer)
    ring = np.full(2 * (ny + 2) + 2 * (nx + 2), -1, dtype=np.int64)
    for mask, base, along in ((wi[:, 0] == 0, 0, 1), (wi[:, 0] == nx + 1, ny + 2, 1),
                              (wi[:, 1] == 0, 2 * (ny + 2), 0),
                              (wi[:, 1] == ny + 1, 2 * (ny + 2) + nx + 2, 0)):
        ring[base + wi[mask, along]] = wid[mask]
    if (ring < 0).any():
        raise ValueError("the grid's wall does not contain the full integer ring")

    def ucoord(ids):
        P = np.empty((ids.size, 2))
        lat = ids < N
        P[lat, 0] = ids[lat] // ny + 1
        P[lat, 1] = ids[lat] % ny + 1
        P[~lat] = W[ids[~lat] - N]
        return P

    def offs(ids):
        return np.where(ids < N, (ids // ny) * pitch + ids % ny, rows * pitch + (ids - N))

    bp = np.asarray(G.band_pairs, dtype=np.int64)
    # rows whose two neighbours are lattice points: integer arithmetic on the ids; only the rows
    # that touch a wall point need the (float) unscaled wall coordinates
    P = bp.shape[0]
    Vi = np.zeros((P, 2), dtype=np.int64)
    exact = np.zeros(P, dtype=bool)
    both = (bp[:, 1] < N) & (bp[:, 2] < N)
    if both.any():
        b = bp[both]
        ci, cj = b[:, 0] // ny, b[:, 0] % ny
        v0 = np.stack([b[:, 1] // ny - ci, b[:, 1] % ny - cj], 1)
        v1 = np.stack([b[:, 2] // ny - ci, b[:, 2] % ny - cj], 1)
        ok = (v1 == -v0).all(axis=1) & (np.abs(v0) <= r).all(axis=1)
        exact[both] = ok
        Vi[both] = np.where(ok[:, None], v0, 0)
    mixed = np.flatnonzero(~both)
    if mixed.size:
        b = bp[mixed]
        PI, P0, P1 = ucoord(b[:, 0]), ucoord(b[:, 1]), ucoord(b[:, 2])
        V0, V1 = P0 - PI, P1 - PI
        ok = ((np.abs(V0 - np.round(V0)) < 1e-9).all(axis=1) & (np.abs(V1 + V0) < 1e-9).all(axis=1)
              & (np.abs(V0) <= r + 1e-9).all(axis=1))
        exact[mixed] = ok
        Vi[mixed] = np.where(ok[:, None], np.round(V0), 0).astype(np.int64)
    lut = np.full((2 * r + 1, 2 * r + 1), -1, dtype=np.int64)
    k = np.arange(D)
    lut[dirs[:, 0] + r, dirs[:, 1] + r] = k
    lut[-dirs[:, 0] + r, -dirs[:, 1] + r] = k
    kk = np.where(exact, lut[Vi[:, 0] + r, Vi[:, 1] + r], -1)
    line = kk >= 0
    flags = np.zeros(band_slots(nx, ny, r), dtype=np.uint64)
    ids = bp[line, 0]
    np.bitwise_or.at(flags, band_slot(nx, ny, r, ids // ny + 1, ids % ny + 1),
                     np.uint64(1) << kk[line].astype(np.uint64))
    rest = bp[~line]
    rest = rest[np.argsort(rest[:, 0], kind="stable")]
    XI = G.coords(rest[:, 0])
    h0 = np.linalg.norm(G.coords(rest[:, 1]) - XI, axis=1)
    h1 = np.linalg.norm(G.coords(rest[:, 2]) - XI, axis=1)
    theta = (1 / h0) / (1 / h0 + 1 / h1)
    centers, counts = np.unique(rest[:, 0], return_counts=True)
    band_ptr = np.zeros(centers.size + 1, dtype=np.int64)
    np.cumsum(counts, out=band_ptr[1:])
    rest_j = np.ascontiguousarray(np.stack([offs(rest[:, 1]), offs(rest[:, 2])], axis=1),
                                  dtype=np.int64)
    return dict(dirs=dirs, ring=ring, flags=flags, band_ptr=band_ptr,
                band_center=np.ascontiguousarray(offs(centers), dtype=np.int64), rest_j=rest_j,
                theta=np.ascontiguousarray(theta, dtype=np.float64))


def flat_tolerance(gmax):
    """Points less than this below a chord count as on it: 8 ulps of max|g| (rounding noise of
    the chord values on planar pieces of the solution)."""
    return 8 * np.finfo(np.float64).eps * float(gmax)


def host_sweep(G, g, tol=1e-11, max_sweeps=400, relax_reps=2, relax_every=0, device=-1,
               flat_tol=None):
    """The HOST builds of ``pde_ce_sweep`` on flat vectors (state layout with pitch = ny):
    ``device=-1`` the serial monotone chain per lattice line under OpenMP (CPU tests, CPU baseline
    of bench.py), ``device=-2`` the thread-by-thread replay of the CUDA kernel's phases (CPU tests).
    Not a product path: ``convex_envelope.solve(..., solver='sweep')`` only runs on the GPU.
    Returns (U, last change, sweeps)."""
    import ctypes as C
    from . import _lib
    nx, ny = G.interior_shape
    t = tables(G, ny, nx)
    U = np.ascontiguousarray(g, dtype=np.float64).copy()       # [lattice | wall] == state layout
    out = (C.c_double * 2)()
    p = lambda a: C.c_void_p(a.ctypes.data)
    _lib.check(_lib.lib().pde_ce_sweep(device, nx, ny, ny, G.num_interior, G.stencil_radius,
                                       t["dirs"].shape[0], p(t["dirs"]), p(t["ring"]), p(t["flags"]),
                                       t["band_center"].shape[0], p(t["band_ptr"]),
                                       p(t["band_center"]), p(t["rest_j"]), p(t["theta"]), p(U),
                                       None, tol,
                                       flat_tolerance(np.abs(U).max()) if flat_tol is None else flat_tol,
                                       max_sweeps, relax_every, relax_reps, out, None))
    return U, out[0], int(out[1])
