"""Device context of a pairs grid: the host side of ``pde_grid`` (C ABI).

A context owns the device copy of the grid description (direction table with
its weights, band pair rows) and offers the raw operations on *state-layout*
tensors that ``ddg`` / ``convex_envelope`` / ``pucci`` / ``solvers`` are written
on.  It is cached on the grid object, the same place the reference caches its
finite-difference matrices (``G._d2cache``, ddi.py:355-356).

PyTorch is used only for device memory and streams; every kernel is reached
through ``libpde_b200.so``.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import pde_jac, check, vp

POLICY_IDENTITY = 0xFFFF
HOST_COPY_LINEAR = True       # see GridContext.d2eigs_host
HOST_CHUNK_ROWS = 128


def _ptr(t):
    return vp(0) if t is None else vp(t.data_ptr())


def _stream():
    return vp(torch.cuda.current_stream().cuda_stream)


def _np_ptr(a):
    return vp(a.ctypes.data)


def require_cuda(device=None):
    if not torch.cuda.is_available():
        raise _lib.PdeError("pyellipticfd_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    if device is None:
        return torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    if device.type != "cuda":
        raise _lib.PdeError("pyellipticfd_b200 computes on CUDA devices only")
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    return device


def to_device(u, device, dtype=torch.float64):
    """numpy / torch (any device) -> contiguous tensor on ``device``."""
    if isinstance(u, torch.Tensor):
        return u.to(device=device, dtype=dtype, non_blocking=True).contiguous()
    a = np.ascontiguousarray(np.asarray(u), dtype={torch.float64: np.float64}[dtype])
    return torch.from_numpy(a).to(device)


def like_input(u, t):
    """Return device tensor ``t`` in the container type of the user's input ``u``:
    CUDA tensor -> stays on the device; host tensor -> pinned host tensor (torch's
    caching host allocator makes the repeated allocation cheap); NumPy -> NumPy."""
    if isinstance(u, torch.Tensor):
        if u.is_cuda:
            return t
        out = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        out.copy_(t, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return out
    return t.cpu().numpy()


def pair_weights(PI, PJ0, PJ1):
    """(w_0, w_1, w0) and X0 with the reference's arithmetic (ddg.py:276-280).  Any
    dimension: the norm is NumPy's ``sqrt(add.reduce(x*x))`` (what ``np.linalg.norm(X, axis=1)``
    evaluates), written out for the 2-D case."""
    X0 = PJ0 - PI
    X1 = PJ1 - PI
    if X0.shape[1] == 2:
        h0 = np.sqrt(X0[:, 0] * X0[:, 0] + X0[:, 1] * X0[:, 1])
        h1 = np.sqrt(X1[:, 0] * X1[:, 0] + X1[:, 1] * X1[:, 1])
    else:
        h0 = np.linalg.norm(X0, axis=1)
        h1 = np.linalg.norm(X1, axis=1)
    w = np.stack([1 / h0, 1 / h1, 2 / (h0 + h1)], axis=1)
    return np.ascontiguousarray(w), np.ascontiguousarray(X0)


# FDRegularGrid(weights="auto"): lattices with inexact coordinates larger than this many interior
# points run the structured kernel with nominal per-direction weights
NOMINAL_ABOVE = 1 << 20
# 3-D lattices run the structured 3-D kernel (False: explicit pair rows for every point, the
# round-1 path; kept for tests that compare the two)
ND_STRUCTURED = True
BAND_CLASSES = True           # 3-D lattices: band rows as translation classes for the operator kernels
MAX_DIRS = 64                 # PDE_MAX_DIRS of include/pde_b200.h


class Slab(object):
    """Rows [x_offset, x_offset + nx_local) of the lattice plus ``halo`` rows each side."""

    def __init__(self, x_offset, nx_local, halo):
        self.x_offset, self.nx_local, self.halo = int(x_offset), int(nx_local), int(halo)


class GridContext(object):
    """Host handle of one ``pde_grid``."""

    def __init__(self, G, device=None, slab=None):
        self.lib = _lib.lib()
        self.device = require_cuda(device)
        self.G = G
        structured = hasattr(G, "stencil_pairs") and hasattr(G, "band_pairs")
        # 3-D lattices: table-driven kernel when the table fits (halo <= 2, D <= 64)
        structured_nd = (ND_STRUCTURED and slab is None and hasattr(G, "stencil_pairs_nd")
                         and hasattr(G, "band_pairs") and getattr(G, "dim", 2) == 3
                         and 1 <= G.halo <= 2 and len(G.stencil_pairs_nd) <= MAX_DIRS
                         and G.num_deep > 0)
        structured = structured or structured_nd
        exact = structured and G.uniform_weights()
        # lattices with inexact coordinates (FDRegularGrid(weights=...)): one weight triple per
        # direction ("nominal", within the documented tolerance) or explicit per-pair rows
        mode = getattr(G, "weights", "auto")
        nominal = structured and not exact and (
            mode == "nominal" or (mode == "auto" and (G.num_interior > NOMINAL_ABOVE or slab is not None)))
        self.nd = False
        if (exact or nominal) and structured_nd:
            self._init_structured_nd(G, exact)
            self.weights = "exact" if exact else "nominal"
        elif exact or nominal:
            self._init_structured(G, slab)
            self.weights = "exact" if exact else "nominal"
        else:
            if slab is not None:
                raise _lib.PdeError("slab decomposition needs a structured grid")
            self._init_generic(G)
            self.weights = "exact"
        self.pitch = int(self.lib.pde_grid_pitch(self.h))
        self.rows = int(self.lib.pde_grid_rows(self.h))
        self.n_state = int(self.lib.pde_grid_state_len(self.h))
        self.n_lattice = self.rows * self.pitch

    # ---- construction --------------------------------------------------------------
    def _create(self, nx_global, ny, nb, radius, x_offset, nx_local, halo, sp, sw, sX,
                band_center, band_ptr, band_j, band_w, band_X):
        sp = np.ascontiguousarray(sp, dtype=np.int32).reshape(-1, 4)
        sw = np.ascontiguousarray(sw, dtype=np.float64).reshape(-1, 3)
        sX = np.ascontiguousarray(sX, dtype=np.float64).reshape(-1, 2)
        band_center = np.ascontiguousarray(band_center, dtype=np.int64)
        band_ptr = np.ascontiguousarray(band_ptr, dtype=np.int64)
        band_j = np.ascontiguousarray(band_j, dtype=np.int64).reshape(-1, 2)
        band_w = np.ascontiguousarray(band_w, dtype=np.float64).reshape(-1, 3)
        band_X = np.ascontiguousarray(band_X, dtype=np.float64).reshape(-1, 2)
        h = vp()
        check(self.lib.pde_grid_create(
            C.byref(h), self.device.index, nx_global, ny, nb, radius, x_offset, nx_local, halo,
            sp.shape[0], _np_ptr(sp), _np_ptr(sw), _np_ptr(sX),
            band_center.shape[0], _np_ptr(band_center), _np_ptr(band_ptr), _np_ptr(band_j),
            _np_ptr(band_w), _np_ptr(band_X)))
        self.h = h
        self.nx_global, self.ny, self.nb = int(nx_global), int(ny), int(nb)
        self.x_offset, self.nx_local, self.halo_rows = int(x_offset), int(nx_local), int(halo)
        self.radius = int(radius)
        self.n_interior_local = self.nx_local * self.ny
        self.stencil_pairs = sp
        self.band_ptr = band_ptr
        self.n_band = band_center.shape[0]

    def _init_structured(self, G, slab):
        nx, ny = G.interior_shape
        r = G.stencil_radius
        if slab is None:
            slab = Slab(0, nx, 0)
        self.slab = slab
        pitch = (ny + 15) // 16 * 16
        rows = slab.nx_local + 2 * slab.halo
        N = G.num_interior
        sp = G.stencil_pairs
        # per-direction weights at a generic deep point (position independent here)
        c = np.array([[r + 1.0, r + 1.0]])
        PI = G._scaled(c)
        w, X0 = pair_weights(np.repeat(PI, len(sp), 0), G._scaled(c + sp[:, 0:2]),
                             G._scaled(c + sp[:, 2:4]))
        # band rows owned by this slab
        bi = G.band_index
        brow = bi // ny
        own = (brow >= slab.x_offset) & (brow < slab.x_offset + slab.nx_local)
        sel = np.flatnonzero(own)
        self.band_sel = sel
        ptr = G.band_ptr
        counts = (ptr[1:] - ptr[:-1])[sel]
        lptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        take = (np.repeat(ptr[sel] - lptr[:-1], counts) + np.arange(lptr[-1])).astype(np.int64)
        bp = G.band_pairs[take]

        def offs(ids):
            ids = np.asarray(ids, dtype=np.int64)
            lat = ids < N
            o = np.empty(ids.shape, np.int64)
            rr = ids[lat] // ny - slab.x_offset + slab.halo
            if lat.any() and ((rr < 0).any() or (rr >= rows).any()):
                raise _lib.PdeError("slab halo too small for the band stencil")
            o[lat] = rr * pitch + ids[lat] % ny
            o[~lat] = rows * pitch + (ids[~lat] - N)
            return o

        bw, bX = pair_weights(G.coords(bp[:, 0]), G.coords(bp[:, 1]), G.coords(bp[:, 2]))
        self.band_global = bi[sel]
        self._create(nx, ny, G.num_bdry, r, slab.x_offset, slab.nx_local, slab.halo,
                     sp, w, X0, offs(bi[sel]), lptr,
                     np.stack([offs(bp[:, 1]), offs(bp[:, 2])], 1) if len(bp) else np.zeros((0, 2)),
                     bw, bX)
        self.band_pairs = bp
        self.structured = True

    def _init_structured_nd(self, G, exact=True):
        """3-D lattice [nx][ny][nz] as nx*ny state-layout rows of nz columns
        (``pde_grid_create_nd``): deep points through the 3-D table kernel, band points as
        explicit rows.  Point ids, policies and the flat vectors are the reference's."""
        nx, ny, nz = G.interior_shape
        dim, r, N = 3, G.stencil_radius, G.num_interior
        pitch = (nz + 15) // 16 * 16
        rows = nx * ny
        sp = np.ascontiguousarray(G.stencil_pairs_nd, dtype=np.int32)
        c = np.full((1, dim), r + 1.0)
        w, X0 = pair_weights(np.repeat(G._scaled(c), len(sp), 0), G._scaled(c + sp[:, :dim]),
                             G._scaled(c + sp[:, dim:]))
        self.stencil_X_nd = X0                      # (D, 3) physical vector of the first neighbour
        bi = np.asarray(G.band_index, dtype=np.int64)
        bp = np.asarray(G.band_pairs, dtype=np.int64)

        def offs(ids):
            ids = np.asarray(ids, dtype=np.int64)
            return np.where(ids < N, (ids // nz) * pitch + ids % nz, rows * pitch + (ids - N))

        bw, bX = pair_weights(G.coords(bp[:, 0]), G.coords(bp[:, 1]), G.coords(bp[:, 2]))
        self.band_X_full = bX
        shape = np.array([nx, ny, nz], dtype=np.int64)
        band_center = np.ascontiguousarray(offs(bi))
        band_ptr = np.ascontiguousarray(G.band_ptr, dtype=np.int64)
        band_j = np.ascontiguousarray(np.stack([offs(bp[:, 1]), offs(bp[:, 2])], 1)
                                      if len(bp) else np.zeros((0, 2), np.int64))
        w = np.ascontiguousarray(w)
        h = vp()
        check(self.lib.pde_grid_create_nd(
            C.byref(h), self.device.index, dim, _np_ptr(shape), G.num_bdry, r, sp.shape[0],
            _np_ptr(sp), _np_ptr(w), bi.shape[0], _np_ptr(band_center), _np_ptr(band_ptr),
            _np_ptr(band_j), _np_ptr(bw)))
        self.h = h
        self.slab = Slab(0, rows, 0)
        self.nx_global, self.ny, self.nb = rows, int(nz), int(G.num_bdry)
        self.x_offset, self.nx_local, self.halo_rows = 0, rows, 0
        self.radius = int(r)
        self.n_interior_local = N
        self.stencil_pairs = sp
        self.band_ptr = band_ptr
        self.n_band = bi.shape[0]
        self.band_pairs = bp
        self.band_global = bi
        self.band_sel = np.arange(bi.shape[0])
        self.structured = True
        self.nd = True
        # compact band rows for the operator kernels (exact lattices: weights are per class)
        self.band_classes = 0
        if BAND_CLASSES and exact:
            t = self._band_classes(G, bi, bp, bw, band_ptr, pitch, rows)
            if t is not None:
                check(self.lib.pde_grid_set_band_classes(
                    self.h, t["n_cls"], _np_ptr(t["cls_ptr"]), _np_ptr(t["cls_ref"]),
                    _np_ptr(t["cls_w"]), _np_ptr(t["cls_id"]), _np_ptr(t["wl_ptr"]), _np_ptr(t["wl"])))
                self.band_classes = t["n_cls"]

    @staticmethod
    def _band_classes(G, bi, bp, bw, band_ptr, pitch, rows):
        """Translation classes of the band rows of a 3-D lattice for ``pde_grid_set_band_classes``:
        band points with the same distances to the six walls (capped at r + 2) have the same rows
        up to translation.  The rows of every class are taken from its first member and EVERY
        member is checked against them (row count, lattice offsets, wall-list positions, weights
        bit for bit); any mismatch -> None, the explicit rows stay in charge."""
        nx, ny, nz = G.interior_shape
        r, N = G.stencil_radius, G.num_interior
        nB, P = bi.shape[0], bp.shape[0]
        if nB == 0 or P == 0 or P >= 2 ** 31:
            return None
        counts = np.diff(band_ptr)
        pt = np.repeat(np.arange(nB), counts)                 # band point number of every row
        kin = np.arange(P) - band_ptr[:-1][pt]                # row number within its point

        def soff(ids):                                        # lattice ids -> state offsets
            return (ids // nz) * pitch + ids % nz

        ctr = soff(bp[:, 0])
        J = bp[:, 1:3]
        lat = J < N
        delta = np.where(lat, soff(np.where(lat, J, 0)) - ctr[:, None], 0)
        # the wall points a band point refers to, in order of first appearance in its rows
        flatJ, flatpt, flatlat = J.ravel(), np.repeat(pt, 2), lat.ravel()
        wsel = np.flatnonzero(~flatlat)
        nbw = int(G.num_bdry)
        key = flatpt[wsel] * nbw + (flatJ[wsel] - N)
        ukey, first = np.unique(key, return_index=True)
        order = np.argsort(first, kind="stable")              # by point, then by first appearance
        upt = ukey[order] // nbw
        wl_counts = np.bincount(upt, minlength=nB)
        wl_ptr = np.concatenate([[0], np.cumsum(wl_counts)])
        pos_sorted = np.arange(order.size) - wl_ptr[:-1][upt]
        pos_of_ukey = np.empty(order.size, np.int64)
        pos_of_ukey[order] = pos_sorted
        wl = rows * pitch + (ukey[order] % nbw)               # state offsets of the wall points
        wpos = np.zeros(2 * P, np.int64)
        wpos[wsel] = pos_of_ukey[np.searchsorted(ukey, key)]
        ref = np.where(flatlat, delta.ravel() * 2, wpos * 2 + 1).reshape(P, 2)
        if np.abs(ref).max() >= 2 ** 31 or wl.max(initial=0) >= 2 ** 31:
            return None
        # classes by wall distances
        idx = np.stack(np.unravel_index(bi, (nx, ny, nz)), 1) + 1
        sig = np.zeros(nB, np.int64)
        for d, n in enumerate((nx, ny, nz)):
            a = np.minimum(idx[:, d] - 1, r + 2)
            b = np.minimum(n - idx[:, d], r + 2)
            sig = (sig * (r + 3) + a) * (r + 3) + b
        _, rep, cls = np.unique(sig, return_index=True, return_inverse=True)
        n_cls = rep.shape[0]
        if n_cls >= 65535:
            return None
        cls_counts = counts[rep]
        cls_ptr = np.concatenate([[0], np.cumsum(cls_counts)])
        if (counts != cls_counts[cls]).any():
            return None
        take = (np.repeat(band_ptr[:-1][rep], cls_counts)
                + np.arange(cls_ptr[-1]) - np.repeat(cls_ptr[:-1], cls_counts))
        cls_ref, cls_w = ref[take], bw[take]
        mine = cls_ptr[:-1][cls[pt]] + kin                    # class row every explicit row maps to
        if not (np.array_equal(ref, cls_ref[mine]) and np.array_equal(bw, cls_w[mine])):
            return None
        return dict(n_cls=int(n_cls), cls_ptr=np.ascontiguousarray(cls_ptr, dtype=np.int32),
                    cls_ref=np.ascontiguousarray(cls_ref, dtype=np.int32),
                    cls_w=np.ascontiguousarray(cls_w, dtype=np.float64),
                    cls_id=np.ascontiguousarray(cls, dtype=np.uint16),
                    wl_ptr=np.ascontiguousarray(wl_ptr, dtype=np.int32),
                    wl=np.ascontiguousarray(wl if wl.size else np.zeros(1), dtype=np.int32))

    def deep_pair_weights(self):
        """(w_0, w_1, w0) of every row of the deep table, at a generic deep point."""
        G, sp = self.G, self.stencil_pairs
        dim = sp.shape[1] // 2
        c = np.full((1, dim), self.radius + 1.0)
        w, _ = pair_weights(np.repeat(G._scaled(c), len(sp), 0), G._scaled(c + sp[:, :dim]),
                            G._scaled(c + sp[:, dim:]))
        return w

    def deep_neighbours(self, deep_ids, rows):
        """Global point ids (J0, J1) of table rows ``rows`` applied at the deep points
        ``deep_ids`` (flat interior ids of this context's lattice)."""
        sp = self.stencil_pairs[rows].astype(np.int64)
        if self.nd:
            nx, ny, nz = self.G.interior_shape
            stride = np.array([ny * nz, nz, 1], dtype=np.int64)
            return (deep_ids + (sp[:, :3] * stride).sum(1), deep_ids + (sp[:, 3:] * stride).sum(1))
        ny = self.ny
        di, dj = deep_ids // ny, deep_ids % ny
        return ((di + sp[:, 0]) * ny + dj + sp[:, 1], (di + sp[:, 2]) * ny + dj + sp[:, 3])

    def _init_generic(self, G):
        """Any object with ``pairs``, ``points``, ``num_interior`` (e.g. a grid built
        by the reference constructor, or a non-dyadic grid): every interior point is
        handled by the explicit pair-row kernel."""
        pairs = np.asarray(G.pairs, dtype=np.int64)
        points = np.asarray(G.points, dtype=np.float64)
        N = int(G.num_interior)
        nb = points.shape[0] - N
        order = np.argsort(pairs[:, 0], kind="stable")
        pairs = pairs[order]
        counts = np.bincount(pairs[:, 0], minlength=N)
        if (counts == 0).any():
            raise TypeError("Point {0} has no pairs".format(int(np.flatnonzero(counts == 0)[0])))
        ptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        pitch = (N + 15) // 16 * 16

        def offs(ids):
            ids = np.asarray(ids, dtype=np.int64)
            return np.where(ids < N, ids, pitch + (ids - N))

        bw, bX = pair_weights(points[pairs[:, 0]], points[pairs[:, 1]], points[pairs[:, 2]])
        self.band_X_full = bX                       # (P, dim): control vectors in any dimension
        if bX.shape[1] != 2:                         # the device copy serves 2-D control only
            bX = np.zeros((bX.shape[0], 2))
        self.slab = Slab(0, 1, 0)
        self._create(1, N, nb, N + 1, 0, 1, 0, np.zeros((0, 4)), np.zeros((0, 3)), np.zeros((0, 2)),
                     offs(np.arange(N)), ptr, np.stack([offs(pairs[:, 1]), offs(pairs[:, 2])], 1),
                     bw, bX)
        self.band_pairs = pairs
        self.band_global = np.arange(N)
        self.pair_order = order
        self.structured = False

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.pde_grid_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # ---- cache on the grid object ----------------------------------------------------------
    @staticmethod
    def get(G, device=None):
        device = require_cuda(device)
        cache = getattr(G, "_b200ctx", None)
        if cache is None:
            cache = {}
            try:
                G._b200ctx = cache
            except AttributeError:
                pass
        key = device.index
        if key not in cache:
            cache[key] = GridContext(G, device)
        return cache[key]

    @staticmethod
    def drop(G):
        """Release the device contexts cached on ``G`` (their HBM goes back to the pool)."""
        cache = getattr(G, "_b200ctx", None)
        if cache:
            cache.clear()

    # ---- buffers ---------------------------------------------------------------------------------
    def zeros(self):
        return torch.zeros(self.n_state, dtype=torch.float64, device=self.device)

    def empty_policy(self):
        return torch.empty(self.n_lattice, dtype=torch.uint16, device=self.device)

    def n_flat(self):
        return self.n_interior_local + self.nb

    def pack(self, u_flat, out=None):
        u_flat = to_device(u_flat, self.device)
        if u_flat.numel() != self.n_flat():
            raise ValueError("expected a vector over all %d points, got %d"
                             % (self.n_flat(), u_flat.numel()))
        S = out if out is not None else torch.empty(self.n_state, dtype=torch.float64,
                                                    device=self.device)
        check(self.lib.pde_grid_pack(self.h, _ptr(u_flat), _ptr(S), _stream()))
        return S

    def unpack(self, S):
        out = torch.empty(self.n_flat(), dtype=torch.float64, device=self.device)
        check(self.lib.pde_grid_unpack(self.h, _ptr(S), _ptr(out), _stream()))
        return out

    def unpad(self, S):
        out = torch.empty(self.n_interior_local, dtype=torch.float64, device=self.device)
        check(self.lib.pde_grid_unpad(self.h, _ptr(S), _ptr(out), _stream()))
        return out

    def unpad_policy(self, P):
        out = torch.empty(self.n_interior_local, dtype=torch.int32, device=self.device)
        check(self.lib.pde_grid_unpad_u16(self.h, _ptr(P), _ptr(out), _stream()))
        return out

    def set_row_window(self, lo=0, hi=None):
        """Restrict the stencil launches to owned rows [lo, hi) (None = all rows)."""
        check(self.lib.pde_grid_set_row_window(self.h, int(lo), int(self.nx_local if hi is None else hi)))

    def set_row_windows2(self, lo, hi, lo2, hi2):
        """Two disjoint row ranges served by one set of launches (the edge strips of a slab)."""
        check(self.lib.pde_grid_set_row_windows2(self.h, int(lo), int(hi), int(lo2), int(hi2)))

    # ---- operators --------------------------------------------------------------------------------
    def d2eigs(self, S, want_min=True, want_max=True, policy=False, mode=0, out=None):
        """Raw ddg.d2eigs on a state tensor -> (lmin, lmax, pmin, pmax) state-layout."""
        if out is not None:
            lmin, lmax, pmin, pmax = out
        else:
            lmin = self.zeros() if want_min else None
            lmax = self.zeros() if want_max else None
            pmin = self.empty_policy() if (policy and want_min) else None
            pmax = self.empty_policy() if (policy and want_max) else None
        check(self.lib.pde_ddg_d2eigs(self.h, _ptr(S), _ptr(lmin), _ptr(lmax), _ptr(pmin),
                                      _ptr(pmax), int(mode), _stream()))
        return lmin, lmax, pmin, pmax

    def d2eigs_host(self, u_host, want_min=True, want_max=True, mode=0, chunk_rows=None,
                    linear=None, halo_hook=None, out=None, native=False):
        """ddg.d2eigs values for a field in PINNED host memory, results in pinned host
        memory: the H2D copy, the stencil kernels and the D2H copy are pipelined over
        chunks of rows on three streams (PCIe is full duplex).  ``linear=False``: the layout
        conversion is done by the 2-D DMA copies themselves; ``linear=True`` (default,
        ``HOST_COPY_LINEAR``): the copy streams carry nothing but back-to-back linear PCIe copies
        to / from flat device staging buffers and the conversions are device-side copies on the
        compute stream (4097^2, 128-row chunks: 3.5 ms against 3.75 ms).  ``native=True``: the same
        pipeline enqueued by ONE C call (``pde_ddg_d2eigs_host``, the C-ABI entry point for host
        buffers; same results, 3.5 - 4.2 ms with a larger run-to-run spread).
        ``halo_hook(S)`` (slab contexts, multigpu.ShardedGrid.d2eigs_host): the slab's first /
        last rows are uploaded first and the hook -- the halo exchange -- runs on the compute
        stream before the first chunk.  ``out``: (pinned lmin, pinned lmax) to reuse."""
        nloc, ny, H = self.nx_local, self.ny, max(self.radius, 1)
        if chunk_rows is None:
            chunk_rows = HOST_CHUNK_ROWS if halo_hook is None else 256
        if native:
            if halo_hook is not None or self.nd or self.halo_rows or not self.structured:
                raise _lib.PdeError("pde_ddg_d2eigs_host serves unsharded 2-D lattices")
            return self._d2eigs_host_native(u_host, want_min, want_max, mode, chunk_rows, out)
        linear = HOST_COPY_LINEAR if linear is None else linear
        if not hasattr(self, "_stage"):
            self._stage = (self.zeros(), self.zeros(), self.zeros(),
                           torch.cuda.Stream(device=self.device), torch.cuda.Stream(device=self.device))
        S, Lmin, Lmax, s_in, s_out = self._stage
        if linear and not hasattr(self, "_stage_flat"):
            self._stage_flat = (torch.empty(self.n_flat(), dtype=torch.float64, device=self.device),
                                torch.empty(nloc * ny, dtype=torch.float64, device=self.device),
                                torch.empty(nloc * ny, dtype=torch.float64, device=self.device))
        main = torch.cuda.current_stream()
        out_min, out_max = out if out is not None else (None, None)
        if want_min and out_min is None:
            out_min = torch.empty(nloc * ny, dtype=torch.float64, pin_memory=True)
        if want_max and out_max is None:
            out_max = torch.empty(nloc * ny, dtype=torch.float64, pin_memory=True)
        edges = list(range(0, nloc, chunk_rows)) + [nloc]
        nchunk = len(edges) - 1
        ev_in = [torch.cuda.Event() for _ in range(nchunk)]
        s_in.wait_stream(main)
        s_out.wait_stream(main)

        def upload(lo, hi, wall):
            """Enqueue (on the current stream) the PCIe copy of rows [lo, hi) of the field."""
            if linear:
                fin = self._stage_flat[0]
                fin[lo * ny:hi * ny].copy_(u_host[lo * ny:hi * ny], non_blocking=True)
                if wall and self.nb:
                    fin[nloc * ny:].copy_(u_host[nloc * ny:], non_blocking=True)
                return
            check(self.lib.pde_grid_h2d_rows(self.h, vp(u_host.data_ptr()), _ptr(S), lo, hi,
                                             int(wall), _stream()))

        def convert_in(lo, hi, wall):
            """linear mode: flat staging -> state layout, a device-side 2-D copy on the COMPUTE
            stream (the copy streams carry nothing but back-to-back linear PCIe copies)."""
            if linear:
                check(self.lib.pde_grid_h2d_rows(self.h, vp(self._stage_flat[0].data_ptr()), _ptr(S),
                                                 lo, hi, int(wall), _stream()))

        with torch.cuda.stream(s_in):
            first = 0
            if halo_hook is not None and self.halo_rows and nloc >= 2 * self.halo_rows:
                # the rows the neighbours need go first; chunk 0 then starts after them
                hr = self.halo_rows
                upload(0, hr, True)
                upload(nloc - hr, nloc, False)
                ev_edge = torch.cuda.Event()
                ev_edge.record()
            else:
                ev_edge = None
            for k in range(nchunk):
                upload(edges[k], edges[k + 1], k == 0 and ev_edge is None)
                ev_in[k].record()
        if halo_hook is not None:
            if ev_edge is not None:
                main.wait_event(ev_edge)
                convert_in(0, self.halo_rows, True)
                convert_in(nloc - self.halo_rows, nloc, False)
            else:
                main.wait_event(ev_in[-1])
                for k in range(nchunk):
                    convert_in(edges[k], edges[k + 1], k == 0)
            halo_hook(S)
        outs = (Lmin if want_min else None, Lmax if want_max else None, None, None)
        converted = nchunk if (linear and halo_hook is not None and ev_edge is None) else 0
        try:
            for k in range(nchunk):
                lo, hi = edges[k], edges[k + 1]
                # rows of chunk k read up to H rows of chunk k+1
                need = min(k + 1, nchunk - 1)
                main.wait_event(ev_in[need])
                while linear and converted <= need:
                    convert_in(edges[converted], edges[converted + 1],
                               converted == 0 and ev_edge is None)
                    converted += 1
                self.set_row_window(lo, hi)
                self.d2eigs(S, want_min=want_min, want_max=want_max, mode=mode, out=outs)
                if linear:      # state layout -> flat staging, still on the compute stream
                    for q, L in enumerate((Lmin, Lmax)):
                        if (want_min, want_max)[q]:
                            check(self.lib.pde_grid_d2h_rows(
                                self.h, _ptr(L), vp(self._stage_flat[1 + q].data_ptr()), lo, hi,
                                _stream()))
                ev = torch.cuda.Event()
                ev.record(main)
                s_out.wait_event(ev)
                with torch.cuda.stream(s_out):
                    for q, (L, o) in enumerate(((Lmin, out_min), (Lmax, out_max))):
                        if o is None or not (want_min, want_max)[q]:
                            continue
                        if linear:
                            o[lo * ny:hi * ny].copy_(self._stage_flat[1 + q][lo * ny:hi * ny],
                                                     non_blocking=True)
                        else:
                            check(self.lib.pde_grid_d2h_rows(self.h, _ptr(L), vp(o.data_ptr()),
                                                             lo, hi, _stream()))
        finally:
            self.set_row_window()
        main.wait_stream(s_out)
        s_out.synchronize()
        return out_min, out_max

    def _d2eigs_host_native(self, u_host, want_min, want_max, mode, chunk_rows, out):
        nloc, ny = self.nx_local, self.ny
        if not hasattr(self, "_stage"):
            self._stage = (self.zeros(), self.zeros(), self.zeros(),
                           torch.cuda.Stream(device=self.device), torch.cuda.Stream(device=self.device))
        if not hasattr(self, "_stage_flat"):
            self._stage_flat = (torch.empty(self.n_flat(), dtype=torch.float64, device=self.device),
                                torch.empty(nloc * ny, dtype=torch.float64, device=self.device),
                                torch.empty(nloc * ny, dtype=torch.float64, device=self.device))
        S, Lmin, Lmax, s_in, s_out = self._stage
        fin, fmin, fmax = self._stage_flat
        out_min, out_max = out if out is not None else (None, None)
        if want_min and out_min is None:
            out_min = torch.empty(nloc * ny, dtype=torch.float64, pin_memory=True)
        if want_max and out_max is None:
            out_max = torch.empty(nloc * ny, dtype=torch.float64, pin_memory=True)
        main = torch.cuda.current_stream()
        check(self.lib.pde_ddg_d2eigs_host(
            self.h, vp(u_host.data_ptr()), _ptr(out_min if want_min else None),
            _ptr(out_max if want_max else None), int(mode), int(chunk_rows), _ptr(S), _ptr(Lmin),
            _ptr(Lmax), _ptr(fin), _ptr(fmin), _ptr(fmax), vp(main.cuda_stream),
            vp(s_in.cuda_stream), vp(s_out.cuda_stream)))
        main.synchronize()
        return (out_min if want_min else None), (out_max if want_max else None)

    def apply_policy(self, policy, X):
        Y = self.zeros()
        check(self.lib.pde_ddg_apply_policy(self.h, _ptr(policy), _ptr(X), _ptr(Y), _stream()))
        return Y

    def control(self, pmin, pmax):
        vmin = torch.empty((self.n_interior_local, 2), dtype=torch.float64, device=self.device)
        vmax = torch.empty_like(vmin)
        check(self.lib.pde_ddg_control(self.h, _ptr(pmin), _ptr(pmax), _ptr(vmin), _ptr(vmax),
                                       _stream()))
        return vmin, vmax

    def jac_struct(self, pmin, pmax=None, cmin=1.0, cmax=0.0):
        J = pde_jac()
        J.d_pmin = pmin.data_ptr()
        J.d_pmax = pmax.data_ptr() if pmax is not None else None
        if isinstance(cmin, torch.Tensor):
            J.d_cmin, J.cmin_s = cmin.data_ptr(), 0.0
        else:
            J.d_cmin, J.cmin_s = None, float(cmin)
        if isinstance(cmax, torch.Tensor):
            J.d_cmax, J.cmax_s = cmax.data_ptr(), 0.0
        else:
            J.d_cmax, J.cmax_s = None, float(cmax)
        return J

    def jac_matvec(self, J, X, transpose=False, out=None):
        Y = out if out is not None else self.zeros()
        check(self.lib.pde_jac_matvec(self.h, C.byref(J), _ptr(X), _ptr(Y), int(transpose),
                                      _stream()))
        return Y

    def jac_diagonal(self, J):
        Dg = self.zeros()
        check(self.lib.pde_jac_diagonal(self.h, C.byref(J), _ptr(Dg), _stream()))
        return Dg

    def ce_residual(self, W, g, policy=True):
        F = self.zeros()
        P = self.empty_policy() if policy else None
        red = torch.zeros(2, dtype=torch.float64, device=self.device)
        check(self.lib.pde_ce_residual(self.h, _ptr(W), _ptr(g), _ptr(F), _ptr(P), _ptr(red),
                                       _stream()))
        return F, P, red

    def pucci_residual(self, W, f, A0, A1, policy=True):
        F = self.zeros()
        pmin = self.empty_policy() if policy else None
        pmax = self.empty_policy() if policy else None
        red = torch.zeros(2, dtype=torch.float64, device=self.device)
        a0 = A0 if isinstance(A0, torch.Tensor) else None
        a1 = A1 if isinstance(A1, torch.Tensor) else None
        check(self.lib.pde_pucci_residual(
            self.h, _ptr(W), _ptr(f), 0.0 if a0 is not None else float(A0),
            0.0 if a1 is not None else float(A1), _ptr(a0), _ptr(a1), _ptr(F), _ptr(pmin),
            _ptr(pmax), _ptr(red), _stream()))
        return F, pmin, pmax, red

    def ce_euler(self, U0, g, dt, solution_tol, operator_tol, max_iters, check_every=64):
        """Device-resident Euler loop.  Returns (U, diff, iters, max|F|, reason)."""
        U1 = self.zeros()
        out = (C.c_double * 5)()
        check(self.lib.pde_ce_euler(self.h, _ptr(U0), _ptr(U1), _ptr(g), float(dt),
                                    float(solution_tol), float(operator_tol), int(max_iters),
                                    int(check_every), out, _stream()))
        iters = int(out[0])
        U = U1 if int(out[1]) == 1 else U0
        return U, float(out[2]), iters, float(out[3]), int(out[4])

    def ce_sweep(self, g, tol, max_sweeps, relax_reps=2, U=None, relax_every=0, flat_tol=None):
        """Convex-hull sweeps (``pde_ce_sweep``, extension): lowers ``U`` (default: a copy of
        ``g``) towards the convex-envelope solution until one sweep changes no value by ``tol``
        or more.  ``relax_every``: direction passes between band relaxations (0: once per
        sweep).  Returns (U, last change, sweeps)."""
        if not self.structured or self.nd or self.halo_rows or self.nx_local != self.nx_global:
            raise _lib.PdeError("the sweep solver needs a structured, unsharded 2-D pairs grid")
        if not hasattr(self, "_sweep"):
            from . import _sweep
            t = _sweep.tables(self.G, self.pitch, self.rows)
            dev = self.device
            nb = int(t["band_center"].shape[0])
            wbytes = int(self.lib.pde_ce_sweep_work_bytes(self.nx_local, self.ny, int(t["dirs"].shape[0]),
                                                          _np_ptr(t["dirs"]), nb))
            if wbytes < 0:
                raise _lib.PdeError("lattice line too long for the sweep kernel")
            self._sweep = dict(
                dirs=t["dirs"], D=int(t["dirs"].shape[0]), n_band=nb,
                ring=torch.from_numpy(t["ring"]).to(dev),
                flags=torch.from_numpy(t["flags"].view(np.int64)).to(dev),
                band_ptr=torch.from_numpy(t["band_ptr"]).to(dev),
                band_center=torch.from_numpy(t["band_center"]).to(dev),
                rest_j=torch.from_numpy(t["rest_j"]).to(dev),
                theta=torch.from_numpy(t["theta"]).to(dev),
                work=torch.zeros((wbytes + 7) // 8, dtype=torch.float64, device=dev))
            t = self._sweep
            # one-off: fixed-vertex masks of every line word of every direction
            check(self.lib.pde_ce_sweep_setup(self.device.index, self.nx_local, self.ny, self.pitch,
                                              self.n_lattice, self.G.stencil_radius, t["D"],
                                              _np_ptr(t["dirs"]), _ptr(t["ring"]), _ptr(t["flags"]),
                                              t["n_band"], _ptr(t["work"]), _stream()))
        t = self._sweep
        if U is None:
            U = g.clone()
        if flat_tol is None:
            if t.get("gmax_of") is not g:
                from . import _sweep
                t["gmax_of"], t["flat_tol"] = g, _sweep.flat_tolerance(float(g.abs().max()))
            flat_tol = t["flat_tol"]
        out = (C.c_double * 2)()
        check(self.lib.pde_ce_sweep(self.device.index, self.nx_local, self.ny, self.pitch,
                                    self.n_lattice, self.G.stencil_radius, t["D"],
                                    _np_ptr(t["dirs"]), _ptr(t["ring"]), _ptr(t["flags"]),
                                    t["n_band"], _ptr(t["band_ptr"]), _ptr(t["band_center"]),
                                    _ptr(t["rest_j"]), _ptr(t["theta"]), _ptr(U),
                                    _ptr(t["work"]), float(tol), float(flat_tol), int(max_sweeps),
                                    int(relax_every), int(relax_reps), out, _stream()))
        return U, float(out[0]), int(out[1])

    def ce_euler_step(self, U, g, dt, out=None):
        """One Euler step -> (Unew, red) with red = [max|F|, max|U-Unew|] on the device."""
        if out is not None:
            Unew, red = out
        else:
            Unew = self.zeros()
            red = torch.zeros(2, dtype=torch.float64, device=self.device)
        check(self.lib.pde_ce_euler_step(self.h, _ptr(U), _ptr(g), float(dt), _ptr(Unew),
                                         _ptr(red), _stream()))
        return Unew, red

    def bicgstab(self, J, b, rtol, atol=0.0, maxiter=0, check_every=32, work=None):
        x = self.zeros()
        if work is None:
            work = torch.empty(8 * self.n_state, dtype=torch.float64, device=self.device)
        info = (C.c_int64 * 2)()
        check(self.lib.pde_bicgstab(self.h, C.byref(J), _ptr(b), _ptr(x), float(rtol), float(atol),
                                    int(maxiter), int(check_every), _ptr(work), info, _stream()))
        return x, int(info[0]), int(info[1])

    def newton_update(self, U, d, F, Jd):
        Unew = torch.empty_like(U)
        out = (C.c_double * 3)()
        check(self.lib.pde_vec_newton_update(self.h, _ptr(U), _ptr(d), _ptr(F), _ptr(Jd),
                                             _ptr(Unew), out, _stream()))
        return Unew, float(out[0]), float(out[1]), float(out[2])

    # ---- policy -> explicit rows (small grids; the `.tocsr()` escape hatch) ------------------------
    def policy_rows(self, policy_flat):
        """(J0, J1, w_0, w_1, w0) global ids / weights of the selected pair of every
        interior point, from a flat int policy (host numpy)."""
        pol = np.asarray(policy_flat, dtype=np.int64)
        N = self.n_interior_local
        J0 = np.empty(N, np.int64)
        J1 = np.empty(N, np.int64)
        W = np.empty((N, 3))
        if self.structured:
            G = self.G
            ny = self.ny
            isband = np.zeros(N, bool)
            isband[self.band_global - self.x_offset * ny] = True
            deep = np.flatnonzero(~isband)
            gid = deep + self.x_offset * ny
            J0[deep], J1[deep] = self.deep_neighbours(gid, pol[deep])
            W[deep] = self.deep_pair_weights()[pol[deep]]
            bl = self.band_global - self.x_offset * ny
            rows = self.band_ptr[:-1] + pol[bl]
            bp = self.band_pairs[rows]
            J0[bl], J1[bl] = bp[:, 1], bp[:, 2]
            bw, _ = pair_weights(G.coords(bp[:, 0]), G.coords(bp[:, 1]), G.coords(bp[:, 2]))
            W[bl] = bw
        else:
            rows = self.band_ptr[:-1] + pol
            bp = self.band_pairs[rows]
            J0, J1 = bp[:, 1], bp[:, 2]
            P = np.asarray(self.G.points)
            W, _ = pair_weights(P[bp[:, 0]], P[bp[:, 1]], P[bp[:, 2]])
        return J0, J1, W
