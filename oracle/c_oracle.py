"""ctypes access to the plain-C oracle (oracle/c/ddg_oracle.c) -- TEST INFRASTRUCTURE."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "c", "liboracle_ddg.so")
_lib = None


def build():
    src = os.path.join(HERE, "c", "ddg_oracle.c")
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", os.path.join(HERE, "c")], stdout=subprocess.DEVNULL)
    return LIB


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
    return _lib


def _p(a):
    return C.c_void_p(a.ctypes.data)


def pairs_d2eigs(points, pairs, u, nint):
    points = np.ascontiguousarray(points, np.float64)
    pairs = np.ascontiguousarray(pairs, np.int64)
    u = np.ascontiguousarray(u, np.float64)
    lmin, lmax = np.empty(nint), np.empty(nint)
    rmin, rmax = np.empty(nint, np.int64), np.empty(nint, np.int64)
    lib().ddg_pairs_d2eigs(_p(points), _p(pairs), C.c_int64(len(pairs)), _p(u), C.c_int64(nint),
                           _p(lmin), _p(lmax), _p(rmin), _p(rmax))
    return lmin, lmax, rmin, rmax


def lattice_d2eigs(G, u, row_lo=0, row_hi=None, out=None):
    """Deep-interior values of a pyellipticfd_b200 FDRegularGrid (or any object with
    interior_shape, scaling, bounds, stencil_radius, stencil_pairs)."""
    nx, ny = G.interior_shape
    u = np.ascontiguousarray(u, np.float64)
    if row_hi is None:
        row_hi = nx
    if out is None:
        out = (np.full(nx * ny, np.nan), np.full(nx * ny, np.nan),
               np.full(nx * ny, -1, np.int32), np.full(nx * ny, -1, np.int32))
    table = np.ascontiguousarray(G.stencil_pairs, np.int32)
    lib().ddg_lattice_d2eigs(C.c_int64(nx), C.c_int64(ny), C.c_double(G.scaling[0]),
                             C.c_double(G.scaling[1]), C.c_double(G.bounds[0][0]),
                             C.c_double(G.bounds[0][1]), C.c_int(G.stencil_radius),
                             C.c_int(len(table)), _p(table), _p(u), C.c_int64(row_lo),
                             C.c_int64(row_hi), _p(out[0]), _p(out[1]), _p(out[2]), _p(out[3]))
    return out
