"""Where the time of a SMALL Newton solve goes (63^2 and 255^2 lattices, r = 2): per-call wall
clock, synchronised, of the pieces of solvers.newton_state."""
import os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyellipticfd_b200.pointclasses import FDRegularGrid
from pyellipticfd_b200._context import GridContext
from pyellipticfd_b200 import convex_envelope, linop

def two_cones(X):
    c = 3 / 7 * np.array([[np.cos(np.pi / 5), np.sin(np.pi / 5)], [-np.cos(np.pi / 5), -np.sin(np.pi / 5)]])
    return np.minimum(np.linalg.norm(X - c[0], axis=1), np.linalg.norm(X - c[1], axis=1))

acc = {}
def wrap(obj, name, key):
    f = getattr(obj, name)
    def g(*a, **k):
        torch.cuda.synchronize(); t = time.perf_counter()
        r = f(*a, **k)
        torch.cuda.synchronize(); d = time.perf_counter() - t
        acc.setdefault(key, [0, 0.0]); acc[key][0] += 1; acc[key][1] += d
        return r
    setattr(obj, name, g)

wrap(GridContext, "ce_residual", "ce_residual")
wrap(GridContext, "newton_update", "newton_update")
wrap(GridContext, "bicgstab", "ctx.bicgstab")
wrap(linop.PolicyJacobian, "matvec_state", "matvec")
for size in (63, 255):
    G = FDRegularGrid([size] * 2, np.array([[-1.0, 1.0], [-1.0, 1.0]]).T, 2, interpolation=False)
    g = two_cones(G.points)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        convex_envelope._solve_grid(G, g, None, "newton", stats={}, max_iters=3)
        acc.clear()
        st = {}
        torch.cuda.synchronize(); t = time.perf_counter()
        U, diff, it, _ = convex_envelope._solve_grid(G, g, None, "newton", stats=st)
        torch.cuda.synchronize(); tot = time.perf_counter() - t
    print(size, "total %.3f s" % tot, st)
    for k, (n, d) in acc.items():
        print("   %-14s calls %4d  total %.4f s  per call %.3f ms" % (k, n, d, d / n * 1e3))
