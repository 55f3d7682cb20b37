import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyellipticfd_b200.pointclasses import FDRegularGrid
from pyellipticfd_b200._context import GridContext
n = 4095
G = FDRegularGrid([n, n], np.array([[0.0, 1.0], [0.0, 1.0]]).T, 4, interpolation=False)
ctx = GridContext.get(G)
host = torch.randn(G.num_points, dtype=torch.float64).pin_memory()
out = torch.empty(G.num_interior, dtype=torch.float64, pin_memory=True)
def t(fn, reps=10):
    fn(); fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3
for lin in (True, False):
    for cr in (128, 256):
        print("linear %s chunk_rows %4d: %.2f ms" % (lin, cr, t(lambda: ctx.d2eigs_host(
            host, want_min=True, want_max=False, chunk_rows=cr, linear=lin, out=(out, None)))), flush=True)
ref = ctx.d2eigs_host(host, want_min=True, want_max=False, linear=False)[0].clone()
for cr in (32, 48, 64, 96, 128, 192, 256):
    ms = t(lambda: ctx.d2eigs_host(host, want_min=True, want_max=False, chunk_rows=cr, out=(out, None)))
    print("native chunk_rows %4d: %.2f ms  equal %s" % (cr, ms, bool(torch.equal(out, ref))), flush=True)
lo, hi = ctx.d2eigs_host(host, chunk_rows=96)
r2 = ctx.d2eigs_host(host, linear=False)
print("both extremes equal:", bool(torch.equal(lo, r2[0])) and bool(torch.equal(hi, r2[1])))
