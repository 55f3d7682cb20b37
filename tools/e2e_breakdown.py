"""Where the end-to-end time of ddg.d2min(G, host_tensor) goes (GPU box)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyellipticfd_b200.pointclasses import FDRegularGrid
from pyellipticfd_b200._context import GridContext
from pyellipticfd_b200 import ddg

n = 4095
G = FDRegularGrid([n, n], np.array([[0.0, 1.0], [0.0, 1.0]]).T, 4, interpolation=False)
ctx = GridContext.get(G)
host = torch.randn(G.num_points, dtype=torch.float64).pin_memory()
out_host = torch.empty(G.num_interior, dtype=torch.float64, pin_memory=True)

def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3

dev = host.cuda()
S = ctx.pack(dev)
lm = ctx.zeros()
flat = ctx.unpad(lm)
print("h2d   %.2f ms (%.1f GB/s)" % ((a := t(lambda: host.to('cuda', non_blocking=True))), host.numel()*8/a/1e6))
print("pack  %.3f ms" % t(lambda: ctx.pack(dev)))
print("eigs  %.3f ms" % t(lambda: ctx.d2eigs(S, want_min=True, want_max=False, out=(lm, None, None, None))))
print("unpad %.3f ms" % t(lambda: ctx.unpad(lm)))
print("d2h   %.2f ms (%.1f GB/s)" % ((b := t(lambda: out_host.copy_(flat, non_blocking=True))), flat.numel()*8/b/1e6))
print("pin alloc+d2h %.2f ms" % t(lambda: torch.empty(flat.shape, dtype=flat.dtype, pin_memory=True).copy_(flat, non_blocking=True)))
print("api   %.2f ms" % t(lambda: ddg.d2min(G, host)))
for lin in (False, True):
    for cr in (128, 256, 512):
        print("linear %s chunk_rows %4d: %.2f ms" % (lin, cr, t(lambda: ctx.d2eigs_host(
            host, want_min=True, want_max=False, chunk_rows=cr, linear=lin), reps=10)))
ref = ctx.d2eigs_host(host, want_min=True, want_max=False, linear=False)[0]
alt = ctx.d2eigs_host(host, want_min=True, want_max=False, linear=True)[0]
print("linear == 2-D result:", bool(torch.equal(ref, alt)))
# duplex check: H2D and D2H of the same size on two streams at once
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def duplex():
    with torch.cuda.stream(s1):
        dev.copy_(host, non_blocking=True)
    with torch.cuda.stream(s2):
        out_host.copy_(flat, non_blocking=True)
print("duplex h2d+d2h %.2f ms" % t(duplex))
