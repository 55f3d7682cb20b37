import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyellipticfd_b200 import multigpu
bind = len(sys.argv) > 1 and sys.argv[1] == "bind"
torch.cuda.init()
node = multigpu.bind_to_gpu_numa_node(0) if bind else None
print("bind", bind, "node", node, "cpus", len(os.sched_getaffinity(0)))
from pyellipticfd_b200.pointclasses import FDRegularGrid
from pyellipticfd_b200._context import GridContext
from pyellipticfd_b200 import ddg
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
dev = host.cuda(); flat = torch.empty(G.num_interior, dtype=torch.float64, device="cuda")
print("h2d %.2f ms" % t(lambda: dev.copy_(host, non_blocking=True)))
print("d2h %.2f ms" % t(lambda: out.copy_(flat, non_blocking=True)))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def duplex():
    with torch.cuda.stream(s1): dev.copy_(host, non_blocking=True)
    with torch.cuda.stream(s2): out.copy_(flat, non_blocking=True)
print("duplex %.2f ms" % t(duplex))
print("api %.2f ms" % t(lambda: ddg.d2min(G, host)))
for cr in (96, 128, 192):
    print("linear chunk %d: %.2f ms" % (cr, t(lambda: ctx.d2eigs_host(host, want_min=True, want_max=False, chunk_rows=cr, out=(out, None)))))
    print("native chunk %d: %.2f ms" % (cr, t(lambda: ctx.d2eigs_host(host, want_min=True, want_max=False, chunk_rows=cr, out=(out, None), native=True))))
