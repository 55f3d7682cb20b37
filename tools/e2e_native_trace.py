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
cr = int(sys.argv[1]) if len(sys.argv) > 1 else 256
for _ in range(4):
    ctx.d2eigs_host(host, want_min=True, want_max=False, chunk_rows=cr, out=(out, None), native=True)
torch.cuda.synchronize()
os.environ["PDE_E2E_TRACE"] = "1"
for _ in range(2):
    t = time.perf_counter()
    ctx.d2eigs_host(host, want_min=True, want_max=False, chunk_rows=cr, out=(out, None), native=True)
    print("wall %.3f ms" % ((time.perf_counter() - t) * 1e3), file=sys.stderr)
