"""Phase ablation of k_hull_pass at the converged state (timing only; PDE_SWEEP_STOP_AFTER=n)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyellipticfd_b200._context import GridContext
from pyellipticfd_b200.pointclasses import FDRegularGrid
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4095
G = FDRegularGrid([n, n], np.array([[-1.0, 1.0], [-1.0, 1.0]]).T, 4, interpolation=False)
C = 3 / 7 * np.array([[np.cos(np.pi / 5), np.sin(np.pi / 5)], [-np.cos(np.pi / 5), -np.sin(np.pi / 5)]])
P = torch.from_numpy(G.points).cuda()
g = torch.minimum(((P + torch.tensor(C[0]).cuda()) ** 2).sum(1).sqrt(), ((P + torch.tensor(C[1]).cuda()) ** 2).sum(1).sqrt())
ctx = GridContext.get(G); gs = ctx.pack(g)
os.environ.pop("PDE_SWEEP_STOP_AFTER", None)
U, c, k = ctx.ce_sweep(gs, 1e-14, 200, relax_every=4)
out = {"sweeps": k}
for stop in (1, 11, 12, 13, 2, 3, 4, 5, 6, 0):
    os.environ["PDE_SWEEP_STOP_AFTER"] = str(stop)
    V = U.clone()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.ce_sweep(gs, 0.0, 1, U=V, relax_every=4)
    torch.cuda.synchronize(); a.record()
    ctx.ce_sweep(gs, 0.0, 4, U=V, relax_every=4)
    b.record(); torch.cuda.synchronize()
    out["stop_after_%d_ms_per_pass" % stop] = round(a.elapsed_time(b) / 4 / 24, 4)
print(json.dumps(out))
