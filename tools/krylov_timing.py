"""Per-iteration cost of the device BiCGStab on the 4097^2 grid (GPU box)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyellipticfd_b200.pointclasses import FDRegularGrid
from pyellipticfd_b200._context import GridContext
from pyellipticfd_b200.linop import PolicyJacobian
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4095
G = FDRegularGrid([n, n], np.array([[-1.0, 1.0], [-1.0, 1.0]]).T, 4, interpolation=False)
ctx = GridContext.get(G)
xs = torch.linspace(-1, 1, n + 2, device='cuda', dtype=torch.float64)[1:-1]
flat = torch.empty(G.num_points, dtype=torch.float64, device='cuda')
r1 = torch.sqrt((xs[:, None] + 0.35) ** 2 + (xs[None, :] + 0.25) ** 2)
r2 = torch.sqrt((xs[:, None] - 0.35) ** 2 + (xs[None, :] - 0.25) ** 2)
flat[:n * n] = torch.minimum(r1, r2).reshape(-1)
wp = torch.from_numpy(G.bdry_points).cuda()
flat[n * n:] = torch.minimum(torch.sqrt((wp[:, 0] + 0.35) ** 2 + (wp[:, 1] + 0.25) ** 2), torch.sqrt((wp[:, 0] - 0.35) ** 2 + (wp[:, 1] - 0.25) ** 2))
g = ctx.pack(flat)
F, P, red = ctx.ce_residual(g, g, policy=True)
J = PolicyJacobian(ctx, P, None, cmin=-1.0)
B = -F
for iters in (64, 256):
    torch.cuda.synchronize(); t = time.perf_counter()
    X, info, its = J.bicgstab_state(B, rtol=1e-30, maxiter=iters, check_every=iters)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    print("n=%d  %d iterations (info %d): %.3f ms/iteration, %.1f GB/s at 192 B/unknown" %
          (n, its, info, dt / its * 1e3, 192.0 * ctx.n_state / (dt / its) / 1e9))
