import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyellipticfd_b200.pointclasses import FDRegularGrid
from pyellipticfd_b200._context import GridContext
from pyellipticfd_b200.linop import PolicyJacobian
n = 4095
G = FDRegularGrid([n, n], np.array([[0.0, 1.0], [0.0, 1.0]]).T, 4, interpolation=False)
ctx = GridContext.get(G)
S = torch.randn(ctx.n_state, dtype=torch.float64, device='cuda')
S.view(-1)[:ctx.n_lattice].view(ctx.rows, ctx.pitch)[:, n:] = 0
g = S + 1.0; g[ctx.n_lattice:] = S[ctx.n_lattice:]
F, P, _ = ctx.ce_residual(S, g, policy=True)
J = PolicyJacobian(ctx, P, None, cmin=-1.0)
B = -F
work = torch.empty(8 * ctx.n_state, dtype=torch.float64, device='cuda')
for iters in (4, 32, 64, 128, 256):
    torch.cuda.synchronize(); t = time.perf_counter()
    x, info, its = ctx.bicgstab(J._J, B, 1e-300, maxiter=iters, check_every=iters, work=work)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    print("iters %4d done %4d info %d: total %.2f ms -> %.3f ms/iter" % (iters, its, info, dt * 1e3, dt / max(its,1) * 1e3), flush=True)
