"""Krylov iteration count of the certifying Newton step on the 4097^2 lattice (sweep solution as U0)
for the launch geometry given by the environment (PDE_ELEM_BLOCKS, PDE_JAC_FUSED)."""
import os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyellipticfd_b200 import convex_envelope
from pyellipticfd_b200.pointclasses import FDRegularGrid
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4095
G = FDRegularGrid([n, n], np.array([[-1.0, 1.0], [-1.0, 1.0]]).T, 4, interpolation=False)
C = 3 / 7 * np.array([[np.cos(np.pi / 5), np.sin(np.pi / 5)], [-np.cos(np.pi / 5), -np.sin(np.pi / 5)]])
P = torch.from_numpy(G.points).cuda()
g = torch.minimum(((P + torch.tensor(C[0]).cuda()) ** 2).sum(1).sqrt(),
                  ((P + torch.tensor(C[1]).cuda()) ** 2).sum(1).sqrt())
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    U, diff, sweeps, _ = convex_envelope.solve(G, g, fdmethod="grid", solver="sweep")
    st = {}
    torch.cuda.synchronize(); t = time.time()
    U2, d2, it, _ = convex_envelope.solve(G, g, U0=U, fdmethod="grid", solver="newton", stats=st)
    torch.cuda.synchronize(); dt = time.time() - t
print("ELEM_BLOCKS", os.environ.get("PDE_ELEM_BLOCKS"), "FUSED", os.environ.get("PDE_JAC_FUSED"),
      "n", n, "newton", it, "krylov", st["krylov_iters"], "fallbacks", st["euler_fallbacks"], "%.3f s" % dt, flush=True)
