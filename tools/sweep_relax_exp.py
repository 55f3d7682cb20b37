"""solver='sweep' on the 4097^2 lattice with different band-relaxation schedules (GPU box)."""
import os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyellipticfd_b200 import convex_envelope
from pyellipticfd_b200._context import GridContext
from pyellipticfd_b200.pointclasses import FDRegularGrid
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4095
G = FDRegularGrid([n, n], np.array([[-1.0, 1.0], [-1.0, 1.0]]).T, 4, interpolation=False)
C = 3 / 7 * np.array([[np.cos(np.pi / 5), np.sin(np.pi / 5)], [-np.cos(np.pi / 5), -np.sin(np.pi / 5)]])
P = torch.from_numpy(G.points).cuda()
g = torch.minimum(((P + torch.tensor(C[0]).cuda()) ** 2).sum(1).sqrt(),
                  ((P + torch.tensor(C[1]).cuda()) ** 2).sum(1).sqrt())
for rep in range(2):
    for re_ in (4, 2, 1):
        st = {}
        torch.cuda.synchronize(); t = time.time()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            U, diff, sweeps, _ = convex_envelope.solve(G, g, fdmethod="grid", solver="sweep", stats=st,
                                                       relax_every=re_)
        torch.cuda.synchronize(); dt = time.time() - t
        print("relax_every", re_, "sweeps", sweeps, "diff %.2e" % diff, "residual %.2e" % st["residual"],
              "%.4f s" % dt, flush=True)
