"""Two sweeps of the convex-hull kernel on the 4097^2 lattice, r = 4 (the program ncu profiles)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from pyellipticfd_b200._context import GridContext
from pyellipticfd_b200.pointclasses import FDRegularGrid

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4095
G = FDRegularGrid([n, n], np.array([[-1.0, 1.0], [-1.0, 1.0]]).T, 4, interpolation=False)
C = 3 / 7 * np.array([[np.cos(np.pi / 5), np.sin(np.pi / 5)], [-np.cos(np.pi / 5), -np.sin(np.pi / 5)]])
P = torch.from_numpy(G.points).cuda()
g = torch.minimum(((P + torch.tensor(C[0]).cuda()) ** 2).sum(1).sqrt(),
                  ((P + torch.tensor(C[1]).cuda()) ** 2).sum(1).sqrt())
ctx = GridContext.get(G)
gs = ctx.pack(g)
sweeps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
U, c, k = ctx.ce_sweep(gs, 0.0, sweeps, relax_every=4)
torch.cuda.synchronize()
print("change", c, "sweeps", k)
