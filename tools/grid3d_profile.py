"""ddg.d2min / d2eigs on an n^3 lattice, r = 2, through the structured 3-D kernel (the program
ncu profiles).  python tools/grid3d_profile.py [n] [reps]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from pyellipticfd_b200._context import GridContext
from pyellipticfd_b200.pointclasses import FDRegularGrid

n = int(sys.argv[1]) if len(sys.argv) > 1 else 129
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
k = int(np.ceil(np.log2(n + 1)))
t = time.time()
G = FDRegularGrid([n] * 3, np.array([[0.0, (n + 1) / 2.0 ** k]] * 3).T, 2, interpolation=False)
ctx = GridContext.get(G)
print("build", time.time() - t, "band", G.band_index.size, "rows", G.band_pairs.shape[0])
S = ctx.pack(torch.randn(G.num_points, dtype=torch.float64, device="cuda"))
lmin = ctx.zeros()
out = (lmin, None, None, None)
for _ in range(reps):
    ctx.d2eigs(S, want_min=True, want_max=False, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    ctx.d2eigs(S, want_min=True, want_max=False, out=out)
e1.record()
torch.cuda.synchronize()
print("d2min ms", e0.elapsed_time(e1) / 20)
