"""3-D convex envelope on the GPU box: fast constructor, ddg.d2min and Euler / Newton solves of
the two-centre distance obstacle on an n^3 lattice (default 65^3, r = 2).
    python tools/grid3d_demo.py [n] [r]
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from pyellipticfd_b200 import ddg, convex_envelope
from pyellipticfd_b200.pointclasses import FDRegularGrid

n = int(sys.argv[1]) if len(sys.argv) > 1 else 63
r = int(sys.argv[2]) if len(sys.argv) > 2 else 2
t = time.time()
G = FDRegularGrid([n] * 3, np.array([[-1.0, 1.0]] * 3).T, r, interpolation=False)
out = {"lattice": "%d^3" % (n + 2), "radius": r, "interior_points": G.num_interior,
       "wall_points": G.num_bdry, "pair_rows": int(G.num_pairs), "grid_build_s": time.time() - t}
C = 3 / 7 * np.array([[np.cos(np.pi / 5), np.sin(np.pi / 5), 0.3],
                      [-np.cos(np.pi / 5), -np.sin(np.pi / 5), -0.3]])
g = np.min([np.sqrt(((G.points + c) ** 2).sum(1)) for c in C], axis=0)
gd = torch.from_numpy(g).cuda()
ddg.d2min(G, gd)
torch.cuda.synchronize()
t = time.time()
for _ in range(10):
    ddg.d2min(G, gd)
torch.cuda.synchronize()
out["d2min_ms"] = (time.time() - t) / 10 * 1e3
out["d2min_Gpair_per_s"] = G.num_pairs / out["d2min_ms"] / 1e6
for solver in ("euler", "newton"):
    import warnings
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        U, diff, it, sec = convex_envelope.solve(G, gd, fdmethod="grid", solver=solver)
    F = convex_envelope.operator(G, U, gd, jacobian=False, fdmethod="grid")[0]
    out[solver] = {"seconds": sec, "iterations": int(it), "diff": float(diff),
                   "residual": float(F.abs().max()), "warnings": [str(x.message) for x in w],
                   "max_U_minus_g": float((U - gd).max()), "min_U_minus_g": float((U - gd).min())}
print(json.dumps(out))
