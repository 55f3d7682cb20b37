"""Time convex_envelope.solve(..., solver='sweep') on growing lattices (GPU box); one JSON line
per size, flushed as soon as it is known.   python tools/sweep_timing.py R N1 N2 ..."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from pyellipticfd_b200 import convex_envelope
from pyellipticfd_b200._context import GridContext
from pyellipticfd_b200.pointclasses import FDRegularGrid

r = int(sys.argv[1]) if len(sys.argv) > 1 else 4
sizes = [int(s) for s in sys.argv[2:]] or [1023, 2047, 4095]
out = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                        "gpurun_out", "sweep_timing.jsonl"), "a") if os.path.isdir("gpurun_out") else sys.stdout
C = 3 / 7 * np.array([[np.cos(np.pi / 5), np.sin(np.pi / 5)], [-np.cos(np.pi / 5), -np.sin(np.pi / 5)]])
for n in sizes:
    t = time.time()
    G = FDRegularGrid([n, n], np.array([[-1.0, 1.0], [-1.0, 1.0]]).T, r, interpolation=False)
    tg = time.time() - t
    P = torch.from_numpy(G.points).cuda()
    g = torch.minimum(((P + torch.tensor(C[0]).cuda()) ** 2).sum(1).sqrt(),
                      ((P + torch.tensor(C[1]).cuda()) ** 2).sum(1).sqrt())
    ctx = GridContext.get(G)
    gs = ctx.pack(g)
    t = time.time()
    ctx.ce_sweep(gs, 0.0, 0)                       # builds and uploads the sweep tables
    torch.cuda.synchronize()
    tt = time.time() - t
    st = {}
    t = time.time()
    U, diff, sweeps, _ = convex_envelope.solve(G, g, fdmethod="grid", solver="sweep", stats=st)
    torch.cuda.synchronize()
    dt = time.time() - t
    print(json.dumps(dict(lattice="%d^2" % (n + 2), radius=r, grid_build_s=round(tg, 3),
                          sweep_tables_s=round(tt, 3), solve_s=round(dt, 4), sweeps=sweeps,
                          diff=diff, residual=st.get("residual"))), file=out, flush=True)
