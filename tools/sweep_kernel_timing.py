"""Time the convex-hull sweep kernel (csrc/ce_sweep.cu) on the GPU box: milliseconds per sweep
(all direction passes + band relaxation) on the first sweeps from g and on the converged state,
for the shared-memory budget given by PDE_SWEEP_SMEM_KB.   python tools/sweep_kernel_timing.py R N ..."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from pyellipticfd_b200._context import GridContext
from pyellipticfd_b200.pointclasses import FDRegularGrid

r = int(sys.argv[1]) if len(sys.argv) > 1 else 4
sizes = [int(s) for s in sys.argv[2:]] or [1023, 4095]
C = 3 / 7 * np.array([[np.cos(np.pi / 5), np.sin(np.pi / 5)], [-np.cos(np.pi / 5), -np.sin(np.pi / 5)]])


def timed(f):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    out = f()
    b.record()
    torch.cuda.synchronize()
    return out, a.elapsed_time(b)


for n in sizes:
    G = FDRegularGrid([n, n], np.array([[-1.0, 1.0], [-1.0, 1.0]]).T, r, interpolation=False)
    P = torch.from_numpy(G.points).cuda()
    g = torch.minimum(((P + torch.tensor(C[0]).cuda()) ** 2).sum(1).sqrt(),
                      ((P + torch.tensor(C[1]).cuda()) ** 2).sum(1).sqrt())
    ctx = GridContext.get(G)
    gs = ctx.pack(g)
    ctx.ce_sweep(gs, 0.0, 0)
    D = ctx._sweep["D"]
    rec = dict(lattice="%d^2" % (n + 2), radius=r, dirs=D, smem_kb=os.environ.get("PDE_SWEEP_SMEM_KB", "110"))
    (U, c, _), ms = timed(lambda: ctx.ce_sweep(gs, 0.0, 1, relax_every=4))
    rec["first_sweep_ms"] = round(ms, 3)
    (U, c, _), ms = timed(lambda: ctx.ce_sweep(gs, 0.0, 4, U=U, relax_every=4))
    rec["sweeps_2_5_ms_each"] = round(ms / 4, 3)
    (U, c, k), ms = timed(lambda: ctx.ce_sweep(gs, 1e-14, 200, U=U, relax_every=4))
    rec["to_1e-14"] = dict(sweeps=k + 5, ms_each=round(ms / max(k, 1), 3), change=c)
    (U, c, _), ms = timed(lambda: ctx.ce_sweep(gs, 0.0, 8, U=U, relax_every=4))
    rec["converged_sweep_ms_each"] = round(ms / 8, 3)
    rec["ms_per_direction_pass_converged"] = round(ms / 8 / D, 4)
    rec["residual"] = float(ctx.ce_residual(U, gs, policy=False)[2][0])
    print(json.dumps(rec), flush=True)
