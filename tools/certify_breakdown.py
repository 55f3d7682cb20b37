"""Where the time of the certifying Newton iteration goes (4097^2 lattice): residual + policy,
BiCGStab, update.   python tools/certify_breakdown.py [N]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from pyellipticfd_b200 import convex_envelope
from pyellipticfd_b200._context import GridContext
from pyellipticfd_b200.linop import PolicyJacobian
from pyellipticfd_b200.pointclasses import FDRegularGrid

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4095
G = FDRegularGrid([n, n], np.array([[-1.0, 1.0], [-1.0, 1.0]]).T, 4, interpolation=False)
g = torch.from_numpy(bench.two_cones(G.points)).cuda()
ctx = GridContext.get(G)
gs = ctx.pack(g)
Us, d, k = ctx.ce_sweep(gs, 1e-14, 200, relax_every=4)
torch.cuda.synchronize()


def T(f):
    torch.cuda.synchronize()
    t = time.perf_counter()
    r = f()
    torch.cuda.synchronize()
    return r, time.perf_counter() - t


out = {"sweeps": k}
for rep in range(2):
    (F, P, red), t_res = T(lambda: ctx.ce_residual(Us, gs, policy=True))
    Jac = PolicyJacobian(ctx, P, None, cmin=-1.0)
    (D, info, kit), t_kry = T(lambda: Jac.bicgstab_state(-F, rtol=1e-6, check_every=32))
    (JD), t_mv = T(lambda: Jac.matvec_state(D))
    (upd), t_up = T(lambda: ctx.newton_update(Us, D, F, JD))
    (_), t_full = T(lambda: convex_envelope.solve(G, g, U0=ctx.unpack(Us), fdmethod="grid", solver="newton"))
    out["rep%d" % rep] = dict(residual_s=t_res, bicgstab_s=t_kry, krylov_iters=kit, info=info,
                              ms_per_krylov_iter=1e3 * t_kry / max(kit, 1), matvec_s=t_mv,
                              update_s=t_up, solve_call_s=t_full, active_fraction=float((P != 0xFFFF).float().mean()))
    for ce in (64, 128):
        (D, info, kit), t_kry = T(lambda: Jac.bicgstab_state(-F, rtol=1e-6, check_every=ce))
        out["rep%d" % rep]["bicgstab_s_check_every_%d" % ce] = t_kry
print(json.dumps(out))
