import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29577")
import numpy as np, torch, torch.distributed as dist
from pyellipticfd_b200.multigpu import ShardedGrid
from pyellipticfd_b200.linop import PolicyJacobian
dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
n = 4095
sh = ShardedGrid([n, n], np.array([[0.0, 1.0], [0.0, 1.0]]).T, 4, 0, 1, torch.device("cuda", 0))
ctx = sh.ctx
flat = sh.local_flat(lambda P: torch.sin(3 * P[:, 0]) * torch.cos(2 * P[:, 1]) + P[:, 0] * P[:, 1])
S = ctx.pack(flat)
g = S + 1.0; g[ctx.n_lattice:] = S[ctx.n_lattice:]
F, Pl, _ = ctx.ce_residual(S, g, policy=True)
print("active fraction", float((Pl.view(ctx.rows, ctx.pitch)[:, :n].to(torch.int32) != 65535).double().mean()))
J = PolicyJacobian(ctx, Pl, None, cmin=-1.0)
B = -F
for name, fn in (("single", lambda it: J.bicgstab_state(B, rtol=1e-300, maxiter=it, check_every=it)),
                 ("stepwise", lambda it: sh.bicgstab(J, B, 1e-300, maxiter=it, check_every=it))):
    fn(4)
    for iters in (64, 128):
        torch.cuda.synchronize(); t = time.perf_counter()
        x, info, its = fn(iters)
        torch.cuda.synchronize(); dt = time.perf_counter() - t
        print("%-9s iters %4d done %4d info %d: %.3f ms/iter" % (name, iters, its, info, dt / max(its, 1) * 1e3), flush=True)
dist.destroy_process_group()
