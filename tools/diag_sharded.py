import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from pyellipticfd_b200 import multigpu
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n, r = 4095, 4
bounds = np.array([[0.0, (world * n + 1) / (n + 1.0)], [0.0, 1.0]]).T
sh = multigpu.ShardedGrid([world * n, n], bounds, r, rank, world, dev)
c = sh.ctx
print(rank, 'ctx', c.nx_local, c.x_offset, c.halo_rows, c.n_band, c.structured, flush=True)
S = torch.randn(c.n_state, dtype=torch.float64, device=dev)
out = (c.zeros(), None, None, None)
def ev(): e = torch.cuda.Event(enable_timing=True); e.record(); return e
for it in range(3):
    sh.exchange_halo(S); c.d2eigs(S, want_min=True, want_max=False, out=out)
torch.cuda.synchronize(); dist.barrier()
for it in range(4):
    t0 = time.perf_counter()
    a = ev(); sh.exchange_halo(S); b = ev(); c.d2eigs(S, want_min=True, want_max=False, out=out); d = ev()
    th = time.perf_counter() - t0
    torch.cuda.synchronize()
    print(rank, 'exchange %.3f ms  d2eigs %.3f ms  host %.3f ms' % (a.elapsed_time(b), b.elapsed_time(d), th * 1e3), flush=True)
# d2eigs alone, repeated
a = ev()
for _ in range(10): c.d2eigs(S, want_min=True, want_max=False, out=out)
b = ev(); torch.cuda.synchronize()
print(rank, 'd2eigs alone %.3f ms' % (a.elapsed_time(b) / 10), flush=True)
dist.destroy_process_group()
