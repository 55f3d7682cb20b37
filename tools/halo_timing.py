"""torchrun worker: time the two halo-exchange methods (GPU box, N ranks)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from pyellipticfd_b200.multigpu import halo_exchange
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
nloc, H, pitch = 4095, 4, 4096
view = torch.randn(nloc + 2 * H, pitch, dtype=torch.float64, device=dev)
for method in ("p2p", "allgather"):
    for _ in range(10):
        halo_exchange(view, nloc, H, rank, world, method=method)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t = time.perf_counter(); e0.record()
    for _ in range(100):
        halo_exchange(view, nloc, H, rank, world, method=method)
    e1.record(); torch.cuda.synchronize()
    if rank == 0:
        print("%-10s device %.1f us/exchange, host %.1f us" % (method, e0.elapsed_time(e1) * 10, (time.perf_counter() - t) * 1e4), flush=True)
dist.destroy_process_group()
