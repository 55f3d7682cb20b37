"""Probe of torch symmetric memory (peer-mapped buffers over NVLink) on N GPUs of one box:
remote halo push + signal latency against the NCCL all-gather it would replace.
   torchrun --nproc-per-node 2 tools/symm_probe.py"""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dev = torch.device("cuda", torch.cuda.current_device())
dist.init_process_group("nccl", device_id=dev)
out = {"rank": rank, "world": world}
try:
    import torch.distributed._symmetric_memory as symm
    pitch, H = 4096, 4
    n = (4095 + 2 * H) * pitch
    S = symm.empty(n, dtype=torch.float64, device=dev)
    hdl = symm.rendezvous(S, dist.group.WORLD)
    out["backend"] = str(symm.get_backend(dev)) if hasattr(symm, "get_backend") else "?"
    out["multicast"] = bool(hdl.has_multicast_support)
    out["ptrs"] = len(hdl.buffer_ptrs)
    S.fill_(float(rank))
    dist.barrier()
    lo, hi = (rank - 1) % world, (rank + 1) % world
    view = S.view(-1, pitch)
    peer_lo = hdl.get_buffer(lo, (n,), torch.float64).view(-1, pitch)
    peer_hi = hdl.get_buffer(hi, (n,), torch.float64).view(-1, pitch)

    def push(step):
        # my last H owned rows -> next rank's leading halo; my first H owned rows -> previous rank's trailing halo
        peer_hi[0:H].copy_(view[4095:4095 + H])
        peer_lo[4095 + H:4095 + 2 * H].copy_(view[H:2 * H])
        hdl.put_signal(hi, 0)
        hdl.put_signal(lo, 1)
        hdl.wait_signal(lo, 0)
        hdl.wait_signal(hi, 1)

    push(0)
    torch.cuda.synchronize()
    dist.barrier()
    ok = bool((view[0:H] == float(lo)).all()) and bool((view[4095 + H:] == float(hi)).all())
    out["push_ok"] = ok

    def timed(fn, reps=200):
        for _ in range(10):
            fn(0)
        torch.cuda.synchronize(); dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for k in range(reps):
            fn(k)
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps * 1e3

    out["push_signal_us"] = timed(push)
    send = torch.empty(2 * H, pitch, dtype=torch.float64, device=dev)
    got = torch.empty(world, 2 * H, pitch, dtype=torch.float64, device=dev)

    def ag(step):
        torch.cat([view[H:2 * H], view[4095:4095 + H]], out=send)
        dist.all_gather_into_tensor(got, send)
        view[0:H].copy_(got[lo, H:2 * H])
        view[4095 + H:].copy_(got[hi, 0:H])

    out["allgather_us"] = timed(ag)
    out["barrier_us"] = timed(lambda k: hdl.barrier(channel=2))
    out["signal_pad_size"] = int(hdl.signal_pad_size)
except Exception as e:
    import traceback
    out["error"] = traceback.format_exc()[-1500:]
print(json.dumps(out), flush=True)
dist.destroy_process_group()
