"""Where does the fixed cost of a slab-sharded operator step come from?  Single-GPU emulation of
the call sequence of multigpu.ShardedGrid.overlapped (GPU box, 1 GPU):
  A  one full-slab launch group
  B  interior window + both edge strips (two launch groups, no communication)
  C  B + the communication-stream choreography with local copies instead of the NCCL all-gather
For each: device time per step (CUDA events) and host enqueue time per step (perf_counter
without synchronisation)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyellipticfd_b200.pointclasses import FDRegularGrid
from pyellipticfd_b200._context import GridContext, Slab

n, r = 4095, 4
G = FDRegularGrid([2 * n, n], np.array([[0.0, (2 * n + 1) / 4096.0], [0.0, 1.0]]).T, r, interpolation=False)
ctx = GridContext(G, torch.device("cuda", 0), slab=Slab(0, n, r))      # rank 0 of 2: halo rows exist
S = ctx.zeros(); S.normal_()
out = (ctx.zeros(), None, None, None)
H = r
view = S[:ctx.rows * ctx.pitch].view(ctx.rows, ctx.pitch)
comm = torch.cuda.Stream()

def launch():
    ctx.d2eigs(S, want_min=True, want_max=False, out=out)

def A():
    launch()

def B():
    ctx.set_row_window(H, ctx.nx_local - H); launch()
    ctx.set_row_windows2(0, H, ctx.nx_local - H, ctx.nx_local); launch()
    ctx.set_row_window()

def C():
    main = torch.cuda.current_stream()
    comm.wait_stream(main)
    with torch.cuda.stream(comm):
        send = torch.cat([view[H:2 * H], view[ctx.nx_local:ctx.nx_local + H]])
        got = torch.empty((2,) + tuple(send.shape), dtype=send.dtype, device=send.device)
        got[0].copy_(send); got[1].copy_(send)                     # stands in for the all-gather
        view[ctx.nx_local + H:ctx.nx_local + 2 * H].copy_(got[1, 0:H])
    ctx.set_row_window(H, ctx.nx_local - H); launch()
    main.wait_stream(comm)
    ctx.set_row_windows2(0, H, ctx.nx_local - H, ctx.nx_local); launch()
    ctx.set_row_window()

edge = torch.cuda.Stream()

def D():
    """C with the edge strips on their own stream: they depend on the exchange only, so they can
    fill the tail of the persistent interior kernel instead of waiting for it."""
    main = torch.cuda.current_stream()
    comm.wait_stream(main)
    with torch.cuda.stream(comm):
        send = torch.cat([view[H:2 * H], view[ctx.nx_local:ctx.nx_local + H]])
        got = torch.empty((2,) + tuple(send.shape), dtype=send.dtype, device=send.device)
        got[0].copy_(send); got[1].copy_(send)
        view[ctx.nx_local + H:ctx.nx_local + 2 * H].copy_(got[1, 0:H])
    edge.wait_stream(comm)
    ctx.set_row_window(H, ctx.nx_local - H); launch()
    with torch.cuda.stream(edge):
        ctx.set_row_windows2(0, H, ctx.nx_local - H, ctx.nx_local); launch()
    ctx.set_row_window()
    main.wait_stream(edge)

def E():
    """D with the edge strips enqueued BEFORE the interior kernel (they start as soon as the
    exchange is done; the interior kernel fills the rest of the machine)."""
    main = torch.cuda.current_stream()
    comm.wait_stream(main)
    with torch.cuda.stream(comm):
        send = torch.cat([view[H:2 * H], view[ctx.nx_local:ctx.nx_local + H]])
        got = torch.empty((2,) + tuple(send.shape), dtype=send.dtype, device=send.device)
        got[0].copy_(send); got[1].copy_(send)
        view[ctx.nx_local + H:ctx.nx_local + 2 * H].copy_(got[1, 0:H])
    edge.wait_stream(comm)
    with torch.cuda.stream(edge):
        ctx.set_row_windows2(0, H, ctx.nx_local - H, ctx.nx_local); launch()
    ctx.set_row_window(H, ctx.nx_local - H); launch()
    ctx.set_row_window()
    main.wait_stream(edge)

ref = None
for name, fn in (("A full slab", A), ("B interior + edges", B), ("C B + comm choreography", C),
                 ("D edges on own stream", D), ("E edges first, own stream", E)):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t = time.perf_counter()
    for _ in range(200):
        fn()
    host = (time.perf_counter() - t) / 200
    e1.record()
    torch.cuda.synchronize()
    got_ = out[0].clone()
    if ref is None:
        ref = got_
    print("%-26s same result %s  device %.4f ms/step   host enqueue %.4f ms/step" % (name, bool(torch.equal(ref, got_)), e0.elapsed_time(e1) / 200, host * 1e3), flush=True)
