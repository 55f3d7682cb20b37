"""Timeline of GridContext.d2eigs_host (H2D / stencil / D2H pipelined over row chunks): CUDA events
at the end of every chunk's upload, kernels and download, printed relative to the first upload."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyellipticfd_b200.pointclasses import FDRegularGrid
from pyellipticfd_b200._context import GridContext, vp, check, _ptr, _stream

n = 4095
G = FDRegularGrid([n, n], np.array([[0.0, 1.0], [0.0, 1.0]]).T, 4, interpolation=False)
ctx = GridContext.get(G)
host = torch.randn(G.num_points, dtype=torch.float64).pin_memory()
out = torch.empty(G.num_interior, dtype=torch.float64, pin_memory=True)
for _ in range(3):
    ctx.d2eigs_host(host, want_min=True, want_max=False, out=(out, None))
torch.cuda.synchronize()
chunk = int(sys.argv[1]) if len(sys.argv) > 1 else 256
S, Lmin, Lmax, s_in, s_out = ctx._stage
main = torch.cuda.current_stream()
edges = list(range(0, n, chunk)) + [n]
nch = len(edges) - 1
ev0 = torch.cuda.Event(enable_timing=True)
e_in = [torch.cuda.Event(enable_timing=True) for _ in range(nch)]
e_k = [torch.cuda.Event(enable_timing=True) for _ in range(nch)]
e_out = [torch.cuda.Event(enable_timing=True) for _ in range(nch)]
torch.cuda.synchronize()
t0 = time.perf_counter()
s_in.wait_stream(main); s_out.wait_stream(main)
with torch.cuda.stream(s_in):
    ev0.record()
    for k in range(nch):
        check(ctx.lib.pde_grid_h2d_rows(ctx.h, vp(host.data_ptr()), _ptr(S), edges[k], edges[k + 1], int(k == 0), _stream()))
        e_in[k].record()
t_up = time.perf_counter() - t0
for k in range(nch):
    lo, hi = edges[k], edges[k + 1]
    main.wait_event(e_in[min(k + 1, nch - 1)])
    ctx.set_row_window(lo, hi)
    ctx.d2eigs(S, want_min=True, want_max=False, out=(Lmin, None, None, None))
    e_k[k].record(main)
    s_out.wait_event(e_k[k])
    with torch.cuda.stream(s_out):
        check(ctx.lib.pde_grid_d2h_rows(ctx.h, _ptr(Lmin), vp(out.data_ptr()), lo, hi, _stream()))
        e_out[k].record()
ctx.set_row_window()
t_host = time.perf_counter() - t0
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
print("host enqueue: uploads %.3f ms, everything %.3f ms; wall %.3f ms" % (t_up * 1e3, t_host * 1e3, t_all * 1e3))
print("chunk  h2d_done  kernel_done  d2h_done   (ms after the first upload started)")
for k in range(nch):
    print("%3d  %8.3f  %8.3f  %8.3f" % (k, ev0.elapsed_time(e_in[k]), ev0.elapsed_time(e_k[k]), ev0.elapsed_time(e_out[k])))
