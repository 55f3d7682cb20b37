"""Time the K1 tuning variants (PDE_K1_VARIANT) of the r = 4 d2min kernel and fingerprint
their output (all variants must produce identical bits).  Run once per variant:
    PDE_K1_VARIANT=6 python tools/k1_variants.py"""
import hashlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyellipticfd_b200.pointclasses import FDRegularGrid
from pyellipticfd_b200._context import GridContext

n = 4095
G = FDRegularGrid([n, n], np.array([[0.0, 1.0], [0.0, 1.0]]).T, 4, interpolation=False)
ctx = GridContext.get(G)
fields = {"random": torch.randn(G.num_points, dtype=torch.float64, device="cuda",
                                generator=torch.Generator("cuda").manual_seed(0))}
x = torch.linspace(0, 1, G.num_points, dtype=torch.float64, device="cuda")
fields["saddle-ish"] = torch.sin(37 * x) - 0.25            # mixed signs, smooth
fields["zeros"] = torch.zeros(G.num_points, dtype=torch.float64, device="cuda")
out = (ctx.zeros(), None, None, None)
for name, f in fields.items():
    S = ctx.pack(f)
    for _ in range(5):
        ctx.d2eigs(S, want_min=True, want_max=False, out=out)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(100):
        ctx.d2eigs(S, want_min=True, want_max=False, out=out)
    b.record(); torch.cuda.synchronize()
    h = hashlib.sha1(out[0].cpu().numpy().tobytes()).hexdigest()[:16]
    print("variant %s %-10s %.4f ms  sha1 %s" % (os.environ.get("PDE_K1_VARIANT", "0"), name,
                                                 a.elapsed_time(b) / 100, h), flush=True)
