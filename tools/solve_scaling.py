"""Time convex_envelope.solve (Newton / policy iteration) on growing grids (GPU)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyellipticfd_b200 import convex_envelope
from pyellipticfd_b200.pointclasses import FDRegularGrid
from oracle.make_golden import two_cones

r = int(sys.argv[1]) if len(sys.argv) > 1 else 2
KW = {}
if os.environ.get("KRYLOV_RTOL"):
    KW["krylov_rtol"] = float(os.environ["KRYLOV_RTOL"])
if os.environ.get("KRYLOV_RESTARTS"):
    KW["krylov_restarts"] = int(os.environ["KRYLOV_RESTARTS"])
if os.environ.get("MAX_ITERS"):          # the reference's cap is 100 (solvers.py:100)
    KW["max_iters"] = int(os.environ["MAX_ITERS"])
sizes = [int(s) for s in sys.argv[2:]] or [63, 127, 255, 511]
for n in sizes:
    t = time.time()
    G = FDRegularGrid([n, n], np.array([[-1.0, 1.0], [-1.0, 1.0]]).T, r, interpolation=False)
    tg = time.time() - t
    g = two_cones(G.points)
    st = {}
    torch.cuda.synchronize()
    t = time.time()
    tr = []
    if os.environ.get("TRACE") == "live":
        t00 = time.time()
        tr = lambda i, fmax, kit, info: print("  it %d  max|F| %.3e  krylov %d  info %d  t %.1fs"
                                              % (i, fmax, kit, info, time.time() - t00), flush=True)
    U, diff, it, _ = convex_envelope._solve_grid(G, g, None, "newton", stats=st, trace=tr, **KW)
    if os.environ.get("TRACE") and not callable(tr):
        print([(a, "%.2e" % b, c, d) for a, b, c, d in tr][:60])
    torch.cuda.synchronize()
    dt = time.time() - t
    print(json.dumps(dict(n=n, r=r, grid_s=round(tg, 2), solve_s=round(dt, 3), newton=it, diff=diff, **st)), flush=True)
