"""Time pucci.solve (Newton, matrix-free BiCGStab) on growing regular grids (GPU).
    python tools/pucci_scaling.py R N1 N2 ...   (Dirichlet problem with u* = x^2 + 2 y^2:
    A = (2, 1) -> 2*lambda_max + lambda_min = 2*4 + 2 = 10)"""
import sys, os, time, json, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyellipticfd_b200 import pucci
from pyellipticfd_b200.pointclasses import FDRegularGrid

r = int(sys.argv[1]) if len(sys.argv) > 1 else 6
sizes = [int(s) for s in sys.argv[2:]] or [255, 511, 1023]
exact = lambda X: X[:, 0] ** 2 + 2 * X[:, 1] ** 2
for n in sizes:
    t = time.time()
    G = FDRegularGrid([n, n], np.array([[0.0, 1.0], [0.0, 1.0]]).T, r, interpolation=False)
    tg = time.time() - t
    st = {}
    tr = []
    g = torch.from_numpy(exact(G.bdry_points)).cuda()
    f = torch.full((G.num_interior,), 10.0, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    t = time.time()
    with warnings.catch_warnings(record=True) as wl:
        warnings.simplefilter("always")
        U, diff, it, _ = pucci._solve_grid(G, f, g, (2, 1), None, "newton", stats=st, trace=tr)
    torch.cuda.synchronize()
    dt = time.time() - t
    xs = torch.arange(1, n + 1, device="cuda", dtype=torch.float64) / (n + 1)
    Ue = (xs[:, None] ** 2 + 2 * xs[None, :] ** 2).reshape(-1)
    err = float((U[:G.num_interior] - Ue).abs().max())
    print(json.dumps(dict(n=n, r=r, D=int(G.stencil_pairs.shape[0]), grid_s=round(tg, 2),
                          solve_s=round(dt, 3), newton=it, diff=diff, err_vs_exact=err,
                          warnings=sorted({str(w.message) for w in wl}),
                          krylov_per_newton=[k for _, _, k, _ in tr][:12], **st)), flush=True)
