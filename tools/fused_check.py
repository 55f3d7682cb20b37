"""Fused vs plain s/t pass of the device BiCGStab: x after k iterations (run twice, with
PDE_JAC_FUSED=0 and =1; the second run compares with the first run's file)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyellipticfd_b200.pointclasses import FDRegularGrid
from pyellipticfd_b200._context import GridContext
from pyellipticfd_b200.linop import PolicyJacobian
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
G = FDRegularGrid([n, n + 37], np.array([[0.0, (n + 1) / 512.0], [0.0, (n + 38) / 512.0]]).T, 4, interpolation=False)
ctx = GridContext.get(G)
rng = np.random.default_rng(0)
X = G.points
g = np.minimum(np.linalg.norm(X - 0.2, axis=1), np.linalg.norm(X - 0.45, axis=1))
assert ctx.structured
U = ctx.pack(torch.from_numpy(g - 0.05 * np.exp(-40 * ((X - 0.3) ** 2).sum(1))).cuda())
gs = ctx.pack(torch.from_numpy(g).cuda())
F, P, red = ctx.ce_residual(U, gs, policy=True)
J = PolicyJacobian(ctx, P, None, cmin=-1.0)
B = -F
out = {}
for k in (1, 2, 3, 8, 40):
    x, info, its = J.bicgstab_state(B, rtol=1e-30, maxiter=k, check_every=k)
    out[k] = ctx.unpack(x).cpu().numpy()
    print(k, info, its, float(np.abs(out[k]).max()))
x, info, its = J.bicgstab_state(B, rtol=1e-8)
print("solve", info, its, float(torch.linalg.vector_norm(J.matvec_state(x) - B)) / float(torch.linalg.vector_norm(B)))
tag = os.environ.get("PDE_JAC_FUSED", "1")
np.savez("gpurun_out/fused_check_%s.npz" % tag, **{str(k): v for k, v in out.items()})
if tag == "1" and os.path.exists("gpurun_out/fused_check_0.npz"):
    z = np.load("gpurun_out/fused_check_0.npz")
    for k in out:
        d = np.abs(out[k] - z[str(k)])
        i = int(np.argmax(d))
        print("k", k, "max |fused - plain|", d.max(), "rel", d.max() / np.abs(z[str(k)]).max(), "at", i, "interior" if i < G.num_interior else "wall", divmod(i, n + 37) if i < G.num_interior else "")
