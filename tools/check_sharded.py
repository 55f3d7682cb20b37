"""torchrun worker: slab-sharded ddg.d2eigs and Euler solve against the oracle / one GPU.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29533 tools/check_sharded.py
Prints one line per check on rank 0; exit code 1 on any mismatch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from pyellipticfd_b200.pointclasses import FDRegularGrid
from pyellipticfd_b200.multigpu import ShardedGrid
from pyellipticfd_b200._context import GridContext
from pyellipticfd_b200 import convex_envelope
from oracle import c_oracle
from oracle.make_golden import two_cones

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ok_all = True

def report(name, ok):
    global ok_all
    t = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    ok_all &= bool(t.item() == 1.0)
    if rank == 0:
        print("%-50s %s" % (name, "OK" if t.item() == 1.0 else "MISMATCH"), flush=True)

for shape, bounds, r in (((515, 300), [[0, 129.0 / 64], [0, 301.0 / 256]], 4),
                         ((255, 255), [[-1, 1], [-1, 1]], 2),
                         ((300, 200), [[0, 301.0 / 64], [0, 201.0 / 64]], 6)):
    G = FDRegularGrid(list(shape), np.array(bounds, dtype=float).T, r, interpolation=False)
    assert G.uniform_weights()
    sh = ShardedGrid(shape, None, r, rank, world, dev, grid=G)
    c = sh.ctx
    u = np.random.default_rng(0).standard_normal(G.num_points)
    N, ny = G.num_interior, shape[1]
    lo, hi = c.x_offset, c.x_offset + c.nx_local
    flat = np.concatenate([u[lo * ny:hi * ny], u[N:]])
    S = c.pack(flat)
    lmin, lmax, pmin, pmax = c.zeros(), c.zeros(), c.empty_policy(), c.empty_policy()
    sh.d2eigs(S, (lmin, lmax, pmin, pmax))
    omin, omax, rmin, rmax = c_oracle.pairs_d2eigs(G.points, G.pairs, u, N)
    a, b = c.unpad(lmin).cpu().numpy(), c.unpad(lmax).cpu().numpy()
    report("d2eigs %sx%s r=%d on %d ranks (bit-exact)" % (shape[0], shape[1], r, world),
           np.array_equal(a, omin[lo * ny:hi * ny]) and np.array_equal(b, omax[lo * ny:hi * ny]))

# Euler solve: sharded loop vs the single-GPU device loop (every rank runs the latter too)
G = FDRegularGrid([127, 127], np.array([[-1.0, 1.0], [-1.0, 1.0]]).T, 2, interpolation=False)
g = two_cones(G.points)
import warnings
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    Uref, dref, itref, _ = convex_envelope.solve(G, g, fdmethod="grid", solver="euler", max_iters=3000)
sh = ShardedGrid((127, 127), None, 2, rank, world, dev, grid=G)
c = sh.ctx
N, ny = G.num_interior, 127
lo, hi = c.x_offset, c.x_offset + c.nx_local
gs = c.pack(np.concatenate([g[lo * ny:hi * ny], g[N:]]))
U, diff, it, fmax = sh.ce_euler(gs.clone(), gs, convex_envelope.cfl_dt(G), max_iters=3000, check_every=64)
mine = c.unpack(U).cpu().numpy()
ref_local = np.concatenate([Uref[lo * ny:hi * ny], Uref[N:]])
report("Euler 127^2 r=2: %d its (1 GPU %d), bit-exact U" % (it, itref),
       it == itref and diff == dref and np.array_equal(mine, ref_local))
# Newton / policy iteration with the slab-parallel BiCGStab vs the single-GPU solve
Uref, dref, itref, _ = convex_envelope.solve(G, g, fdmethod="grid", solver="newton")
st = {}
U, diff, it = sh.ce_newton(gs, stats=st)
mine = c.unpack(U).cpu().numpy()
ref_local = np.concatenate([Uref[lo * ny:hi * ny], Uref[N:]])
err = float(np.abs(mine - ref_local).max())
report("Newton 127^2 r=2: %d its / %d Krylov (1 GPU %d its), |dU| %.1e" %
       (it, st["krylov_iters"], itref, err), err < 2e-6 and it < 100 and itref < 100)
# (the Krylov solves stop at rtol 1e-6 and their dot products are summed in a different order on
# every partition, so the active sets of the policy iteration may flip at different steps: the
# iteration COUNT is not reproducible across partitions, the converged solution is)
# point cloud sharded by point range: ddi / ddf operators on every rank's slice, re-assembled
# by one all-gather, against the single-GPU table
from scipy.spatial import Delaunay
from oracle import disc
from pyellipticfd_b200.pointclasses import FDTriMesh
from pyellipticfd_b200.multigpu import ShardedCloud
from pyellipticfd_b200 import ddi, ddf
h0 = 0.02
theta = np.pi * h0 ** (1 / 3)
p, nint = disc.disc_points(h0, theta)
Gc = FDTriMesh(p, Delaunay(p).simplices, num_interior=nint, angular_resolution=theta)
uc = torch.from_numpy(np.random.default_rng(3).standard_normal(len(p))).to(dev)
for name, mod in (("ddi", ddi), ("ddf", ddf)):
    sc = ShardedCloud(Gc, rank, world, dev, scheme=name)
    lo_all, hi_all = sc.d2eigs_all(uc)
    Gc._d2cache = None
    (rmin, _, _), (rmax, _, _) = mod.d2eigs(Gc, uc)
    report("%s.d2eigs on %d points, %d point ranges (bit-exact vs 1 GPU)" % (name, len(p), world),
           bool(torch.equal(lo_all, rmin)) and bool(torch.equal(hi_all, rmax))
           and lo_all.shape[0] == nint)
# sharded point-cloud SOLVES (round 2): convex envelope of the two cones of
# examples/convex_envelope.py:8-18 and a Pucci problem on the disc, policy iteration + Euler on
# the sharded table against the single-GPU solve
from pyellipticfd_b200 import convex_envelope, pucci
import time
P = torch.from_numpy(p).to(dev)
cen = 3 / 7 * np.array([[np.cos(np.pi / 5), np.sin(np.pi / 5)], [-np.cos(np.pi / 5), -np.sin(np.pi / 5)]])
gc = torch.minimum(((P - torch.tensor(cen[0], device=dev)) ** 2).sum(1).sqrt(),
                   ((P - torch.tensor(cen[1], device=dev)) ** 2).sum(1).sqrt())
sc = ShardedCloud(Gc, rank, world, dev, scheme="ddi")
for solver, kw in (("newton", dict(solution_tol=1e-10)), ("euler", dict(max_iters=300))):
    st = {}
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        t = time.time()
        Us, diff, it, _ = sc.ce_solve(gc, solver=solver, stats=st, **kw)
        torch.cuda.synchronize()
        ts = time.time() - t
        Gc._d2cache = None
        t = time.time()
        U1, diff1, it1, _ = convex_envelope.solve(Gc, gc, fdmethod="interpolate", solver=solver, **kw)
        torch.cuda.synchronize()
        t1 = time.time() - t
    err = float((Us - U1).abs().max())
    # identical on every rank: replicated vectors, no reduction
    chk = torch.tensor([float(Us.sum()), float(it)], device=dev, dtype=torch.float64)
    lo_, hi_ = chk.clone(), chk.clone()
    dist.all_reduce(lo_, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi_, op=dist.ReduceOp.MAX)
    report("sharded cloud convex_envelope.solve(%s): %d its (1 GPU: %d), |U - U_1gpu| = %.2e, "
           "%.2f s vs %.2f s" % (solver, it, it1, err, ts, t1),
           err < (1e-8 if solver == "newton" else 1e-12) and abs(it - it1) <= 2
           and bool(torch.equal(lo_, hi_)))
fP = 2.0 * torch.ones(nint, device=dev, dtype=torch.float64)
gP = (P[nint:] ** 2).sum(1)
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    Us, diff, it, _ = sc.pucci_solve(fP, gP, (1.0, 1.0), solver="newton")
    Gc._d2cache = None
    U1, diff1, it1, _ = pucci.solve(Gc, fP, gP, (1.0, 1.0), fdmethod="interpolate", solver="newton")
err = float((Us - U1).abs().max())
report("sharded cloud pucci.solve(newton): %d its (1 GPU: %d), |U - U_1gpu| = %.2e" % (it, it1, err),
       err < 1e-6 and abs(it - it1) <= 2)
dist.destroy_process_group()
sys.exit(0 if ok_all else 1)
