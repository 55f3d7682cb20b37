"""GPU: convex_envelope.solve(..., solver='sweep') (extension, csrc/ce_sweep.cu) against the
converged solutions of the unmodified reference (tests/golden/solve_*.npz), and as the starting
point of the reference's own Newton iteration."""
import glob
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SOLVE = sorted(glob.glob(os.path.join(GOLD, "solve_*.npz")))


@pytest.mark.parametrize("path", SOLVE, ids=[os.path.basename(p) for p in SOLVE])
def test_sweep_solver_reaches_the_reference_solution(path):
    from pyellipticfd_b200 import convex_envelope
    from pyellipticfd_b200.pointclasses import FDRegularGrid
    z = np.load(path)
    G = FDRegularGrid([int(s) for s in z["shape"]], z["bounds"], int(z["radius"]), interpolation=False)
    g = z["g"]
    U, diff, sweeps, _ = convex_envelope.solve(G, g, fdmethod="grid", solver="sweep")
    assert diff < 1e-6 and sweeps < 200
    assert np.abs(U - z["newton_U"]).max() < 2e-6            # the reference's solver tolerance
    F = convex_envelope.operator(G, U, g, jacobian=False, fdmethod="grid")[0]
    assert np.abs(F).max() < 1e-6 and (U <= g).all()
    # tight tolerances: the discrete solution itself
    U, diff, sweeps, _ = convex_envelope.solve(G, g, fdmethod="grid", solver="sweep",
                                               solution_tol=1e-12, operator_tol=1e-9)
    assert np.abs(U - z["newton_U"]).max() < 1e-8
    # ... from which the reference's Newton iteration has nothing left to do
    Un, dn, itn, _ = convex_envelope.solve(G, g, U0=U, fdmethod="grid", solver="newton")
    assert itn <= 5 and np.abs(Un - z["newton_U"]).max() < 2e-6


@pytest.mark.parametrize("shape,r", [((40, 31), 3), ((70, 130), 4), ((257, 300), 4), ((1100, 64), 6)])
def test_device_sweep_equals_the_host_replay(shape, r):
    """One sweep of k_hull_pass on the device against the thread-by-thread host replay of the same
    phases (pde_ce_sweep, device = -2) and the serial chain (device = -1)."""
    from oracle.make_golden import two_cones
    from pyellipticfd_b200 import _sweep
    from pyellipticfd_b200._context import GridContext
    from pyellipticfd_b200.pointclasses import FDRegularGrid
    nx, ny = shape            # dyadic spacing on both axes: the structured (K1) context
    G = FDRegularGrid(list(shape), np.array([[-1.0, -1.0 + (nx + 1) / 64], [-0.5, -0.5 + (ny + 1) / 64]]).T,
                      r, interpolation=False)
    rng = np.random.default_rng(r)
    g = two_cones(G.points) + 0.02 * rng.standard_normal(G.num_points)
    ctx = GridContext.get(G)
    gs = ctx.pack(g)
    for sweeps in (1, 3):
        Ud, cd, nd = ctx.ce_sweep(gs, 0.0, sweeps, relax_every=4)
        Ud = ctx.unpack(Ud).cpu().numpy()
        for dev in (-2, -1):
            Uh, ch, _ = _sweep.host_sweep(G, g, tol=0.0, max_sweeps=sweeps, relax_every=4, device=dev)
            assert nd == sweeps and np.abs(Ud - Uh).max() < 1e-13, (dev, np.abs(Ud - Uh).max())
            assert abs(cd - ch) < 1e-13
