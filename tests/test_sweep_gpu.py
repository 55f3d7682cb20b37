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
