"""CPU: evidence for bench.py's statement that BASELINE configs[2] (Pucci, radius-6 stencil, Newton)
has no converged solve to time -- the reference's own Newton loop (solvers.py:168-233 on
pucci.py:105-160, restated in oracle/solvers_oracle.py and pinned against the unmodified
reference on the solve fixtures) does not settle for wide stencils: with the reference's defaults
it hits its 100-iteration cap on a 65^2 lattice at r = 4 with a residual of order 1, while the
same problem converges in a handful of iterations at r = 2.  (GPU runs of the same loop:
profiles/r02_pucci_newton_scaling.jsonl -- 257^2, r = 6: 101 iterations, residual 3.0.)"""
import warnings

import numpy as np

from oracle import solvers_oracle as SO
from pyellipticfd_b200.pointclasses import FDRegularGrid


def exact(X):
    return X[:, 0] ** 2 + 2 * X[:, 1] ** 2            # 2 lambda_max + lambda_min = 2*4 + 2 = 10


def run(n, r):
    G = FDRegularGrid([n, n], np.array([[0.0, 1.0], [0.0, 1.0]]).T, r, interpolation=False)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with np.errstate(all="ignore"):
            U, diff, it, _ = SO.pucci_solve(G, lambda X: 10 + 0 * X[:, 0], exact, (2, 1))
    N = G.num_interior
    Fv = np.concatenate([np.full(N, 10.0), exact(G.points[N:])])
    R, _ = SO.pucci_residual(G, U, Fv, (2, 1), False)
    return it, diff, float(np.abs(R).max())


def test_reference_newton_converges_for_the_narrow_stencil():
    it, diff, res = run(63, 2)
    assert it < 15 and diff < 1e-6 and res < 1e-6


def test_reference_newton_does_not_settle_for_wide_stencils():
    it, diff, res = run(63, 4)
    assert it == 101 and res > 0.1          # "Maximum iterations reached", residual O(1)
