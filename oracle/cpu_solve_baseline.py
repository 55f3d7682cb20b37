"""CPU baseline of ONE policy-iteration (Newton) step of convex_envelope.solve at full size.

TEST / BENCH INFRASTRUCTURE.  The reference's Newton loop cannot be run at
4097^2 (its constructor is O(N^2) and `Jac[rows,:] = ...` is super-linear), so its cost is
measured component-wise on the host, with the reference's own building blocks, and
multiplied by the iteration counts of the GPU run:

  t_op     ddg.d2min(G, U, jacobian=True) of the NumPy port (ddg.py:251-327) on a bounded
           sample of deep rows, extrapolated linearly to all interior points (the operator
           is a flat gather + lexsort: linear in the number of pairs);
  t_jac    assembling the CSR Jacobian of a policy at FULL size (identity rows + 3-nnz
           rows, convex_envelope.py:48-54) by direct COO->CSR construction -- the same
           matrix as the reference's row assignment, built the cheapest way SciPy offers;
  t_kit    one iteration of scipy.sparse.linalg.bicgstab with the Jacobi preconditioner
           (solvers.py:179-181) on that full-size matrix.

    python -m oracle.cpu_solve_baseline N R [ROWS] [KRYLOV_ITS]
prints one JSON line.  SciPy's SpMV and NumPy's gathers are single-threaded; `cores` = 1.
"""
import json
import sys
import time

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as sla

from . import ddg_oracle
from .cpu_baseline import lattice_chunk


def main(argv=None):
    from . import grid_oracle
    a = argv or sys.argv[1:]
    n, r = int(a[0]), int(a[1])
    rows = int(a[2]) if len(a) > 2 else 48
    kits = int(a[3]) if len(a) > 3 else 4
    table = grid_oracle.stencil_table(r)       # no product code in the CPU arm
    D = table.shape[0]
    H = int(np.abs(table).max())
    h = 1.0 / (n + 1)

    # t_op: operator + pair Jacobian on a sample of deep rows
    G = lattice_chunk(n, r, table, (h, h), (0.0, 0.0), r + H, rows)
    u = np.random.default_rng(0).standard_normal(G.num_points)
    ddg_oracle.d2min(G, u, jacobian=True, control=False)
    t = time.perf_counter()
    ddg_oracle.d2min(G, u, jacobian=True, control=False)
    t_sample = time.perf_counter() - t
    t_op = t_sample * (n * n) / G.num_interior

    # full-size policy Jacobian: every interior row active with a random pair of the table
    N = n * n
    rng = np.random.default_rng(1)
    ix, iy = np.divmod(np.arange(N, dtype=np.int64), n)
    k = rng.integers(0, D, N)
    a0, b0, a1, b1 = (table[k, c] for c in range(4))
    inside = ((ix + a0 >= 0) & (ix + a0 < n) & (iy + b0 >= 0) & (iy + b0 < n) &
              (ix + a1 >= 0) & (ix + a1 < n) & (iy + b1 >= 0) & (iy + b1 < n))
    active = inside & (rng.random(N) < 0.7)
    I = np.flatnonzero(active)
    J0 = (ix[I] + a0[I]) * n + iy[I] + b0[I]
    J1 = (ix[I] + a1[I]) * n + iy[I] + b1[I]
    hh = h * np.sqrt(a0[I] ** 2.0 + b0[I] ** 2.0)
    w = 1.0 / hh ** 2
    t = time.perf_counter()
    ident = np.ones(N)
    ident[I] = 0.0
    rows_ = np.concatenate([I, I, I, np.arange(N)])
    cols_ = np.concatenate([J0, J1, I, np.arange(N)])
    vals_ = np.concatenate([-w, -w, 2 * w, ident])
    Jac = sp.coo_matrix((vals_, (rows_, cols_)), shape=(N, N)).tocsr()
    Jac.eliminate_zeros()
    t_jac = time.perf_counter() - t

    b = rng.standard_normal(N)
    P = sp.diags(1.0 / Jac.diagonal(), format="csr")
    sla.bicgstab(Jac, b, M=P, rtol=1e-300, maxiter=1)
    t = time.perf_counter()
    sla.bicgstab(Jac, b, M=P, rtol=1e-300, maxiter=kits)
    t_kit = (time.perf_counter() - t) / kits

    print(json.dumps(dict(
        n=n, r=r, D=D, cores=1, t_op=t_op, t_jac=t_jac, t_kit=t_kit,
        sample="operator: %d deep rows x %d cols (%d pairs) of the NumPy port, extrapolated x%.0f; "
               "Jacobian assembly and %d SciPy bicgstab iterations at full size (%d unknowns)"
               % (rows, n - 2 * r, len(G.pairs), (n * n) / G.num_interior, kits, N))))


if __name__ == "__main__":
    main()
