"""Matrix-free stand-ins for the SciPy CSR matrices the reference returns.

The reference hands back ``scipy.sparse.csr_matrix`` objects: the selected-pair
finite-difference matrices ``M_min`` / ``M_max`` of shape (num_interior,
num_points) (ddg.py:313-323) and the square Jacobians of
``convex_envelope.operator`` (convex_envelope.py:48-54) and ``pucci.solve``
(pucci.py:136-139).  Everything the solvers and the reference tests do with
them is ``.dot(x)``, ``.transpose().dot(x)``, ``.diagonal()`` and ``.shape``
(solvers.py:179-192, tests/test_ddg.py:75-78).  These classes offer exactly that,
backed by the per-point policy (two bytes per point) on the device, plus
``.tocsr()`` for small grids.
"""
import numpy as np
import scipy.sparse as sp
import torch

from ._context import to_device, like_input


class PolicyMatrix(object):
    """``coef * M_policy`` with M the (num_interior, num_points) matrix whose row i
    is the 3-point row of the pair selected at interior point i."""

    def __init__(self, ctx, policy, coef=1.0, transposed=False):
        self.ctx = ctx
        self.policy = policy            # state-layout uint16 tensor on the device
        self.coef = coef
        self.transposed = transposed

    @property
    def shape(self):
        s = (self.ctx.n_interior_local, self.ctx.n_flat())
        return s[::-1] if self.transposed else s

    def __neg__(self):
        return PolicyMatrix(self.ctx, self.policy, -self.coef, self.transposed)

    def __mul__(self, a):
        return PolicyMatrix(self.ctx, self.policy, self.coef * a, self.transposed)

    __rmul__ = __mul__

    def transpose(self):
        return PolicyMatrix(self.ctx, self.policy, self.coef, not self.transposed)

    T = property(transpose)

    def dot(self, x):
        ctx = self.ctx
        if not self.transposed:
            X = ctx.pack(x)
            Y = ctx.apply_policy(self.policy, X)
            y = ctx.unpad(Y)
        else:
            # M^T z: z over interior points -> all points
            z = to_device(x, ctx.device)
            Z = ctx.pack(torch.cat([z, torch.zeros(ctx.nb, dtype=z.dtype, device=z.device)]))
            J = ctx.jac_struct(self.policy)
            Y = ctx.jac_matvec(J, Z, transpose=True)
            y = ctx.unpack(Y)      # wall rows of J are identity on a zero vector: no contribution
        if self.coef != 1.0:
            y = y * self.coef
        return like_input(x, y)

    __matmul__ = dot

    def tocsr(self):
        """Explicit SciPy CSR, entries exactly as ddg.py:313-323 builds them."""
        ctx = self.ctx
        pol = ctx.unpad_policy(self.policy).cpu().numpy()
        J0, J1, W = ctx.policy_rows(pol)
        N = ctx.n_interior_local
        I = np.arange(N)
        w0 = W[:, 2]
        centre = (-W[:, 0]) + (-W[:, 1])
        M = sp.coo_matrix((np.concatenate([w0 * W[:, 0], w0 * W[:, 1], w0 * centre]) * self.coef,
                           (np.concatenate([I, I, I]), np.concatenate([J0, J1, I]))),
                          shape=(N, ctx.n_flat())).tocsr()
        return M.transpose().tocsr() if self.transposed else M


class PolicyJacobian(object):
    """Square Jacobian over all points:
    interior row i = cmax_i * M_pmax[i] + cmin_i * M_pmin[i]  (or e_i when the policy
    marks the row inactive), wall rows = identity."""

    def __init__(self, ctx, pmin, pmax=None, cmin=1.0, cmax=0.0, transposed=False):
        self.ctx = ctx
        self.pmin, self.pmax, self.cmin, self.cmax = pmin, pmax, cmin, cmax
        self.transposed = transposed
        self._J = ctx.jac_struct(pmin, pmax, cmin, cmax)

    @property
    def shape(self):
        n = self.ctx.n_flat()
        return (n, n)

    def transpose(self):
        return PolicyJacobian(self.ctx, self.pmin, self.pmax, self.cmin, self.cmax,
                              not self.transposed)

    T = property(transpose)

    # state-layout fast paths used by the solvers
    def matvec_state(self, X, out=None):
        return self.ctx.jac_matvec(self._J, X, transpose=self.transposed, out=out)

    def dot(self, x):
        ctx = self.ctx
        Y = self.matvec_state(ctx.pack(x))
        return like_input(x, ctx.unpack(Y))

    __matmul__ = dot

    def diagonal(self):
        return self.ctx.unpack(self.ctx.jac_diagonal(self._J))

    def bicgstab_state(self, B, rtol, atol=0.0, maxiter=0, check_every=32):
        """Jacobi-preconditioned BiCGStab (solvers.py:179-181) on state tensors."""
        if self.transposed:
            raise NotImplementedError("Krylov solve with the transposed Jacobian")
        return self.ctx.bicgstab(self._J, B, rtol, atol, maxiter, check_every)

    def tocsr(self):
        ctx = self.ctx
        N, n = ctx.n_interior_local, ctx.n_flat()
        pm = ctx.unpad_policy(self.pmin).cpu().numpy()
        active = pm != 0xFFFF

        def coefs(c):
            if isinstance(c, torch.Tensor):
                return ctx.unpad(c).cpu().numpy()
            return np.full(N, float(c))

        def rows(pol, c):
            pol = np.where(active, pol, 0)
            J0, J1, W = ctx.policy_rows(pol)
            I = np.arange(N)
            w0 = W[:, 2]
            centre = (-W[:, 0]) + (-W[:, 1])
            v = np.concatenate([w0 * W[:, 0], w0 * W[:, 1], w0 * centre]) * np.tile(c, 3)
            m = np.tile(active, 3)
            return sp.coo_matrix((v[m], (np.concatenate([I, I, I])[m],
                                         np.concatenate([J0, J1, I])[m])), shape=(n, n))

        M = rows(pm, coefs(self.cmin))
        if self.pmax is not None:
            M = M + rows(ctx.unpad_policy(self.pmax).cpu().numpy(), coefs(self.cmax))
        ident = np.ones(n)
        ident[:N][active] = 0.0
        M = (M + sp.diags(ident)).tocsr()
        return M.transpose().tocsr() if self.transposed else M
