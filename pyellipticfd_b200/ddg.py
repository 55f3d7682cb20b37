"""Finite differences on regular grids with antipodal stencil pairs -- CUDA.

Drop-in for the second-derivative part of /root/reference/ddg.py: same function
names, positional/keyword arguments, return conventions and exceptions
(ddg.py:177-339).  ``u`` / ``v`` may be NumPy arrays or torch tensors (host or
device); results come back in the container type of ``u``.  The Jacobians are
:class:`~pyellipticfd_b200.linop.PolicyMatrix` objects (``.dot``,
``.transpose()``, ``.tocsr()``) instead of SciPy CSR matrices.

The first-derivative functions ``d1`` / ``d1n`` / ``d1grad`` (ddg.py:8-175) work on the
grid's neighbour list (``_nbops``).
"""
import numpy as np
import torch

from . import _ddutils
from ._context import GridContext, like_input, to_device
from .linop import PolicyMatrix
from ._nbops import d1, d1n, d1grad          # noqa: F401  (ddg.py:8-175)

# 0 = reference operation order (bit-exact values and selection); 1 = re-associated fast mode
ARITHMETIC_MODE = 0


def _pinned_host(u):
    return (isinstance(u, torch.Tensor) and not u.is_cuda and u.is_pinned()
            and u.dtype == torch.float64 and u.is_contiguous())


def _require_pairs(G):
    # structured grids answer without touching `pairs` (a lazily materialised property there)
    if (not hasattr(G, "stencil_pairs") and not hasattr(G, "stencil_pairs_nd")
            and getattr(G, "pairs", None) is None):
        raise TypeError('The grid was has no finite difference pairs. '
                        'Construct the grid with "interpolation=False"')


def d2eigs(G, u, jacobian=False, control=False):
    """Smallest and largest directional second derivative over the stencil pairs.

    Same contract as ddg.d2eigs (ddg.py:251-327):
    ``((lambda_min, M_min, v_min), (lambda_max, M_max, v_max))`` with values over
    the interior points; ``M_*`` and ``v_*`` are None unless requested.  Tie rule:
    min -> last pair attaining it, max -> first, in the grid's pair order.
    """
    _require_pairs(G)
    ctx = GridContext.get(G)
    need_policy = jacobian or control
    if _pinned_host(u) and not need_policy and ctx.structured and not ctx.nd and ctx.nx_local >= 1024:
        lo, hi = ctx.d2eigs_host(u, mode=ARITHMETIC_MODE)      # pipelined over row chunks
        return (lo, None, None), (hi, None, None)
    S = ctx.pack(u)
    lmin, lmax, pmin, pmax = ctx.d2eigs(S, policy=need_policy, mode=ARITHMETIC_MODE)
    lam_min = like_input(u, ctx.unpad(lmin))
    lam_max = like_input(u, ctx.unpad(lmax))
    M_min = M_max = v_min = v_max = None
    if jacobian:
        M_min, M_max = PolicyMatrix(ctx, pmin), PolicyMatrix(ctx, pmax)
    if control:
        if getattr(G, "dim", 2) == 2:
            vmn, vmx = ctx.control(pmin, pmax)
        else:
            vmn, vmx = _control_nd(ctx, pmin, pmax)
        v_min, v_max = like_input(u, vmn), like_input(u, vmx)
    return (lam_min, M_min, v_min), (lam_max, M_max, v_max)


def _control_nd(ctx, pmin, pmax):
    """Control directions in dimension != 2 (3-D pairs grids run through the explicit-row
    kernel): v_min = X[min pair] * w[min pair], v_max = X[max pair] * w[MIN pair]
    (ddg.py:297-300, quirk Q1 replicated), gathered on the host from the selected rows."""
    kmin = ctx.unpad_policy(pmin).cpu().numpy().astype(np.int64)
    kmax = ctx.unpad_policy(pmax).cpu().numpy().astype(np.int64)
    X = ctx.band_X_full
    if ctx.nd:       # structured 3-D: deep points index the direction table, band points their rows
        bl = ctx.band_global
        Xmin, Xmax = ctx.stencil_X_nd[kmin % len(ctx.stencil_X_nd)], ctx.stencil_X_nd[kmax % len(ctx.stencil_X_nd)]
        Xmin[bl], Xmax[bl] = X[ctx.band_ptr[:-1] + kmin[bl]], X[ctx.band_ptr[:-1] + kmax[bl]]
    else:
        Xmin, Xmax = X[ctx.band_ptr[:-1] + kmin], X[ctx.band_ptr[:-1] + kmax]
    h = np.linalg.norm(Xmin, axis=1)
    w = (1 / h)[:, None]
    dev = ctx.device
    return torch.from_numpy(Xmin * w).to(dev), torch.from_numpy(Xmax * w).to(dev)


def _one_extreme(G, u, which, jacobian=False, control=False):
    """d2min / d2max without the other half of the work (the reference computes both
    and discards one, ddg.py:329-339)."""
    if control:                      # quirk Q1 couples v_max to the min pair: need both
        return d2eigs(G, u, jacobian=jacobian, control=True)[which]
    _require_pairs(G)
    ctx = GridContext.get(G)
    if _pinned_host(u) and not jacobian and ctx.structured and not ctx.nd and ctx.nx_local >= 1024:
        lo, hi = ctx.d2eigs_host(u, want_min=(which == 0), want_max=(which == 1),
                                 mode=ARITHMETIC_MODE)
        return (lo if which == 0 else hi), None, None
    S = ctx.pack(u)
    lmin, lmax, pmin, pmax = ctx.d2eigs(S, want_min=(which == 0), want_max=(which == 1),
                                        policy=jacobian, mode=ARITHMETIC_MODE)
    lam = like_input(u, ctx.unpad(lmin if which == 0 else lmax))
    M = PolicyMatrix(ctx, pmin if which == 0 else pmax) if jacobian else None
    return lam, M, None


def d2min(G, u, **kwargs):
    """ddg.d2min (ddg.py:329-333)."""
    return _one_extreme(G, u, 0, **kwargs)


def d2max(G, u, **kwargs):
    """ddg.d2max (ddg.py:335-339)."""
    return _one_extreme(G, u, 1, **kwargs)


def _aligned_policy(G, ctx, v):
    """Per interior point, the pair best aligned with direction v: largest
    cos^2(v, x_0) + cos^2(v, x_1), first pair among ties (ddg.py:223-231).
    Geometry-only pre-processing, done on the host per call like the reference."""
    v = _ddutils.process_v(G, v)
    N = ctx.n_interior_local
    pol = np.zeros(N, dtype=np.int64)

    def score(X0, X1, vv):
        def c2(X):          # any dimension; identical arithmetic to the 2-D expressions
            h = np.sqrt(np.add.reduce(X * X, axis=-1))
            Xs = X / h[..., None]
            return np.add.reduce(Xs * vv, axis=-1) ** 2
        return c2(X0) + c2(X1)

    if ctx.structured:
        ny = ctx.ny
        isband = np.zeros(N, bool)
        isband[ctx.band_global] = True
        deep = np.flatnonzero(~isband)
        sp = ctx.stencil_pairs
        dim = sp.shape[1] // 2
        c = np.array([[ctx.radius + 1.0] * dim])
        X0 = G._scaled(c + sp[:, :dim]) - G._scaled(c)
        X1 = G._scaled(c + sp[:, dim:]) - G._scaled(c)
        s = score(X0[None], X1[None], v[deep][:, None, :])
        pol[deep] = np.argmax(s, axis=1)
        bp = ctx.band_pairs
        PI, P0, P1 = G.coords(bp[:, 0]), G.coords(bp[:, 1]), G.coords(bp[:, 2])
        sb = score(P0 - PI, P1 - PI, v[bp[:, 0]])
        centres = ctx.band_global
    else:
        bp = ctx.band_pairs
        P = np.asarray(G.points)
        PI, P0, P1 = P[bp[:, 0]], P[bp[:, 1]], P[bp[:, 2]]
        sb = score(P0 - PI, P1 - PI, v[bp[:, 0]])
        centres = ctx.band_global
    ptr = ctx.band_ptr
    if len(centres):
        smax = np.maximum.reduceat(sb, ptr[:-1])
        seg = np.repeat(np.arange(len(centres)), ptr[1:] - ptr[:-1])
        pos = np.arange(sb.size)
        first = np.minimum.reduceat(np.where(sb == smax[seg], pos, sb.size), ptr[:-1])
        pol[centres] = first - ptr[:-1]
    full = torch.zeros(ctx.n_lattice, dtype=torch.int32, device=ctx.device)
    view = full.view(ctx.rows, ctx.pitch)[ctx.halo_rows:ctx.halo_rows + ctx.nx_local, :ctx.ny]
    view.copy_(torch.from_numpy(pol.reshape(ctx.nx_local, ctx.ny).astype(np.int32)).to(ctx.device))
    return full.to(torch.uint16)


def d2(G, v, u=None, jacobian=False):
    """Second derivative in direction ``v`` (ddg.d2, ddg.py:177-249): the pair
    best aligned with ``v`` at every interior point, same 3-point formula.
    Returns ``(d2u, M)``; ``u=None`` yields the matrix only."""
    _require_pairs(G)
    if u is None:
        jacobian = True
    ctx = GridContext.get(G)
    policy = _aligned_policy(G, ctx, v)
    M = PolicyMatrix(ctx, policy)
    d2u = None
    if u is not None:
        # the policy mat-vec evaluates w0*((u0-uc)*w_0 + (u1-uc)*w_1) in the reference's
        # operation order (ddg.py:236), so M.dot(u) IS the operator value, bit for bit
        d2u = like_input(u, ctx.unpad(ctx.apply_policy(policy, ctx.pack(u))))
    return d2u, (M if jacobian else None)
