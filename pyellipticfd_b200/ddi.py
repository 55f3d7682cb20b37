"""Finite differences with linear interpolation on point clouds -- CUDA.

Drop-in for the second-derivative part of /root/reference/ddi.py (``d2``
ddi.py:223-301, ``d2eigs`` / ``d2min`` / ``d2max`` ddi.py:304-427): same names,
arguments, return conventions and exceptions.  The per-direction finite
difference matrices of the reference are one ELL table on the device (see
``_ell.py``); Jacobians are :class:`~pyellipticfd_b200._ell.EllMatrix` objects.
``d1`` / ``d1n`` / ``d1grad`` (ddi.py:12-221) use the same table machinery with
two-point rows (forward simplex only).
"""
import numpy as np

from . import _ell

eps = _ell.EPS


def _num_directions(G):
    """Number of search directions (ddi.py:120-122)."""
    return np.max([int(2 * np.ceil(np.pi ** 2 / G.angular_resolution ** 2)), 4])


def d1(G, v, u=None, jacobian=False, domain="interior"):
    """First derivative in direction ``v`` by linear interpolation inside the simplex whose
    cone contains ``v`` (ddi.py:12-92).  Returns ``(d1u, M)``."""
    return _ell.d1(G, v, u, jacobian, domain)


def d1n(G, u=None, **kwargs):
    """Derivative along the inward normal on the boundary points (ddi.py:94-117)."""
    if getattr(G, "bdry_normals", None) is None:
        raise TypeError("The grid has no boundary normals")
    return d1(G, -np.asarray(G.bdry_normals), u=u, domain='boundary', **kwargs)


def d1grad(G, u, jacobian=False, control=False, cache_fdmatrices=True):
    """First derivative in the direction of the gradient: maximum over
    ``_num_directions(G)`` directions (ddi.py:124-221).  Returns ``(d1_max, M, v)``."""
    if G.dim == 3:
        raise NotImplementedError("Gradient not yet implemented in 3d.")        # ddi.py:157-158
    return _ell.d1grad(G, u, __name__, int(_num_directions(G)), jacobian, control,
                       cache_fdmatrices)


def d2(G, v, u=None, jacobian=False):
    """Second derivative in direction ``v`` by linear interpolation inside the
    forward and backward simplices (ddi.py:223-301).  Returns ``(d2u, M)``."""
    return _ell.d2(G, v, u, jacobian, froese=False)


def d2eigs(G, u, jacobian=False, control=False, cache_fdmatrices=True):
    """Min / max directional second derivative over ``_num_directions(G)`` directions
    (ddi.py:304-414): ``((lmin, M_min, v_min), (lmax, M_max, v_max))``."""
    if G.dim == 3:
        raise NotImplementedError("Eigenvalues not yet implemented in 3d.")      # ddi.py:330-331
    return _ell.d2eigs(G, u, __name__, int(_num_directions(G)), False, jacobian, control,
                       cache_fdmatrices)


def d2min(G, u, **kwargs):
    return d2eigs(G, u, **kwargs)[0]


def d2max(G, u, **kwargs):
    return d2eigs(G, u, **kwargs)[1]


# hooks used by convex_envelope / pucci for fdmethod='interpolate'
def _ce_operator(G, W, g, jacobian):
    return _ell.ce_operator(G, W, g, jacobian, __name__, int(_num_directions(G)), False)


def _ce_solve(G, g, U0, solver, **kwargs):
    return _ell.ce_solve(G, g, U0, solver, __name__, int(_num_directions(G)), False, **kwargs)


def _pucci_operator(G, W, A, jacobian):
    return _ell.pucci_operator(G, W, A, jacobian, __name__, int(_num_directions(G)), False)


def _pucci_solve(G, f, g, A, U0, solver, **kwargs):
    return _ell.pucci_solve(G, f, g, A, U0, solver, __name__, int(_num_directions(G)), False, **kwargs)
