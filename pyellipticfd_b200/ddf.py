"""Froese's generalised finite differences in 2-D -- CUDA.

Drop-in for /root/reference/ddf.py (``d2`` ddf.py:8-131, ``d2eigs`` / ``d2min`` /
``d2max`` ddf.py:137-265).  As in the reference, grids that carry antipodal pairs
are delegated to ``ddg`` (ddf.py:36-37, :165-166).  Deviation: the
backward simplex indices are taken from the backward match; the reference indexes
them with the *forward* selection array (ddf.py:84), which only matters when a
direction lies exactly on a simplex edge.
"""
import numpy as np

from . import _ell, ddg


def _has_pairs(G):
    return hasattr(G, "stencil_pairs") or getattr(G, "pairs", None) is not None


def _num_directions(G):
    """Number of search directions (ddf.py:133-135)."""
    return np.max([int(np.ceil(2 * np.pi / G.angular_resolution)), 4])


def d2(G, v, u=None, jacobian=False):
    if _has_pairs(G):
        return ddg.d2(G, v, u=u, jacobian=jacobian)
    if G.dim != 2:
        raise TypeError('Dimensions other than two not supported.')
    return _ell.d2(G, v, u, jacobian, froese=True)


def d2eigs(G, u, jacobian=False, control=False, cache_fdmatrices=True):
    if _has_pairs(G):
        return ddg.d2eigs(G, u, jacobian=jacobian, control=control)
    if G.dim == 3:
        raise NotImplementedError("Eigenvalues not implemented in 3d.")
    return _ell.d2eigs(G, u, __name__, int(_num_directions(G)), True, jacobian, control,
                       cache_fdmatrices)


def d2min(G, u, **kwargs):
    return d2eigs(G, u, **kwargs)[0]


def d2max(G, u, **kwargs):
    return d2eigs(G, u, **kwargs)[1]


# hooks used by convex_envelope / pucci for fdmethod='froese'
def _ce_operator(G, W, g, jacobian):
    return _ell.ce_operator(G, W, g, jacobian, __name__, int(_num_directions(G)), True)


def _ce_solve(G, g, U0, solver, **kwargs):
    return _ell.ce_solve(G, g, U0, solver, __name__, int(_num_directions(G)), True, **kwargs)


def _pucci_operator(G, W, A, jacobian):
    return _ell.pucci_operator(G, W, A, jacobian, __name__, int(_num_directions(G)), True)


def _pucci_solve(G, f, g, A, U0, solver, **kwargs):
    return _ell.pucci_solve(G, f, g, A, U0, solver, __name__, int(_num_directions(G)), True, **kwargs)
