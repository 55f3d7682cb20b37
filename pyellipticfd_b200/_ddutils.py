"""Argument normalisation for the directional-derivative routines
(same contract as /root/reference/_ddutils.py:6-47)."""
import numpy as np

_COUNT = {"interior": "num_interior", "boundary": "num_bdry"}


def _components(angles, dim):
    """Cartesian components of the direction given by polar angle(s), in the reference's
    convention: (cos a, sin a) in 2-D; (sin t cos p, sin t sin p, cos p) in 3-D."""
    if dim == 2:
        a, = angles
        return [np.cos(a), np.sin(a)]
    t, p = angles
    return [np.sin(t) * np.cos(p), np.sin(t) * np.sin(p), np.cos(p)]


def process_v(G, v, domain="interior"):
    """Turn the user's direction argument into an (n, dim) array of unit vectors, one per
    point of ``domain``.

    Accepted forms, tried in this order: constant angle(s) (one in 2-D, two in 3-D), one
    constant vector, angle(s) per point, one vector per point.  Anything else is returned
    unchanged, as the reference does."""
    if domain not in _COUNT:
        raise ValueError("domain must be 'interior' or 'boundary'")
    n, dim = getattr(G, _COUNT[domain]), G.dim
    if hasattr(v, "detach"):          # torch tensor
        v = v.detach().cpu().numpy()
    v = np.array(v, dtype=np.float64)
    angular = dim in (2, 3)
    if angular and v.size == dim - 1:
        const = [float(c) for c in _components(v.reshape(dim - 1), dim)]
        return np.broadcast_to(const, (n, dim))
    if v.size == dim:
        return np.broadcast_to(v / np.linalg.norm(v), (n, dim))
    if angular and (v.size == n if dim == 2 else v.shape == (n, 2)):
        return np.array(_components(v.reshape(n, dim - 1).T, dim)).T
    if v.shape == (n, dim):
        lengths = np.linalg.norm(v, axis=1)
        return v / lengths[:, None]
    return v
