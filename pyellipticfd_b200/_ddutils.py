"""Argument normalisation for the directional-derivative routines
(mirror of /root/reference/_ddutils.py:6-47)."""
import numpy as np


def process_v(G, v, domain="interior"):
    """Turn the user's direction argument into an (N, dim) array of unit vectors.

    Accepted forms, as in the reference: a scalar angle (2-D), a pair of angles
    (3-D), one constant vector, one angle per point (2-D), two angles per point
    (3-D), or one vector per point."""
    if domain == "interior":
        N = G.num_interior
    elif domain == "boundary":
        N = G.num_bdry
    else:
        raise ValueError("domain must be 'interior' or 'boundary'")
    dim = G.dim
    if hasattr(v, "detach"):          # torch tensor
        v = v.detach().cpu().numpy()
    v = np.array(v, dtype=np.float64)
    if v.size == 1 and dim == 2:
        return np.broadcast_to([np.cos(v).item(), np.sin(v).item()], (N, dim))
    if v.size == 2 and dim == 3:
        return np.broadcast_to([np.sin(v[0]) * np.cos(v[1]), np.sin(v[0]) * np.sin(v[1]),
                                np.cos(v[1])], (N, dim))
    if v.size == dim:
        return np.broadcast_to(v / np.linalg.norm(v), (N, dim))
    if v.size == N and dim == 2:
        v = v.reshape(N)
        return np.array([np.cos(v), np.sin(v)]).T
    if v.shape == (N, 2) and dim == 3:
        return np.array([np.sin(v[:, 0]) * np.cos(v[:, 1]), np.sin(v[:, 0]) * np.sin(v[:, 1]),
                         np.cos(v[:, 1])]).T
    if v.shape == (N, dim):
        return v / np.linalg.norm(v, axis=1)[:, None]
    return v
