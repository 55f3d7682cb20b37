"""Unit-disc point cloud for the point-cloud tests -- TEST INFRASTRUCTURE.

Follows the recipe of /root/reference/tests/setup_discs.py:17-39 (interior points at
spacing h0 inside radius 1-c, fine boundary ring at spacing hB, Delaunay) with the
absent ``distmesh`` generator replaced by a jittered hexagonal lattice (a plain hex
lattice puts search directions exactly on simplex edges, where ddf has no tolerance).
"""
import numpy as np


def disc_points(h0, angular_resolution, seed=0, jitter=0.1):
    hB = h0 / 2 * np.tan(angular_resolution / 2)
    c = np.max([hB * (-1 + 2 / np.sin(angular_resolution / 2)), 2 * hB / np.tan(angular_resolution / 2)])
    R = 1.0 - c
    rng = np.random.default_rng(seed)
    nx = int(np.ceil(2 / h0)) + 2
    ys = np.arange(-nx, nx + 1) * h0 * np.sqrt(3) / 2
    pts = []
    for j, y in enumerate(ys):
        xs = np.arange(-nx, nx + 1) * h0 + (h0 / 2 if j % 2 else 0.0)
        pts.append(np.stack([xs, np.full(xs.size, y)], 1))
    p = np.concatenate(pts)
    p = p + rng.uniform(-jitter * h0, jitter * h0, p.shape)
    p = p[np.sqrt((p ** 2).sum(1)) <= R]
    th = np.arange(0, 2 * np.pi, hB)
    ring = np.stack([np.cos(th), np.sin(th)], 1)
    return np.concatenate([p, ring]), p.shape[0]
