"""CPU: the NumPy restatement of ddi / ddf against the reference outputs in the fixtures."""
import glob
import os

import numpy as np
import pytest

from oracle import ddi_oracle
from pyellipticfd_b200.pointclasses import FDPointCloud

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CLOUD = sorted(glob.glob(os.path.join(GOLD, "cloud_*.npz")))


@pytest.mark.parametrize("path", CLOUD[:1], ids=[os.path.basename(p) for p in CLOUD[:1]])
@pytest.mark.parametrize("tag", ["ddi", "ddf"])
def test_oracle_matches_reference(path, tag):
    z = np.load(path)
    G = FDPointCloud(z["points"], angular_resolution=float(z["angular_resolution"]),
                     num_interior=int(z["num_interior"]),
                     spatial_resolution=float(z["spatial_resolution"]))
    G.simplices = z["simplices"].astype(np.int64)
    froese = tag == "ddf"
    assert ddi_oracle.num_directions(G, froese) == int(z[tag + "_nds"])
    table = ddi_oracle.d2eigs_table(G, froese)
    X, Y = G.points.T
    for name, u in (("saddle", X ** 2 - Y ** 2),
                    ("random", np.random.default_rng(0).standard_normal(len(X)))):
        lmin, lmax, _ = ddi_oracle.d2eigs(G, u, froese, table)
        rmin, rmax = z["%s_lmin_%s" % (tag, name)], z["%s_lmax_%s" % (tag, name)]
        tol = 1e-9 * max(np.abs(rmin).max(), np.abs(rmax).max())
        assert np.abs(lmin - rmin).max() < tol and np.abs(lmax - rmax).max() < tol
