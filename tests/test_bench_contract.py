"""CPU: the parts of bench.py that do not need a GPU -- argument defaults of the driver
contract, the synthetic disc generator, the obstacle of the reference example."""
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_defaults_follow_the_contract(monkeypatch):
    b = _bench()
    monkeypatch.setattr(sys, "argv", ["bench.py"])
    a = b.parse()
    assert a.gpus == 1 and a.impl == "b200" and a.warmup >= 3 and a.steps > 0
    assert a.n == 4095 and a.radius == 4            # BASELINE configs[1]: 4097^2 lattice, r = 4
    monkeypatch.setattr(sys, "argv", ["bench.py", "--gpus", "8", "--rows", "4096", "--cols", "32766",
                                      "--steps", "7", "--warmup", "3", "--impl", "reference"])
    a = b.parse()
    assert (a.gpus, a.n, a.cols, a.steps, a.warmup, a.impl) == (8, 4096, 32766, 7, 3, "reference")
    assert "Gpoint.dir/s" in b.METRIC and b.UNIT == "Gpoint.dir/s"


def test_disc_cloud_generator():
    b = _bench()
    h0 = 0.05
    theta = np.pi * h0 ** (1 / 3)
    p, nint = b.disc_cloud(h0, theta)
    r = np.sqrt((p ** 2).sum(1))
    assert nint > 1000 and (r[:nint] < 1).all()
    assert np.allclose(r[nint:], 1.0)                # boundary ring on the unit circle
    hB = h0 / 2 * np.tan(theta / 2)
    assert len(p) - nint == len(np.arange(0, 2 * np.pi, hB))
    # same recipe as the test-side generator (oracle/disc.py); the two draw their jitter from
    # differently sized lattices, so only the point density and the ring agree exactly
    from oracle import disc
    q, nq = disc.disc_points(h0, theta)
    assert abs(nq - nint) <= 0.02 * nint and len(q) - nq == len(p) - nint
    assert np.array_equal(q[nq:], p[nint:])


def test_obstacle_is_the_reference_example():
    b = _bench()
    from oracle.make_golden import two_cones
    X = np.random.default_rng(0).uniform(-1, 1, (50, 2))
    assert np.array_equal(b.two_cones(X), two_cones(X))
