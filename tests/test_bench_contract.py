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
    # 0 = "by GPU count": 4095 (BASELINE configs[1], 4097^2 lattice) on one GPU; on N > 1 GPUs one
    # 4096 x 32766 slab each (N = 8: the 32768^2 lattice of configs[4]) -- resolved in main()
    assert a.n == 0 and a.cols == 0 and a.radius == 4
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


def _line(name):
    import json
    path = os.path.join(ROOT, "profiles", name)
    return [json.loads(l) for l in open(path) if l.startswith("{")][-1]


def test_committed_bench_lines_carry_the_contract_keys():
    """profiles/r01_bench*.json are bench.py's own output on the B200 box: every key of the
    driver contract is present and consistent."""
    d = _line("r01_bench.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step",
              "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
              "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in d, k
    assert d["unit"] == "Gpoint.dir/s" and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s"
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["traffic"] is None or r["traffic"] > 0
    c = d["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["sample"]
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert e["value"] < d["value"]                      # host copies inside the timed region
    assert d["gpu_launches"] >= d["steps"]
    assert d["warmup"] >= 3 and not set(d["clocks"]["reasons"]) & {
        "hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    # value = interior points x pairs / time
    assert abs(d["value"] - 4095 * 4095 * 24 / d["ms_per_step"] / 1e6) < 1e-6 * d["value"]
    ref = _line("r01_bench_reference.json")
    assert ref["impl"] == "reference" and ref["unit"] == d["unit"]
    assert ref["e2e"]["h2d_bytes_per_step"] == 0 and ref["cpu_baseline"]["value"] == ref["value"]
