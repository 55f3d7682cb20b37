"""CPU: the convex-hull sweep solver (csrc/ce_sweep.cu, an extension beyond the reference)
through its HOST build (pde_ce_sweep with device = -1: the same line routine the GPU runs, one
OpenMP thread per lattice line) against the reference's converged solutions
(tests/golden/solve_*.npz) and the NumPy restatement of the residual."""
import glob
import os

import numpy as np
import pytest

from oracle import solvers_oracle as SO
from oracle.make_golden import two_cones
from pyellipticfd_b200 import _lib, _sweep
from pyellipticfd_b200.pointclasses import FDRegularGrid

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SOLVE = sorted(glob.glob(os.path.join(GOLD, "solve_*.npz")))


host_sweep = _sweep.host_sweep     # device=-1: serial chain per line; -2: replay of the CUDA kernel


@pytest.mark.parametrize("path", SOLVE, ids=[os.path.basename(p) for p in SOLVE])
def test_sweeps_reach_the_reference_solution(path):
    z = np.load(path)
    G = FDRegularGrid([int(s) for s in z["shape"]], z["bounds"], int(z["radius"]), interpolation=False)
    U, change, sweeps = host_sweep(G, z["g"])
    assert change < 1e-11 and sweeps < 120
    # the reference's Newton solution solves the discrete equation to 1e-13
    assert np.abs(U - z["newton_U"]).max() < 1e-8
    F, _ = SO.ce_operator(G, U.copy(), z["g"], False)
    assert np.abs(F).max() < 1e-6                                  # the reference's operator_tol
    assert (U <= z["g"]).all()


@pytest.mark.parametrize("shape,r", [((15, 15), 2), ((40, 31), 3), ((70, 130), 4), ((33, 200), 5),
                                     ((129, 129), 6)])
def test_kernel_replay_equals_the_serial_chain(shape, r):
    """The CUDA kernel's phases (bitmap divide-and-conquer hull, wrapped line bundles), replayed
    on the host thread by thread, against the serial monotone chain: one sweep to rounding, and
    the same converged solution."""
    G = FDRegularGrid(list(shape), np.array([[-1.0, 1.0], [-1.0, 1.0]]).T, r, interpolation=False)
    rng = np.random.default_rng(r)
    for g in (two_cones(G.points), two_cones(G.points) + 0.05 * rng.standard_normal(G.num_points)):
        U1, c1, _ = host_sweep(G, g, max_sweeps=1)
        U2, c2, _ = host_sweep(G, g, max_sweeps=1, device=-2)
        assert np.abs(U1 - U2).max() < 1e-14 and abs(c1 - c2) < 1e-14
        Ua, _, sa = host_sweep(G, g, tol=1e-13, relax_every=4)
        Ub, _, sb = host_sweep(G, g, tol=1e-13, relax_every=4, device=-2)
        assert np.abs(Ua - Ub).max() < 1e-13 and abs(sa - sb) <= 2
        assert (Ub <= g).all()


@pytest.mark.parametrize("shape,r", [((2300, 20), 2), ((20, 2300), 2), ((1500, 40), 3)])
def test_kernel_replay_on_long_lines_with_planar_pieces(shape, r):
    """Lines longer than one 1024-point group, obstacles with planar pieces: exercises the
    flat-word / flat-group shortcuts and the merge levels across warps of the kernel."""
    nx, ny = shape
    Lx, Ly = (nx + 1) / 64.0, (ny + 1) / 64.0
    G = FDRegularGrid([nx, ny], np.array([[0.0, Lx], [0.0, Ly]]).T, r, interpolation=False)
    X, Y = G.points.T
    s, c = (X / Lx, Y / Ly) if nx > ny else (Y / Ly, X / Lx)
    for g in (np.abs(s - 0.5) * 3 + 0.3 * np.sin(40 * s) * (s > 0.8) + 0.2 * c, 0 * X + 1.0,
              two_cones(np.stack([2 * X / Lx - 1, 2 * Y / Ly - 1], 1))):
        U1, _, _ = host_sweep(G, g, max_sweeps=1)
        U2, _, _ = host_sweep(G, g, max_sweeps=1, device=-2)
        assert np.abs(U1 - U2).max() < 1e-14
        Ua, _, sa = host_sweep(G, g, tol=1e-13, relax_every=4)
        Ub, _, sb = host_sweep(G, g, tol=1e-13, relax_every=4, device=-2)
        assert np.abs(Ua - Ub).max() < 1e-13 and abs(sa - sb) <= 2 and (Ub <= g).all()
        F, _ = SO.ce_operator(G, Ub.copy(), g, False)
        assert np.abs(F).max() < 1e-8


def test_sweep_counts_do_not_grow_with_the_lattice():
    counts = []
    for n in (31, 63, 127):
        G = FDRegularGrid([n, n], np.array([[-1.0, 1.0], [-1.0, 1.0]]).T, 3, interpolation=False)
        g = two_cones(G.points)
        U, change, sweeps = host_sweep(G, g, tol=1e-9)
        F, _ = SO.ce_operator(G, U.copy(), g, False)
        assert np.abs(F).max() < 1e-5
        counts.append(sweeps)
    assert max(counts) <= 2 * min(counts) + 8, counts


def test_convex_obstacle_is_a_fixed_point():
    G = FDRegularGrid([21, 17], np.array([[0.0, 1.0], [0.0, 2.0]]).T, 2, interpolation=False)
    X, Y = G.points.T
    g = (X - 0.3) ** 2 + 2 * (Y - 1.1) ** 2 + X * Y
    U, change, sweeps = host_sweep(G, g, tol=1e-13)
    assert sweeps <= 2 and np.abs(U - g).max() < 1e-13


from hypothesis import given, settings, strategies as st


@settings(max_examples=10, deadline=None)
@given(nx=st.integers(7, 22), ny=st.integers(7, 22), r=st.integers(1, 3), seed=st.integers(0, 10 ** 6),
       sx=st.sampled_from([1.0, 0.5, 0.37]), rough=st.sampled_from([0.0, 0.05, 0.5]))
def test_random_obstacles_against_the_reference_algorithm(nx, ny, r, seed, sx, rough):
    """Random boxes (anisotropic, non-dyadic spacing) and obstacles: the sweeps end on a function
    that solves the discrete equation (NumPy restatement of convex_envelope.operator) and that
    the restated Newton iteration of the reference, started from it, leaves where it is."""
    if min(nx, ny) < 2 * r + 1:
        return
    bounds = np.array([[-0.3, -0.3 + sx * (nx + 1) / 8.0], [0.0, (ny + 1) / 8.0]]).T
    G = FDRegularGrid([nx, ny], bounds, r, interpolation=False)
    X, Y = G.points.T
    rng = np.random.default_rng(seed)
    g = np.sin(5 * X + rng.uniform(0, 3)) * np.cos(4 * Y) + rough * rng.standard_normal(X.size)
    U, change, sweeps = host_sweep(G, g, tol=1e-13)
    assert change < 1e-13 and sweeps < 300
    F, _ = SO.ce_operator(G, U.copy(), g, False)
    scale = 1.0 / G.scaling.min() ** 2
    assert np.abs(F).max() < 1e-11 * scale and (U <= g).all()
    with np.errstate(all="ignore"):
        Un, diff, it, _ = SO.ce_solve(G, g, U0=U.copy(), solver="newton")
    assert it <= 2 and np.abs(Un - U).max() < 1e-9


class _HostCtx(object):
    """Stand-in for GridContext in the host logic of convex_envelope._solve_grid(solver='sweep'):
    flat torch CPU vectors, the host build of the line routine, the NumPy residual."""

    def __init__(self, G):
        import torch
        self.G, self.torch = G, torch
        self.calls = []

    def pack(self, u):
        return self.torch.as_tensor(np.asarray(u, dtype=np.float64)).clone()

    def unpack(self, S):
        return S

    def ce_sweep(self, g, tol, max_sweeps, relax_reps=2, U=None, relax_every=0):
        self.calls.append(int(max_sweeps))
        start = g if U is None else U
        out, change, n = host_sweep(self.G, start.numpy(), tol=tol, max_sweeps=max_sweeps,
                                    relax_reps=relax_reps, relax_every=relax_every)
        return self.torch.from_numpy(out), change, n

    def ce_residual(self, W, g, policy=True):
        F, _ = SO.ce_operator(self.G, W.numpy().copy(), g.numpy(), False)
        return self.torch.from_numpy(F), None, self.torch.tensor([np.abs(F).max(), 0.0])


def test_solver_option_host_logic(monkeypatch):
    """Stopping rule, iteration cap, warning and timeout of solver='sweep' (Python side)."""
    import warnings
    from pyellipticfd_b200 import convex_envelope as CE
    z = np.load(os.path.join(GOLD, "solve_n31_r2.npz"))
    G = FDRegularGrid([int(s) for s in z["shape"]], z["bounds"], int(z["radius"]), interpolation=False)
    ctx = _HostCtx(G)
    monkeypatch.setattr(CE.GridContext, "get", staticmethod(lambda Grid, device=None: ctx))
    st = {}
    U, diff, sweeps, secs = CE.solve(G, z["g"], fdmethod="grid", solver="sweep", stats=st)
    assert diff < 1e-6 and st["residual"] < 1e-6 and 0 < sweeps <= 8 * len(ctx.calls)
    assert np.abs(U - z["newton_U"]).max() < 2e-6 and isinstance(U, np.ndarray)
    assert all(c == 8 for c in ctx.calls)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        U, diff, sweeps, _ = CE.solve(G, z["g"], fdmethod="grid", solver="sweep", max_iters=3)
    assert sweeps == 3 and any("Maximum iterations" in str(x.message) for x in w)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        U, diff, sweeps, _ = CE.solve(G, z["g"], fdmethod="grid", solver="sweep", check_every=1,
                                      solution_tol=1e-300, timeout=0.0)
    assert sweeps == 1 and len(w) == 1
    with pytest.raises(ValueError):
        CE.solve(G, z["g"], fdmethod="grid", solver="nope")
