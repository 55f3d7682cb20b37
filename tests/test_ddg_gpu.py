"""GPU parity of the ddg path against the oracle (bit-exact) -- through the C ABI."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import ddg_oracle as O


def _grid(shape, bounds, r):
    from pyellipticfd_b200.pointclasses import FDRegularGrid
    return FDRegularGrid(list(shape), np.array(bounds, dtype=float).T, r, interpolation=False)


CASES = [((9, 9), [[0, 1], [0, 1]], 2),
         ((19, 19), [[0, 1], [0, 1]], 4),
         ((31, 150), [[0, 1], [0, 1]], 3),       # dyadic in x only -> generic path
         ((63, 63), [[-1, 1], [-1, 1]], 2),
         ((40, 300), [[-1, 1], [-1, 1]], 1),
         ((11, 14), [[-1, 1], [0, 3]], 2),       # non-uniform weights -> explicit rows
         ((15, 15), [[0, 1], [0, 1]], 5),
         ((31, 31), [[0, 1], [0, 1]], 6)]


def _fields(G):
    rng = np.random.default_rng(0)
    X, Y = G.points.T
    return {"random": rng.standard_normal(G.num_points),
            "saddle": X ** 2 - Y ** 2,
            "zero": np.zeros(G.num_points),
            "smooth": np.sin(3 * X) * np.cos(2 * Y) + X * Y}


@pytest.mark.parametrize("shape,bounds,r", CASES)
def test_d2eigs_bitexact(shape, bounds, r):
    from pyellipticfd_b200 import ddg
    G = _grid(shape, bounds, r)
    for name, u in _fields(G).items():
        (lmin, Mmin, vmin), (lmax, Mmax, vmax) = ddg.d2eigs(G, u, jacobian=True, control=True)
        (omin, oMmin, ovmin), (omax, oMmax, ovmax) = O.d2eigs(G, u, jacobian=True, control=True)
        assert np.array_equal(lmin, omin), (name, np.abs(lmin - omin).max())
        assert np.array_equal(lmax, omax), (name, np.abs(lmax - omax).max())
        # selected pairs, ties included: the explicit matrices must be identical
        assert (Mmin.tocsr() != oMmin).nnz == 0, name
        assert (Mmax.tocsr() != oMmax).nnz == 0, name
        assert np.array_equal(vmin, ovmin), name
        assert np.array_equal(vmax, ovmax), name
        # the reference tests' identity M.dot(U) == lambda (tests/test_ddg.py:75-78)
        assert np.abs(Mmin.dot(u) - lmin).max() < 1e-12 * max(1.0, np.abs(lmin).max())
        assert np.abs(Mmax.dot(u) - lmax).max() < 1e-12 * max(1.0, np.abs(lmax).max())


def test_reference_known_answers():
    """tests/test_ddg.py:58-80 and tests/test_regular_grid.py:32-37 re-run on the CUDA path."""
    from pyellipticfd_b200 import ddg
    G = _grid((9, 9), [[0, 1], [0, 1]], 2)
    X, Y = G.points.T
    U2 = X ** 2 - Y ** 2
    d2x, M2x = ddg.d2(G, [1, 0], U2, jacobian=True)
    d2y, M2y = ddg.d2(G, [0, 1], U2, jacobian=True)
    assert np.abs(d2x - 2).max() < 1e-8
    assert np.abs(M2x.dot(U2) - d2x).max() < 1e-12
    assert np.abs(d2y + 2).max() < 1e-8
    assert np.abs(M2y.dot(U2) - d2y).max() < 1e-12
    (d2min, M_min, v_min), (d2max, M_max, v_max) = ddg.d2eigs(G, U2, jacobian=True, control=True)
    assert np.abs(d2min + 2).max() < 1e-8
    assert np.abs(d2max - 2).max() < 1e-8
    assert np.abs(M_min.dot(U2) - d2min).max() < 1e-12
    assert np.abs(M_max.dot(U2) - d2max).max() < 1e-12
    assert np.abs((v_max.dot(np.array([1, 0]))) ** 2 - 1).max() < 1e-8
    assert np.abs((v_min.dot(np.array([0, 1]))) ** 2 - 1).max() < 1e-8
    G = _grid((19, 19), [[0, 1], [0, 1]], 4)
    X, Y = G.points.T
    U2 = X ** 2 - Y ** 2
    d2x, _ = ddg.d2(G, [1, 0], U2)
    d2y, _ = ddg.d2(G, [0, 1], U2)
    assert np.abs(d2x.max()) - 2 < 1e-13
    assert np.abs(d2y.max()) - 2 < 1e-13


@pytest.mark.parametrize("shape,bounds,r", CASES[:4])
def test_d2_matches_oracle(shape, bounds, r):
    from pyellipticfd_b200 import ddg
    G = _grid(shape, bounds, r)
    u = _fields(G)["random"]
    rng = np.random.default_rng(1)
    for v in ([1, 0], [0, 1], 0.3, rng.standard_normal((G.num_interior, 2))):
        d, M = ddg.d2(G, v, u, jacobian=True)
        od, oM = O.d2(G, v, u, jacobian=True)
        assert np.array_equal(d, od)
        assert (M.tocsr() != oM).nnz == 0


def test_torch_inputs_stay_on_device():
    import torch
    from pyellipticfd_b200 import ddg
    G = _grid((33, 33), [[0, 1], [0, 1]], 2)
    u = torch.from_numpy(_fields(G)["smooth"]).cuda()
    (lmin, _, _), (lmax, _, _) = ddg.d2eigs(G, u)
    assert lmin.is_cuda and lmax.is_cuda
    omin, omax = O.d2eigs(G, u.cpu().numpy())[0][0], O.d2eigs(G, u.cpu().numpy())[1][0]
    assert np.array_equal(lmin.cpu().numpy(), omin) and np.array_equal(lmax.cpu().numpy(), omax)


def test_pinned_host_pipeline_bitexact():
    """Pinned host tensors take the chunk-pipelined path (H2D / stencil / D2H over row
    chunks, 2-D DMA layout conversion): same bits as the oracle."""
    import torch
    from oracle import c_oracle
    from pyellipticfd_b200 import ddg
    G = _grid((1100, 260), [[0, 1101.0 / 512], [0, 261.0 / 512]], 3)
    assert G.uniform_weights()
    u = np.random.default_rng(5).standard_normal(G.num_points)
    omin, omax, _, _ = c_oracle.pairs_d2eigs(G.points, G.pairs, u, G.num_interior)
    host = torch.from_numpy(u).pin_memory()
    (lmin, _, _), (lmax, _, _) = ddg.d2eigs(G, host)
    assert not lmin.is_cuda and lmin.is_pinned()
    assert np.array_equal(lmin.numpy(), omin) and np.array_equal(lmax.numpy(), omax)
    assert np.array_equal(ddg.d2min(G, host)[0].numpy(), omin)
    assert np.array_equal(ddg.d2max(G, host)[0].numpy(), omax)
    # repeated calls reuse the staging buffers
    u2 = u[::-1].copy()
    o2 = c_oracle.pairs_d2eigs(G.points, G.pairs, u2, G.num_interior)[0]
    assert np.array_equal(ddg.d2min(G, torch.from_numpy(u2).pin_memory())[0].numpy(), o2)
