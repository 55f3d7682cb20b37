"""CPU: host-side logic and error behaviour of the reference-facing modules (no GPU work)."""
import numpy as np
import pytest

from oracle import ddg_oracle
from pyellipticfd_b200 import _ddutils, ddg, ddi, ddf, pucci, solvers, stencils
from pyellipticfd_b200.pointclasses import FDPointCloud, FDRegularGrid


class _G(object):
    def __init__(self, n, nb=3, dim=2):
        self.num_interior, self.num_bdry, self.dim = n, nb, dim


@pytest.mark.parametrize("v", [0.3, [1.0, 2.0], np.linspace(0, 1, 7),
                               np.random.default_rng(0).standard_normal((7, 2))])
def test_process_v_matches_reference_restatement(v):
    G = _G(7)
    a = _ddutils.process_v(G, v)
    b = ddg_oracle.process_v(G, v)
    assert a.shape == (7, 2) and np.array_equal(np.asarray(a), np.asarray(b))
    assert np.allclose(np.linalg.norm(a, axis=1), 1)


def test_process_v_boundary_domain():
    G = _G(7, nb=4)
    assert _ddutils.process_v(G, [0.0, 1.0], domain="boundary").shape == (4, 2)


def test_errors_follow_the_reference():
    cloud = FDPointCloud(np.zeros((4, 2)), num_interior=2)       # no pairs, no simplices
    with pytest.raises(TypeError):                                # ddg.py:200-201
        ddg.d2(cloud, [1, 0], np.zeros(4))
    with pytest.raises(TypeError):                                # pointclasses.py:158-159
        FDPointCloud(np.zeros((4, 2)))
    with pytest.raises(TypeError):                                # pointclasses.py:161-166
        FDPointCloud(np.zeros((4, 2)), num_interior=2, angular_resolution=4.0)
    c3 = FDPointCloud(np.zeros((4, 3)), num_interior=2)
    with pytest.raises(NotImplementedError):                      # ddi.py:330-331
        ddi.d2eigs(c3, np.zeros(4))
    with pytest.raises(NotImplementedError):                      # ddf.py:168-169
        ddf.d2eigs(c3, np.zeros(4))
    with pytest.raises(NotImplementedError):                      # pucci.py:33-34
        pucci.operator(c3, np.zeros(4), (2, 1))
    with pytest.raises(ValueError):                               # solvers.py:150-151
        solvers.NewtonEulerLS(np.zeros(3), lambda U, jacobian=True: None, euler_ratio=0)
    with pytest.raises(ValueError):                               # stencils.py:69-70
        stencils.create(2, 1)
    with pytest.raises(ValueError):                               # stencils.py:31-32
        stencils.create(0, 2)


def test_generic_euler_on_numpy_closure():
    """solvers.euler with a plain NumPy operator: stopping rule and iteration count of
    solvers.py:63-91 (U <- U - dt*F, stop when both tolerances hold, i counted from 1)."""
    target = np.array([1.0, -2.0, 0.5])
    U, diff, i, t = solvers.euler(np.zeros(3), lambda U: (U - target, 0.5), max_iters=200)
    assert np.abs(U - target).max() < 1e-5 and diff < 1e-6 and 1 < i < 200
    with pytest.warns(UserWarning):          # default max_iters = U.size (solvers.py:56-57)
        U2, _, i2, _ = solvers.euler(np.zeros(3), lambda U: (U - target, 0.5))
    assert i2 == 3


def test_grid_metadata_and_lazy_pairs():
    G = FDRegularGrid([12, 9], np.array([[0, 13 / 16.0], [0, 10 / 16.0]]).T, 2, interpolation=False)
    assert G.uniform_weights() and G.dim == 2
    assert G.num_points == G.num_interior + G.num_bdry
    assert G.pairs.shape[1] == 3 and (np.diff(G.pairs[:, 0]) >= 0).all()
    assert G.num_pairs == G.pairs.shape[0]
    assert np.array_equal(G.interior, np.arange(G.num_interior))
    assert np.array_equal(G.bdry_points, G.points[G.num_interior:])
    assert G.min_radius == G.scaling.min() and G.min_interior_nb_dist == G.scaling.min()
    with pytest.raises(NotImplementedError):
        FDRegularGrid([5, 5], np.array([[0, 1], [0, 1.0]]).T, 2, interpolation=True)


def test_public_signatures_match_the_reference():
    """Names, parameter names and defaults of the reference's public functions
    (ddg.py, ddi.py, ddf.py, convex_envelope.py, pucci.py, solvers.py, pointclasses.py)."""
    import inspect
    from pyellipticfd_b200 import convex_envelope
    from pyellipticfd_b200.pointclasses import FDTriMesh

    def params(f):
        return [(p.name, p.default if p.default is not inspect._empty else None)
                for p in inspect.signature(f).parameters.values()
                if p.kind in (p.POSITIONAL_OR_KEYWORD, p.KEYWORD_ONLY)]

    for mod in (ddg, ddi):
        assert params(mod.d1) == [("G", None), ("v", None), ("u", None), ("jacobian", False),
                                  ("domain", "interior")]                 # ddg.py:8, ddi.py:12
        assert [n for n, _ in params(mod.d1n)][:2] == ["G", "u"]             # ddg.py:86, ddi.py:94
    assert params(ddg.d1grad) == [("G", None), ("u", None), ("jacobian", False), ("control", False)]
    assert params(ddi.d1grad) == [("G", None), ("u", None), ("jacobian", False), ("control", False),
                                  ("cache_fdmatrices", True)]                # ddi.py:124
    for mod in (ddg, ddi, ddf):
        assert params(mod.d2) == [("G", None), ("v", None), ("u", None), ("jacobian", False)]
    assert params(ddg.d2eigs) == [("G", None), ("u", None), ("jacobian", False), ("control", False)]
    for mod in (ddi, ddf):
        assert params(mod.d2eigs) == [("G", None), ("u", None), ("jacobian", False),
                                      ("control", False), ("cache_fdmatrices", True)]
    assert params(convex_envelope.operator) == [("Grid", None), ("W", None), ("g", None),
                                                ("jacobian", True), ("fdmethod", "interpolate")]
    assert params(convex_envelope.solve)[:5] == [("Grid", None), ("g", None), ("U0", None),
                                                 ("fdmethod", "interpolate"), ("solver", "newton")]
    assert params(pucci.solve)[:7] == [("Grid", None), ("f", None), ("dirichlet", None),
                                       ("A", (2, 1)), ("U0", None), ("fdmethod", "interpolate"),
                                       ("solver", "newton")]
    assert [n for n, _ in params(solvers.euler)] == [
        "U", "operator", "solution_tol", "operator_tol", "max_iters", "timeout", "zeromean",
        "zeromax", "plotter"]                                                # solvers.py:13-15
    assert [n for n, _ in params(solvers.NewtonEulerLS)] == [
        "U", "operator", "solution_tol", "operator_tol", "max_iters", "euler_ratio", "plotter",
        "max_euler_iters", "verbose", "force_euler"]                         # solvers.py:98-101
    assert [n for n, _ in params(FDTriMesh.__init__)][1:4] == ["p", "t", "angular_resolution"]
    assert [n for n, _ in params(FDRegularGrid.__init__)][1:5] == [
        "interior_shape", "bounds", "stencil_radius", "interpolation"]     # pointclasses.py:480


def test_first_derivative_errors():
    cloud = FDPointCloud(np.zeros((4, 2)), num_interior=2)
    with pytest.raises(TypeError):                       # no boundary normals
        ddg.d1n(cloud, np.zeros(4))
    with pytest.raises(TypeError):
        ddi.d1n(cloud, np.zeros(4))
    c3 = FDPointCloud(np.zeros((4, 3)), num_interior=2)
    with pytest.raises(NotImplementedError):             # ddi.py:157-158
        ddi.d1grad(c3, np.zeros(4))


def test_public_helpers_match_the_reference():
    """pointclasses.compute_pairs / remove_colinear_neighbours / compute_simplices (:12-121),
    the FDPointCloud properties adjacency / min_dist_to_bdry / maximum_bdry_res (:239-280) and
    stencils.stern_brocot (stencils.py:6-23) against outputs of the unmodified reference
    (tests/golden/helpers.npz, oracle/make_golden.py::helpers_case)."""
    import os
    from pyellipticfd_b200 import pointclasses as PC, stencils
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "helpers.npz"))
    for N in (1, 2, 4):
        assert np.array_equal(stencils.stern_brocot(N), z["stern_brocot_%d" % N])
    th = float(z["angular_resolution"])
    O = PC.FDPointCloud(z["points"], angular_resolution=th, num_interior=int(z["num_interior"]),
                        spatial_resolution=float(z["spatial_resolution"]),
                        bdry_resolution=float(z["bdry_resolution"]), dist_to_bdry=float(z["dist_to_bdry"]))
    assert O.min_dist_to_bdry == float(z["min_dist_to_bdry"])
    assert O.maximum_bdry_res == float(z["maximum_bdry_res"])
    for prefer in ("min", "max"):
        O.neighbours = z["neighbours_in"].astype(np.int64)
        PC.remove_colinear_neighbours(O, 1 - np.cos(th / 8), prefer=prefer)
        assert np.array_equal(O.neighbours, z["neighbours_" + prefer])
    assert O.adjacency.nnz == int(z["adjacency_nnz"]) and O.adjacency.shape == (O.num_points,) * 2
    PC.compute_simplices(O)
    canon = lambda S: sorted((int(a),) + tuple(sorted((int(b), int(c)))) for a, b, c in S)
    assert canon(O.simplices) == canon(z["simplices_max"])       # facet orientation is Qhull's own
    Q = PC.FDPointCloud(z["grid_points"], num_interior=int(z["grid_num_interior"]))
    Q.neighbours = z["grid_neighbours"].astype(np.int64)
    PC.compute_pairs(Q, 1e-3)
    assert np.array_equal(Q.pairs, z["grid_pairs"])
