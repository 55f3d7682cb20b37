"""TEST INFRASTRUCTURE ONLY -- import shim for the *unmodified* upstream reference.

This module makes the read-only tree at /root/reference importable as the
package ``pyellipticfd`` inside THIS container so that
``oracle/make_golden.py`` can generate golden vectors from the real reference
and so that the numpy/C restatements under ``oracle/`` can be pinned against it.
It is never imported by the product package or by the ``-m gpu`` tests.  On the GPU box
/root/reference does not exist; ``bench.py --impl reference`` (the CPU reference arm) imports
the byte-for-byte copy under the git-ignored ``oracle/_ref/`` instead (oracle/ref_install.py).

The four mechanical patches paper over API drift between
the reference's 2018-era NumPy/SciPy and the versions installed here; none of
them changes arithmetic:

* ``scipy.sparse.linalg.dsolve.linsolve.MatrixRankWarning`` (solvers.py:10)
* ``bicgstab(tol=...)`` -> ``rtol`` (solvers.py:181)
* ``np.linalg.solve`` with a stack of vectors (ddi.py:63,265; ddf.py:60)
* ``np.in1d`` removed in NumPy 2 (ddi.py:51,253; ddf.py:47; pointclasses.py)
"""
import os
import sys
import types
import tempfile
import warnings

import numpy as np
import scipy.sparse.linalg as sla

def _root():
    """/root/reference in the build container; on the GPU box the byte-for-byte copy that
    oracle/ref_install.py left in the git-ignored oracle/_ref/ (it travels with the snapshot)."""
    env = os.environ.get("PYELLIPTICFD_REFERENCE")
    if env:
        return env
    if os.path.exists("/root/reference/ddg.py"):
        return "/root/reference"
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "pyellipticfd")


REFERENCE_ROOT = _root()

_installed = False


def available():
    return os.path.isdir(REFERENCE_ROOT) and os.path.exists(
        os.path.join(REFERENCE_ROOT, "ddg.py"))


def install():
    """Make ``import pyellipticfd`` resolve to the reference tree. Idempotent."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)

    link_dir = tempfile.mkdtemp(prefix="pyellipticfd_ref_")
    os.symlink(REFERENCE_ROOT, os.path.join(link_dir, "pyellipticfd"))
    sys.path.insert(0, link_dir)
    sys.dont_write_bytecode = True          # the reference tree is read-only

    m = types.ModuleType("scipy.sparse.linalg.dsolve.linsolve")
    m.MatrixRankWarning = sla.MatrixRankWarning
    pkg = types.ModuleType("scipy.sparse.linalg.dsolve")
    pkg.linsolve = m
    sys.modules.setdefault("scipy.sparse.linalg.dsolve", pkg)
    sys.modules["scipy.sparse.linalg.dsolve.linsolve"] = m

    _bicgstab = sla.bicgstab

    def bicgstab(A, b, x0=None, tol=None, **kw):
        if tol is not None:
            kw["rtol"] = tol
        return _bicgstab(A, b, x0=x0, **kw)

    sla.bicgstab = bicgstab

    _solve = np.linalg.solve

    def solve(a, b):
        a = np.asarray(a)
        b = np.asarray(b)
        if b.ndim == a.ndim - 1:
            return _solve(a, b[..., None])[..., 0]
        return _solve(a, b)

    np.linalg.solve = solve
    # NumPy 2 deprecates in1d, and NewtonEulerLS promotes warnings to errors (solvers.py:174)
    np.in1d = lambda a, b, **k: np.isin(a, b, **k).ravel()
    warnings.filterwarnings("ignore", category=DeprecationWarning)
    _installed = True


def modules():
    """Return the reference modules as a namespace (after install())."""
    install()
    from pyellipticfd import (ddg, ddi, ddf, solvers, convex_envelope, pucci,
                              pointclasses, stencils, _ddutils)
    return types.SimpleNamespace(ddg=ddg, ddi=ddi, ddf=ddf, solvers=solvers,
                                 convex_envelope=convex_envelope, pucci=pucci,
                                 pointclasses=pointclasses, stencils=stencils,
                                 _ddutils=_ddutils)
