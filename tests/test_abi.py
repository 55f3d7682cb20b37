"""CPU: the C-ABI library loads and exports every symbol include/pde_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "pde_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pde_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_path():
    syms = declared_symbols()
    for must in ("pde_grid_create", "pde_ddg_d2eigs", "pde_jac_matvec", "pde_bicgstab",
                 "pde_ce_euler", "pde_ce_residual", "pde_pucci_residual", "pde_ell_d2eigs"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    from pyellipticfd_b200 import _lib
    L = _lib.lib()
    missing = [s for s in declared_symbols() if not hasattr(L, s)]
    assert not missing, missing
    assert L.pde_abi_version() == 1


def test_ctypes_signatures_cover_the_header():
    from pyellipticfd_b200 import _lib
    missing = [s for s in declared_symbols() if s not in _lib.SIGNATURES]
    assert not missing, missing


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import numpy as np
    from pyellipticfd_b200 import ddg, _lib
    from pyellipticfd_b200.pointclasses import FDRegularGrid
    G = FDRegularGrid([9, 9], np.array([[0, 1], [0, 1]]).T, 2, interpolation=False)
    with pytest.raises(_lib.PdeError):
        ddg.d2eigs(G, np.zeros(G.num_points))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "pyellipticfd_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
