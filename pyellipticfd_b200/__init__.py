"""pyellipticfd hot path, B200-native (sm_100a).

Drop-in for the wide-stencil operator modules of cfinlay/pyellipticfd
(``ddg``, ``ddi``, ``ddf``), the ``convex_envelope`` / ``pucci`` problems and
the ``solvers`` module, backed by hand-written CUDA reached through the C ABI
declared in ``include/pde_b200.h``.  There is no CPU fallback: importing the
operator modules without the built library raises.
"""
__version__ = "0.1.0"
