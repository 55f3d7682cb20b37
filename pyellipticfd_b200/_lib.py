"""ctypes binding of ``libpde_b200.so`` (the C ABI in ``include/pde_b200.h``).

There is no CPU fallback: if the library is missing, ``lib()`` raises.  The
library is built in-tree by ``pyellipticfd_b200._build`` /
``__graft_entry__.build()``.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libpde_b200.so")

_lib = None

c_i64 = C.c_int64
c_int = C.c_int
c_dbl = C.c_double
vp = C.c_void_p


class pde_ell_jac(C.Structure):
    _fields_ = [("d_kmin", vp), ("d_kmax", vp), ("cmin", c_dbl), ("cmax", c_dbl)]


class pde_jac(C.Structure):
    _fields_ = [("d_pmin", vp), ("d_pmax", vp), ("cmin_s", c_dbl), ("cmax_s", c_dbl),
                ("d_cmin", vp), ("d_cmax", vp)]


# name -> (restype, argtypes); mirrors include/pde_b200.h one to one
SIGNATURES = {
    "pde_abi_version": (c_int, []),
    "pde_last_error": (C.c_char_p, []),
    "pde_kernel_launches": (c_i64, []),
    "pde_grid_create": (c_int, [C.POINTER(vp), c_int, c_i64, c_i64, c_i64, c_int,
                                c_i64, c_i64, c_i64, c_int, vp, vp, vp,
                                c_i64, vp, vp, vp, vp, vp]),
    "pde_grid_create_nd": (c_int, [C.POINTER(vp), c_int, c_int, vp, c_i64, c_int, c_int, vp, vp,
                                   c_i64, vp, vp, vp, vp]),
    "pde_grid_set_band_classes": (c_int, [vp, c_int, vp, vp, vp, vp, vp, vp]),
    "pde_grid_destroy": (c_int, [vp]),
    "pde_ddg_d2eigs_host": (c_int, [vp, vp, vp, vp, c_int, c_i64, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    "pde_grid_set_row_window": (c_int, [vp, c_i64, c_i64]),
    "pde_grid_set_row_windows2": (c_int, [vp, c_i64, c_i64, c_i64, c_i64]),
    "pde_grid_set_dynamic_schedule": (c_int, [vp, c_int]),
    "pde_grid_pitch": (c_i64, [vp]),
    "pde_grid_rows": (c_i64, [vp]),
    "pde_grid_state_len": (c_i64, [vp]),
    "pde_grid_policy_len": (c_i64, [vp]),
    "pde_grid_pack": (c_int, [vp, vp, vp, vp]),
    "pde_grid_unpack": (c_int, [vp, vp, vp, vp]),
    "pde_grid_h2d_rows": (c_int, [vp, vp, vp, c_i64, c_i64, c_int, vp]),
    "pde_grid_d2h_rows": (c_int, [vp, vp, vp, c_i64, c_i64, vp]),
    "pde_grid_unpad": (c_int, [vp, vp, vp, vp]),
    "pde_grid_unpad_u16": (c_int, [vp, vp, vp, vp]),
    "pde_ddg_d2eigs": (c_int, [vp, vp, vp, vp, vp, vp, c_int, vp]),
    "pde_ddg_apply_policy": (c_int, [vp, vp, vp, vp, vp]),
    "pde_ddg_control": (c_int, [vp, vp, vp, vp, vp, vp]),
    "pde_jac_matvec": (c_int, [vp, C.POINTER(pde_jac), vp, vp, c_int, vp]),
    "pde_jac_diagonal": (c_int, [vp, C.POINTER(pde_jac), vp, vp]),
    "pde_ce_residual": (c_int, [vp, vp, vp, vp, vp, vp, vp]),
    "pde_ce_euler": (c_int, [vp, vp, vp, vp, c_dbl, c_dbl, c_dbl, c_i64, c_i64, vp, vp]),
    "pde_ce_euler_step": (c_int, [vp, vp, vp, c_dbl, vp, vp, vp]),
    "pde_pucci_residual": (c_int, [vp, vp, vp, c_dbl, c_dbl, vp, vp, vp, vp, vp, vp, vp]),
    "pde_bicgstab": (c_int, [vp, C.POINTER(pde_jac), vp, vp, c_dbl, c_dbl, c_i64, c_i64,
                             vp, vp, vp]),
    "pde_bicgstab_scal_bytes": (c_i64, []),
    "pde_bicgstab_dist_step": (c_int, [vp, C.POINTER(pde_jac), c_int, vp, vp, vp, vp, vp, vp,
                                       c_dbl, c_dbl, c_i64, vp]),
    "pde_bicgstab_poll": (c_int, [vp, vp, vp]),
    "pde_vec_newton_update": (c_int, [vp, vp, vp, vp, vp, vp, vp, vp]),
    "pde_ell_create": (c_int, [C.POINTER(vp), c_int, c_i64, c_i64, c_int, vp, vp]),
    "pde_ell_destroy": (c_int, [vp]),
    "pde_ell_set_row_offset": (c_int, [vp, c_i64]),
    "pde_ell_d2eigs": (c_int, [vp, vp, vp, vp, vp, vp, vp]),
    "pde_ell_apply_policy": (c_int, [vp, vp, vp, vp, vp]),
    "pde_ell_build": (c_int, [c_int, c_i64, vp, vp, vp, c_int, vp, c_int, c_dbl, c_int, c_i64,
                              c_i64, vp, vp, vp, vp]),
    "pde_nb_d1": (c_int, [c_int, c_i64, c_i64, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    "pde_rows_apply": (c_int, [c_int, c_i64, c_i64, vp, vp, vp, vp, vp]),
    "pde_ell_residual": (c_int, [vp, vp, vp, c_int, c_dbl, c_dbl, vp, vp, vp, vp, vp]),
    "pde_ell_jac_matvec": (c_int, [vp, C.POINTER(pde_ell_jac), vp, vp, vp]),
    "pde_ell_jac_diagonal": (c_int, [vp, C.POINTER(pde_ell_jac), vp, vp]),
    "pde_ell_bicgstab": (c_int, [vp, C.POINTER(pde_ell_jac), vp, vp, c_dbl, c_dbl, c_i64, c_i64,
                                 vp, vp, vp]),
    "pde_ell_newton_update": (c_int, [vp, vp, vp, vp, vp, vp, vp, vp]),
    "pde_fp64_peak": (c_int, [c_int, C.POINTER(c_dbl)]),
    "pde_fp64_probe": (c_int, [c_int, c_int, C.POINTER(c_dbl)]),
    "pde_band_build": (c_int, [c_i64, c_i64, c_int, c_dbl, c_dbl, c_dbl, c_dbl, c_i64, vp,
                               c_i64, vp, c_dbl, C.POINTER(vp), C.POINTER(vp), C.POINTER(c_dbl),
                               C.POINTER(vp), C.POINTER(vp)]),
    "pde_halo_stage_doubles": (c_i64, [c_i64, c_i64]),
    "pde_halo_push": (c_int, [vp, c_i64, c_i64, c_i64, vp, vp, vp, C.c_uint64, vp]),
    "pde_halo_wait_copy": (c_int, [vp, c_i64, c_i64, c_i64, vp, c_int, c_int, C.c_uint64, vp]),
    "pde_ce_sweep_work_bytes": (c_i64, [c_i64, c_i64, c_int, vp, c_i64]),
    "pde_ce_sweep_setup": (c_int, [c_int, c_i64, c_i64, c_i64, c_i64, c_int, c_int, vp, vp, vp, c_i64,
                                   vp, vp]),
    "pde_ce_sweep": (c_int, [c_int, c_i64, c_i64, c_i64, c_i64, c_int, c_int, vp, vp, vp, c_i64, vp,
                             vp, vp, vp, vp, vp, c_dbl, c_dbl, c_i64, c_int, c_int, vp, vp]),
    "pde_band_build_nd": (c_int, [c_int, vp, c_int, vp, vp, c_i64, vp, c_i64, vp, c_dbl,
                                  C.POINTER(vp), C.POINTER(vp), C.POINTER(c_dbl),
                                  C.POINTER(vp), C.POINTER(vp)]),
    "pde_free": (None, [vp]),
    "pde_set_host_threads": (c_int, [c_int]),
    "pde_cloud_build": (c_int, [c_int, c_i64, c_i64, vp, vp, vp, c_int, c_dbl, c_dbl, c_dbl, c_int,
                                c_int, vp, vp, vp, vp, vp]),
}


class PdeError(RuntimeError):
    pass


def lib():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PdeError(
            "libpde_b200.so is not built (%s). Run `python -m pyellipticfd_b200._build`; "
            "pyellipticfd_b200 has no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(L, name)
        except AttributeError:
            continue            # optional symbol groups are checked by the caller / tests
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def check(rc):
    if rc != 0:
        msg = lib().pde_last_error()
        raise PdeError((msg or b"unknown error").decode() + " [rc=%d]" % rc)


def kernel_launches():
    return int(lib().pde_kernel_launches())
