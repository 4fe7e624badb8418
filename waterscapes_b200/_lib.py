"""ctypes binding of libmpet_b200.so (the C-ABI in include/mpet_b200.h).

There is deliberately NO fallback: if the CUDA library is missing or fails to load, importing the
product path raises.  PyTorch is used only to own device buffers and streams.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libmpet_b200.so")

_c_ctx = C.c_void_p
_p = C.c_void_p
_i64 = C.c_int64
_f64 = C.c_double
_int = C.c_int

# name -> (restype, argtypes); kept in the same order as include/mpet_b200.h
PROTOTYPES = {
    "mpet_create": (_int, [_int, C.POINTER(_c_ctx)]),
    "mpet_destroy": (None, [_c_ctx]),
    "mpet_last_error": (C.c_char_p, [_c_ctx]),
    "mpet_abi_version": (_int, []),
    "mpet_set_mesh": (_int, [_c_ctx, _p, _p, _i64, _i64, _int, _p]),
    "mpet_get_sizes": (_int, [_c_ctx, C.POINTER(_i64)]),
    "mpet_get_edges": (_int, [_c_ctx, _p, _p]),
    "mpet_get_cell_dofs": (_int, [_c_ctx, _p, _p]),
    "mpet_get_pattern": (_int, [_c_ctx, _p, _p, _p]),
    "mpet_set_params": (_int, [_c_ctx, _f64, _f64, C.POINTER(_f64), C.POINTER(_f64), C.POINTER(_f64),
                               C.POINTER(_f64), _f64, _f64]),
    "mpet_set_params_total_pressure": (_int, [_c_ctx, _f64, _f64, C.POINTER(_f64), C.POINTER(_f64), C.POINTER(_f64),
                                              C.POINTER(_f64), _f64, _f64]),
    "mpet_set_cell_coefficient": (_int, [_c_ctx, _int, _p, _p]),
    "mpet_set_dof_permutation": (_int, [_c_ctx, _p, _p]),
    "mpet_assemble_lhs": (_int, [_c_ctx, _p]),
    "mpet_add_entries": (_int, [_c_ctx, _p, _p, _p, _i64, _p]),
    "mpet_assemble_prec": (_int, [_c_ctx, _p]),
    "mpet_get_values": (_int, [_c_ctx, _int, _p, _p]),
    "mpet_set_dirichlet_dofs": (_int, [_c_ctx, _p, _i64, _p]),
    "mpet_set_dirichlet_values": (_int, [_c_ctx, _p, _p]),
    "mpet_rhs_prev": (_int, [_c_ctx, _p, _p, _p]),
    "mpet_mass_apply": (_int, [_c_ctx, _int, _f64, _p, _p, _p]),
    "mpet_lumped": (_int, [_c_ctx, _int, _p, _p]),
    "mpet_apply_dirichlet_rhs": (_int, [_c_ctx, _p, _p]),
    "mpet_spmv": (_int, [_c_ctx, _p, _p, _p]),
    "mpet_csr_spmv": (_int, [_c_ctx, _i64, _p, _p, _p, _p, _p, _f64, _p]),
    "mpet_krylov_setup": (_int, [_c_ctx, _int, _int, _f64, _f64, _int, _int]),
    "mpet_krylov_reference_norm": (_int, [_c_ctx, _int]),
    "mpet_pc_setup": (_int, [_c_ctx, _p]),
    "mpet_solve": (_int, [_c_ctx, _p, _p, C.POINTER(_f64), _p]),
    "mpet_pc_apply": (_int, [_c_ctx, _p, _p, _p]),
    "mpet_set_border": (_int, [_c_ctx, _int, _p, _p]),
    "mpet_set_prec_shift": (_int, [_c_ctx, _f64, C.POINTER(_f64)]),
    "mpet_attach_comm": (_int, [_c_ctx, _p, _int, _int]),
    "mpet_nccl_unique_id": (_int, [_p]),
    "mpet_set_halo": (_int, [_c_ctx, _int, C.POINTER(_int), C.POINTER(_i64), _p, C.POINTER(_i64), _p, _p, _p]),
    "mpet_launch_count": (_i64, [_c_ctx, _int]),
    "mpet_profile": (_int, [_c_ctx, _int, C.POINTER(_f64)]),
    "mpet_device_bytes": (_i64, [_c_ctx]),
    "mpet_pc_bytes": (_i64, [_c_ctx]),
    "mpet_comm_kind": (_int, [_c_ctx]),
}

_lib = None


class MpetLibraryError(RuntimeError):
    pass


def load():
    """Load the shared library (once) and attach prototypes.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MpetLibraryError(
            "%s not found: build it with `python -m waterscapes_b200.build` "
            "(there is no CPU fallback for the MPET hot path)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)          # AttributeError if a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(ctx, rc):
    if rc != 0:
        msg = load().mpet_last_error(ctx)
        raise MpetLibraryError(msg.decode() if msg else "libmpet_b200 error %d" % rc)
