"""ctypes binding of libhtf_b200.so (include/htf_b200.h).

This is the reference-side stub a maintainer would add in place of
``from hoomd.htf import _htf`` (/root/reference htf/tensorflowcompute.py:3) and of
``load_htf_op_library`` (htf/simmodel.py:696-711).  It fails loudly when the library is
missing: there is no CPU fallback.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HTF_B200_LIB") or os.path.join(os.path.dirname(_HERE), "lib", "libhtf_b200.so")
ABI_VERSION = 18

OK, EINVAL, ECUDA, ENOMEM, ESTATE, ESKEW, EARCH = 0, -1, -2, -3, -4, -5, -6
FLAG_DETERMINISTIC = 1
FLAG_ANY_ARCH = 2

# every symbol include/htf_b200.h declares: name -> (restype, argtypes)
_vp, _i64, _i32, _f32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_float
_fp3 = ctypes.POINTER(ctypes.c_float)
SYMBOLS = {
    "htf_abi_version": (_i32, []),
    "htf_create": (_i32, [ctypes.POINTER(_vp), _i32, _i64, _i32, _f32, _i32]),
    "htf_destroy": (None, [_vp]),
    "htf_last_error": (ctypes.c_char_p, [_vp]),
    "htf_set_box": (_i32, [_vp, _fp3, _fp3, _fp3]),
    "htf_set_roi": (_i32, [_vp, _fp3, _fp3]),
    "htf_pack_halo": (_i32, [_vp, _vp, _i64, _i32, _f32, _i32, _vp, _i64, _vp, _vp, _vp]),
    "htf_eds_step": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _f32, _f32, _vp]),
    "htf_integrate_half": (_i32, [_vp, _i32, _vp, _vp, _vp, _i64, _f32, _f32, _f32, _i32, ctypes.c_uint64, ctypes.c_uint64, _vp]),
    "htf_skin_configure": (_i32, [_vp, _f32, _i32]),
    "htf_skin_rebuild": (_i32, [_vp, _vp, _i64, _i64, _i64, _vp]),
    "htf_skin_nlist": (_i32, [_vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp]),
    "htf_skin_status": (_i32, [_vp, _vp, _i32, _vp]),
    "htf_pack_halo_pair": (_i32, [_vp, _vp, _i64, _i32, _f32, _f32, _vp, _vp, _i64, _vp, _vp, _vp]),
    "htf_set_mapped_nlist": (_i32, [_vp, _i32]),
    "htf_set_cutoff": (_i32, [_vp, _f32, _i32]),
    "htf_bin_particles": (_i32, [_vp, _vp, _i64, _vp]),
    "htf_build_nlist": (_i32, [_vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp]),
    "htf_lj_forces": (_i32, [_vp, _vp, _i64, _i32, _vp, _vp, _vp, _i32, _vp]),
    "htf_lj_forces_rdf": (_i32, [_vp, _vp, _i64, _i32, _vp, _vp, _vp, _i32, _vp, _f32, _f32, _i32, _vp]),
    "htf_lj_cv_forces": (_i32, [_vp, _vp, _i64, _i32, _vp, _f32, _vp, _vp, _i32, _vp, _vp, _vp, _f32, _f32, _i32, _vp]),
    "htf_mlp_param_sizes": (_i32, [ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]),
    "htf_mlp_pack": (_i32, [_vp, _vp, _vp, _vp]),
    "htf_mlp_forces": (_i32, [_vp, _vp, _i64, _i32, _vp, _vp, _f32, _vp, _vp]),
    "htf_rdf_hist": (_i32, [_vp, _vp, _i64, _i32, _vp, _i64, _f32, _f32, _i32, _i32, _i32, _vp, _vp]),
    "htf_lj_step": (_i32, [_vp, _vp, _i64, _i64, _i64, _vp, _vp, _vp, _i32, _vp, _vp, _f32, _f32, _i32, _vp]),
    "htf_unstuff4": (_i32, [_vp, _vp, _vp, _i64, _vp]),
    "htf_lj_rows": (_i32, [_vp, _i64, _i64, _i64, _vp, _vp, _vp, _i32, _vp, _vp, _f32, _f32, _i32, _vp]),
    "htf_lj_cv_step": (_i32, [_vp, _vp, _i64, _i64, _i64, _vp, _f32, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _f32, _f32, _i32, _vp]),
    "htf_set_pipeline": (_i32, [_vp, _i32]),
    "htf_mlp_train_grads": (_i32, [_vp, _vp, _i64, _i32, _vp, _f32, _vp, _i64, _vp, _vp, _vp, _vp]),
    "htf_adam_step": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _f32, _f32, _f32, _f32, _vp]),
    "htf_comm_create": (_i32, [_vp, _i32, _i32, _i64, ctypes.c_char_p]),
    "htf_comm_connect": (_i32, [_vp, ctypes.c_char_p]),
    "htf_comm_exchange_halo": (_i32, [_vp, _vp, _i64, _i32, _f32, _f32, _vp, _vp]),
    "htf_comm_allreduce_i64": (_i32, [_vp, _vp, _i32, _vp]),
    "htf_comm_allreduce_f64": (_i32, [_vp, _vp, _i32, _vp]),
    "htf_comm_allreduce_f32": (_i32, [_vp, _vp, _i32, _vp]),
    "htf_comm_status": (_i32, [_vp, ctypes.POINTER(ctypes.c_int32), _vp]),
    "htf_comm_destroy": (_i32, [_vp]),
    "htf_launch_count": (_i64, [_vp]),
    "htf_get_cell_grid": (_i32, [_vp, ctypes.POINTER(ctypes.c_int)]),
}

_lib = None


class HtfError(RuntimeError):
    """A libhtf_b200 call returned a negative status (the C ABI never throws)."""

    def __init__(self, code, message):
        super().__init__("libhtf_b200 error %d: %s" % (code, message))
        self.code = code


def load(path=None):
    """dlopen the library and type every exported symbol.  Raises if it is absent or stale."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise ImportError(
            "libhtf_b200.so not found at %s -- build it with `python hoomd-tf_b200/build.py` "
            "(there is no CPU fallback for the hot path)" % p)
    L = ctypes.CDLL(p)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(L, name)          # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    if L.htf_abi_version() != ABI_VERSION:
        raise ImportError("libhtf_b200.so ABI %d != expected %d; rebuild" % (L.htf_abi_version(), ABI_VERSION))
    if path is None:
        _lib = L
    return L


def check(ctx, rc):
    if rc != OK:
        msg = load().htf_last_error(ctx)
        raise HtfError(rc, msg.decode("utf-8", "replace") if msg else "?")
    return rc
