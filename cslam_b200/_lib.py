"""ctypes binding of libcslam_b200.so (the C ABI declared in include/cslam_b200.h).

There is deliberately no fallback: if the shared object is missing, or a
compute entry point is called without a CUDA device, an exception is raised.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcslam_b200.so")

OK = 0
ERR_INVALID, ERR_CUDA, ERR_OOM, ERR_LIMIT, ERR_SINGULAR, ERR_NOCONV = -1, -2, -3, -4, -5, -6
DTYPE_F32, DTYPE_F64 = 0, 1


class CslamError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"libcslam_b200 error {status}: {message}")
        self.status = status


class SingularLaplacianError(CslamError):
    """Raised where the reference's SuperLU factorisation raises
    RuntimeError('Factor is exactly singular') (cslam/mac/mac.py:52-58)."""


_c = ctypes
_vp, _i, _i64, _f, _d = _c.c_void_p, _c.c_int, _c.c_int64, _c.c_float, _c.c_double
_P = _c.POINTER

# name -> (restype, argtypes).  Must list every symbol declared in include/cslam_b200.h
# (tests/test_abi.py checks the two against each other).
SIGNATURES = {
    "cslam_version": (_c.c_char_p, []),
    "cslam_last_error": (_c.c_char_p, []),
    "cslam_device_count": (_i, []),
    "cslam_launch_count": (_i64, []),
    # A6 NNS
    "cslam_nns_create": (_i, [_i, _i, _P(_vp)]),
    "cslam_nns_destroy": (_i, [_vp]),
    "cslam_nns_add_host": (_i, [_vp, _vp, _i, _i64]),
    "cslam_nns_add_device": (_i, [_vp, _vp, _i64, _vp]),
    "cslam_nns_size": (_i64, [_vp]),
    "cslam_nns_dim": (_i, [_vp]),
    "cslam_nns_read_rows": (_i, [_vp, _i64, _i64, _vp]),
    "cslam_nns_search_host": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp]),
    "cslam_nns_search_device": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "cslam_nns_set_mode": (_i, [_vp, _i]),
    "cslam_nns_set_sample_rows": (_i, [_vp, _i]),
    "cslam_nns_last_timing": (_i, [_vp, _P(_f), _P(_i), _P(_f)]),
    # A11-A15 MAC
    "cslam_mac_create": (_i, [_i, _i64, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _i, _P(_vp)]),
    "cslam_mac_destroy": (_i, [_vp]),
    "cslam_mac_set_options": (_i, [_vp, _d, _i, _i]),
    "cslam_mac_fiedler": (_i, [_vp, _vp, _P(_d), _vp, _P(_i)]),
    "cslam_mac_grad": (_i, [_vp, _vp, _vp]),
    "cslam_mac_fw_subset": (_i, [_vp, _vp, _i, _i, _d, _vp, _vp, _P(_d), _P(_i), _vp, _vp]),
    "cslam_mac_fw_subset_sparse": (_i, [_vp, _i64, _vp, _vp, _i, _i, _d, _vp, _i64, _vp, _vp, _P(_i64),
                                        _P(_d), _P(_i), _vp, _vp]),
    "cslam_mac_stats": (_i, [_vp, _P(_i64), _P(_i64), _P(_i)]),
    "cslam_mac_solver_timing": (_i, [_vp, _P(_d), _P(_i64), _P(_i64), _P(_i64)]),
    "cslam_keymap_create": (_i, [_i64, _i, _P(_vp)]),
    "cslam_keymap_destroy": (_i, [_vp]),
    "cslam_keymap_size": (_i64, [_vp]),
    "cslam_keymap_lookup": (_i, [_vp, _vp, _i64, _vp]),
    "cslam_keymap_insert": (_i, [_vp, _vp, _vp, _i64]),
    "cslam_keymap_erase": (_i, [_vp, _vp, _i64, _vp]),
    "cslam_swarm_hits": (_i, [_i, _i, _i, _vp, _vp, _vp, _d, _vp, _i, _vp]),
    "cslam_swarm_intra": (_i, [_i, _i, _i, _i64, _vp, _vp, _vp, _vp, _vp]),
    "cslam_debug_grid_barrier": (_i, [_i, _i, _i, _i, _i, _i, _vp]),
    "cslam_debug_rayleigh_ritz": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _P(_i), _vp]),
    "cslam_fiedler_csr": (_i, [_i, _vp, _vp, _vp, _d, _i, _i, _P(_d), _vp, _P(_i)]),
    # A1-A5 descriptor extraction
    "cslam_preproc_create": (_i, [_i, _i, _i, _i, _i, _P(_vp)]),
    "cslam_preproc_destroy": (_i, [_vp]),
    "cslam_preproc_run": (_i, [_vp, _vp, _i, _vp, _vp]),
    "cslam_vlad_forward": (_i, [_vp, _i, _i, _i, _vp, _vp, _i, _vp, _vp]),
    "cslam_pca_workspace_floats": (_i64, [_i, _i]),
    "cslam_pca_project_l2": (_i, [_vp, _i, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp]),
    "cslam_gem_head_forward": (_i, [_vp, _i, _i, _i, _f, _f, _vp, _vp, _i, _vp, _vp]),
    # (f4) Scan Context matching
    "cslam_sc_create": (_i, [_i, _i, _i, _i, _P(_vp)]),
    "cslam_sc_destroy": (_i, [_vp]),
    "cslam_sc_size": (_i64, [_vp]),
    "cslam_sc_capacity": (_i64, [_vp]),
    "cslam_sc_add_host": (_i, [_vp, _vp, _i, _i64]),
    "cslam_sc_read": (_i, [_vp, _i64, _i64, _vp, _vp]),
    "cslam_sc_search_host": (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "cslam_sc_last_timing": (_i, [_vp, _P(_f), _P(_f)]),
}

_lib = None


def load():
    """Load the shared library (once) and declare every signature."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing. Build it with `python -m cslam_b200.build` "
            "(nvcc, sm_100a). cslam_b200 has no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error():
    return load().cslam_last_error().decode(errors="replace")


def check(status):
    if status == OK:
        return
    msg = last_error()
    if status == ERR_SINGULAR:
        raise SingularLaplacianError(status, msg)
    raise CslamError(status, msg)


def device_count():
    return int(load().cslam_device_count())


def require_device():
    if device_count() <= 0:
        raise CslamError(ERR_CUDA, "no CUDA device visible; cslam_b200 has no CPU fallback")


def launch_count():
    return int(load().cslam_launch_count())


def ptr(a):
    """Raw data pointer of a numpy array / torch tensor / int."""
    if a is None:
        return None
    if isinstance(a, int):
        return ctypes.c_void_p(a)
    if hasattr(a, "data_ptr"):
        return ctypes.c_void_p(a.data_ptr())
    return ctypes.c_void_p(a.ctypes.data)
