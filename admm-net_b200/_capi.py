"""ctypes loader for libadmmnet_b200.so (include/admmnet_b200.h).  Fails loudly: there is no CPU or
PyTorch fallback behind this boundary."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libadmmnet_b200.so")

_vp, _i, _d, _sz = C.c_void_p, C.c_int, C.c_double, C.c_size_t
_PROTOS = {
    "admmnet_last_error": (C.c_char_p, []),
    "admmnet_version": (_i, []),
    "admmnet_param_stride": (_i, [_i]),
    "admmnet_forward_workspace_bytes": (_i, [_i, _i, _i, _i, _i, C.POINTER(_sz)]),
    "admmnet_forward": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _sz, _i, _vp]),
    "admmnet_layer_chunk": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _sz, _i, _vp]),
    "admmnet_layer": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _sz, _i, _vp]),
    "admmnet_layer_rsum": (_i, [_vp, _sz, _i, _i, _i, _i, _i, _i, _vp]),
    "admmnet_ws_scalars": (_i, [_vp, _sz, _i, _i, _i, _i, _i, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp)]),
    "admmnet_set_mean": (_i, [_vp, _sz, _i, _i, _i, _i, _i, _i, _d, _vp]),
    "admmnet_final_phi": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _sz, _i, _vp]),
    "admmnet_reset_status": (_i, [_vp, _sz, _i, _i, _i, _i, _i, _vp]),
    "admmnet_status": (_i, [_vp, _sz, _i, _i, _i, _i, _i, _vp, C.POINTER(_i)]),
    "admmnet_eigh_workspace_bytes": (_i, [_i, _i, _i, C.POINTER(_sz)]),
    "admmnet_eigh_batched": (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _sz, _i, _vp, _vp]),
    "admm_classic_forward": (_i, [_vp, _vp, _i, _i, _i, _d, _i, _vp, _vp]),
    "peak_search_full": (_i, [_vp, _i, _i, _i, _i, _vp, _i, _vp, _i, _d, _d, _d, _d, _d, _d, _d, _i, _i, _vp, _vp, _i,
                              _vp, _vp, _vp, _vp]),
    "peak_surface_maxima": (_i, [_vp, _i, _vp, _i, _vp, _i, _i, _vp, _vp, _vp, _vp]),
    "peak_search_points": (_i, [_vp, _i, _i, _i, _vp, _vp, _i, _vp, _vp]),
    "admmnet_head_param_count": (_i, [_i, _i]),
    "admmnet_peak_head": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "admmnet_arrow_eigh": (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp]),
    "admmnet_generate": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _d, _d, C.c_ulonglong, _vp]),
    "admmnet_generate_dataset": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _d, _d, _d, C.c_ulonglong, _vp]),
    "admmnet_profile_begin": (_i, []),
    "admmnet_profile_kinds": (_i, []),
    "admmnet_profile_kind_name": (C.c_char_p, [_i]),
    "admmnet_profile_end": (_i, [C.POINTER(_d), C.POINTER(C.c_longlong)]),
    "admmnet_tail_tc_smem_bytes": (_i, [_i]),
    "admmnet_tail_tc_mma_flops": (_d, [_i]),
    "admmnet_tail_tc_profile_read": (_i, [_vp]),
    "admmnet_dc_profile_read": (_i, [_vp]),
    "admmnet_tc_gemm_probe": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "admmnet_fp32_peak_launch": (_i, [_vp, _i, _i, C.POINTER(_d), _vp]),
}
EXPORTS = tuple(_PROTOS)

_lib = None


class AdmmnetError(RuntimeError):
    pass


def lib():
    """The loaded library; raises if it has not been built (python admm-net_b200/build.py)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise AdmmnetError(
                f"{LIB_PATH} is missing: the CUDA extension has not been built. Run "
                "`python -c 'import __graft_entry__ as g; g.build()'` (nvcc, sm_100a). There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _PROTOS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(code):
    if code != 0:
        raise AdmmnetError(f"admmnet_b200 error {code}: {lib().admmnet_last_error().decode()}")


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise AdmmnetError("admmnet_b200 needs a CUDA device (sm_100a); there is no CPU fallback on the product path")
