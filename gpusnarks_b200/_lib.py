"""ctypes loader for libgpusnarks_b200.so.  Fails loudly: a missing library is an ImportError,
a missing GPU is an error from gsn_ctx_create -- there is no CPU path behind this module."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GSN_LIB") or os.path.join(HERE, "libgpusnarks_b200.so")  # GSN_LIB: experiment builds

# every symbol include/gpusnarks_b200.h declares (tests check that the .so exports them all)
SYMBOLS = [
    "gsn_ctx_create", "gsn_ctx_destroy", "gsn_last_error", "gsn_set_field768", "gsn_ctx_trim", "gsn_launch_count",
    "gsn_ntt768_host", "gsn_ntt768_host_batch", "gsn_ntt768_device", "gsn_ntt768_prepare", "gsn_ntt768_strided_device",
    "gsn_ntt768_device_ex", "gsn_fourstep_table768", "gsn_ntt768_device_scatter", "gsn_peer_barrier", "gsn_ipc_export", "gsn_ipc_import", "gsn_ipc_close", "gsn_fp768_binop_host", "gsn_fp768_binop_device", "gsn_fp768_powers_device", "gsn_fp768_twiddle_table_device", "gsn_fp768_inner_product_device", "gsn_fp768_inner_product_host", "gsn_g1_multiexp_host", "gsn_g1_multiexp_multi_host", "gsn_g1_multiexp_device", "gsn_g1_multiexp_device_ex", "gsn_fp2_binop_host", "gsn_fp2_binop_device", "gsn_ntt32_host", "gsn_ntt32_device",
    "gsn_device_count", "gsn_host_alloc", "gsn_host_free", "gsn_device_alloc", "gsn_device_free",
    "gsn_memcpy_h2d", "gsn_memcpy_d2h", "gsn_ctx_synchronize", "gsn_int32_issue_rates",
    "gsn_ntt768_time_device", "gsn_ntt32_time_device",
    "gsn_ctx_set_option", "gsn_ntt768_plan_info", "gsn_coset_ntt768_device", "gsn_coset_ntt768_host",
    "gsn_fourstep_create", "gsn_fourstep_destroy", "gsn_fourstep_info", "gsn_fourstep_buffers", "gsn_fourstep_connect",
    "gsn_fourstep_forward", "gsn_fourstep_inverse", "gsn_fourstep_phase_ms", "gsn_fourstep_set_timing",
    "gsn_multi_create", "gsn_multi_destroy", "gsn_multi_ntt768_host", "gsn_multi_device_buffers", "gsn_multi_ntt768_device",
    "gsn_multi_synchronize",
]

_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH) and not os.environ.get("GSN_LIB"):
        try:  # a fresh checkout: build in-tree once (nvcc, sm_100a); failure falls through to the loud error below
            from . import build as _build
            _build.build()
        except Exception:
            pass
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m gpusnarks_b200.build` "
            "(nvcc, sm_100a).  gpusnarks_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, u32p, sz, i, u32, u64p = C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_uint32, C.POINTER(C.c_uint64)
    L.gsn_last_error.restype = C.c_char_p
    L.gsn_ctx_create.argtypes = [C.POINTER(vp), i]
    L.gsn_ctx_destroy.argtypes = [vp]
    L.gsn_set_field768.argtypes = [vp, i]
    L.gsn_ctx_trim.argtypes = [vp]
    L.gsn_launch_count.argtypes = [vp, u64p]
    L.gsn_ntt768_host.argtypes = [vp, u32p, sz, u32p, i]
    L.gsn_ntt768_host_batch.argtypes = [vp, C.POINTER(vp), sz, sz, u32p, i]
    L.gsn_ntt768_device.argtypes = [vp, vp, sz, sz, u32p, i, vp]
    L.gsn_ntt768_prepare.argtypes = [vp, sz, sz, u32p, i]
    L.gsn_ntt768_strided_device.argtypes = [vp, vp, sz, sz, C.c_uint, u32p, i, vp]
    L.gsn_ntt768_device_ex.argtypes = [vp, vp, sz, sz, C.c_uint, u32p, C.c_uint, vp, vp]
    L.gsn_ntt768_device_scatter.argtypes = [vp, vp, sz, sz, C.c_uint, u32p, C.c_uint, vp, C.POINTER(vp), C.c_uint, C.c_uint, C.c_uint, C.c_uint, vp]
    L.gsn_peer_barrier.argtypes = [vp, C.POINTER(vp), C.c_uint, C.c_uint, C.c_uint, vp]
    L.gsn_ipc_export.argtypes = [vp, vp, C.c_char_p]
    L.gsn_ipc_import.argtypes = [vp, C.c_char_p, C.POINTER(vp)]
    L.gsn_ipc_close.argtypes = [vp, vp]
    L.gsn_fourstep_table768.argtypes = [vp, vp, sz, sz, sz, sz, sz, u32p, C.c_uint, vp]
    L.gsn_fp768_binop_host.argtypes = [vp, i, u32p, u32p, u32p, sz]
    L.gsn_fp768_binop_device.argtypes = [vp, i, vp, vp, vp, sz, vp]
    L.gsn_fp768_powers_device.argtypes = [vp, vp, sz, u32p, u32p, vp]
    L.gsn_fp768_twiddle_table_device.argtypes = [vp, vp, vp, sz, vp]
    L.gsn_fp768_inner_product_device.argtypes = [vp, vp, vp, vp, sz, vp]
    L.gsn_fp768_inner_product_host.argtypes = [vp, u32p, u32p, u32p, sz]
    L.gsn_g1_multiexp_host.argtypes = [vp, u32p, u32p, u32p, sz]
    L.gsn_g1_multiexp_multi_host.argtypes = [C.POINTER(C.c_int), C.c_uint, u32p, u32p, u32p, sz]
    L.gsn_g1_multiexp_device.argtypes = [vp, vp, vp, vp, sz, vp]
    L.gsn_fp2_binop_host.argtypes = [vp, i, u32p, u32p, u32p, sz]
    L.gsn_fp2_binop_device.argtypes = [vp, i, vp, vp, vp, sz, vp]
    L.gsn_g1_multiexp_device_ex.argtypes = [vp, vp, vp, vp, sz, C.c_uint, C.c_uint, vp]
    L.gsn_ntt32_host.argtypes = [vp, u32p, sz, u32, u32, i]
    L.gsn_ntt32_device.argtypes = [vp, vp, sz, sz, u32, u32, i, vp]
    L.gsn_device_count.argtypes = [C.POINTER(i)]
    L.gsn_host_alloc.argtypes = [C.POINTER(vp), sz]
    L.gsn_host_free.argtypes = [vp]
    L.gsn_device_alloc.argtypes = [vp, C.POINTER(vp), sz]
    L.gsn_device_free.argtypes = [vp, vp]
    L.gsn_memcpy_h2d.argtypes = [vp, vp, vp, sz]
    L.gsn_memcpy_d2h.argtypes = [vp, vp, vp, sz]
    L.gsn_ctx_synchronize.argtypes = [vp]
    L.gsn_int32_issue_rates.argtypes = [vp, C.POINTER(C.c_double), i, C.POINTER(i), C.POINTER(i), C.POINTER(i)]
    L.gsn_ntt768_time_device.argtypes = [vp, vp, sz, sz, u32p, i, i, C.POINTER(C.c_float)]
    L.gsn_ntt32_time_device.argtypes = [vp, vp, sz, sz, u32, u32, i, i, C.POINTER(C.c_float)]
    u64 = C.c_uint64
    pu = C.POINTER(C.c_uint)
    L.gsn_ctx_set_option.argtypes = [vp, i, u64]
    L.gsn_ntt768_plan_info.argtypes = [vp, sz, u32p, i, u64p, pu, pu, u64p, u64p]
    L.gsn_coset_ntt768_device.argtypes = [vp, vp, sz, sz, u32p, u32p, i, vp]
    L.gsn_coset_ntt768_host.argtypes = [vp, u32p, sz, u32p, u32p, i]
    L.gsn_fourstep_create.argtypes = [vp, C.POINTER(vp), C.c_uint, u32p, C.c_uint, C.c_uint, C.c_uint]
    L.gsn_fourstep_destroy.argtypes = [vp]
    L.gsn_fourstep_info.argtypes = [vp, pu, pu, pu, u64p, pu]
    L.gsn_fourstep_buffers.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    L.gsn_fourstep_connect.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    L.gsn_fourstep_forward.argtypes = [vp, vp, C.POINTER(vp)]
    L.gsn_fourstep_inverse.argtypes = [vp, vp, C.POINTER(vp)]
    L.gsn_fourstep_phase_ms.argtypes = [vp, C.POINTER(C.c_float), u64p]
    L.gsn_fourstep_set_timing.argtypes = [vp, i]
    L.gsn_multi_create.argtypes = [C.POINTER(vp), C.POINTER(i), C.c_uint, sz, u32p, C.c_uint]
    L.gsn_multi_destroy.argtypes = [vp]
    L.gsn_multi_ntt768_host.argtypes = [vp, u32p, i]
    L.gsn_multi_device_buffers.argtypes = [vp, C.c_uint, C.POINTER(vp), C.POINTER(vp)]
    L.gsn_multi_ntt768_device.argtypes = [vp, i]
    L.gsn_multi_synchronize.argtypes = [vp]
    for s in SYMBOLS:
        if s != "gsn_last_error":
            getattr(L, s).restype = i
    _lib = L
    return L
