"""ctypes face of oracle/liboracle.so (CPU oracle) and oracle/_ref/*.so (the reference
compiled from /root/reference).  TEST INFRASTRUCTURE: only tests/, bench.py's cpu_baseline
leg and __graft_entry__.smoke() may import this."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")


def build_oracle(force=False):
    so = os.path.join(ORACLE_DIR, "liboracle.so")
    if force or not os.path.exists(so):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "liboracle.so"])
    return so


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build_oracle())
        L.oracle_set_field768.argtypes = [C.c_int]
        L.oracle_set_mod32.argtypes = [C.c_uint32]
        L.oracle_set_threads.argtypes = [C.c_int]
        L.oracle_max_threads.restype = C.c_int
        L.oracle_fp768_binop.argtypes = [C.c_int, _u32p, _u32p, _u32p, C.c_size_t]
        L.oracle_fp768_pow.argtypes = [_u32p, _u32p, C.c_uint64]
        L.oracle_fp768_inverse.argtypes = [_u32p, _u32p]
        L.oracle_fft768.argtypes = [_u32p, C.c_size_t, _u32p, C.c_int]
        L.oracle_ifft768.argtypes = [_u32p, C.c_size_t, _u32p, C.c_int]
        L.oracle_naive_dft768.argtypes = [_u32p, C.c_size_t, _u32p]
        L.oracle_dft_points768.argtypes = [_u32p, _u32p, C.c_size_t, _u32p, _u64p, C.c_size_t]
        L.oracle_dft_points768_mt.argtypes = [_u32p, _u32p, C.c_size_t, _u32p, _u64p, C.c_size_t]
        L.oracle_fft32.argtypes = [_u32p, C.c_size_t, C.c_uint32, C.c_int]
        L.oracle_ifft32.argtypes = [_u32p, C.c_size_t, C.c_uint32, C.c_int]
        L.oracle_naive_dft32.argtypes = [_u32p, C.c_size_t, C.c_uint32]
        L.oracle_dft_points32.argtypes = [_u32p, _u32p, C.c_size_t, C.c_uint32, _u64p, C.c_size_t]
        L.oracle_time_fft768.argtypes = [_u32p, C.c_size_t, _u32p, C.c_int]
        L.oracle_time_fft768.restype = C.c_double
        L.oracle_time_fft32.argtypes = [_u32p, C.c_size_t, C.c_uint32, C.c_int]
        L.oracle_time_fft32.restype = C.c_double
        _lib = L
    return _lib


def _c(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def set_field768(which):
    assert lib().oracle_set_field768({"fr": 0, "fq": 1}[which]) == 0


def fp768_binop(op, a, b):
    a, b = _c(a).reshape(-1, 24), _c(b).reshape(-1, 24)
    out = np.empty_like(a)
    lib().oracle_fp768_binop({"mul": 0, "add": 1, "sub": 2}[op], out, a, b, a.shape[0])
    return out


def fp768_pow(a, e):
    out = np.empty(24, dtype=np.uint32)
    lib().oracle_fp768_pow(out, _c(a), int(e))
    return out


def fp768_inverse(a):
    out = np.empty(24, dtype=np.uint32)
    lib().oracle_fp768_inverse(out, _c(a))
    return out


def fft768(a, omega, log_cpus=-1, inverse=False):
    v = _c(a).reshape(-1, 24).copy()
    (lib().oracle_ifft768 if inverse else lib().oracle_fft768)(v, v.shape[0], _c(omega), log_cpus)
    return v


def naive_dft768(a, omega):
    v = _c(a).reshape(-1, 24).copy()
    lib().oracle_naive_dft768(v, v.shape[0], _c(omega))
    return v


def dft_points768(a, omega, ks):
    v = _c(a).reshape(-1, 24)
    ks = np.ascontiguousarray(ks, dtype=np.uint64)
    out = np.empty((len(ks), 24), dtype=np.uint32)
    lib().oracle_dft_points768(out, v, v.shape[0], _c(omega), ks, len(ks))
    return out


def dft_points768_mt(a, omega, ks):
    """dft_points768 with the index range of every evaluation split over all host threads"""
    v = _c(a).reshape(-1, 24)
    ks = np.ascontiguousarray(ks, dtype=np.uint64)
    out = np.empty((len(ks), 24), dtype=np.uint32)
    lib().oracle_dft_points768_mt(out, v, v.shape[0], _c(omega), ks, len(ks))
    return out


def fft32(a, omega, mod, log_cpus=-1, inverse=False):
    lib().oracle_set_mod32(int(mod))
    v = _c(a).copy()
    (lib().oracle_ifft32 if inverse else lib().oracle_fft32)(v, v.shape[0], int(omega), log_cpus)
    return v


def naive_dft32(a, omega, mod):
    lib().oracle_set_mod32(int(mod))
    v = _c(a).copy()
    lib().oracle_naive_dft32(v, v.shape[0], int(omega))
    return v


def dft_points32(a, omega, mod, ks):
    lib().oracle_set_mod32(int(mod))
    v = _c(a)
    ks = np.ascontiguousarray(ks, dtype=np.uint64)
    out = np.empty(len(ks), dtype=np.uint32)
    lib().oracle_dft_points32(out, v, v.shape[0], int(omega), ks, len(ks))
    return out
