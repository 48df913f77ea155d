"""Host-side Python face of the C ABI (include/gpusnarks_b200.h).

Mirrors the reference's operator interface for the FFT path: `best_fft(a, omega)` transforms
a vector of field elements in place given a root of unity (reference cuda/fft_kernel.h:24-25,
test/main.cpp:57).  Elements are numpy uint32 arrays: shape (n, 24) for the 768-bit field
(`fields::Scalar::im_rep`, little-endian limbs, Montgomery form) and shape (n,) for the 32-bit
field (`dummy_fields::Field::im_rep`).  All compute happens in the CUDA library.
"""
import ctypes as C

import numpy as np

from . import _lib

FIELD_FR = 0
FIELD_FQ = 1
NL = 24
FLAG_INVERSE_ROOT, FLAG_NO_SCALE, FLAG_SCALE_TABLE = 1, 2, 4

ERRORS = {1: "INVALID_ARG", 2: "NOT_POW2", 3: "TOO_LARGE", 4: "BAD_OMEGA", 5: "CUDA", 6: "NO_DEVICE", 7: "BAD_MODULUS"}


class GsnError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"GSN_ERR_{ERRORS.get(code, code)}: {msg}")
        self.code = code


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _limbs(a, shape_tail=(NL,)):
    a = np.ascontiguousarray(a, dtype=np.uint32)
    return a


class Context:
    """gsn_ctx: one stream, one workspace, cached twiddle plans."""

    def __init__(self, device=0):
        self.L = _lib.load()
        h = C.c_void_p()
        self._h = None
        self._check(self.L.gsn_ctx_create(C.byref(h), int(device)))
        self._h = h
        self.device = device

    def _check(self, rc):
        if rc != 0:
            raise GsnError(rc, self.L.gsn_last_error().decode())

    def close(self):
        if self._h is not None:
            self.L.gsn_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- configuration
    def set_field768(self, field):
        self._check(self.L.gsn_set_field768(self._h, int(field)))

    def trim(self):
        self._check(self.L.gsn_ctx_trim(self._h))

    def launch_count(self):
        c = C.c_uint64()
        self._check(self.L.gsn_launch_count(self._h, C.byref(c)))
        return c.value

    def synchronize(self):
        self._check(self.L.gsn_ctx_synchronize(self._h))

    # ---- memory helpers
    def device_alloc(self, nbytes):
        p = C.c_void_p()
        self._check(self.L.gsn_device_alloc(self._h, C.byref(p), int(nbytes)))
        return p.value

    def device_free(self, dptr):
        self._check(self.L.gsn_device_free(self._h, C.c_void_p(dptr)))

    def h2d(self, dptr, arr):
        arr = np.ascontiguousarray(arr)
        self._check(self.L.gsn_memcpy_h2d(self._h, C.c_void_p(dptr), _ptr(arr), arr.nbytes))

    def d2h(self, arr, dptr):
        assert arr.flags["C_CONTIGUOUS"]
        self._check(self.L.gsn_memcpy_d2h(self._h, _ptr(arr), C.c_void_p(dptr), arr.nbytes))

    # ---- 768-bit field
    def best_fft768(self, a, omega, inverse=False):
        """in-place transform of a host (n, 24) uint32 array -- the reference's best_fft<Scalar>"""
        assert a.dtype == np.uint32 and a.flags["C_CONTIGUOUS"] and a.ndim == 2 and a.shape[1] == NL
        omega = _limbs(omega)
        self._check(self.L.gsn_ntt768_host(self._h, _ptr(a), a.shape[0], _ptr(omega), int(bool(inverse))))
        return a

    def best_fft768_batch(self, arrays, omega, inverse=False):
        """in-place transforms of several host (n, 24) uint32 arrays with overlapped copies (gsn_ntt768_host_batch)"""
        n = arrays[0].shape[0]
        for a in arrays:
            assert a.dtype == np.uint32 and a.flags["C_CONTIGUOUS"] and a.shape == (n, NL)
        omega = _limbs(omega)
        ptrs = (C.c_void_p * len(arrays))(*[a.ctypes.data for a in arrays])
        self._check(self.L.gsn_ntt768_host_batch(self._h, ptrs, len(arrays), n, _ptr(omega), int(bool(inverse))))
        return arrays

    def ntt768(self, a, omega, inverse=False):
        out = np.ascontiguousarray(a, dtype=np.uint32).reshape(-1, NL).copy()
        return self.best_fft768(out, omega, inverse)

    def ntt768_device(self, dptr, n, omega, inverse=False, batch=1, log_r=0, stream=None):
        omega = _limbs(omega)
        self._check(self.L.gsn_ntt768_strided_device(self._h, C.c_void_p(dptr), int(n), int(batch), int(log_r), _ptr(omega),
                                                     int(bool(inverse)), C.c_void_p(stream or 0)))

    def ntt768_device_ex(self, dptr, n, omega, batch=1, log_r=0, inverse_root=False, no_scale=False, pre_table=None, stream=None):
        omega = _limbs(omega)
        flags = (FLAG_INVERSE_ROOT if inverse_root else 0) | (FLAG_NO_SCALE if no_scale else 0)
        self._check(self.L.gsn_ntt768_device_ex(self._h, C.c_void_p(dptr), int(n), int(batch), int(log_r), _ptr(omega), flags,
                                                C.c_void_p(pre_table or 0), C.c_void_p(stream or 0)))

    def ntt768_device_scatter(self, dptr, n, omega, peers, my_rank, rank_shift, ins_shift, batch=1, log_r=0, inverse_root=False, no_scale=False,
                              pre_table=None, stream=None):
        omega = _limbs(omega)
        flags = (FLAG_INVERSE_ROOT if inverse_root else 0) | (FLAG_NO_SCALE if no_scale else 0)
        arr = (C.c_void_p * len(peers))(*[C.c_void_p(p) for p in peers])
        self._check(self.L.gsn_ntt768_device_scatter(self._h, C.c_void_p(dptr), int(n), int(batch), int(log_r), _ptr(omega), flags,
                                                     C.c_void_p(pre_table or 0), arr, len(peers), int(my_rank), int(rank_shift), int(ins_shift),
                                                     C.c_void_p(stream or 0)))

    def peer_barrier(self, peer_flags, my_rank, epoch, stream=None):
        arr = (C.c_void_p * len(peer_flags))(*[C.c_void_p(p) for p in peer_flags])
        self._check(self.L.gsn_peer_barrier(self._h, arr, len(peer_flags), int(my_rank), int(epoch) & 0xFFFFFFFF, C.c_void_p(stream or 0)))

    def ipc_export(self, dptr):
        buf = C.create_string_buffer(64)
        self._check(self.L.gsn_ipc_export(self._h, C.c_void_p(dptr), buf))
        return buf.raw

    def ipc_import(self, handle):
        p = C.c_void_p()
        self._check(self.L.gsn_ipc_import(self._h, C.c_char_p(handle), C.byref(p)))
        return p.value

    def ipc_close(self, dptr):
        self._check(self.L.gsn_ipc_close(self._h, C.c_void_p(dptr)))

    def fourstep_table768(self, dptr, rows, cols, row0, col0, n_total, omega, inverse_root=False, scale=False, stream=None):
        omega = _limbs(omega)
        flags = (FLAG_INVERSE_ROOT if inverse_root else 0) | (FLAG_SCALE_TABLE if scale else 0)
        self._check(self.L.gsn_fourstep_table768(self._h, C.c_void_p(dptr), int(rows), int(cols), int(row0), int(col0), int(n_total),
                                                 _ptr(omega), flags, C.c_void_p(stream or 0)))

    def prepare768(self, n, omega, inverse=False, batch=1):
        omega = _limbs(omega)
        self._check(self.L.gsn_ntt768_prepare(self._h, int(n), int(batch), _ptr(omega), int(bool(inverse))))

    def time_ntt768(self, dptr, n, omega, inverse=False, batch=1, reps=10):
        omega = _limbs(omega)
        ms = (C.c_float * reps)()
        self._check(self.L.gsn_ntt768_time_device(self._h, C.c_void_p(dptr), int(n), int(batch), _ptr(omega), int(bool(inverse)), reps, ms))
        return list(ms)

    def fp768_binop(self, op, a, b):
        a = np.ascontiguousarray(a, dtype=np.uint32).reshape(-1, NL)
        b = np.ascontiguousarray(b, dtype=np.uint32).reshape(-1, NL)
        out = np.empty_like(a)
        self._check(self.L.gsn_fp768_binop_host(self._h, {"mul": 0, "add": 1, "sub": 2}[op], _ptr(out), _ptr(a), _ptr(b), a.shape[0]))
        return out

    def fp768_binop_device(self, op, d_out, d_a, d_b, count, stream=None):
        self._check(self.L.gsn_fp768_binop_device(self._h, {"mul": 0, "add": 1, "sub": 2}[op], C.c_void_p(d_out), C.c_void_p(d_a), C.c_void_p(d_b),
                                                  int(count), C.c_void_p(stream or 0)))

    def fp768_powers_device(self, d_table, count, base, scale=None, stream=None):
        base = _limbs(base)
        sc = _limbs(scale) if scale is not None else None
        self._check(self.L.gsn_fp768_powers_device(self._h, C.c_void_p(d_table), int(count), _ptr(base), _ptr(sc) if sc is not None else None,
                                                   C.c_void_p(stream or 0)))

    def fp768_twiddle_table_device(self, d_table, d_elems, count, stream=None):
        """count Montgomery-form elements (96 B each) -> pre-twiddle table of the transform kernels (192 B per entry)"""
        self._check(self.L.gsn_fp768_twiddle_table_device(self._h, C.c_void_p(d_table), C.c_void_p(d_elems), int(count), C.c_void_p(stream or 0)))

    def multiexp768(self, a, b):
        """sum_i a[i] * b[i] -- the reference's multiexp<Scalar, Scalar> (cuda/multi_exp.h:24-25) on host arrays"""
        a = np.ascontiguousarray(a, dtype=np.uint32).reshape(-1, NL)
        b = np.ascontiguousarray(b, dtype=np.uint32).reshape(-1, NL)
        assert a.shape == b.shape
        out = np.empty(NL, dtype=np.uint32)
        self._check(self.L.gsn_fp768_inner_product_host(self._h, _ptr(out), _ptr(a), _ptr(b), a.shape[0]))
        return out

    def fp2_binop(self, op, a, b):
        """element-wise Fq2 = Fq[u]/(u^2 - 13) arithmetic on (count, 2, 24) arrays (the reference's fp2)"""
        a = np.ascontiguousarray(a, dtype=np.uint32).reshape(-1, 2, NL)
        b = np.ascontiguousarray(b, dtype=np.uint32).reshape(-1, 2, NL)
        out = np.empty_like(a)
        self._check(self.L.gsn_fp2_binop_host(self._h, {"mul": 0, "add": 1, "sub": 2}[op], _ptr(out), _ptr(a), _ptr(b), a.shape[0]))
        return out

    def g1_multiexp(self, points, scalars, method="auto", window_bits=0):
        """sum_i scalars[i] * points[i] on MNT4-753 G1 -- the reference's multiexp<mnt4753_G1, Scalar>.  points: (n, 3, 24)
        projective Montgomery limbs over Fq; scalars: (n, 24) raw integers.  (Fq is built in: any context field works.)
        method: "auto", "naive" (the reference's algorithm) or "bucket" (Pippenger); window_bits 0 = automatic."""
        points = np.ascontiguousarray(points, dtype=np.uint32).reshape(-1, 3, NL)
        scalars = np.ascontiguousarray(scalars, dtype=np.uint32).reshape(-1, NL)
        assert points.shape[0] == scalars.shape[0]
        n = points.shape[0]
        out = np.empty((3, NL), dtype=np.uint32)
        if method == "auto" and window_bits == 0:
            self._check(self.L.gsn_g1_multiexp_host(self._h, _ptr(out), _ptr(points), _ptr(scalars), n))
            return out
        dp, ds, do = self.device_alloc(max(n, 1) * 288), self.device_alloc(max(n, 1) * 96), self.device_alloc(288)
        try:
            if n:
                self.h2d(dp, points)
                self.h2d(ds, scalars)
            self._check(self.L.gsn_g1_multiexp_device_ex(self._h, C.c_void_p(do), C.c_void_p(dp), C.c_void_p(ds), n,
                                                         {"auto": 0, "naive": 1, "bucket": 2}[method], int(window_bits), None))
            self.synchronize()
            self.d2h(out, do)
        finally:
            for p in (dp, ds, do):
                self.device_free(p)
        return out

    def g1_multiexp_device(self, d_out, d_points, d_scalars, n, method="auto", window_bits=0, stream=None):
        """device-resident form (gsn_g1_multiexp_device_ex): 288-byte result at d_out"""
        self._check(self.L.gsn_g1_multiexp_device_ex(self._h, C.c_void_p(d_out), C.c_void_p(d_points), C.c_void_p(d_scalars), int(n),
                                                     {"auto": 0, "naive": 1, "bucket": 2}[method], int(window_bits), C.c_void_p(stream or 0)))

    def coset_ntt768(self, a, omega, shift, inverse=False):
        """forward: evaluations of the polynomial with coefficients a on the coset shift * <omega>;
        inverse: coefficients from such evaluations.  Host arrays in/out (gsn_coset_ntt768_host)."""
        out = np.ascontiguousarray(a, dtype=np.uint32).reshape(-1, NL).copy()
        omega, shift = _limbs(omega), _limbs(shift)
        self._check(self.L.gsn_coset_ntt768_host(self._h, _ptr(out), out.shape[0], _ptr(omega), _ptr(shift), int(bool(inverse))))
        return out

    def coset_ntt768_device(self, dptr, n, omega, shift, inverse=False, batch=1, stream=None):
        """device-resident coset transform: shift^i fused into the first pass (forward) / n^-1 shift^-i into the last (inverse)"""
        omega, shift = _limbs(omega), _limbs(shift)
        self._check(self.L.gsn_coset_ntt768_device(self._h, C.c_void_p(dptr), int(n), int(batch), _ptr(omega), _ptr(shift), int(bool(inverse)),
                                                   C.c_void_p(stream or 0)))

    # ---- tuning / introspection
    def set_option(self, option, value):
        opt = {"flat_table_limit": 1, "plan_cache_bytes": 2, "kernel_variant": 3}[option]
        self._check(self.L.gsn_ctx_set_option(self._h, opt, C.c_uint64(int(value) & 0xFFFFFFFFFFFFFFFF)))

    def plan_info768(self, n, omega, inverse=False):
        omega = _limbs(omega)
        tb, cp, cb = C.c_uint64(), C.c_uint64(), C.c_uint64()
        passes, two = C.c_uint(), C.c_uint()
        self._check(self.L.gsn_ntt768_plan_info(self._h, int(n), _ptr(omega), int(bool(inverse)), C.byref(tb), C.byref(passes), C.byref(two),
                                                C.byref(cp), C.byref(cb)))
        return {"table_bytes": tb.value, "passes": passes.value, "two_level_boundaries": two.value, "cached_plans": cp.value,
                "cached_bytes": cb.value}

    # ---- 32-bit field
    def best_fft32(self, a, omega, mod, inverse=False):
        assert a.dtype == np.uint32 and a.flags["C_CONTIGUOUS"] and a.ndim == 1
        self._check(self.L.gsn_ntt32_host(self._h, _ptr(a), a.shape[0], int(omega), int(mod), int(bool(inverse))))
        return a

    def ntt32(self, a, omega, mod, inverse=False):
        out = np.ascontiguousarray(a, dtype=np.uint32).reshape(-1).copy()
        return self.best_fft32(out, omega, mod, inverse)

    def ntt32_device(self, dptr, n, omega, mod, inverse=False, batch=1, stream=None):
        self._check(self.L.gsn_ntt32_device(self._h, C.c_void_p(dptr), int(n), int(batch), int(omega), int(mod), int(bool(inverse)),
                                            C.c_void_p(stream or 0)))

    def time_ntt32(self, dptr, n, omega, mod, inverse=False, batch=1, reps=10):
        ms = (C.c_float * reps)()
        self._check(self.L.gsn_ntt32_time_device(self._h, C.c_void_p(dptr), int(n), int(batch), int(omega), int(mod), int(bool(inverse)),
                                                 reps, ms))
        return list(ms)

    # ---- measurement
    def int32_issue_rates(self):
        rates = (C.c_double * 16)()
        nm, sm, khz = C.c_int(), C.c_int(), C.c_int()
        self._check(self.L.gsn_int32_issue_rates(self._h, rates, 16, C.byref(nm), C.byref(sm), C.byref(khz)))
        names = ["imad_lo", "imad_hi", "imad_wide", "imad_wide_shared_operands", "imad_wide_x_chain", "iadd3_x_chain", "dfma_f64",
                 "imad_wide_plus_dfma_interleaved"]
        return {"rates": {names[k]: rates[k] for k in range(nm.value)}, "sm_count": sm.value, "sm_clock_khz": khz.value}


class FourStepPlan:
    """gsn_fourstep: one rank's share of a sharded transform (see include/gpusnarks_b200.h)."""

    def __init__(self, ctx, logn, omega, n_ranks, my_rank, directions=("forward", "inverse")):
        self.ctx = ctx
        self.L = ctx.L
        omega = _limbs(omega)
        d = (1 if "forward" in directions else 0) | (2 if "inverse" in directions else 0)
        h = C.c_void_p()
        self._h = None
        ctx._check(self.L.gsn_fourstep_create(ctx._h, C.byref(h), int(logn), _ptr(omega), int(n_ranks), int(my_rank), d))
        self._h = h
        a, b, c, e = C.c_uint(), C.c_uint(), C.c_uint(), C.c_uint()
        tb = C.c_uint64()
        ctx._check(self.L.gsn_fourstep_info(h, C.byref(a), C.byref(b), C.byref(c), C.byref(tb), C.byref(e)))
        self.log_n1, self.log_n2, self.rank_bit, self.table_bytes, self.per_source = a.value, b.value, c.value, tb.value, bool(e.value)
        px, py0, py1, pf = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        ctx._check(self.L.gsn_fourstep_buffers(h, C.byref(px), C.byref(py0), C.byref(py1), C.byref(pf)))
        self.x, self.y0, self.y1, self.flags = px.value, py0.value, py1.value, pf.value

    def connect(self, peer_x, peer_y0, peer_y1, peer_flags):
        def arr(ps):
            return (C.c_void_p * len(ps))(*[C.c_void_p(p) for p in ps])
        self.ctx._check(self.L.gsn_fourstep_connect(self._h, arr(peer_x), arr(peer_y0), arr(peer_y1), arr(peer_flags)))

    def forward(self, stream=None):
        y = C.c_void_p()
        self.ctx._check(self.L.gsn_fourstep_forward(self._h, C.c_void_p(stream or 0), C.byref(y)))
        return y.value

    def inverse(self, stream=None):
        x = C.c_void_p()
        self.ctx._check(self.L.gsn_fourstep_inverse(self._h, C.c_void_p(stream or 0), C.byref(x)))
        return x.value

    def set_timing(self, on):
        self.ctx._check(self.L.gsn_fourstep_set_timing(self._h, int(bool(on))))

    def phase_ms(self):
        ms = (C.c_float * 3)()
        calls = C.c_uint64()
        self.ctx._check(self.L.gsn_fourstep_phase_ms(self._h, ms, C.byref(calls)))
        return {"column+scatter": ms[0], "signal": ms[1], "row": ms[2], "calls": calls.value}

    def close(self):
        if self._h is not None:
            self.L.gsn_fourstep_destroy(self._h)
            self._h = None


class MultiGpu:
    """gsn_multi: a sharded transform driven from ONE process over several devices (peer access, no NCCL)."""

    def __init__(self, devices, n, omega, directions=("forward", "inverse")):
        self.L = _lib.load()
        omega = _limbs(omega)
        d = (1 if "forward" in directions else 0) | (2 if "inverse" in directions else 0)
        h = C.c_void_p()
        self._h = None
        devs = (C.c_int * len(devices))(*devices)
        rc = self.L.gsn_multi_create(C.byref(h), devs, len(devices), int(n), _ptr(omega), d)
        if rc:
            raise GsnError(rc, self.L.gsn_last_error().decode())
        self._h = h
        self.n = n

    def _check(self, rc):
        if rc:
            raise GsnError(rc, self.L.gsn_last_error().decode())

    def ntt_host(self, a, inverse=False):
        assert a.dtype == np.uint32 and a.flags["C_CONTIGUOUS"] and a.shape == (self.n, NL)
        self._check(self.L.gsn_multi_ntt768_host(self._h, _ptr(a), int(bool(inverse))))
        return a

    def ntt_device(self, inverse=False):
        self._check(self.L.gsn_multi_ntt768_device(self._h, int(bool(inverse))))

    def synchronize(self):
        self._check(self.L.gsn_multi_synchronize(self._h))

    def close(self):
        if self._h is not None:
            self.L.gsn_multi_destroy(self._h)
            self._h = None


def g1_multiexp_multi(points, scalars, devices=None):
    """gsn_g1_multiexp_multi_host: sum_i scalars[i] * points[i] with the points cut into one slice per device
    (devices=None: every visible device); returns the (3, 24) projective result"""
    L = _lib.load()
    points = np.ascontiguousarray(points, dtype=np.uint32).reshape(-1, 3, NL)
    scalars = np.ascontiguousarray(scalars, dtype=np.uint32).reshape(-1, NL)
    assert points.shape[0] == scalars.shape[0]
    out = np.empty((3, NL), dtype=np.uint32)
    devs = list(devices) if devices is not None else []
    arr = (C.c_int * max(len(devs), 1))(*devs)
    rc = L.gsn_g1_multiexp_multi_host(arr, len(devs), _ptr(out), _ptr(points), _ptr(scalars), points.shape[0])
    if rc:
        raise GsnError(rc, L.gsn_last_error().decode())
    return out


def device_count():
    L = _lib.load()
    c = C.c_int()
    rc = L.gsn_device_count(C.byref(c))
    return c.value if rc == 0 else 0
