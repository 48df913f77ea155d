"""Four-step (Bailey) 768-bit NTT sharded across the GPUs of one box, one process per GPU.

Same decomposition as the reference's host "parallel FFT" (`_basic_parallel_radix2_FFT_inner`,
reference test/fft_host.h:56-117: short-side DFTs times omega^(j*i), row FFTs, transposed
write-out), with the short side done as a fast NTT and the shared-memory `tmp` matrix replaced
by ONE all-to-all over NVLink (torch.distributed / NCCL):

    n = n1 * n2,  input index i = i1*n2 + i2,  output index k = k1 + n1*k2
    1. column NTTs   rank g owns columns i2 in [g*C, (g+1)*C), C = n2/G: n1-point transforms
                     along i1 with root omega^n2                      (gsn_ntt768_device_ex, log_r)
    2. all-to-all    rank g sends rows k1 in [h*R, (h+1)*R), R = n1/G, to rank h
    3. row NTTs      rank h owns rows k1: multiply by omega^(k1*i2) (table fused into the first
                     pass) and transform along i2 with root omega^n1

Layouts (both are views of the natural-order vector, no data is ever bit-reversed):
    column-block  x[i1, c]  = a[i1*n2 + g*C + c]            shape (n1, C, 24)   forward input
    row-block     y[r, k2]  = A[(h*R + r) + n1*k2]          shape (R, n2, 24)   forward output
The inverse runs the three steps backwards (inverse row NTTs, all-to-all, conjugate twiddles
* n^-1 fused into the inverse column NTTs) and maps row-block back to column-block.

The numerical work is done by a backend object; the product backend is `CudaBackend` (the C
ABI).  tests/ inject a CPU backend to exercise the exchange logic under gloo.
"""
import os

import numpy as np
import torch
import torch.distributed as dist

from . import field as F


def _ilog2(x):
    assert x > 0 and x & (x - 1) == 0, f"{x} is not a power of two"
    return x.bit_length() - 1


class CudaBackend:
    """Local transforms through libgpusnarks_b200.so on torch-owned device memory (NCCL-exchange comparison path)."""

    def __init__(self, ctx, device):
        self.ctx = ctx
        self.device = device

    def _stream(self):
        # torch's legacy default stream has handle 0, which the C ABI reads as "use the context's
        # own stream"; pass cudaStreamLegacy (0x1) instead so the kernels stay ordered with torch ops
        return torch.cuda.current_stream(self.device).cuda_stream or 1

    def ntt(self, t, n, batch, log_r, omega, inverse_root=False, no_scale=False, pre_table=None):
        assert t.is_cuda and t.is_contiguous() and t.dtype == torch.int32
        self.ctx.ntt768_device_ex(t.data_ptr(), n, omega, batch=batch, log_r=log_r, inverse_root=inverse_root, no_scale=no_scale,
                                  pre_table=pre_table.data_ptr() if pre_table is not None else None, stream=self._stream())

    def table(self, rows, cols, row0, col0, n_total, omega, inverse_root=False, scale=False):
        t = torch.empty((rows, cols, 2 * F.NL), dtype=torch.int32, device=self.device)  # (w, w'') per entry
        self.ctx.fourstep_table768(t.data_ptr(), rows, cols, row0, col0, n_total, omega, inverse_root=inverse_root, scale=scale,
                                   stream=self._stream())
        return t


class FourStepNTT768:
    def __init__(self, backend, logn, omega, group=None, modulus=F.FR, directions=("forward", "inverse"), log_n1=None):
        self.be = backend
        self.group = group
        self.G = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.logn = logn
        # the short side is one shared-memory pass (<= 2^10) when the transform is large enough:
        # 2^24 = 2^10 x 2^14 is three passes in all, 2^12 x 2^12 would be four
        self.log_n1 = split_log_n1(logn) if log_n1 is None else log_n1
        self.log_n2 = logn - self.log_n1
        self.n1, self.n2, self.n = 1 << self.log_n1, 1 << self.log_n2, 1 << logn
        assert self.n1 % self.G == 0 and self.n2 % self.G == 0, "both sides of the matrix must split across the ranks"
        self.C, self.R = self.n2 // self.G, self.n1 // self.G
        self.omega = np.ascontiguousarray(omega, dtype=np.uint32)
        self.w_col = F.mont_pow(self.omega, self.n2, modulus)  # n1-th root
        self.w_row = F.mont_pow(self.omega, self.n1, modulus)  # n2-th root
        self.tw_fwd = self.tw_inv = None
        self._timing_nccl = {} if os.environ.get("GSN_FOURSTEP_TIMING") else None
        if "forward" in directions:   # rows (rank*R + r), all columns i2
            self.tw_fwd = self.be.table(self.R, self.n2, self.rank * self.R, 0, self.n, self.omega)
        if "inverse" in directions:   # all rows k1, columns rank*C + c; carries n^-1
            self.tw_inv = self.be.table(self.n1, self.C, 0, self.rank * self.C, self.n, self.omega, inverse_root=True, scale=True)

    # shapes of the two layouts on this rank
    def column_block_shape(self):
        return (self.n1, self.C, F.NL)

    def row_block_shape(self):
        return (self.R, self.n2, F.NL)

    def _all_to_all(self, send):
        if self.G == 1:
            return send
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv, send, group=self.group)
        return recv

    def forward(self, x):
        """x: column-block (n1, C, 24) int32, overwritten.  Returns row-block (R, n2, 24)."""
        assert tuple(x.shape) == self.column_block_shape()
        t = getattr(self, "_timing_nccl", None)
        if t is not None:
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
            ev[0].record()
        self.be.ntt(x, self.n1, 1, _ilog2(self.C), self.w_col)
        if t is not None:
            ev[1].record()
        recv = self._all_to_all(x.view(self.G, self.R, self.C, F.NL))        # [g, r, c]: rows of this rank from every g
        if t is not None:
            ev[2].record()
        y = recv.permute(1, 0, 2, 3).contiguous().view(self.R, self.n2, F.NL)  # [r, i2 = g*C + c]
        if t is not None:
            ev[3].record()
        self.be.ntt(y, self.n2, self.R, 0, self.w_row, pre_table=self.tw_fwd)
        if t is not None:
            ev[4].record()
            torch.cuda.synchronize()
            for i, name in enumerate(("column", "all_to_all", "repack", "row")):
                t[name] = t.get(name, 0.0) + ev[i].elapsed_time(ev[i + 1])
            t["calls"] = t.get("calls", 0) + 1
        return y

    def inverse(self, y):
        """y: row-block (R, n2, 24) int32, overwritten.  Returns column-block (n1, C, 24)."""
        assert tuple(y.shape) == self.row_block_shape()
        self.be.ntt(y, self.n2, self.R, 0, self.w_row, inverse_root=True, no_scale=True)
        send = y.view(self.R, self.G, self.C, F.NL).permute(1, 0, 2, 3).contiguous()  # [g, r, c]
        x = self._all_to_all(send).view(self.n1, self.C, F.NL)                         # [k1 = h*R + r, c]
        self.be.ntt(x, self.n1, 1, _ilog2(self.C), self.w_col, inverse_root=True, no_scale=True, pre_table=self.tw_inv)
        return x


def split_log_n1(logn):
    return min(10, logn // 2)


# ---- helpers to move between a natural-order vector and the two layouts (tests, examples)
def to_column_block(a, logn, G, rank, log_n1=None):
    """a: (n, 24) natural order -> this rank's (n1, C, 24) column block"""
    log_n1 = split_log_n1(logn) if log_n1 is None else log_n1
    n1, n2 = 1 << log_n1, 1 << (logn - log_n1)
    C = n2 // G
    return a.reshape(n1, n2, F.NL)[:, rank * C:(rank + 1) * C].clone() if torch.is_tensor(a) else \
        np.ascontiguousarray(a.reshape(n1, n2, F.NL)[:, rank * C:(rank + 1) * C])


def from_row_blocks(blocks, logn, log_n1=None):
    """list over ranks of (R, n2, 24) row blocks -> (n, 24) natural order: A[k1 + n1*k2] = y[k1][k2]"""
    log_n1 = split_log_n1(logn) if log_n1 is None else log_n1
    n1, n2 = 1 << log_n1, 1 << (logn - log_n1)
    y = np.concatenate([np.asarray(b) for b in blocks], axis=0)  # (n1, n2, 24) indexed [k1][k2]
    return np.ascontiguousarray(y.transpose(1, 0, 2)).reshape(n1 * n2, F.NL)


class _DeviceArray:
    """wraps library-owned device memory for torch.as_tensor (CUDA array interface)"""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<i4", "data": (int(ptr), False), "version": 3}


# ---- layouts of the fused plan (gsn_fourstep, include/gpusnarks_b200.h)
def fused_rank_bit(logn, G, max_pass_log=10):
    """bit of the column index i2 at which the owning rank sits: the top of the low digit of the row transform's
    first pass (so that every tile of that pass reads one source rank); log2(C) for a one-pass row transform"""
    log_n2 = logn - split_log_n1(logn)
    logG = _ilog2(G)
    npass = -(-log_n2 // max_pass_log) if log_n2 else 1
    base, extra = divmod(log_n2, npass) if log_n2 else (0, 0)
    digits = [base + (1 if i < extra else 0) for i in range(npass)]
    low = sum(digits[1:])
    if npass < 2 or low < logG:
        low = log_n2
    return low - logG


def to_column_layout(a, logn, G, rank, rank_bit=None):
    """a: (n, 24) natural order -> this rank's (n1, C, 24) column layout: x[i1][c] = a[i1*n2 + i2(c)], i2(c) = c with the
    rank inserted at bit rank_bit"""
    log_n1 = split_log_n1(logn)
    n1, n2 = 1 << log_n1, 1 << (logn - log_n1)
    C = n2 // G
    rb = fused_rank_bit(logn, G) if rank_bit is None else rank_bit
    v = a.reshape(n1, C >> rb, G, 1 << rb, F.NL)[:, :, rank]
    return v.reshape(n1, C, F.NL).clone() if torch.is_tensor(a) else np.ascontiguousarray(v.reshape(n1, C, F.NL))


def from_column_layouts(blocks, logn, rank_bit=None):
    """inverse of to_column_layout over all ranks: list of (n1, C, 24) -> (n, 24) natural order"""
    G = len(blocks)
    log_n1 = split_log_n1(logn)
    n1, n2 = 1 << log_n1, 1 << (logn - log_n1)
    C = n2 // G
    rb = fused_rank_bit(logn, G) if rank_bit is None else rank_bit
    out = np.empty((n1, C >> rb, G, 1 << rb, F.NL), dtype=np.uint32)
    for g, b in enumerate(blocks):
        out[:, :, g] = np.asarray(b).reshape(n1, C >> rb, 1 << rb, F.NL)
    return out.reshape(n1 * n2, F.NL)


def from_row_layouts(blocks, logn):
    """list over ranks of (n2, R, 24) row layouts -> (n, 24) natural order: A[(h*R + r) + n1*k2] = y_h[k2][r]"""
    y = np.stack([np.asarray(b) for b in blocks], axis=1)  # (n2, G, R, 24)
    return np.ascontiguousarray(y).reshape(-1, F.NL)


def to_row_layout(A, logn, G, rank):
    """A: (n, 24) natural order -> this rank's (n2, R, 24) row layout"""
    log_n1 = split_log_n1(logn)
    n1, n2 = 1 << log_n1, 1 << (logn - log_n1)
    R = n1 // G
    v = A.reshape(n2, G, R, F.NL)[:, rank]
    return v.clone() if torch.is_tensor(A) else np.ascontiguousarray(v)


class FusedFourStepNTT768:
    """Four-step NTT whose exchange is fused into the transform kernels (the product path): a thin
    torch.distributed front end of the C plan object gsn_fourstep.  The last pass of the column
    transforms stores each output element straight into the destination rank's row buffer over
    NVLink (CUDA-IPC mapped peer memory), and the first row pass starts, source rank by source rank,
    as soon as that rank's columns have arrived (arrival flags in peer memory).  torch.distributed
    is used once, to exchange the IPC handles.

    The plan owns the buffers: `self.x` (column layout, (n1, C, 24) int32) and two row buffers
    (row layout, (n2, R, 24)); forward(): fill self.x, call, read the returned tensor.
    """

    def __init__(self, ctx, device, logn, omega, group=None, modulus=F.FR, directions=("forward", "inverse")):
        from .ntt import FourStepPlan
        self.ctx, self.device, self.group = ctx, device, group
        self.G = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.logn = logn
        self.plan = FourStepPlan(ctx, logn, omega, self.G, self.rank, directions)
        self.log_n1, self.log_n2 = self.plan.log_n1, self.plan.log_n2
        self.n1, self.n2, self.n = 1 << self.log_n1, 1 << self.log_n2, 1 << logn
        self.C, self.R = self.n2 // self.G, self.n1 // self.G
        self.rank_bit = self.plan.rank_bit
        self.x = torch.as_tensor(_DeviceArray(self.plan.x, self.column_block_shape()), device=device)
        self.y0 = torch.as_tensor(_DeviceArray(self.plan.y0, self.row_block_shape()), device=device)
        self.y1 = torch.as_tensor(_DeviceArray(self.plan.y1, self.row_block_shape()), device=device)
        self._cur = self.y0
        self._imported = []
        px, py0, py1, pf = (self._exchange_handles(p) for p in (self.plan.x, self.plan.y0, self.plan.y1, self.plan.flags))
        if self.G > 1:
            self.plan.connect(px, py0, py1, pf)
            dist.barrier(group=self.group)  # every rank has zeroed and published its flags

    def column_block_shape(self):
        return (self.n1, self.C, F.NL)

    def row_block_shape(self):
        return (self.n2, self.R, F.NL)

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream or 1  # 0 would mean "the context's own stream"

    def _exchange_handles(self, ptr):
        if self.G == 1:
            return [ptr]
        mine = torch.frombuffer(bytearray(self.ctx.ipc_export(ptr)), dtype=torch.uint8).to(self.device)
        allh = [torch.empty_like(mine) for _ in range(self.G)]
        dist.all_gather(allh, mine, group=self.group)
        out = []
        for r, h in enumerate(allh):
            if r == self.rank:
                out.append(ptr)
            else:
                p = self.ctx.ipc_import(bytes(h.cpu().numpy().tobytes()))
                self._imported.append(p)
                out.append(p)
        return out

    def forward(self, x=None):
        """column layout self.x -> row layout (returns the row buffer just written)"""
        if x is not None and x.data_ptr() != self.x.data_ptr():
            self.x.copy_(x)
        yp = self.plan.forward(self._stream())
        self._cur = self.y0 if yp == self.plan.y0 else self.y1
        return self._cur

    def inverse(self, y=None):
        """row layout (the buffer the last forward() filled) -> column layout self.x (returns self.x)"""
        if y is not None and y.data_ptr() != self._cur.data_ptr():
            self._cur.copy_(y)
        self.plan.inverse(self._stream())
        return self.x

    def phase_ms(self):
        return self.plan.phase_ms()

    def close(self):
        torch.cuda.synchronize(self.device)
        if self.G > 1:
            dist.barrier(group=self.group)
            for p in self._imported:
                self.ctx.ipc_close(p)
            dist.barrier(group=self.group)
        self.x = self.y0 = self.y1 = self._cur = None
        self.plan.close()
