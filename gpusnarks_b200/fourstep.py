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
    """Local transforms through libgpusnarks_b200.so on torch-owned device memory."""

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


class FusedFourStepNTT768(FourStepNTT768):
    """Four-step NTT whose exchange is fused into the transform kernels: the last pass of the
    column (forward) or row (inverse) transforms stores each output element straight into the
    destination rank's buffer over NVLink (gsn_ntt768_device_scatter, CUDA-IPC mapped peer
    memory), so there is no all-to-all and no repacking copy.  NCCL is used only for two
    stream-ordered barriers per transform (a one-element all-reduce) and for the one-time
    exchange of the IPC handles.

    The plan owns the two buffers: `self.x` (column-block, (n1, C, 24) int32) and `self.y`
    (row-block, (R, n2, 24)).  forward(): fill self.x, call, read self.y.  inverse(): the reverse.
    """

    def __init__(self, ctx, device, logn, omega, group=None, modulus=F.FR, directions=("forward", "inverse"), log_n1=None):
        super().__init__(CudaBackend(ctx, device), logn, omega, group=group, modulus=modulus, directions=directions, log_n1=log_n1)
        self.ctx = ctx
        self.device = device
        nbytes = self.n1 * self.C * F.NL * 4
        self._xp = ctx.device_alloc(nbytes)
        self._yp = ctx.device_alloc(nbytes)
        self.x = torch.as_tensor(_DeviceArray(self._xp, self.column_block_shape()), device=device)
        self.y = torch.as_tensor(_DeviceArray(self._yp, self.row_block_shape()), device=device)
        self._yp2 = ctx.device_alloc(nbytes)  # second receive buffer (forward calls alternate)
        self.y2 = torch.as_tensor(_DeviceArray(self._yp2, self.row_block_shape()), device=device)
        self._flip = 1
        self._timing = {} if os.environ.get("GSN_FOURSTEP_TIMING") else None
        self._flag = torch.zeros(1, dtype=torch.int32, device=device)
        # flag array for the peer-memory barrier (8 slots), zeroed before anyone can signal it
        self._fp = ctx.device_alloc(256)
        ctx.h2d(self._fp, np.zeros(64, dtype=np.uint32))
        self._epoch = 0
        self._nccl_barrier = bool(os.environ.get("GSN_FOURSTEP_NCCL_BARRIER"))
        self.peers_x, self.peers_y = self._exchange_handles(self._xp), self._exchange_handles(self._yp)
        self.peers_y2 = self._exchange_handles(self._yp2)
        self.peer_flags = self._exchange_handles(self._fp)
        if self.G > 1:
            dist.barrier(group=self.group)  # every rank has zeroed and published its flags
        self.logG, self.logC, self.logR = _ilog2(self.G), _ilog2(self.C), _ilog2(self.R)

    def _exchange_handles(self, ptr):
        if self.G == 1:
            return [ptr]
        mine = torch.frombuffer(bytearray(self.ctx.ipc_export(ptr)), dtype=torch.uint8).to(self.device)
        allh = [torch.empty_like(mine) for _ in range(self.G)]
        dist.all_gather(allh, mine, group=self.group)
        out = []
        for r, h in enumerate(allh):
            out.append(ptr if r == self.rank else self.ctx.ipc_import(bytes(h.cpu().numpy().tobytes())))
        return out

    def _barrier(self):
        """stream-ordered barrier across the ranks: later kernels on this stream start after every rank got here"""
        if self.G == 1:
            return
        if self._nccl_barrier:
            dist.all_reduce(self._flag, group=self.group)
        else:
            self._epoch += 1
            self.ctx.peer_barrier(self.peer_flags, self.rank, self._epoch, stream=self.be._stream())

    def forward(self, x=None):
        """column-block self.x -> row-block self.y (returns self.y)"""
        if x is not None and x.data_ptr() != self.x.data_ptr():
            self.x.copy_(x)
        be = self.be
        t = self._timing
        if t is not None:
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            ev[0].record()
        # Receive buffers alternate between calls, so the only hazard left is "stores have landed":
        # a rank that has passed the barrier of call i+1 knows every rank finished the row pass of
        # call i, hence its buffer of call i may be overwritten by call i+2 without a second barrier.
        self._flip ^= 1
        peers_y, yp, y = (self.peers_y, self._yp, self.y) if self._flip == 0 else (self.peers_y2, self._yp2, self.y2)
        # column NTTs; output element (k1, c) of this rank goes to rank k1 / R, position [k1 % R][rank*C + c]
        self.ctx.ntt768_device_scatter(self._xp, self.n1, self.w_col, peers_y, self.rank, rank_shift=self.logR + self.logC,
                                       ins_shift=self.logC, batch=1, log_r=self.logC, stream=be._stream())
        if t is not None:
            ev[1].record()
        self._barrier()  # all stores into y have landed
        if t is not None:
            ev[2].record()
        be.ntt(y, self.n2, self.R, 0, self.w_row, pre_table=self.tw_fwd)
        if t is not None:
            ev[3].record()
            torch.cuda.synchronize(self.device)
            for i, name in enumerate(("column+scatter", "barrier", "row")):
                t[name] = t.get(name, 0.0) + ev[i].elapsed_time(ev[i + 1])
            t["calls"] = t.get("calls", 0) + 1
        return y

    def inverse(self, y=None):
        """row-block self.y -> column-block self.x (returns self.x)"""
        if y is None:
            y = self.y if self._flip == 0 else self.y2  # the buffer the last forward() filled
        if y.data_ptr() != self.y.data_ptr():
            self.y.copy_(y)
        be = self.be
        self._barrier()
        # inverse row NTTs; output element (r, i2) goes to rank i2 / C, position [rank*R + r][i2 % C]
        self.ctx.ntt768_device_scatter(self._yp, self.n2, self.w_row, self.peers_x, self.rank, rank_shift=self.logC,
                                       ins_shift=self.logR + self.logC, batch=self.R, log_r=0, inverse_root=True, no_scale=True,
                                       stream=be._stream())
        self._barrier()
        be.ntt(self.x, self.n1, 1, self.logC, self.w_col, inverse_root=True, no_scale=True, pre_table=self.tw_inv)
        return self.x

    def close(self):
        torch.cuda.synchronize(self.device)
        if self.G > 1:
            dist.barrier(group=self.group)
            for r in range(self.G):
                if r != self.rank:
                    self.ctx.ipc_close(self.peers_x[r])
                    self.ctx.ipc_close(self.peers_y[r])
                    self.ctx.ipc_close(self.peers_y2[r])
                    self.ctx.ipc_close(self.peer_flags[r])
            dist.barrier(group=self.group)
        self.x = self.y = self.y2 = None
        self.ctx.device_free(self._xp)
        self.ctx.device_free(self._yp)
        self.ctx.device_free(self._yp2)
        self.ctx.device_free(self._fp)
