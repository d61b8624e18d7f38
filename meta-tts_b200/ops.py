"""Thin tensor-level wrappers over the libmtts C ABI (device pointers + current stream).

PyTorch is plumbing here: it owns device memory and streams.  Every function below launches
hand-written sm_100a kernels through ctypes; nothing falls back to torch ops.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional, Sequence

import torch

from . import lib as L

launch_count = 0          # number of libmtts kernel launches issued from this process (bench: gpu_launches)


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise L.MttsError("libmtts ops need CUDA tensors: there is no CPU path")


# ---------------------------------------------------------------------------------------------
# bf16 split helpers (host-side utilities used by tests; kernels produce hi/lo in their epilogues)
# ---------------------------------------------------------------------------------------------
def split_bf16(x: torch.Tensor):
    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16)
    return hi, lo


# ---------------------------------------------------------------------------------------------
# generic tcgen05 GEMM
# ---------------------------------------------------------------------------------------------
@dataclass
class Opnd:
    """One GEMM operand: a 4-D strided view of a bf16 buffer (dims[0] contiguous)."""
    hi: torch.Tensor
    lo: Optional[torch.Tensor]
    major: int                      # L.MAJOR_K / L.MAJOR_MN
    dims: Sequence[int]             # up to 4 extents, dims[0] contiguous
    strides: Sequence[int]          # element strides, strides[0] == 1
    src2: int = L.SRC_ZERO
    src3: int = L.SRC_ZERO
    shift_src: int = L.SRC_ZERO
    shift_base: int = 0
    shift_step: int = 0
    offset: int = 0                 # element offset into hi/lo

    def fill(self, o: L.Operand) -> None:
        esz = 2
        o.hi = self.hi.data_ptr() + self.offset * esz
        o.lo = (self.lo.data_ptr() + self.offset * esz) if self.lo is not None else None
        o.major = self.major
        o.src2, o.src3 = self.src2, self.src3
        o.shift_src, o.shift_base, o.shift_step = self.shift_src, self.shift_base, self.shift_step
        d = list(self.dims) + [1] * (4 - len(self.dims))
        s = list(self.strides) + [0] * (4 - len(self.strides))
        for i in range(4):
            o.dims[i] = int(d[i])
            o.strides[i] = int(s[i])


def gemm(a: Opnd, b: Opnd, M: int, N: int, K: int, *,
         c_f32: Optional[torch.Tensor] = None, c_hi: Optional[torch.Tensor] = None,
         c_lo: Optional[torch.Tensor] = None, ldc: int, c_off: int = 0, c_sz0: int = 0, c_sz1: int = 0,
         alpha: float = 1.0, bias: Optional[torch.Tensor] = None, bias_sz0: int = 0,
         gate: Optional[torch.Tensor] = None, flags: int = 0,
         ntaps: int = 1, nkb: int = 1, nz0: int = 1, nz1: int = 1,
         split: int = 1, block_n: int = 0, ksplit: int = 1) -> None:
    global launch_count
    _need_cuda(a.hi, b.hi, c_f32, c_hi, c_lo, bias, gate)
    d = L.GemmDesc()
    d.M, d.N, d.K = M, N, K
    d.ntaps, d.nkb, d.nz0, d.nz1 = ntaps, nkb, nz0, nz1
    d.split, d.block_n, d.ksplit, d.flags = split, block_n, ksplit, flags
    d.alpha = alpha
    a.fill(d.a)
    b.fill(d.b)
    d.c_f32 = (c_f32.data_ptr() + 4 * c_off) if c_f32 is not None else None
    d.c_hi = (c_hi.data_ptr() + 2 * c_off) if c_hi is not None else None
    d.c_lo = (c_lo.data_ptr() + 2 * c_off) if c_lo is not None else None
    d.ldc, d.c_sz0, d.c_sz1 = ldc, c_sz0, c_sz1
    d.bias = _ptr(bias)
    d.bias_sz0 = bias_sz0
    d.gate = (gate.data_ptr() + 2 * c_off) if gate is not None else None
    L.call("mtts_gemm", C.byref(d), _stream())
    launch_count += 1


# ---------------------------------------------------------------------------------------------
# LengthRegulator
# ---------------------------------------------------------------------------------------------
def length_regulate_index(dur: torch.Tensor, T: int):
    """dur [B,L] int64 (targets) or float32 (predicted) -> (idx [B,T] int32, mel_len [B] int64)."""
    global launch_count
    _need_cuda(dur)
    B, Lp = dur.shape
    dur = dur.contiguous()
    idx = torch.empty((B, T), dtype=torch.int32, device=dur.device)
    mel_len = torch.empty((B,), dtype=torch.int64, device=dur.device)
    if dur.dtype == torch.int64:
        di, df = dur.data_ptr(), None
    elif dur.dtype == torch.float32:
        di, df = None, dur.data_ptr()
    else:
        raise L.MttsError(f"durations must be int64 or float32, got {dur.dtype}")
    L.call("mtts_length_regulate_index", di, df, B, Lp, T, idx.data_ptr(), mel_len.data_ptr(), _stream())
    launch_count += 1
    return idx, mel_len


def length_regulate_fwd(x: torch.Tensor, idx: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    global launch_count
    _need_cuda(x, idx)
    B, Lp, Cc = x.shape
    T = idx.shape[1]
    if out is None:
        out = torch.empty((B, T, Cc), dtype=torch.float32, device=x.device)
    L.call("mtts_length_regulate_fwd", x.data_ptr(), idx.data_ptr(), B, Lp, T, Cc, out.data_ptr(), _stream())
    launch_count += 1
    return out


def length_regulate_bwd(dy: torch.Tensor, dur: torch.Tensor, Lp: int, out: Optional[torch.Tensor] = None):
    global launch_count
    _need_cuda(dy, dur)
    B, T, Cc = dy.shape
    if out is None:
        out = torch.empty((B, Lp, Cc), dtype=torch.float32, device=dy.device)
    di, df = (dur.data_ptr(), None) if dur.dtype == torch.int64 else (None, dur.data_ptr())
    L.call("mtts_length_regulate_bwd", dy.data_ptr(), di, df, B, Lp, T, Cc, out.data_ptr(), _stream())
    launch_count += 1
    return out
