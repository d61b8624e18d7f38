"""Thin tensor-level wrappers over the libmtts C ABI (device pointers + current stream).

PyTorch is plumbing here: it owns device memory and streams.  Every function below launches
hand-written sm_100a kernels through ctypes; nothing falls back to torch ops.
"""
from __future__ import annotations

import ctypes as C
import os
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
    hi2: Optional[torch.Tensor] = None   # second product term (same geometry / offset), see mtts_gemm_desc.a2_hi
    lo2: Optional[torch.Tensor] = None

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
         split: int = 1, block_n: int = 0, ksplit: int = 1, pair: bool = False, ln: Optional[dict] = None) -> None:
    """ln = dict(res, gamma, beta, lens, T, z, stats, pre, salt[, eps]): the GEMM's epilogue is dropout -> + residual -> LayerNorm ->
    pad-row zeroing (mtts_gemm_ln, include/mtts.h; N == 256): c_f32 / c_hi / c_lo receive the LayerNorm output."""
    global launch_count
    _need_cuda(a.hi, b.hi, c_f32, c_hi, c_lo, bias, gate)
    d = L.GemmDesc()
    d.M, d.N, d.K = M, N, K
    d.ntaps, d.nkb, d.nz0, d.nz1 = ntaps, nkb, nz0, nz1
    d.split, d.block_n, d.ksplit, d.flags = split, block_n, ksplit, flags
    d.pair = 1 if pair else 0
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
    if a.hi2 is not None or b.hi2 is not None:
        assert a.hi2 is not None and b.hi2 is not None, "second product term needs both operands"
        d.a2_hi = a.hi2.data_ptr() + 2 * a.offset
        d.a2_lo = (a.lo2.data_ptr() + 2 * a.offset) if a.lo2 is not None else None
        d.b2_hi = b.hi2.data_ptr() + 2 * b.offset
        d.b2_lo = (b.lo2.data_ptr() + 2 * b.offset) if b.lo2 is not None else None
    if ln is not None:
        _need_cuda(ln.get("res"), ln["gamma"], ln["beta"], ln.get("lens"), ln.get("z"), ln.get("stats"))
        e = L.LnEpilogue()
        e.res, e.gamma, e.beta, e.lens = _ptr(ln.get("res")), _ptr(ln["gamma"]), _ptr(ln["beta"]), _ptr(ln.get("lens"))
        e.T, e.eps = int(ln["T"]), float(ln.get("eps", 1e-5))
        e.z_out, e.stats = _ptr(ln.get("z")), _ptr(ln.get("stats"))
        e.drop_thr, e.drop_seed, e.drop_scale = ln.get("pre", NO_DROP)
        e.drop_salt = _ptr(ln.get("salt"))
        d.block_n = 64
        L.call("mtts_gemm_ln", C.byref(d), C.byref(e), _stream())
    else:
        L.call("mtts_gemm", C.byref(d), _stream())
    launch_count += 1


# ---------------------------------------------------------------------------------------------
# LengthRegulator
# ---------------------------------------------------------------------------------------------
def length_regulate_index(dur: torch.Tensor, T: int, idx: Optional[torch.Tensor] = None,
                          mel_len: Optional[torch.Tensor] = None):
    """dur [B,L] int64 (targets) or float32 (predicted) -> (idx [B,T] int32, mel_len [B] int64)."""
    global launch_count
    _need_cuda(dur)
    B, Lp = dur.shape
    assert dur.is_contiguous()
    if idx is None:
        idx = torch.empty((B, T), dtype=torch.int32, device=dur.device)
    if mel_len is None:
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


# ---------------------------------------------------------------------------------------------
# CudaOps: the op set the engine is written against.  Every method is one (or a fixed handful of)
# libmtts kernel launch(es) on the current stream; tensors are preallocated by the caller.
# ---------------------------------------------------------------------------------------------
def _p(t):
    return None if t is None else t.data_ptr()


NO_DROP = (0, 0, 1.0)


def drop_site(p: float, seed: int):
    """(thr, seed, scale) of one dropout site (include/mtts.h): thr = floor(p * 2^24), scale = 1/(1-p)."""
    if p is None or p <= 0.0 or seed is None:
        return NO_DROP
    return (int(p * (1 << 24)), int(seed) & 0xFFFFFFFF, 1.0 / (1.0 - p))


class _SideCtx:
    """Fork: everything issued so far on the current stream happens-before the side work."""

    def __init__(self, be):
        self.be = be
        self.ctx = None

    def __enter__(self):
        be = self.be
        if not be.use_side_stream:
            return self
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        be._side.wait_event(ev)
        be._side_used = True
        self.ctx = torch.cuda.stream(be._side)
        self.ctx.__enter__()
        return self

    def __exit__(self, *a):
        if self.ctx is not None:
            self.ctx.__exit__(*a)
        return False


class _BranchCtx:
    """Fork a named auxiliary stream from the current stream (work issued so far happens-before the branch).  Inside the
    branch the weight-gradient side stream is not used (its join on the main stream would otherwise wait for the branch)."""

    def __init__(self, be, name, local=False):
        self.be, self.name, self.ctx, self.saved, self.local = be, name, None, None, local

    def __enter__(self):
        be = self.be
        if not be.use_branches or (self.local and not be.use_local_branches):
            return self
        if self.local:                         # a branch private to the chain (main or another branch) that forks it
            self.name = f"{self.name}@{be._cur_branch or 'main'}"
        st = be._branches.get(self.name)
        if st is None:
            st = be._branches[self.name] = torch.cuda.Stream(device=be.device)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        st.wait_event(ev)
        be._branch_used.add(self.name)
        self.saved = (be.use_side_stream, be._cur_branch)
        be.use_side_stream = False
        be._cur_branch = self.name
        self.ctx = torch.cuda.stream(st)
        self.ctx.__enter__()
        return self

    def __exit__(self, *a):
        if self.ctx is not None:
            self.ctx.__exit__(*a)
            self.be.use_side_stream, self.be._cur_branch = self.saved
        return False


class CudaOps:
    """sm_100a backend.  `split` = 1 (bf16) or 3 (bf16x3 hi/lo, fp32-grade)."""

    name = "cuda"

    def __init__(self, split: int = 3, device: str = "cuda:0"):
        if not torch.cuda.is_available():
            raise L.MttsError("CudaOps needs a CUDA device (B200): there is no CPU fallback")
        L.call("mtts_check_device")
        assert split in (1, 3)
        self.split = split
        self.device = torch.device(device)
        self._side = torch.cuda.Stream(device=self.device)      # weight-gradient GEMMs / bias column sums (off the critical path)
        self._side_used = False
        self.use_side_stream = True
        self._branches = {}            # named auxiliary streams (encoder work that nothing on the main chain waits for)
        self._branch_used = set()
        self._cur_branch = None
        self.use_local_branches = os.environ.get("MTTS_NO_ATT", "0") != "1"
        self.use_branches = os.environ.get("MTTS_NO_BRANCH", "0") != "1"
        self.drop_salt = None          # device int32[1] (uint32 bits) mixed into every dropout seed; None = 0

    # ---- side stream: work that nothing on the critical path waits for (captured as a parallel graph branch) ----
    def side(self):
        return _SideCtx(self)

    def branch(self, name: str, local: bool = False):
        return _BranchCtx(self, name, local)

    def join(self, name: str, local: bool = False):
        """The current stream waits for everything issued on branch `name`."""
        if local:
            name = f"{name}@{self._cur_branch or 'main'}"
        if name in self._branch_used:
            ev = torch.cuda.Event()
            ev.record(self._branches[name])
            torch.cuda.current_stream().wait_event(ev)
            self._branch_used.discard(name)

    def join_side(self):
        if self._side_used:
            ev = torch.cuda.Event()
            ev.record(self._side)
            torch.cuda.current_stream().wait_event(ev)
            self._side_used = False

    # ---- allocation helpers (plumbing) ----
    def empty(self, shape, dtype=torch.float32):
        return torch.empty(shape, dtype=dtype, device=self.device)

    def zeros(self, shape, dtype=torch.float32):
        return torch.zeros(shape, dtype=dtype, device=self.device)

    def zero_(self, t):
        assert t.is_contiguous()
        L.call("mtts_zero", t.data_ptr(), t.numel() * t.element_size(), _stream())     # cudaMemsetAsync: no fill kernel

    # kernels launched per C-ABI call (memsets not counted)
    KERNELS = {"mtts_bn_fwd": 2, "mtts_bn_bwd": 2, "mtts_bn_tfwd": 2, "mtts_bn_tbwd": 2, "mtts_loss_fwd": 2, "mtts_sumsq": 2, "mtts_dot": 2}
    SCALAR_WS = 2048        # floats behind a sumsq / dot result (include/mtts.h MTTS_SCALAR_WS): [0] = the value, [1..] = per-CTA partials

    def _call(self, name, *args):
        global launch_count
        L.call(name, *args, _stream())
        launch_count += self.KERNELS.get(name, 1)

    # ---- GEMM ----
    def gemm(self, a: Opnd, b: Opnd, M, N, K, **kw):
        kw.setdefault("split", self.split)
        if kw.get("ln") is not None:
            kw["ln"] = dict(kw["ln"], salt=self.drop_salt)
        gemm(a, b, M, N, K, **kw)

    # ---- collate on the device: ragged -> padded ----
    def pack_rows(self, src, row_off, B, Lmax, row_bytes, dst):
        """dst[b, t] = src[row_off[b] + t] (t < len_b) else 0; byte-exact (include/mtts.h: mtts_pack_rows)."""
        self._call("mtts_pack_rows", _p(src), _p(row_off), B, Lmax, row_bytes, _p(dst))

    # ---- length regulator ----
    def lr_index(self, dur, T, idx=None, mel_len=None):
        return length_regulate_index(dur, T, idx, mel_len)

    def lr_fwd(self, x, idx, out):
        return length_regulate_fwd(x, idx, out)

    def lr_bwd(self, dy, dur, Lp, out):
        return length_regulate_bwd(dy, dur, Lp, out)

    # ---- row ops ----
    def ln_fwd(self, y, res, gamma, beta, lens, T, R, C, z_out, stats, out, out_hi, out_lo, eps=1e-5, pre=NO_DROP, post=NO_DROP):
        self._call("mtts_ln_fwd", _p(y), _p(res), _p(gamma), _p(beta), _p(lens), T, R, C, eps, _p(z_out), _p(stats),
                   _p(out), _p(out_hi), _p(out_lo), *pre, *post, _p(self.drop_salt))

    def ln_bwd(self, dy, z, stats, gamma, lens, T, R, C, relu_gate, dz, dz_hi, dz_lo, dgamma, dbeta, dbias, pre=NO_DROP,
               post=NO_DROP):
        self._call("mtts_ln_bwd", _p(dy), _p(z), _p(stats), _p(gamma), _p(lens), T, R, C, int(relu_gate), _p(dz),
                   _p(dz_hi), _p(dz_lo), _p(dgamma), _p(dbeta), _p(dbias), *pre, *post, _p(self.drop_salt))

    def ln_tfwd(self, ydot, resdot, z, stats, gamma, gdot, bdot, lens, T, R, C, zdot_out, out, out_hi, out_lo, pre=NO_DROP,
                post=NO_DROP):
        self._call("mtts_ln_tfwd", _p(ydot), _p(resdot), _p(z), _p(stats), _p(gamma), _p(gdot), _p(bdot), _p(lens),
                   T, R, C, _p(zdot_out), _p(out), _p(out_hi), _p(out_lo), *pre, *post, _p(self.drop_salt))

    def ln_tbwd(self, dy, ddy, z, zdot, stats, gamma, gdot, lens, T, R, C, relu_gate, ddz, ddz_hi, ddz_lo,
                ddgamma, ddbeta, ddbias, pre=NO_DROP, post=NO_DROP):
        self._call("mtts_ln_tbwd", _p(dy), _p(ddy), _p(z), _p(zdot), _p(stats), _p(gamma), _p(gdot), _p(lens), T, R, C,
                   int(relu_gate), _p(ddz), _p(ddz_hi), _p(ddz_lo), _p(ddgamma), _p(ddbeta), _p(ddbias), *pre, *post, _p(self.drop_salt))

    def rowdot_fwd(self, h, hdot, w, wdot, b, bdot, lens, T, R, C, out):
        self._call("mtts_rowdot_fwd", _p(h), _p(hdot), _p(w), _p(wdot), _p(b), _p(bdot), _p(lens), T, R, C, _p(out))

    def rowdot_bwd(self, dout, ddout, h, hdot, w, wdot, lens, T, R, C, dh, dw, db):
        self._call("mtts_rowdot_bwd", _p(dout), _p(ddout), _p(h), _p(hdot), _p(w), _p(wdot), _p(lens), T, R, C,
                   _p(dh), _p(dw), _p(db))

    def softmax(self, mode, A, Bm, p_hi, p_lo, pd_hi, pd_lo, klens, nz, H, Lq, Lk, ld, o_hi, o_lo):
        self._call("mtts_softmax", mode, _p(A), _p(Bm), _p(p_hi), _p(p_lo), _p(pd_hi), _p(pd_lo), _p(klens), nz, H,
                   Lq, Lk, ld, _p(o_hi), _p(o_lo))

    # ---- fused attention (csrc/mtts_attn.cu) ----
    def _attn_desc(self, qkv_hi, qkv_lo, klens, B, H, T, dk, lse, split=None):
        d = L.AttnDesc()
        d.B, d.H, d.T, d.dk = B, H, T, dk
        d.Tl = lse.shape[-1]
        d.split = split or self.split
        d.scale = 1.0 / (dk ** 0.5)
        d.qkv_hi, d.qkv_lo, d.klens, d.lse = _p(qkv_hi), _p(qkv_lo), _p(klens), _p(lse)
        return d

    def attn_fwd(self, qkv_hi, qkv_lo, klens, B, H, T, dk, o_hi, o_lo, lse, p_hi=None, p_lo=None, Tp=0, split=None):
        """o = softmax(q k^T / sqrt(dk), keys >= klens masked) v; lse [B,H,Tl] (log2 domain); optional emit of P [B,H,T,Tp]."""
        global launch_count
        _need_cuda(qkv_hi, o_hi, lse)
        d = self._attn_desc(qkv_hi, qkv_lo, klens, B, H, T, dk, lse, split)
        d.o_hi, d.o_lo, d.p_hi, d.p_lo, d.Tp = _p(o_hi), _p(o_lo), _p(p_hi), _p(p_lo), Tp
        L.call("mtts_attn_fwd", C.byref(d), _stream())
        launch_count += 1

    def attn_bwd(self, parts, qkv_hi, qkv_lo, klens, B, H, T, dk, o_hi, o_lo, lse, do_hi, do_lo, dvec, dqkv_hi, dqkv_lo,
                 dp=None, ds_hi=None, ds_lo=None, Tp=0, split=None):
        """parts: mask of L.ATTN_PREP (dvec = rowsum(do*o)), L.ATTN_DQ (dq block of dqkv; optional emit of dP / dS),
        L.ATTN_DK, L.ATTN_DV (dk / dv blocks of dqkv)."""
        global launch_count
        _need_cuda(qkv_hi, do_hi, lse, dvec)
        d = self._attn_desc(qkv_hi, qkv_lo, klens, B, H, T, dk, lse, split)
        d.o_hi, d.o_lo, d.do_hi, d.do_lo, d.dvec = _p(o_hi), _p(o_lo), _p(do_hi), _p(do_lo), _p(dvec)
        d.dqkv_hi, d.dqkv_lo, d.dp, d.ds_hi, d.ds_lo, d.Tp = _p(dqkv_hi), _p(dqkv_lo), _p(dp), _p(ds_hi), _p(ds_lo), Tp
        L.call("mtts_attn_bwd", C.byref(d), int(parts), _stream())
        launch_count += bin(int(parts)).count("1")

    # ---- gathers / broadcasts / sums ----
    def embed_fwd(self, idx, table, base, pos, T, R, C, out, hi, lo):
        self._call("mtts_embed_fwd", _p(idx), _p(table), _p(base), _p(pos), T, R, C, _p(out), _p(hi), _p(lo))

    def embed_bwd(self, idx, dy, R, C, skip_idx, scale, dtable):
        self._call("mtts_embed_bwd", _p(idx), _p(dy), R, C, skip_idx, scale, _p(dtable))

    def bucketize(self, v, bins, nb, R, out):
        self._call("mtts_bucketize", _p(v), _p(bins), nb, R, _p(out))

    def add_rowvec(self, x, vec, vec_bstride, pos, B, T, C, out, hi, lo):
        self._call("mtts_add_rowvec", _p(x), _p(vec), vec_bstride, _p(pos), B, T, C, _p(out), _p(hi), _p(lo))

    def spk_embed(self, ids, table, n, C, average, n_out, out):
        self._call("mtts_spk_embed", _p(ids), _p(table), n, C, int(average), n_out, _p(out))

    def spk_embed_bwd(self, ids, dspk, n, C, average, n_out, scale, dtable):
        self._call("mtts_spk_embed_bwd", _p(ids), _p(dspk), n, C, int(average), n_out, scale, _p(dtable))

    def colsum(self, f32, hi, lo, nb, R, C, out):
        self._call("mtts_colsum", _p(f32), _p(hi), _p(lo), nb, R, C, _p(out))

    # ---- batch norm ----
    def bn_fwd(self, x, gamma, beta, R, C, tanh_flag, running_mean, running_var, ws, stats, out, hi, lo,
               eps=1e-5, momentum=0.1, drop=NO_DROP):
        self._call("mtts_bn_fwd", _p(x), _p(gamma), _p(beta), R, C, eps, momentum, int(tanh_flag), _p(running_mean),
                   _p(running_var), _p(ws), _p(stats), _p(out), _p(hi), _p(lo), *drop, _p(self.drop_salt))

    def bn_bwd(self, dout, o, x, stats, gamma, R, C, tanh_flag, ws, dx, hi, lo, dgamma, dbeta, beta=None, drop=NO_DROP):
        self._call("mtts_bn_bwd", _p(dout), _p(o), _p(x), _p(stats), _p(gamma), R, C, int(tanh_flag), _p(ws), _p(dx),
                   _p(hi), _p(lo), _p(dgamma), _p(dbeta), *drop, _p(self.drop_salt))

    def bn_tfwd(self, xdot, x, stats, gamma, gdot, bdot, o, R, C, tanh_flag, ws, tsums, odot, hi, lo, beta=None,
                drop=NO_DROP):
        self._call("mtts_bn_tfwd", _p(xdot), _p(x), _p(stats), _p(gamma), _p(gdot), _p(bdot), _p(o), R, C,
                   int(tanh_flag), _p(ws), _p(tsums), _p(odot), _p(hi), _p(lo), *drop, _p(self.drop_salt))

    def bn_tbwd(self, dout, ddout, o, odot, x, xdot, stats, tsums, gamma, gdot, R, C, tanh_flag, ws, ddx, hi, lo,
                ddgamma, ddbeta, beta=None, bdot=None, drop=NO_DROP):
        # beta / bdot are only used by the CPU restatement (the kernels use the saved tanh output o / odot)
        self._call("mtts_bn_tbwd", _p(dout), _p(ddout), _p(o), _p(odot), _p(x), _p(xdot), _p(stats), _p(tsums),
                   _p(gamma), _p(gdot), R, C, int(tanh_flag), _p(ws), _p(ddx), _p(hi), _p(lo), _p(ddgamma), _p(ddbeta),
                   *drop, _p(self.drop_salt))

    def bn_eval(self, x, gamma, beta, running_mean, running_var, R, C, tanh_flag, out, hi, lo, eps=1e-5):
        self._call("mtts_bn_eval", _p(x), _p(gamma), _p(beta), _p(running_mean), _p(running_var), R, C, eps, int(tanh_flag),
                   _p(out), _p(hi), _p(lo))

    # ---- free-running synthesis / vocoder-side decode ----
    def duration_round(self, logd, d_control, out):
        self._call("mtts_duration_round", _p(logd), float(d_control), logd.numel(), _p(out))

    def unary(self, op, x, a, b, out, hi=None, lo=None):
        """op 0: log(max(x,a)*b); 1: exp(x)*a; 2: x*a (include/mtts.h MTTS_UN_*)."""
        self._call("mtts_unary", int(op), _p(x), x.numel(), float(a), float(b), _p(out), _p(hi), _p(lo))

    def reflect_pad(self, x, B, N, pad, ld, out, hi, lo):
        self._call("mtts_reflect_pad", _p(x), B, N, pad, ld, _p(out), _p(hi), _p(lo))

    def stft_polar(self, ri, R, nb, ld, im_off, ldm, mag, phase, energy, mag_hi=None, mag_lo=None):
        self._call("mtts_stft_polar", _p(ri), R, nb, ld, im_off, ldm, _p(mag), _p(phase), _p(energy), _p(mag_hi), _p(mag_lo))

    def stft_recombine(self, mag, phase, ri, R, nb, ld, im_off, ldm, hi, lo):
        self._call("mtts_stft_recombine", _p(mag), _p(phase), _p(ri), R, nb, ld, im_off, ldm, _p(hi), _p(lo))

    def istft_finish(self, ola, wsum, tiny, scale, B, n, trim, out):
        self._call("mtts_istft_finish", _p(ola), _p(wsum), float(tiny), float(scale), B, n, trim, _p(out))

    # ---- loss ----
    def loss_fwd(self, mel, post, mel_tgt, mel_lens, p, p_tgt, e, e_tgt, logd, dur, src_lens, B, T, Lp, NM, ws, out6,
                 counts):
        self._call("mtts_loss_fwd", _p(mel), _p(post), _p(mel_tgt), _p(mel_lens), _p(p), _p(p_tgt), _p(e), _p(e_tgt),
                   _p(logd), _p(dur), _p(src_lens), B, T, Lp, NM, _p(ws), _p(out6), _p(counts))

    def loss_bwd(self, mel, post, mel_tgt, mel_lens, p, p_tgt, e, e_tgt, logd, dur, src_lens, B, T, Lp, NM, counts,
                 scale, tangent, dmel, dpost, dp, de, dlogd):
        self._call("mtts_loss_bwd", _p(mel), _p(post), _p(mel_tgt), _p(mel_lens), _p(p), _p(p_tgt), _p(e), _p(e_tgt),
                   _p(logd), _p(dur), _p(src_lens), B, T, Lp, NM, _p(counts), scale, int(tangent), _p(dmel), _p(dpost),
                   _p(dp), _p(de), _p(dlogd))

    # ---- flat elementwise ----
    def split_(self, src, hi, lo):
        self._call("mtts_split", _p(src), _p(hi), _p(lo), src.numel())

    def sgd_split(self, theta, g, lr, out, hi, lo):
        self._call("mtts_sgd_split", _p(theta), _p(g), lr, _p(out), _p(hi), _p(lo), theta.numel())

    def axpby(self, a, x, b, y):
        self._call("mtts_axpby", a, _p(x), b, _p(y), x.numel())

    def sumsq(self, x, out):
        assert out.numel() >= self.SCALAR_WS, "sumsq: the result buffer carries the per-CTA partials (SCALAR_WS floats)"
        self._call("mtts_sumsq", _p(x), x.numel(), _p(out))

    def dot(self, x, y, out):
        assert out.numel() >= self.SCALAR_WS, "dot: the result buffer carries the per-CTA partials (SCALAR_WS floats)"
        self._call("mtts_dot", _p(x), _p(y), x.numel(), _p(out))

    def adam_clip(self, p, g, m, v, sumsq, gscale, max_norm, hyper, beta1, beta2, eps, hi, lo):
        self._call("mtts_adam_clip", _p(p), _p(g), _p(m), _p(v), _p(sumsq), gscale, max_norm, _p(hyper), beta1, beta2,
                   eps, _p(hi), _p(lo), p.numel())
