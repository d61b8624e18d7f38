"""Per-op CPU restatement of the libmtts op set — TEST INFRASTRUCTURE ONLY.

`RefOps` implements, with plain PyTorch on the CPU, exactly the op interface the engine
(`meta-tts_b200/engine.py`) is written against (`meta-tts_b200/ops.py: CudaOps`).  It exists so that
  * `-m "not gpu"` tests can validate the host logic (descriptor construction, 4-pass orchestration,
    MAML recursion) against the autograd oracle without a GPU, and
  * `-m gpu` tests can check every CUDA kernel in isolation against an independently derived result
    (the tangent / tangent-backward forms here come from torch.func.jvp / vjp, not from the
    hand-derived formulas the kernels implement).
The product never imports this module; CudaOps raises without a GPU (no fallback).
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn.functional as F
from torch.func import jvp, vjp

SRC_ZERO, SRC_Z0, SRC_Z1, SRC_TAP, SRC_KB = 0, 1, 2, 3, 4
MAJOR_K, MAJOR_MN = 0, 1
EPI_RELU, EPI_ACCUM, EPI_GATE, EPI_BIAS_ROW, EPI_ADD_C = 1, 2, 4, 8, 16


def _split(v: torch.Tensor):
    hi = v.to(torch.bfloat16)
    lo = (v - hi.float()).to(torch.bfloat16)
    return hi, lo


def _put(dst: Optional[torch.Tensor], v: torch.Tensor):
    if dst is not None:
        dst.copy_(v.reshape(dst.shape).to(dst.dtype))


def _put_split(hi, lo, v):
    if hi is not None:
        h, l = _split(v.float())
        hi.copy_(h.reshape(hi.shape))
        if lo is not None:
            lo.copy_(l.reshape(lo.shape))


def _val(hi, lo):
    v = hi.float()
    if lo is not None:
        v = v + lo.float()
    return v


def _row_mask(lens, T, R):
    if lens is None:
        return torch.ones(R, dtype=torch.bool)
    r = torch.arange(R)
    return (r % T) < lens[(r // T)]


NO_DROP = (0, 0, 1.0)


class RefOps:
    name = "ref"

    def __init__(self, split: int = 3, dtype=torch.float64):
        self.split = split
        self.device = torch.device("cpu")
        self.acc = dtype
        self.n_calls = 0

    def empty(self, shape, dtype=torch.float32):
        return torch.full(shape, float("nan") if dtype.is_floating_point else 0, dtype=dtype)

    def zeros(self, shape, dtype=torch.float32):
        return torch.zeros(shape, dtype=dtype)

    def zero_(self, t):
        t.zero_()

    def side(self):
        import contextlib
        return contextlib.nullcontext()

    def join_side(self):
        pass

    def branch(self, name, local=False):
        import contextlib
        return contextlib.nullcontext()

    def join(self, name, local=False):
        pass

    # ------------------------------------------------------------------------------------------
    # GEMM descriptor emulator (semantics of mtts_gemm: TMA coordinates, OOB zero fill, epilogue)
    # ------------------------------------------------------------------------------------------
    def _fetch(self, op, n_mn, K, z0, z1, tap, kb, split, term=0):
        pick = lambda s: {SRC_ZERO: 0, SRC_Z0: z0, SRC_Z1: z1, SRC_TAP: tap, SRC_KB: kb}[s]  # noqa: E731
        d = list(op.dims) + [1] * (4 - len(op.dims))
        s = list(op.strides) + [0] * (4 - len(op.strides))
        c2, c3 = pick(op.src2), pick(op.src3)
        shift = op.shift_base + op.shift_step * pick(op.shift_src)
        t_hi, t_lo = (op.hi, op.lo) if term == 0 else (op.hi2, op.lo2)
        flat_hi = t_hi.reshape(-1)
        flat = flat_hi.float()
        if split == 3:
            assert t_lo is not None, "split=3 operand without lo"
            flat = flat + t_lo.reshape(-1).float()
        mn = torch.arange(n_mn)
        kk = torch.arange(K)
        if c2 >= d[2] or c3 >= d[3]:
            return torch.zeros(n_mn, K, dtype=self.acc)
        base = op.offset + c2 * s[2] + c3 * s[3]
        if op.major == MAJOR_K:
            rows = mn + shift
            valid = ((rows >= 0) & (rows < d[1]))[:, None] & (kk < d[0])[None, :]
            idx = base + rows.clamp(0, d[1] - 1)[:, None] * s[1] + kk.clamp(max=d[0] - 1)[None, :]
        else:
            rows = kk + shift
            valid = (mn < d[0])[:, None] & ((rows >= 0) & (rows < d[1]))[None, :]
            idx = base + mn.clamp(max=d[0] - 1)[:, None] + rows.clamp(0, d[1] - 1)[None, :] * s[1]
        assert int(idx.max()) < flat.numel(), "operand view exceeds its buffer"
        v = flat[idx.reshape(-1)].reshape(n_mn, K).to(self.acc)
        return torch.where(valid, v, torch.zeros((), dtype=self.acc))

    def gemm(self, a, b, M, N, K, *, c_f32=None, c_hi=None, c_lo=None, ldc, c_off=0, c_sz0=0, c_sz1=0, alpha=1.0,
             bias=None, bias_sz0=0, gate=None, flags=0, ntaps=1, nkb=1, nz0=1, nz1=1, split=None, block_n=0, ksplit=1, pair=False, ln=None):
        self.n_calls += 1
        if ln is not None:      # mtts_gemm_ln: the same product, then dropout -> + residual -> LayerNorm -> pad rows as the epilogue
            assert N == 256 and ldc == N and nz0 == nz1 == 1 and ksplit == 1 and not pair and flags == 0 and c_off == 0
            y = torch.zeros(M, N)
            self.gemm(a, b, M, N, K, c_f32=y, ldc=N, alpha=alpha, bias=bias, ntaps=ntaps, nkb=nkb, split=split)
            self.n_calls -= 2
            self.ln_fwd(y, ln.get("res"), ln["gamma"], ln["beta"], ln.get("lens"), ln["T"], M, N, ln.get("z"), ln.get("stats"),
                        c_f32, c_hi, c_lo, eps=ln.get("eps", 1e-5), pre=ln.get("pre", NO_DROP))
            return
        split = split or self.split
        assert c_f32 is not None or c_hi is not None
        if flags & EPI_ADD_C:
            assert c_f32 is not None and ksplit == 1 and not (flags & EPI_ACCUM)
        if ksplit > 1:
            assert (flags & EPI_ACCUM) and c_hi is None
        # TMA legality (what cuTensorMapEncodeTiled would reject)
        for op in (a, b):
            st = list(op.strides)
            assert st[0] == 1 and all(x % 8 == 0 for x in st[1:] if x), f"illegal strides {st}"
            assert op.offset % 8 == 0, "operand base not 16B aligned"
        mi = torch.arange(M)[:, None]
        ni = torch.arange(N)[None, :]
        for z1 in range(nz1):
            for z0 in range(nz0):
                acc = torch.zeros(M, N, dtype=self.acc)
                nterms = 2 if (getattr(a, "hi2", None) is not None or getattr(b, "hi2", None) is not None) else 1
                for term in range(nterms):
                    for tap in range(ntaps):
                        for kb in range(nkb):
                            A = self._fetch(a, M, K, z0, z1, tap, kb, split, term)
                            Bm = self._fetch(b, N, K, z0, z1, tap, kb, split, term)
                            acc += A @ Bm.t()
                v = acc * alpha
                if bias is not None:
                    bv = bias.reshape(-1)[z0 * bias_sz0:]
                    v = v + (bv[:M, None] if (flags & EPI_BIAS_ROW) else bv[None, :N]).to(self.acc)
                idx = (c_off + z0 * c_sz0 + z1 * c_sz1 + mi * ldc + ni).reshape(-1)
                if flags & EPI_ADD_C:
                    v = v + c_f32.reshape(-1)[idx].reshape(M, N).to(self.acc)
                if flags & EPI_RELU:
                    v = v.clamp_min(0)
                if flags & EPI_GATE:
                    g = gate.reshape(-1)[idx].reshape(M, N).float()
                    v = torch.where(g > 0, v, torch.zeros((), dtype=self.acc))
                v32 = v.float()
                if c_f32 is not None:
                    flat = c_f32.reshape(-1)
                    if flags & EPI_ACCUM:
                        flat[idx] = flat[idx] + v32.reshape(-1)
                    else:
                        flat[idx] = v32.reshape(-1)
                if c_hi is not None:
                    h, l = _split(v32)
                    c_hi.reshape(-1)[idx] = h.reshape(-1)
                    if c_lo is not None:
                        c_lo.reshape(-1)[idx] = l.reshape(-1)

    # ------------------------------------------------------------------------------------------
    # ragged -> padded pack (utils/tools.py:270-301 pad_1D / pad_2D semantics, bytes copied unchanged)
    # ------------------------------------------------------------------------------------------
    def pack_rows(self, src, row_off, B, Lmax, row_bytes, dst):
        self.n_calls += 1
        s8 = src.contiguous().view(torch.uint8).reshape(-1, row_bytes)
        d8 = dst.view(torch.uint8).reshape(B, Lmax, row_bytes)
        d8.zero_()
        for b in range(B):
            r0, r1 = int(row_off[b]), int(row_off[b + 1])
            d8[b, :r1 - r0] = s8[r0:r1]

    # ------------------------------------------------------------------------------------------
    # LengthRegulator
    # ------------------------------------------------------------------------------------------
    def lr_index(self, dur, T, idx_out=None, len_out=None):
        d = dur.to(torch.float64).trunc().clamp_min(0).long() if dur.is_floating_point() else dur.clamp_min(0)
        cum = d.cumsum(1)
        t = torch.arange(T)[None, :].expand(dur.shape[0], -1).contiguous()
        idx = torch.searchsorted(cum, t, right=True).int()
        mel_len = cum[:, -1].clone()
        idx = torch.where(t < mel_len[:, None], idx, torch.full_like(idx, -1))
        if idx_out is not None:
            idx_out.copy_(idx)
            len_out.copy_(mel_len)
            return idx_out, len_out
        return idx, mel_len

    def lr_fwd(self, x, idx, out):
        B, T = idx.shape
        g = torch.gather(x, 1, idx.clamp_min(0).long()[..., None].expand(-1, -1, x.shape[2]))
        out.copy_(torch.where((idx >= 0)[..., None], g, torch.zeros(())))
        return out

    def lr_bwd(self, dy, dur, Lp, out):
        idx, _ = self.lr_index(dur, dy.shape[1])
        out.zero_()
        m = (idx >= 0)[..., None].float()
        out.scatter_add_(1, idx.clamp_min(0).long()[..., None].expand(-1, -1, dy.shape[2]), dy * m)
        return out

    # ------------------------------------------------------------------------------------------
    # LayerNorm family: forward by F.layer_norm, the rest by autograd / functorch (independent of
    # the hand-derived kernel formulas)
    # ------------------------------------------------------------------------------------------
    @staticmethod
    def _ln(z, gamma, beta, mask, eps=1e-5):
        return F.layer_norm(z, (z.shape[-1],), gamma, beta, eps) * mask[:, None]

    # dropout sites: (thr, seed, scale) as in include/mtts.h; the same integer hash evaluated on the host
    drop_salt = None

    def _dm(self, site, R, C):
        thr, seed, scale = site
        if thr == 0:
            return 1.0
        from oracle.fs2_oracle import drop_keep
        salt = (int(self.drop_salt.reshape(-1)[0].item()) & 0xFFFFFFFF) if self.drop_salt is not None else 0
        eff = (int(seed) + salt * 0x632BE5AB) & 0xFFFFFFFF
        return drop_keep(thr, eff, R * C).reshape(R, C).double() * float(torch.tensor(scale, dtype=torch.float32))

    def ln_fwd(self, y, res, gamma, beta, lens, T, R, C, z_out, stats, out, out_hi, out_lo, eps=1e-5, pre=NO_DROP,
               post=NO_DROP):
        self.n_calls += 1
        z = y.reshape(R, C).double() * self._dm(pre, R, C) + (res.reshape(R, C).double() if res is not None else 0)
        mask = _row_mask(lens, T, R).double()
        o = self._ln(z, gamma.double(), beta.double(), mask, eps) * self._dm(post, R, C)
        if stats is not None:
            mean = z.mean(1)
            rstd = 1.0 / torch.sqrt(z.var(1, unbiased=False) + eps)
            stats.copy_(torch.stack([mean, rstd], 1).reshape(stats.shape).float())
        _put(z_out, z)
        _put(out, o)
        _put_split(out_hi, out_lo, o)

    def ln_bwd(self, dy, z, stats, gamma, lens, T, R, C, relu_gate, dz, dz_hi, dz_lo, dgamma, dbeta, dbias, pre=NO_DROP,
               post=NO_DROP):
        self.n_calls += 1
        zz = z.reshape(R, C).double()
        mask = _row_mask(lens, T, R).double()
        g = gamma.double()
        beta0 = torch.zeros(C, dtype=torch.float64)
        _, fn = vjp(lambda a, b, c: self._ln(a, b, c, mask), zz, g, beta0)
        dzz, dg, db = fn(dy.reshape(R, C).double() * self._dm(post, R, C))
        if relu_gate:
            dzz = dzz * (zz > 0)
        _put(dz, dzz)                               # residual path
        dzz = dzz * self._dm(pre, R, C)             # branch path
        _put_split(dz_hi, dz_lo, dzz)
        if dgamma is not None:
            dgamma += dg.float()
        if dbeta is not None:
            dbeta += db.float()
        if dbias is not None:
            dbias += dzz.sum(0).float()

    def ln_tfwd(self, ydot, resdot, z, stats, gamma, gdot, bdot, lens, T, R, C, zdot_out, out, out_hi, out_lo, pre=NO_DROP,
                post=NO_DROP):
        self.n_calls += 1
        zz = z.reshape(R, C).double()
        zd = ydot.reshape(R, C).double() * self._dm(pre, R, C) + (resdot.reshape(R, C).double() if resdot is not None else 0)
        mask = _row_mask(lens, T, R).double()
        g = gamma.double()
        gd = gdot.double() if gdot is not None else torch.zeros_like(g)
        bd = bdot.double() if bdot is not None else torch.zeros_like(g)
        _, od = jvp(lambda a, b, c: self._ln(a, b, c, mask), (zz, g, torch.zeros_like(g)), (zd, gd, bd))
        od = od * self._dm(post, R, C)
        _put(zdot_out, zd)
        _put(out, od)
        _put_split(out_hi, out_lo, od)

    def ln_tbwd(self, dy, ddy, z, zdot, stats, gamma, gdot, lens, T, R, C, relu_gate, ddz, ddz_hi, ddz_lo, ddgamma,
                ddbeta, ddbias, pre=NO_DROP, post=NO_DROP):
        self.n_calls += 1
        zz, zd = z.reshape(R, C).double(), zdot.reshape(R, C).double()
        mpost = self._dm(post, R, C)
        d, dd = dy.reshape(R, C).double() * mpost, ddy.reshape(R, C).double() * mpost
        mask = _row_mask(lens, T, R).double()
        g = gamma.double()
        gd = gdot.double() if gdot is not None else torch.zeros_like(g)
        b0 = torch.zeros_like(g)

        def bwd(a, gg, bb, cot):
            _, fn = vjp(lambda p, q, r: self._ln(p, q, r, mask), a, gg, bb)
            return fn(cot)

        _, (tz, tg, tb) = jvp(bwd, (zz, g, b0, d), (zd, gd, torch.zeros_like(g), dd))
        if relu_gate:
            tz = tz * (zz > 0)
        _put(ddz, tz)
        tz = tz * self._dm(pre, R, C)
        _put_split(ddz_hi, ddz_lo, tz)
        if ddgamma is not None:
            ddgamma += tg.float()
        if ddbeta is not None:
            ddbeta += tb.float()
        if ddbias is not None:
            ddbias += tz.sum(0).float()

    # ------------------------------------------------------------------------------------------
    # rowdot (Linear(C,1) head + row mask)
    # ------------------------------------------------------------------------------------------
    def rowdot_fwd(self, h, hdot, w, wdot, b, bdot, lens, T, R, C, out):
        self.n_calls += 1
        mask = _row_mask(lens, T, R).double()
        hh, ww = h.reshape(R, C).double(), w.reshape(C).double()
        if hdot is None:
            v = hh @ ww + b.double().reshape(())
        else:
            v = hdot.reshape(R, C).double() @ ww
            if wdot is not None:
                v = v + hh @ wdot.reshape(C).double()
            if bdot is not None:
                v = v + bdot.double().reshape(())
        _put(out, v * mask)

    def rowdot_bwd(self, dout, ddout, h, hdot, w, wdot, lens, T, R, C, dh, dw, db):
        self.n_calls += 1
        mask = _row_mask(lens, T, R).double()
        hh, ww = h.reshape(R, C).double(), w.reshape(C).double()
        s = dout.reshape(R).double() * mask
        if ddout is None:
            _put(dh, s[:, None] * ww[None, :])
            if dw is not None:
                dw += (s @ hh).reshape(dw.shape).float()
            if db is not None:
                db += s.sum().float()
        else:
            sd = ddout.reshape(R).double() * mask
            wd = wdot.reshape(C).double() if wdot is not None else torch.zeros_like(ww)
            _put(dh, sd[:, None] * ww[None, :] + s[:, None] * wd[None, :])
            if dw is not None:
                dw += (sd @ hh + s @ hdot.reshape(R, C).double()).reshape(dw.shape).float()
            if db is not None:
                db += sd.sum().float()

    # ------------------------------------------------------------------------------------------
    # masked softmax family
    # ------------------------------------------------------------------------------------------
    def softmax(self, mode, A, Bm, p_hi, p_lo, pd_hi, pd_lo, klens, nz, H, Lq, Lk, ld, o_hi, o_lo):
        self.n_calls += 1
        Av = A.reshape(nz, Lq, ld)[..., :Lk].double()
        if klens is not None:
            kl = klens[(torch.arange(nz) // H)].clamp(max=Lk)
        else:
            kl = torch.full((nz,), Lk)
        km = (torch.arange(Lk)[None, :] < kl[:, None])[:, None, :]          # [nz,1,Lk] valid keys
        sm = lambda s: torch.softmax(s.masked_fill(~km, -math.inf), -1)  # noqa: E731
        if mode == 0:
            o = sm(Av)
        else:
            P = _val(p_hi, p_lo).reshape(nz, Lq, ld)[..., :Lk].double()
            if mode == 1:
                o = P * (Av - (P * Av).sum(-1, keepdim=True)) * km
            else:
                Pd = _val(pd_hi, pd_lo).reshape(nz, Lq, ld)[..., :Lk].double()
                Bv = Bm.reshape(nz, Lq, ld)[..., :Lk].double()
                d = (P * Av).sum(-1, keepdim=True)
                dd = (Pd * Av + P * Bv).sum(-1, keepdim=True)
                o = (Pd * (Av - d) + P * (Bv - dd)) * km
        full = torch.zeros(nz, Lq, ld, dtype=torch.float64)
        full[..., :Lk] = o
        _put_split(o_hi, o_lo, full)

    # ------------------------------------------------------------------------------------------
    # fused attention (mtts_attn_fwd / mtts_attn_bwd): ScaledDotProductAttention, Modules.py:14-25, and its autograd,
    # restated with torch ops.  lse is kept in the log2 domain like the kernels do.
    # ------------------------------------------------------------------------------------------
    def _attn_parts(self, qkv_hi, qkv_lo, klens, B, H, T, dk):
        x = _val(qkv_hi, qkv_lo if self.split == 3 else None).reshape(B, T, 3, H, dk).to(self.acc)
        q, k, v = (x[:, :, i].permute(0, 2, 1, 3) for i in range(3))                     # [B,H,T,dk]
        kl = klens.clamp(max=T) if klens is not None else torch.full((B,), T)
        km = (torch.arange(T)[None, :] < kl[:, None])[:, None, None, :]                    # [B,1,1,T] valid keys
        s = (q @ k.transpose(-1, -2)) / math.sqrt(dk)
        return q, k, v, s.masked_fill(~km, -math.inf), km

    def attn_fwd(self, qkv_hi, qkv_lo, klens, B, H, T, dk, o_hi, o_lo, lse, p_hi=None, p_lo=None, Tp=0, split=None):
        self.n_calls += 1
        q, k, v, s, km = self._attn_parts(qkv_hi, qkv_lo, klens, B, H, T, dk)
        P = torch.softmax(s, -1)
        lse[..., :T] = (torch.logsumexp(s, -1) / math.log(2.0)).float()
        _put_split(o_hi, o_lo if self.split == 3 else None, (P @ v).permute(0, 2, 1, 3).reshape(B * T, H * dk))
        if p_hi is not None:
            full = torch.zeros(B, H, T, Tp, dtype=self.acc)
            full[..., :T] = P
            _put_split(p_hi, p_lo if self.split == 3 else None, full)

    def attn_bwd(self, parts, qkv_hi, qkv_lo, klens, B, H, T, dk, o_hi, o_lo, lse, do_hi, do_lo, dvec, dqkv_hi, dqkv_lo,
                 dp=None, ds_hi=None, ds_lo=None, Tp=0, split=None):
        self.n_calls += 1
        lo = (lambda t: t if self.split == 3 else None)
        dO = _val(do_hi, lo(do_lo)).reshape(B, T, H, dk).permute(0, 2, 1, 3).to(self.acc)
        if parts & 1:
            O = _val(o_hi, lo(o_lo)).reshape(B, T, H, dk).permute(0, 2, 1, 3).to(self.acc)
            dvec[..., :T] = (dO * O).sum(-1).float()
        if not parts & 14:
            return
        q, k, v, s, km = self._attn_parts(qkv_hi, qkv_lo, klens, B, H, T, dk)
        P = torch.exp(s - (lse[..., :T].to(self.acc) * math.log(2.0))[..., None])          # recomputed from the saved log-sum-exp
        dP = dO @ v.transpose(-1, -2)
        dS = P * (dP - dvec[..., :T].to(self.acc)[..., None])
        sc = 1.0 / math.sqrt(dk)
        cur = _val(dqkv_hi, lo(dqkv_lo)).reshape(B, T, 3, H, dk).to(self.acc).clone()
        if parts & 2:
            cur[:, :, 0] = (sc * (dS @ k)).permute(0, 2, 1, 3)
            if ds_hi is not None:
                full = torch.zeros(B, H, T, Tp, dtype=self.acc)
                full[..., :T] = dS
                _put_split(ds_hi, lo(ds_lo), full)
                full[..., :T] = dP
                _put(dp, full)
        if parts & 4:
            cur[:, :, 1] = (sc * (dS.transpose(-1, -2) @ q)).permute(0, 2, 1, 3)
        if parts & 8:
            cur[:, :, 2] = (P.transpose(-1, -2) @ dO).permute(0, 2, 1, 3)
        _put_split(dqkv_hi, lo(dqkv_lo), cur.reshape(B * T, 3 * H * dk))

    # ------------------------------------------------------------------------------------------
    # gathers / broadcasts / sums
    # ------------------------------------------------------------------------------------------
    def embed_fwd(self, idx, table, base, pos, T, R, C, out, hi, lo):
        self.n_calls += 1
        v = table.reshape(-1, C)[idx.reshape(R)].double()
        if base is not None:
            v = v + base.reshape(R, C).double()
        if pos is not None:
            v = v + pos.reshape(-1, C)[torch.arange(R) % T].double()
        _put(out, v)
        _put_split(hi, lo, v)

    def embed_bwd(self, idx, dy, R, C, skip_idx, scale, dtable):
        self.n_calls += 1
        ii = idx.reshape(R)
        keep = (ii != skip_idx).float()[:, None]
        dtable.reshape(-1, C).index_add_(0, ii, dy.reshape(R, C) * keep * scale)

    def bucketize(self, v, bins, nb, R, out):
        self.n_calls += 1
        out.copy_(torch.bucketize(v.reshape(R), bins.reshape(nb)).reshape(out.shape))

    def add_rowvec(self, x, vec, vec_bstride, pos, B, T, C, out, hi, lo):
        self.n_calls += 1
        v = x.reshape(B, T, C).double()
        if vec is not None:
            vv = torch.stack([vec.reshape(-1)[b * vec_bstride: b * vec_bstride + C] for b in range(B)])
            v = v + vv[:, None, :].double()
        if pos is not None:
            v = v + pos.reshape(-1, C)[:T][None].double()
        _put(out, v)
        _put_split(hi, lo, v)

    def spk_embed(self, ids, table, n, C, average, n_out, out):
        self.n_calls += 1
        e = table.reshape(-1, C)[ids.reshape(n)]
        if average:
            e = e.mean(0, keepdim=True).expand(n_out, -1)
        _put(out, e)

    def spk_embed_bwd(self, ids, dspk, n, C, average, n_out, scale, dtable):
        self.n_calls += 1
        d = dspk.reshape(n_out, C)
        tab = dtable.reshape(-1, C)
        if average:
            s = d.sum(0) * scale / n
            for i in range(n):
                tab[ids[i]] += s
        else:
            for i in range(n):
                tab[ids[i]] += scale * d[i]

    def colsum(self, f32, hi, lo, nb, R, C, out):
        self.n_calls += 1
        v = f32.reshape(nb, R, C).double() if f32 is not None else _val(hi, lo).reshape(nb, R, C).double()
        out += v.sum(1).reshape(out.shape).float()

    # ------------------------------------------------------------------------------------------
    # BatchNorm (train) + tanh
    # ------------------------------------------------------------------------------------------
    @staticmethod
    def _bn(x, gamma, beta, tanh_flag, eps=1e-5, m=1.0):
        mean = x.mean(0)
        var = x.var(0, unbiased=False)
        y = (x - mean) / torch.sqrt(var + eps) * gamma + beta
        return (torch.tanh(y) if tanh_flag else y) * m

    def bn_fwd(self, x, gamma, beta, R, C, tanh_flag, running_mean, running_var, ws, stats, out, hi, lo, eps=1e-5,
               momentum=0.1, drop=NO_DROP):
        self.n_calls += 1
        xx = x.reshape(R, C).double()
        o = self._bn(xx, gamma.double(), beta.double(), tanh_flag, eps, self._dm(drop, R, C))
        mean, var = xx.mean(0), xx.var(0, unbiased=False)
        stats.copy_(torch.cat([mean, 1.0 / torch.sqrt(var + eps)]).float().reshape(stats.shape))
        if running_mean is not None:
            running_mean.mul_(1 - momentum).add_(momentum * mean.float())
            running_var.mul_(1 - momentum).add_(momentum * (var * R / max(R - 1, 1)).float())
        _put(out, o)
        _put_split(hi, lo, o)

    def bn_bwd(self, dout, o, x, stats, gamma, R, C, tanh_flag, ws, dx, hi, lo, dgamma, dbeta, beta=None, drop=NO_DROP):
        self.n_calls += 1
        xx, g = x.reshape(R, C).double(), gamma.double()
        bt = beta.double() if beta is not None else torch.zeros_like(g)
        assert beta is not None or not tanh_flag, "RefOps needs beta for the tanh layers"
        m = self._dm(drop, R, C)
        _, fn = vjp(lambda a, b, c: self._bn(a, b, c, tanh_flag, m=m), xx, g, bt)
        dxx, dg, db = fn(dout.reshape(R, C).double())
        _put(dx, dxx)
        _put_split(hi, lo, dxx)
        if dgamma is not None:
            dgamma += dg.float()
        if dbeta is not None:
            dbeta += db.float()

    def bn_tfwd(self, xdot, x, stats, gamma, gdot, bdot, o, R, C, tanh_flag, ws, tsums, odot, hi, lo, beta=None,
                drop=NO_DROP):
        self.n_calls += 1
        xx, xd, g = x.reshape(R, C).double(), xdot.reshape(R, C).double(), gamma.double()
        gd = gdot.double() if gdot is not None else torch.zeros_like(g)
        bd = bdot.double() if bdot is not None else torch.zeros_like(g)
        bt = beta.double() if beta is not None else torch.zeros_like(g)
        assert beta is not None or not tanh_flag
        m = self._dm(drop, R, C)
        _, od = jvp(lambda a, b, c: self._bn(a, b, c, tanh_flag, m=m), (xx, g, bt), (xd, gd, bd))
        mean, rstd = stats.reshape(2, C)[0].double(), stats.reshape(2, C)[1].double()
        xh = (xx - mean) * rstd
        tsums.copy_(torch.cat([xd.mean(0), (xd * xh).mean(0)]).float().reshape(tsums.shape))
        _put(odot, od)
        _put_split(hi, lo, od)

    def bn_tbwd(self, dout, ddout, o, odot, x, xdot, stats, tsums, gamma, gdot, R, C, tanh_flag, ws, ddx, hi, lo,
                ddgamma, ddbeta, beta=None, bdot=None, drop=NO_DROP):
        self.n_calls += 1
        m = self._dm(drop, R, C)
        xx, xd, g = x.reshape(R, C).double(), xdot.reshape(R, C).double(), gamma.double()
        gd = gdot.double() if gdot is not None else torch.zeros_like(g)
        assert beta is not None or not tanh_flag
        b0 = beta.double() if beta is not None else torch.zeros_like(g)
        bd = bdot.double() if bdot is not None else torch.zeros_like(g)

        def bwd(a, gg, bb, cot):
            _, fn = vjp(lambda p, q, r: self._bn(p, q, r, tanh_flag, m=m), a, gg, bb)
            return fn(cot)

        _, (tx, tg, tb) = jvp(bwd, (xx, g, b0, dout.reshape(R, C).double()),
                              (xd, gd, bd, ddout.reshape(R, C).double()))
        _put(ddx, tx)
        _put_split(hi, lo, tx)
        if ddgamma is not None:
            ddgamma += tg.float()
        if ddbeta is not None:
            ddbeta += tb.float()

    # ------------------------------------------------------------------------------------------
    # loss
    # ------------------------------------------------------------------------------------------
    @staticmethod
    def _loss(mel, post, p, e, logd, mel_tgt, p_tgt, e_tgt, dur, mel_lens, src_lens, B, T, Lp, NM):
        mm = (torch.arange(T)[None, :] < mel_lens[:, None])
        sm = (torch.arange(Lp)[None, :] < src_lens[:, None])
        tgt = mel_tgt.reshape(B, T, NM).double()
        n_mel = mm.sum() * NM
        n_src = sm.sum()
        l_mel = ((mel.reshape(B, T, NM) - tgt).abs() * mm[..., None]).sum() / n_mel
        l_post = ((post.reshape(B, T, NM) - tgt).abs() * mm[..., None]).sum() / n_mel
        l_p = (((p.reshape(B, Lp) - p_tgt.reshape(B, Lp).double()) ** 2) * sm).sum() / n_src
        l_e = (((e.reshape(B, Lp) - e_tgt.reshape(B, Lp).double()) ** 2) * sm).sum() / n_src
        l_d = (((logd.reshape(B, Lp) - torch.log(dur.reshape(B, Lp).double() + 1)) ** 2) * sm).sum() / n_src
        return l_mel + l_post + l_d + l_p + l_e, (l_mel, l_post, l_p, l_e, l_d), n_mel, n_src

    def loss_fwd(self, mel, post, mel_tgt, mel_lens, p, p_tgt, e, e_tgt, logd, dur, src_lens, B, T, Lp, NM, ws, out6, counts):
        self.n_calls += 1
        tot, parts, n_mel, n_src = self._loss(mel.double(), post.double(), p.double(), e.double(), logd.double(), mel_tgt,
                                              p_tgt, e_tgt, dur, mel_lens, src_lens, B, T, Lp, NM)
        out6.copy_(torch.stack([tot, *parts]).float())
        counts.copy_(torch.tensor([float(n_mel), float(n_src)]))

    def loss_bwd(self, mel, post, mel_tgt, mel_lens, p, p_tgt, e, e_tgt, logd, dur, src_lens, B, T, Lp, NM, counts, scale,
                 tangent, dmel, dpost, dp, de, dlogd):
        self.n_calls += 1
        n_src = counts[1].double()
        sm = (torch.arange(Lp)[None, :] < src_lens[:, None]).double()
        if tangent:
            dmel.zero_()
            dpost.zero_()
            k = 2.0 * scale / n_src
            _put(dp, k * p.reshape(B, Lp).double() * sm)
            _put(de, k * e.reshape(B, Lp).double() * sm)
            _put(dlogd, k * logd.reshape(B, Lp).double() * sm)
            return
        ins = [t.double().clone().requires_grad_(True) for t in (mel, post, p, e, logd)]
        tot, _, _, _ = self._loss(*ins, mel_tgt, p_tgt, e_tgt, dur, mel_lens, src_lens, B, T, Lp, NM)
        gs = torch.autograd.grad(tot * scale, ins)
        for dst, g in zip((dmel, dpost, dp, de, dlogd), gs):
            _put(dst, g)

    # ------------------------------------------------------------------------------------------
    # free-running synthesis / vocoder-side decode (include/mtts.h: mtts_duration_round ... mtts_istft_finish)
    # ------------------------------------------------------------------------------------------
    def bn_eval(self, x, gamma, beta, running_mean, running_var, R, C, tanh_flag, out, hi, lo, eps=1e-5):
        self.n_calls += 1
        xx = x.reshape(R, C).double()
        y = (xx - running_mean.double()) / torch.sqrt(running_var.double() + eps) * gamma.double() + beta.double()
        o = torch.tanh(y) if tanh_flag else y
        _put(out, o)
        _put_split(hi, lo, o)

    def duration_round(self, logd, d_control, out):
        self.n_calls += 1
        out.copy_(torch.clamp(torch.round(torch.exp(logd) - 1) * d_control, min=0).reshape(out.shape))

    def unary(self, op, x, a, b, out, hi=None, lo=None):
        self.n_calls += 1
        v = x.float()
        v = torch.log(torch.clamp(v, min=a) * b) if op == 0 else (torch.exp(v) * a if op == 1 else v * a)
        _put(out, v)
        _put_split(hi, lo, v)

    def reflect_pad(self, x, B, N, pad, ld, out, hi, lo):
        self.n_calls += 1
        xp = F.pad(x.reshape(B, 1, N).float(), (pad, pad), mode="reflect").reshape(B, N + 2 * pad)
        v = torch.zeros(B, ld)
        n = min(ld, N + 2 * pad)
        v[:, :n] = xp[:, :n]
        _put(out, v)
        _put_split(hi, lo, v)

    def stft_polar(self, ri, R, nb, ld, im_off, ldm, mag, phase, energy, mag_hi=None, mag_lo=None):
        self.n_calls += 1
        r = ri.reshape(R, ld).float()
        re, im = r[:, :nb], r[:, im_off:im_off + nb]
        m = torch.zeros(R, ldm)
        ph = torch.zeros(R, ldm)
        m[:, :nb] = torch.sqrt(re ** 2 + im ** 2)
        ph[:, :nb] = torch.atan2(im, re)
        _put(mag, m)
        _put(phase, ph)
        _put_split(mag_hi, mag_lo, m)
        if energy is not None:
            energy.copy_(torch.norm(m[:, :nb], dim=1).reshape(energy.shape))

    def stft_recombine(self, mag, phase, ri, R, nb, ld, im_off, ldm, hi, lo):
        self.n_calls += 1
        m = mag.reshape(R, ldm)[:, :nb].float()
        if phase is not None:
            ph = phase.reshape(R, ldm)[:, :nb].float()
        else:
            r = ri.reshape(R, ld).float()
            ph = torch.atan2(r[:, im_off:im_off + nb], r[:, :nb])
        v = torch.zeros(R, ld)
        v[:, :nb] = m * torch.cos(ph)
        v[:, im_off:im_off + nb] = m * torch.sin(ph)
        _put_split(hi, lo, v)

    def istft_finish(self, ola, wsum, tiny, scale, B, n, trim, out):
        self.n_calls += 1
        v = ola.reshape(B, n).float().clone()
        nz = wsum > tiny
        v[:, nz] = v[:, nz] / wsum[nz]
        out.copy_((v * scale)[:, trim:n - trim].reshape(out.shape))

    # ------------------------------------------------------------------------------------------
    # flat elementwise
    # ------------------------------------------------------------------------------------------
    def split_(self, src, hi, lo):
        self.n_calls += 1
        _put_split(hi, lo, src)

    def sgd_split(self, theta, g, lr, out, hi, lo):
        self.n_calls += 1
        v = theta - lr * g
        out.copy_(v)
        _put_split(hi, lo, v)

    def axpby(self, a, x, b, y):
        self.n_calls += 1
        y.mul_(b).add_(a * x)

    def sumsq(self, x, out):
        self.n_calls += 1
        out.reshape(-1)[0] = (x.double() ** 2).sum().float()       # [0] = the value ([1..]: the kernels' per-CTA partials)

    def dot(self, x, y, out):
        self.n_calls += 1
        out.reshape(-1)[0] = (x.double() * y.double()).sum().float()

    def adam_clip(self, p, g, m, v, sumsq, gscale, max_norm, hyper, beta1, beta2, eps, hi, lo):
        self.n_calls += 1
        norm = torch.sqrt(sumsq.reshape(-1)[0].double()).item() * gscale
        coef = min(1.0, max_norm / (norm + 1e-6)) if max_norm > 0 else 1.0
        gg = g * (coef * gscale)
        lr, bc1, bc2 = [float(x) for x in hyper[:3]]
        m.mul_(beta1).add_((1 - beta1) * gg)
        v.mul_(beta2).add_((1 - beta2) * gg * gg)
        denom = v.sqrt() / math.sqrt(bc2) + eps
        p.sub_((lr / bc1) * m / denom)
        _put_split(hi, lo, p)
