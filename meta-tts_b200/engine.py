"""Functional FastSpeech2 engine over the libmtts op set: forward, backward, tangent-forward (JVP)
and tangent-backward (JVP of the backward) passes, written without autograd.

This is the host side of the hot path `forward_learner -> FastSpeech2 -> FastSpeech2Loss`
(lightning/systems/base_adaptor.py:41-95, lightning/model/fastspeech2.py:40-112, loss.py:19-92 of
the reference).  The four passes are what the MAML engine (`maml.py`) composes:

    inner step k :  forward(theta_k) ; backward -> g_k ; theta_{k+1} = theta_k - lr*g_k
    query        :  forward(theta_K) ; backward -> lambda_K (adapted params), grad_phi (encoder)
    second order :  for k = K-1..0:  tfwd(theta_k; v = lambda_{k+1}) ; tbwd -> H_k v
                    lambda_k = lambda_{k+1} - lr * (H v)[adapted] ; grad_phi -= lr * (H v)[encoder]

All tensors live in preallocated `Tape`s (stable addresses => the whole step is CUDA-graph
capturable); every op is one libmtts kernel launch through the backend `be` (ops.CudaOps).
Activations are token-major [B, T, C]; anything consumed by a GEMM is kept as bf16 hi(/lo).

Internal parameter layout (`ParamLayout`): one flat fp32 arena + bf16 hi/lo arenas of the same
geometry; Conv1d weights are stored [k, out, in] (K-major GEMM operand per tap), w_qs/w_ks/w_vs are
adjacent so the QKV projection is one N=768 GEMM.  `pack`/`unpack` convert from/to the reference's
state_dict (same keys, Conv1d [out, in, k]).
"""
from __future__ import annotations

import os

import math
import zlib
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import lib as L
from .ops import NO_DROP, Opnd

N_SYMBOLS = 360      # len(text.symbols.symbols), text/symbols.py:21-29
N_MEL = 80
POSTNET_CH = [N_MEL, 512, 512, 512, 512, N_MEL]


def _rup(x: int, m: int) -> int:
    return (x + m - 1) // m * m


# =================================================================================================
# parameters
# =================================================================================================
def param_specs(cfg, n_speaker: int) -> List[Tuple[str, Tuple[int, ...]]]:
    """Trainable parameters (reference state_dict names and shapes), in arena order."""
    tr = cfg["transformer"]
    d = tr["encoder_hidden"]
    di = tr["conv_filter_size"]
    k1, k2 = tr["conv_kernel_size"]
    specs: List[Tuple[str, Tuple[int, ...]]] = []

    def fft(prefix, dm, nh):
        for n in ("w_qs", "w_ks", "w_vs"):
            specs.append((f"{prefix}.slf_attn.{n}.weight", (dm, dm)))
        for n in ("w_qs", "w_ks", "w_vs"):
            specs.append((f"{prefix}.slf_attn.{n}.bias", (dm,)))
        specs.extend([(f"{prefix}.slf_attn.layer_norm.weight", (dm,)), (f"{prefix}.slf_attn.layer_norm.bias", (dm,)),
                      (f"{prefix}.slf_attn.fc.weight", (dm, dm)), (f"{prefix}.slf_attn.fc.bias", (dm,)),
                      (f"{prefix}.pos_ffn.w_1.weight", (di, dm, k1)), (f"{prefix}.pos_ffn.w_1.bias", (di,)),
                      (f"{prefix}.pos_ffn.w_2.weight", (dm, di, k2)), (f"{prefix}.pos_ffn.w_2.bias", (dm,)),
                      (f"{prefix}.pos_ffn.layer_norm.weight", (dm,)), (f"{prefix}.pos_ffn.layer_norm.bias", (dm,))])

    specs.append(("encoder.src_word_emb.weight", (N_SYMBOLS + 1, d)))
    for i in range(tr["encoder_layer"]):
        fft(f"encoder.layer_stack.{i}", d, tr["encoder_head"])
    f = cfg["variance_predictor"]["filter_size"]
    kv = cfg["variance_predictor"]["kernel_size"]
    for p in ("duration_predictor", "pitch_predictor", "energy_predictor"):
        pre = f"variance_adaptor.{p}"
        specs.extend([(f"{pre}.conv_layer.conv1d_1.conv.weight", (f, d, kv)), (f"{pre}.conv_layer.conv1d_1.conv.bias", (f,)),
                      (f"{pre}.conv_layer.layer_norm_1.weight", (f,)), (f"{pre}.conv_layer.layer_norm_1.bias", (f,)),
                      (f"{pre}.conv_layer.conv1d_2.conv.weight", (f, f, kv)), (f"{pre}.conv_layer.conv1d_2.conv.bias", (f,)),
                      (f"{pre}.conv_layer.layer_norm_2.weight", (f,)), (f"{pre}.conv_layer.layer_norm_2.bias", (f,)),
                      (f"{pre}.linear_layer.weight", (1, f)), (f"{pre}.linear_layer.bias", (1,))])
    nb = cfg["variance_embedding"]["n_bins"]
    specs.append(("variance_adaptor.pitch_embedding.weight", (nb, d)))
    specs.append(("variance_adaptor.energy_embedding.weight", (nb, d)))
    dd = tr["decoder_hidden"]
    for i in range(tr["decoder_layer"]):
        fft(f"decoder.layer_stack.{i}", dd, tr["decoder_head"])
    specs.extend([("mel_linear.weight", (N_MEL, dd)), ("mel_linear.bias", (N_MEL,))])
    for i in range(5):
        specs.extend([(f"postnet.convolutions.{i}.0.conv.weight", (POSTNET_CH[i + 1], POSTNET_CH[i], 5)),
                      (f"postnet.convolutions.{i}.0.conv.bias", (POSTNET_CH[i + 1],)),
                      (f"postnet.convolutions.{i}.1.weight", (POSTNET_CH[i + 1],)),
                      (f"postnet.convolutions.{i}.1.bias", (POSTNET_CH[i + 1],))])
    specs.append(("speaker_emb.model.weight", (n_speaker, d)))
    return specs


def const_names(cfg) -> List[str]:
    names = ["encoder.position_enc", "decoder.position_enc", "variance_adaptor.pitch_bins", "variance_adaptor.energy_bins"]
    for i in range(5):
        names += [f"postnet.convolutions.{i}.1.running_mean", f"postnet.convolutions.{i}.1.running_var",
                  f"postnet.convolutions.{i}.1.num_batches_tracked"]
    return names


@dataclass
class PEntry:
    name: str
    sd_shape: Tuple[int, ...]       # reference state_dict shape
    shape: Tuple[int, ...]          # internal shape ([k,out,in] for Conv1d weights)
    offset: int
    numel: int
    adapted: bool


class ParamLayout:
    """Flat arena layout: [ non-adapted trainable | adapted trainable ], 64-element aligned entries."""

    def __init__(self, cfg, n_speaker: int, adapt_modules: Sequence[str]):
        self.cfg = cfg
        self.n_speaker = n_speaker
        self.adapt_modules = tuple(adapt_modules)
        specs = param_specs(cfg, n_speaker)
        self.entries: Dict[str, PEntry] = {}
        off = 0
        self.order: List[str] = []
        for want_adapted in (False, True):
            if want_adapted:
                self.adapt_begin = off
            for name, shp in specs:
                ad = name.split(".")[0] in self.adapt_modules
                if ad != want_adapted:
                    continue
                ishape = (shp[2], shp[0], shp[1]) if len(shp) == 3 else shp
                n = int(math.prod(shp))
                self.entries[name] = PEntry(name, tuple(shp), tuple(ishape), off, n, ad)
                self.order.append(name)
                off = _rup(off + n, 64)
        self.total = off
        self.n_adapt = self.total - self.adapt_begin
        self.n_trainable = sum(e.numel for e in self.entries.values())
        self.n_adapted_params = sum(e.numel for e in self.entries.values() if e.adapted)

    def is_adapted_module(self, prefix: str) -> bool:
        return prefix.split(".")[0] in self.adapt_modules

    # ---- conversion from / to the reference state_dict (host-side plumbing, not on the hot path) ----
    def pack(self, state_dict, flat: torch.Tensor) -> None:
        flat.zero_()
        for name, e in self.entries.items():
            t = state_dict[name].detach().to(torch.float32)
            assert tuple(t.shape) == e.sd_shape, f"{name}: {tuple(t.shape)} != {e.sd_shape}"
            if len(e.sd_shape) == 3:
                t = t.permute(2, 0, 1).contiguous()
            flat[e.offset:e.offset + e.numel].copy_(t.reshape(-1))

    def unpack(self, flat: torch.Tensor) -> Dict[str, torch.Tensor]:
        out = {}
        for name, e in self.entries.items():
            t = flat[e.offset:e.offset + e.numel].reshape(e.shape)
            if len(e.sd_shape) == 3:
                t = t.permute(1, 2, 0)
            out[name] = t.contiguous().clone()
        return out


class Wt:
    """One parameter's views: fp32 master and bf16 hi/lo (GEMM operand layout)."""
    __slots__ = ("f32", "hi", "lo", "shape")

    def __init__(self, f32, hi, lo, shape):
        self.f32, self.hi, self.lo, self.shape = f32, hi, lo, shape


class ParamSet:
    """Name -> views over (flat f32, hi, lo) arenas; the adapted region may come from separate
    (fast-weight) arenas of length layout.n_adapt.  hi/lo may be None (gradient arenas)."""

    def __init__(self, layout: ParamLayout, base_f32, base_hi=None, base_lo=None, fast_f32=None, fast_hi=None,
                 fast_lo=None, only_adapted: bool = False):
        self.layout = layout
        self.base = (base_f32, base_hi, base_lo)
        self.fast = (fast_f32, fast_hi, fast_lo)
        self.only_adapted = only_adapted
        self._cache: Dict[str, Wt] = {}

    def has(self, name: str) -> bool:
        e = self.layout.entries[name]
        return e.adapted or not self.only_adapted

    def get(self, name: str) -> Optional[Wt]:
        if name in self._cache:
            return self._cache[name]
        e = self.layout.entries[name]
        if self.only_adapted and not e.adapted:
            return None
        if e.adapted and self.fast[0] is not None:
            arenas, off = self.fast, e.offset - self.layout.adapt_begin
        else:
            arenas, off = self.base, e.offset
        v = [a[off:off + e.numel].view(e.shape) if a is not None else None for a in arenas]
        w = Wt(v[0], v[1], v[2], e.shape)
        self._cache[name] = w
        return w

    def span(self, first: str, last: str, shape) -> Optional[Wt]:
        """A view covering adjacent entries first..last (e.g. w_qs|w_ks|w_vs -> [768, 256])."""
        key = f"{first}|{last}"
        if key in self._cache:
            return self._cache[key]
        e0, e1 = self.layout.entries[first], self.layout.entries[last]
        if self.only_adapted and not e0.adapted:
            return None
        n = e1.offset + e1.numel - e0.offset
        assert n == math.prod(shape), "span is not contiguous"
        if e0.adapted and self.fast[0] is not None:
            arenas, off = self.fast, e0.offset - self.layout.adapt_begin
        else:
            arenas, off = self.base, e0.offset
        v = [a[off:off + n].view(shape) if a is not None else None for a in arenas]
        w = Wt(v[0], v[1], v[2], tuple(shape))
        self._cache[key] = w
        return w


# =================================================================================================
# buffers
# =================================================================================================
class Act:
    """An activation: fp32 and/or bf16 hi(/lo) tensors of logical shape [B, T, C]."""
    __slots__ = ("f32", "hi", "lo", "B", "T", "C")

    def __init__(self, f32, hi, lo, B, T, C):
        self.f32, self.hi, self.lo, self.B, self.T, self.C = f32, hi, lo, B, T, C


class Tape:
    """Named, lazily allocated, persistent device buffers (stable addresses across replays)."""

    def __init__(self, be, split: int):
        self.be = be
        self.split = split
        self.t: Dict[str, torch.Tensor] = {}

    def buf(self, name, shape, dtype=torch.float32, zero=False):
        t = self.t.get(name)
        if t is None:
            t = self.be.zeros(tuple(shape), dtype) if zero else self.be.empty(tuple(shape), dtype)
            self.t[name] = t
        assert tuple(t.shape) == tuple(shape), f"tape buffer {name}: {tuple(t.shape)} vs {tuple(shape)}"
        return t

    def f32(self, name, shape):
        return self.buf(name + ":f", shape)

    def bf(self, name, shape):
        hi = self.buf(name + ":h", shape, torch.bfloat16)
        lo = self.buf(name + ":l", shape, torch.bfloat16) if self.split == 3 else None
        return hi, lo

    def act(self, name, B, T, C, f32=True, bf=True) -> Act:
        f = self.f32(name, (B, T, C)) if f32 else None
        h, l = self.bf(name, (B, T, C)) if bf else (None, None)
        return Act(f, h, l, B, T, C)

    def scratch(self, name, shape, dtype=torch.float32):
        """Shape-keyed scratch buffer (support / query / encoder / decoder shapes differ)."""
        return self.buf(f"{name}[{'x'.join(str(int(v)) for v in shape)}]:s", shape, dtype)

    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in self.t.values())


class TapeView:
    """Utterances [b0, b1) of a batch-major Tape: every buffer whose leading extent is B or B*T rows is allocated for the FULL
    batch in the parent tape and handed out as the contiguous slice of this group, so a pass may fill (or read) the tape
    group by group on concurrent chains while every other pass keeps seeing ordinary full-batch buffers."""

    def __init__(self, parent: "Tape", b0: int, b1: int, B: int, rows_per_item: int):
        self.p, self.b0, self.b1, self.B = parent, b0, b1, B
        self.split = parent.split
        self.row0 = b0 * rows_per_item           # first token row of the group (dropout element indices are global)

    @property
    def drop_pass(self):
        return getattr(self.p, "drop_pass", None)

    def buf(self, name, shape, dtype=torch.float32, zero=False):
        nb = self.b1 - self.b0
        assert shape[0] % nb == 0, f"tape view: {name} {tuple(shape)} is not batch-major"
        k = shape[0] // nb
        full = self.p.buf(name, (k * self.B,) + tuple(shape[1:]), dtype, zero)
        return full[self.b0 * k:self.b1 * k]

    def f32(self, name, shape):
        return self.buf(name + ":f", shape)

    def bf(self, name, shape):
        hi = self.buf(name + ":h", shape, torch.bfloat16)
        lo = self.buf(name + ":l", shape, torch.bfloat16) if self.split == 3 else None
        return hi, lo

    def act(self, name, B, T, C, f32=True, bf=True) -> Act:
        f = self.f32(name, (B, T, C)) if f32 else None
        h, l = self.bf(name, (B, T, C)) if bf else (None, None)
        return Act(f, h, l, B, T, C)


def _act_rows(a: Optional[Act], b0: int, b1: int) -> Optional[Act]:
    if a is None:
        return None
    sl = lambda t: None if t is None else t[b0:b1]  # noqa: E731
    return Act(sl(a.f32), sl(a.hi), sl(a.lo), b1 - b0, a.T, a.C)


# EXPERIMENT (off): forward-type passes (forward, tangent forward) of the FFT stacks as TWO concurrent chains over half of the
# utterances each.  Idea: per-kernel latency dominates the under-filled GEMMs of a 4-utterance batch, so two half-size chains
# might finish in about the time of one.  Measured on B200 (configs[1]): parity green (24 GPU tests with MTTS_SPLIT_FWD=1) but
# 12.18 ms/step vs 11.80 (+234 launches): the step does not get faster by running more kernels side by side — the extra launches
# cost more than the concurrency returns, the same outcome as the narrow split-K tiles (SMALL_SPLITK_MODEL).  Kept because the
# mechanism (TapeView slices, per-chain scratch, seed-folded dropout offsets) is what a fused per-block kernel will need.
SPLIT_FWD = os.environ.get("MTTS_SPLIT_FWD", "0") == "1"
SPLIT_FWD_MAX_ROWS = 8192
_HASH_G_INV = pow(0x9E3779B9, -1, 1 << 32)       # the dropout hash adds element_index to seed * 0x9E3779B9 (mod 2^32)


@dataclass
class Batch:
    """Device-resident teacher-forced batch (the reference 12-tuple's tensor fields, collate.py:47-60)."""
    spk_ids: torch.Tensor        # i64 [n_spk_ids]  ids fed to speaker_emb (support ids for the query pass)
    average_spk: bool
    texts: torch.Tensor          # i64 [B, L]
    src_lens: torch.Tensor       # i64 [B]
    mels: torch.Tensor           # f32 [B, T, 80]
    mel_lens: torch.Tensor       # i64 [B]
    pitches: torch.Tensor        # f32 [B, L]
    energies: torch.Tensor       # f32 [B, L]
    durations: torch.Tensor      # i64 [B, L]
    B: int
    L: int
    T: int


# =================================================================================================
# GEMM builders
# =================================================================================================
@dataclass
class BMat:
    """Batched matrix X[b,h][r,c] = buf[off + b*sb + h*sh + r*sr + c]."""
    hi: Optional[torch.Tensor]
    lo: Optional[torch.Tensor]
    off: int
    sb: int
    sh: int
    sr: int
    rows: int
    cols: int
    f32: Optional[torch.Tensor] = None


# Narrow tiles + deeper split-K for SMALL accumulate GEMMs: measured on B200 12.00 ms/step vs 11.87 with the widest-tile rule
# (the extra CTAs of the off-critical-path weight gradients take SMs from the main chain), so it stays off.
SMALL_SPLITK_MODEL = os.environ.get("MTTS_SMALL_SPLITK", "0") == "1"
# Fused attention (csrc/mtts_attn.cu) for the forward / backward passes: scores -> softmax -> P V (and the recomputing backward)
# as ONE kernel per pass, S / dP never written.  MTTS_FUSED_ATTN=0 restores the three-launch chain (A/B measurements).
FUSED_ATTN = os.environ.get("MTTS_FUSED_ATTN", "1") == "1"
# dropout -> + residual -> LayerNorm -> pad-row zeroing as the EPILOGUE of the out-projection / conv k=1 GEMM that feeds it (mtts_gemm_ln:
# the 256-wide row is spread over a 4-CTA cluster that exchanges row statistics through distributed shared memory) instead of a
# LayerNorm launch of its own.  MTTS_FUSED_LN=0 restores GEMM + mtts_ln_fwd (A/B measurements).
FUSED_LN = os.environ.get("MTTS_FUSED_LN", "1") == "1"
USE_PAIR = True        # 2-CTA (cta_group::2) tiles; set False to fall back to the 1-CTA kernel everywhere


def _pick_cfg(mt: int, nz: int, N: int, k_iters: int, can_splitk: bool):
    """(block_n, pair, ksplit) for an output of nz x mt row tiles (128 rows) by N columns.
    Bigger / paired tiles cut L2->SMEM operand bytes per flop (the 1-CTA 128x128 tile is L2-throughput bound
    in bf16x3, profiles/r01_ncu_summary.md); small grids are filled by split-K when the epilogue is a pure
    fp32 accumulation, otherwise by narrower tiles."""
    opts = []
    if USE_PAIR and N >= 192:
        opts.append((256, True))
    if USE_PAIR and N >= 96:
        opts.append((128, True))
    if N >= 192:
        opts.append((256, False))
    if N > 64:
        opts.append((128, False))
    opts.append((64, False))

    def ctas(o):
        bn, pair = o
        rows = 2 * ((mt + 1) // 2) if pair else mt
        return nz * rows * ((N + bn - 1) // bn)

    # no split-K: in-graph launch-time model fitted to tools/gemm_floor.py on B200 (profiles/r01_gemm_floor_in_graph.txt):
    # a wave costs  fixed(tile) + k_iters * per_iter(tile)  and a launch runs ceil(ctas / 148) waves.  (The former
    # ">= 96 CTAs" rule picked 128-wide pairs for the QKV projection: 23.9 us where the 256-wide pair takes 16.2.)
    fixed = {(64, False): 5.9, (128, False): 7.2, (128, True): 7.8, (256, False): 11.0, (256, True): 11.0}
    # us per k-block: narrow tiles and the 128-wide pair run two k-blocks per stage fill (KD = 2)
    per_it = {(64, False): 0.59, (128, False): 0.80, (128, True): 0.55, (256, False): 1.5, (256, True): 0.87}
    if can_splitk:
        o = opts[0]                      # widest tile; fill the machine along the contraction instead
        ks = max(1, min(148 // max(ctas(o), 1), max(1, k_iters // 6), 16))
        if SMALL_SPLITK_MODEL and ctas(o) * ks < 48:
            # a small accumulate GEMM (weight gradients of the 256-wide layers on 512 ... 3456 rows): the widest tile leaves
            # it on a handful of CTAs that each pay the 11 us fixed cost of a 256-wide tile; narrower tiles split further
            # along the contraction reach more SMs with a cheaper prologue / epilogue
            def cost_k(o2):
                c = max(ctas(o2), 1)
                k2 = max(1, min(148 // c, max(1, k_iters // 2), 16))
                return fixed[o2] + -(-k_iters // k2) * per_it[o2] + 0.15 * k2, k2     # + red traffic of k2 partial tiles

            o = min(opts, key=lambda o2: cost_k(o2)[0])
            ks = cost_k(o)[1]
        return o[0], o[1], ks
    def cost(o):
        waves = -(-ctas(o) // 148)
        return (fixed[o] + k_iters * per_it[o]) * (1.0 + 0.7 * (waves - 1))   # later waves overlap the previous one's tail

    o = min(opts, key=cost)
    return o[0], o[1], 1


# Per-class precision policy.  Classes: p.* = forward / backward passes, t.* = Hessian-vector (tangent) passes;
# fwd = Linear / Conv forward products, dgrad = data gradients, wgrad = weight gradients, bmm = attention products.  Measured on
# BASELINE configs[1] against the fp32 oracle (tools/precision_budget.py, profiles/r02_precision_budget.md): running the four
# classes below single-pass bf16 (variance predictors exempt) leaves the outputs untouched (7e-5) and moves the outer gradient by
# 3.1e-4 of its norm (2.2e-4 with everything bf16x3), while p.fwd / p.dgrad / p.bmm / t.dgrad each cost > 2e-3 and stay bf16x3.
DEFAULT_SPLIT_POLICY = {"t.fwd": 1, "t.bmm": 1, "t.wgrad": 1, "p.wgrad": 1}


def split_policy_from_env() -> Dict[str, int]:
    """MTTS_SPLIT_POLICY: unset = DEFAULT_SPLIT_POLICY; "strict" (or empty) = every product bf16x3; else "class=1,class=3,..."."""
    v = os.environ.get("MTTS_SPLIT_POLICY")
    if v is None:
        return dict(DEFAULT_SPLIT_POLICY)
    if v.strip() in ("", "strict"):
        return {}
    return dict((k.strip(), int(x)) for k, x in (kv.split("=") for kv in v.split(",") if kv.strip()))


def _strict_precision(fn):
    """Engine methods whose products ignore the per-class policy.  The variance predictors: their outputs feed MSE losses (non-zero
    curvature, so tangent-forward errors reach the outer gradient directly), their GEMMs see only B*L = 512 rows (little averaging
    of rounding noise in the weight gradients) and cost nothing (0.4 GFLOP each) — measured with dropout on, single-pass products
    there moved the outer gradient by 4.5e-3 of its norm (2.6e-3 from one predictor's conv weight)."""
    def wrapped(self, *a, **kw):
        self.g.strict_depth += 1
        try:
            return fn(self, *a, **kw)
        finally:
            self.g.strict_depth -= 1
    wrapped.__name__, wrapped.__doc__ = fn.__name__, fn.__doc__
    return wrapped


class Gemm:
    """Descriptor construction for the contractions of the model (all through be.gemm).
    `x2/w2`-style arguments add a SECOND product term in the same launch (tangent passes)."""

    def __init__(self, be):
        self.be = be
        self.split_override = None      # engine.hvp() sets 1: the lr-scaled curvature terms run single-pass bf16
        self.tangent = False            # inside a Hessian-vector pass (engine.hvp)
        self.policy = split_policy_from_env()
        self.strict_depth = 0           # > 0: inside a block whose products all stay at the engine's split (_strict_precision)

    def split_of(self, kind: str):
        """Operand split (1 = single-pass bf16, 3 = bf16x3) of GEMM class `kind` in the current pass type, or None = backend default."""
        s = self.policy.get(("t." if self.tangent else "p.") + kind) if (self.be.split == 3 and self.strict_depth == 0) else None
        return s if s is not None else self.split_override

    def _gemm(self, kind, *a, **kw):
        s = self.split_of(kind)
        if s is not None:
            kw["split"] = s
        self.be.gemm(*a, **kw)

    # y[b,t,:] = sum_j x[b,t+j-p,:] W_j^T (+bias) [+ sum_j x2[b,t+j-p,:] W2_j^T] ; W: [k, N, Cin]
    def conv_fwd(self, x: Act, w: Wt, bias, out_f32, out_hi, out_lo, relu=False, gate=None, add_c=False,
                 x2: Optional[Act] = None, w2: Optional[Wt] = None, ln: Optional[dict] = None):
        """ln (k = 1, N = 256 only): the outputs receive LayerNorm(dropout(y) + res) with pad rows zeroed — see ops.gemm."""
        k, N, Cin = w.shape if len(w.shape) == 3 else (1,) + tuple(w.shape)
        assert Cin == x.C and (x2 is None) == (w2 is None)
        if ln is not None:
            assert k == 1 and N == 256 and x2 is None and not (relu or add_c) and gate is None
            a = Opnd(x.hi, x.lo, L.MAJOR_K, (Cin, x.B * x.T), (1, Cin))
            wop = Opnd(w.hi, w.lo, L.MAJOR_K, (Cin, N, k), (1, Cin, N * Cin), src2=L.SRC_TAP)
            self._gemm("fwd", a, wop, x.B * x.T, N, Cin, c_f32=out_f32, c_hi=out_hi, c_lo=out_lo, ldc=N, bias=bias, block_n=64, ln=ln)
            return
        p = (k - 1) // 2
        flags = (L.EPI_RELU if relu else 0) | (L.EPI_GATE if gate is not None else 0) | (L.EPI_ADD_C if add_c else 0)
        h2 = (lambda t: (t.hi, t.lo)) if x2 is not None else (lambda t: (None, None))
        wop = Opnd(w.hi, w.lo, L.MAJOR_K, (Cin, N, k), (1, Cin, N * Cin), src2=L.SRC_TAP)
        if w2 is not None:
            wop.hi2, wop.lo2 = w2.hi, w2.lo
        if k == 1:
            a = Opnd(x.hi, x.lo, L.MAJOR_K, (Cin, x.B * x.T), (1, Cin))
            if x2 is not None:
                a.hi2, a.lo2 = x2.hi, x2.lo
            bn, pair, _ = _pick_cfg((x.B * x.T + 127) // 128, 1, N, 0, False)
            self._gemm("fwd", a, wop, x.B * x.T, N, Cin, c_f32=out_f32, c_hi=out_hi, c_lo=out_lo, ldc=N, bias=bias,
                         gate=gate, flags=flags, block_n=bn, pair=pair)
        else:
            a = Opnd(x.hi, x.lo, L.MAJOR_K, (Cin, x.T, x.B), (1, Cin, x.T * Cin), src2=L.SRC_Z0,
                     shift_src=L.SRC_TAP, shift_base=-p, shift_step=1)
            if x2 is not None:
                a.hi2, a.lo2 = x2.hi, x2.lo
            bn, pair, _ = _pick_cfg((x.T + 127) // 128, x.B, N, 0, False)
            self._gemm("fwd", a, wop, x.T, N, Cin, c_f32=out_f32, c_hi=out_hi, c_lo=out_lo, ldc=N, c_sz0=x.T * N, bias=bias,
                         gate=gate, flags=flags, ntaps=k, nz0=x.B, block_n=bn, pair=pair)

    # dx[b,t,:] = sum_j dy[b,t-j+p,:] W_j  [+ sum_j dy2[b,t-j+p,:] W2_j]
    def conv_dgrad(self, dy: Act, w: Wt, out_f32, out_hi, out_lo, gate=None, add_c=False,
                   dy2: Optional[Act] = None, w2: Optional[Wt] = None):
        k, N, Cin = w.shape if len(w.shape) == 3 else (1,) + tuple(w.shape)
        assert N == dy.C and (dy2 is None) == (w2 is None)
        p = (k - 1) // 2
        wop = Opnd(w.hi, w.lo, L.MAJOR_MN, (Cin, N, k), (1, Cin, N * Cin), src2=L.SRC_TAP)
        if w2 is not None:
            wop.hi2, wop.lo2 = w2.hi, w2.lo
        nt = 2 if dy2 is not None else 1
        # "+= into an fp32 buffer" epilogues may be split along the contraction (red.global.add instead of +C)
        can_split = add_c and out_hi is None and gate is None
        mt = (dy.B * dy.T + 127) // 128 if k == 1 else (dy.T + 127) // 128
        bn, pair, ks = _pick_cfg(mt, 1 if k == 1 else dy.B, Cin, nt * k * ((N + 63) // 64), can_split)
        if ks > 1:
            flags = L.EPI_ACCUM
        else:
            flags = (L.EPI_GATE if gate is not None else 0) | (L.EPI_ADD_C if add_c else 0)
        if k == 1:
            a = Opnd(dy.hi, dy.lo, L.MAJOR_K, (N, dy.B * dy.T), (1, N))
            if dy2 is not None:
                a.hi2, a.lo2 = dy2.hi, dy2.lo
            self._gemm("dgrad", a, wop, dy.B * dy.T, Cin, N, c_f32=out_f32, c_hi=out_hi, c_lo=out_lo, ldc=Cin, gate=gate,
                         flags=flags, block_n=bn, pair=pair, ksplit=ks)
        else:
            a = Opnd(dy.hi, dy.lo, L.MAJOR_K, (N, dy.T, dy.B), (1, N, dy.T * N), src2=L.SRC_Z0,
                     shift_src=L.SRC_TAP, shift_base=p, shift_step=-1)
            if dy2 is not None:
                a.hi2, a.lo2 = dy2.hi, dy2.lo
            self._gemm("dgrad", a, wop, dy.T, Cin, N, c_f32=out_f32, c_hi=out_hi, c_lo=out_lo, ldc=Cin, c_sz0=dy.T * Cin,
                         gate=gate, flags=flags, ntaps=k, nz0=dy.B, block_n=bn, pair=pair, ksplit=ks)

    # dW_j[n,c] += sum_{b,t} dy[b,t,n] x[b,t+j-p,c]  [+ dy2 (x) x2]
    def conv_wgrad(self, dy: Act, x: Act, dw_f32: torch.Tensor, scale: float = 1.0,
                   dy2: Optional[Act] = None, x2: Optional[Act] = None):
        shp = tuple(dw_f32.shape)
        k, N, Cin = shp if len(shp) == 3 else (1,) + shp
        assert N == dy.C and Cin == x.C and dy.B == x.B and dy.T == x.T and (dy2 is None) == (x2 is None)
        p = (k - 1) // 2
        nt = 2 if dy2 is not None else 1
        iters = nt * ((dy.B * dy.T + 63) // 64) if k == 1 else nt * dy.B * ((dy.T + 63) // 64)
        bn, pair, ks = _pick_cfg((N + 127) // 128, k, Cin, iters, True)
        if k == 1:
            R = dy.B * dy.T
            a = Opnd(dy.hi, dy.lo, L.MAJOR_MN, (N, R), (1, N))
            b = Opnd(x.hi, x.lo, L.MAJOR_MN, (Cin, R), (1, Cin))
            if dy2 is not None:
                a.hi2, a.lo2, b.hi2, b.lo2 = dy2.hi, dy2.lo, x2.hi, x2.lo
            self._gemm("wgrad", a, b, N, Cin, R, c_f32=dw_f32, ldc=Cin, flags=L.EPI_ACCUM, alpha=scale, ksplit=ks, block_n=bn,
                         pair=pair)
        else:
            a = Opnd(dy.hi, dy.lo, L.MAJOR_MN, (N, dy.T, dy.B), (1, N, dy.T * N), src2=L.SRC_KB)
            b = Opnd(x.hi, x.lo, L.MAJOR_MN, (Cin, x.T, x.B), (1, Cin, x.T * Cin), src2=L.SRC_KB,
                     shift_src=L.SRC_Z0, shift_base=-p, shift_step=1)
            if dy2 is not None:
                a.hi2, a.lo2, b.hi2, b.lo2 = dy2.hi, dy2.lo, x2.hi, x2.lo
            self._gemm("wgrad", a, b, N, Cin, dy.T, c_f32=dw_f32, ldc=Cin, c_sz0=N * Cin, flags=L.EPI_ACCUM, alpha=scale,
                         nkb=dy.B, nz0=k, ksplit=ks, block_n=bn, pair=pair)

    # C[b,h] = alpha * ( op(A) op(B)^T [+ op(A2) op(B2)^T] )   (op = identity or transpose, see BMat)
    def bmm(self, A: BMat, a_t: bool, Bm: BMat, b_t: bool, Cm: BMat, nb: int, nh: int, alpha=1.0, add_c=False,
            A2: Optional[BMat] = None, B2: Optional[BMat] = None):
        M = A.cols if a_t else A.rows
        K = A.rows if a_t else A.cols
        N = Bm.cols if b_t else Bm.rows
        assert K == (Bm.rows if b_t else Bm.cols) and (A2 is None) == (B2 is None)
        a = Opnd(A.hi, A.lo, L.MAJOR_MN if a_t else L.MAJOR_K, (A.cols, A.rows, nh, nb), (1, A.sr, A.sh, A.sb),
                 src2=L.SRC_Z0, src3=L.SRC_Z1, offset=A.off)
        b = Opnd(Bm.hi, Bm.lo, L.MAJOR_MN if b_t else L.MAJOR_K, (Bm.cols, Bm.rows, nh, nb), (1, Bm.sr, Bm.sh, Bm.sb),
                 src2=L.SRC_Z0, src3=L.SRC_Z1, offset=Bm.off)
        if A2 is not None:
            assert (A2.off, A2.sr, A2.sh, A2.sb) == (A.off, A.sr, A.sh, A.sb) and (B2.off, B2.sr, B2.sh, B2.sb) == (Bm.off, Bm.sr, Bm.sh, Bm.sb)
            a.hi2, a.lo2, b.hi2, b.lo2 = A2.hi, A2.lo, B2.hi, B2.lo
        bn, pair, _ = _pick_cfg((M + 127) // 128, nb * nh, N, 0, False)
        self._gemm("bmm", a, b, M, N, K, c_f32=Cm.f32, c_hi=Cm.hi, c_lo=Cm.lo, ldc=Cm.sr, c_off=Cm.off, c_sz0=Cm.sh,
                     c_sz1=Cm.sb, alpha=alpha, flags=(L.EPI_ADD_C if add_c else 0), nz0=nh, nz1=nb,
                     block_n=bn, pair=pair)


# =================================================================================================
# the model
# =================================================================================================
class FS2Engine:
    """Four-pass functional FastSpeech2 (+ loss) over a backend."""

    def __init__(self, be, cfg, layout: ParamLayout, consts: Dict[str, torch.Tensor]):
        self.be = be
        self.cfg = cfg
        self.layout = layout
        self.consts = consts
        self.g = Gemm(be)
        self.split = be.split
        tr = cfg["transformer"]
        self.d = tr["encoder_hidden"]
        self.n_enc, self.n_dec = tr["encoder_layer"], tr["decoder_layer"]
        self.h_enc, self.h_dec = tr["encoder_head"], tr["decoder_head"]
        self.d_inner = tr["conv_filter_size"]
        self.nbins = cfg["variance_embedding"]["n_bins"]
        self.scr = Tape(be, self.split)        # shared scratch (never read across passes)
        self._scr_main = self.scr
        self._scr_branch = Tape(be, self.split)   # scratch of work issued on an auxiliary stream (may run concurrently)
        # Precision of the Hessian-vector pass.  Its result enters the outer gradient multiplied by the inner lr
        # (1e-3); running it single-pass bf16 (hvp_split = 1, operand hi halves only) moves individual gradient
        # tensors by up to ~3e-3 relative (CPU emulation, tests/test_engine_cpu.py), so the default keeps the
        # engine's precision and the faster setting is opt-in.
        self.hvp_split = int(os.environ.get("MTTS_HVP_SPLIT", self.split))
        assert self.d % 128 == 0

    def new_tape(self) -> Tape:
        return Tape(self.be, self.split)

    # ---------------------------------------------------------------------------------------------
    # FFT block
    # ---------------------------------------------------------------------------------------------
    def _attn_geom(self, B, T, H):
        d = self.d
        dk = d // H
        Tp = _rup(T, 8)
        row = 3 * d
        q = lambda buf_hi, buf_lo, which, f32=None: BMat(buf_hi, buf_lo, which * d, T * row, dk, row, T, dk, f32)  # noqa: E731
        pm = lambda hi, lo, f32=None: BMat(hi, lo, 0, H * T * Tp, T * Tp, Tp, T, T, f32)  # noqa: E731
        om = lambda hi, lo, f32=None: BMat(hi, lo, 0, T * d, dk, d, T, dk, f32)  # noqa: E731
        return dk, Tp, q, pm, om

    def _fused_attn(self, dk: int) -> bool:
        return FUSED_ATTN and dk == 128 and hasattr(self.be, "attn_fwd")

    def _attn_emit(self, tp, pf: str) -> bool:
        """Write P / dP / dS out of the fused kernels?  Only tapes that a Hessian-vector pass re-reads need them (second-order
        support passes: `tp.attn_emit`, unknown = keep), and only layers that have a tangent pass (adapted modules)."""
        return bool(getattr(tp, "attn_emit", True)) and self.layout.is_adapted_module(pf)

    @staticmethod
    def _attn_vec(tp, name: str, B: int, H: int, T: int):
        """Per-row attention statistics [B, H, Tl] (log-sum-exp / softmax-backward row term), Tl = T rounded up to 128."""
        return tp.buf(name + ":f", (B, H, _rup(T, 128)), zero=True)

    # ---- dropout sites (include/mtts.h): launch scalar = crc32(site) ^ (pass_index * 2654435761); the device adds the
    #      per-step salt.  The pass index lives on the primal tape, so every pass over it draws the same mask. ----
    def _site(self, tp: Tape, site: str, p: float, C: int = 0):
        idx = getattr(tp, "drop_pass", None)
        if idx is None or p <= 0.0:
            return NO_DROP
        scalar = (zlib.crc32(site.encode()) ^ ((int(idx) * 2654435761) & 0xFFFFFFFF)) & 0xFFFFFFFF
        row0 = getattr(tp, "row0", 0)
        if row0:
            # a group of utterances launched on its own: the kernel numbers its elements from 0, the mask is defined on the index
            # in the full batch.  keep(i + off) hashes (i + off + seed_eff * G) mod 2^32 = (i + (seed_eff + off * G^-1) * G), so the
            # element offset folds into the launch-time seed.
            scalar = (scalar + row0 * C * _HASH_G_INV) & 0xFFFFFFFF
        return (int(p * (1 << 24)), scalar, 1.0 / (1.0 - p))

    def _fft_drop(self, tp: Tape, pf: str):
        p = self.cfg["transformer"]["encoder_dropout" if pf.startswith("encoder") else "decoder_dropout"]
        return self._site(tp, f"{pf}.slf_attn", p, self.d), self._site(tp, f"{pf}.pos_ffn", p, self.d)

    def _vp_drop(self, tp: Tape, pf: str):
        p = self.cfg["variance_predictor"]["dropout"]
        return self._site(tp, f"{pf}.1", p), self._site(tp, f"{pf}.2", p)

    def _fft_names(self, pf):
        a, f = f"{pf}.slf_attn", f"{pf}.pos_ffn"
        return a, f

    def _qkv(self, P: ParamSet, pf):
        a = f"{pf}.slf_attn"
        d = self.d
        w = P.span(f"{a}.w_qs.weight", f"{a}.w_vs.weight", (3 * d, d))
        b = P.span(f"{a}.w_qs.bias", f"{a}.w_vs.bias", (3 * d,))
        return w, b

    def fft_fwd(self, P: ParamSet, pf: str, tp: Tape, x: Act, lens, H: int) -> Act:
        be, g, scr = self.be, self.g, self.scr
        B, T, d = x.B, x.T, self.d
        R = B * T
        a_, f_ = self._fft_names(pf)
        dk, Tp, qm, pm, om = self._attn_geom(B, T, H)
        wqkv, bqkv = self._qkv(P, pf)
        qkv_h, qkv_l = tp.bf(f"{pf}.qkv", (R, 3 * d))
        g.conv_fwd(x, wqkv, bqkv.f32, None, qkv_h, qkv_l)
        o = tp.act(f"{pf}.o", B, T, d, f32=False)
        if self._fused_attn(dk):
            # P is only written out for tapes a Hessian-vector pass will re-read (tp.attn_emit; unknown = keep it)
            p_h, p_l = tp.bf(f"{pf}.P", (B, H, T, Tp)) if self._attn_emit(tp, pf) else (None, None)
            be.attn_fwd(qkv_h, qkv_l, lens, B, H, T, dk, o.hi, o.lo, self._attn_vec(tp, f"{pf}.lse", B, H, T), p_h, p_l, Tp,
                        split=g.split_of("bmm"))
        else:
            S = scr.scratch("S", (B, H, T, Tp))
            g.bmm(qm(qkv_h, qkv_l, 0), False, qm(qkv_h, qkv_l, 1), False, pm(None, None, S), B, H, alpha=1.0 / math.sqrt(dk))
            p_h, p_l = tp.bf(f"{pf}.P", (B, H, T, Tp))
            be.softmax(0, S, None, None, None, None, None, lens, B * H, H, T, T, Tp, p_h, p_l)
            g.bmm(pm(p_h, p_l), False, qm(qkv_h, qkv_l, 2), True, om(o.hi, o.lo), B, H)
        y1 = tp.act(f"{pf}.y1", B, T, d)
        fused_ln = FUSED_LN and d == 256
        if fused_ln:     # out-projection with dropout + residual + LayerNorm + pad-row zeroing as its epilogue (SubLayers.py:54-55)
            g.conv_fwd(o, P.get(f"{a_}.fc.weight"), P.get(f"{a_}.fc.bias").f32, y1.f32, y1.hi, y1.lo,
                       ln=dict(res=x.f32, gamma=P.get(f"{a_}.layer_norm.weight").f32, beta=P.get(f"{a_}.layer_norm.bias").f32, lens=lens, T=T,
                               z=tp.f32(f"{pf}.z1", (B, T, d)), stats=tp.f32(f"{pf}.st1", (R, 2)), pre=self._fft_drop(tp, pf)[0]))
        else:
            y0 = scr.scratch("y0", (B, T, d))
            g.conv_fwd(o, P.get(f"{a_}.fc.weight"), P.get(f"{a_}.fc.bias").f32, y0, None, None)
            be.ln_fwd(y0, x.f32, P.get(f"{a_}.layer_norm.weight").f32, P.get(f"{a_}.layer_norm.bias").f32, lens, T, R, d,
                      tp.f32(f"{pf}.z1", (B, T, d)), tp.f32(f"{pf}.st1", (R, 2)), y1.f32, y1.hi, y1.lo,
                      pre=self._fft_drop(tp, pf)[0])
        h = tp.act(f"{pf}.h", B, T, self.d_inner, f32=False)
        g.conv_fwd(y1, P.get(f"{f_}.w_1.weight"), P.get(f"{f_}.w_1.bias").f32, None, h.hi, h.lo, relu=True)
        out = tp.act(f"{pf}.out", B, T, d)
        if fused_ln and P.get(f"{f_}.w_2.weight").shape[0] == 1:     # conv k=1 (w_2) with the same epilogue (SubLayers.py:88-91)
            g.conv_fwd(h, P.get(f"{f_}.w_2.weight"), P.get(f"{f_}.w_2.bias").f32, out.f32, out.hi, out.lo,
                       ln=dict(res=y1.f32, gamma=P.get(f"{f_}.layer_norm.weight").f32, beta=P.get(f"{f_}.layer_norm.bias").f32, lens=lens, T=T,
                               z=tp.f32(f"{pf}.z2", (B, T, d)), stats=tp.f32(f"{pf}.st2", (R, 2)), pre=self._fft_drop(tp, pf)[1]))
        else:
            y2 = scr.scratch("y0", (B, T, d))
            g.conv_fwd(h, P.get(f"{f_}.w_2.weight"), P.get(f"{f_}.w_2.bias").f32, y2, None, None)
            be.ln_fwd(y2, y1.f32, P.get(f"{f_}.layer_norm.weight").f32, P.get(f"{f_}.layer_norm.bias").f32, lens, T, R, d,
                      tp.f32(f"{pf}.z2", (B, T, d)), tp.f32(f"{pf}.st2", (R, 2)), out.f32, out.hi, out.lo,
                      pre=self._fft_drop(tp, pf)[1])
        return out

    def fft_bwd(self, P: ParamSet, G: ParamSet, pf: str, tp: Tape, x: Act, lens, H: int, dout: torch.Tensor,
                dx_out: torch.Tensor):
        """dout: dL/d(out) fp32 [B,T,d] (kept: the tangent-backward pass re-reads it);
        dx_out: destination for dL/d(x)."""
        be, g, scr = self.be, self.g, self.scr
        B, T, d = x.B, x.T, self.d
        R = B * T
        a_, f_ = self._fft_names(pf)
        dk, Tp, qm, pm, om = self._attn_geom(B, T, H)
        wqkv, _ = self._qkv(P, pf)
        gwqkv, gbqkv = self._qkv(G, pf)
        qkv_h, qkv_l = tp.bf(f"{pf}.qkv", (R, 3 * d))
        o = tp.act(f"{pf}.o", B, T, d, f32=False)
        y1 = tp.act(f"{pf}.y1", B, T, d)
        h = tp.act(f"{pf}.h", B, T, self.d_inner, f32=False)
        # LN2
        dz2 = tp.act(f"{pf}.dz2", B, T, d)          # f32 part becomes dL/dy1 (total) after the ADD_C below
        be.ln_bwd(dout, tp.f32(f"{pf}.z2", (B, T, d)), tp.f32(f"{pf}.st2", (R, 2)), P.get(f"{f_}.layer_norm.weight").f32,
                  lens, T, R, d, 0, dz2.f32, dz2.hi, dz2.lo, G.get(f"{f_}.layer_norm.weight").f32,
                  G.get(f"{f_}.layer_norm.bias").f32, G.get(f"{f_}.w_2.bias").f32, pre=self._fft_drop(tp, pf)[1])
        # conv k=1 (w_2), ReLU gate, conv k=9 (w_1)
        dh = tp.act(f"{pf}.dh", B, T, self.d_inner, f32=False)
        g.conv_dgrad(dz2, P.get(f"{f_}.w_2.weight"), None, dh.hi, dh.lo, gate=h.hi)
        with be.side():
            g.conv_wgrad(dz2, h, G.get(f"{f_}.w_2.weight").f32)
            be.colsum(None, dh.hi, dh.lo, 1, R, self.d_inner, G.get(f"{f_}.w_1.bias").f32)
            g.conv_wgrad(dh, y1, G.get(f"{f_}.w_1.weight").f32)
        g.conv_dgrad(dh, P.get(f"{f_}.w_1.weight"), dz2.f32, None, None, add_c=True)       # dz2.f32 := dL/dy1
        # LN1  (f32 result goes straight into dx_out; the QKV dgrad below adds onto it)
        dz1 = Act(dx_out, *tp.bf(f"{pf}.dz1", (B, T, d)), B, T, d)
        be.ln_bwd(dz2.f32, tp.f32(f"{pf}.z1", (B, T, d)), tp.f32(f"{pf}.st1", (R, 2)), P.get(f"{a_}.layer_norm.weight").f32,
                  lens, T, R, d, 0, dz1.f32, dz1.hi, dz1.lo, G.get(f"{a_}.layer_norm.weight").f32,
                  G.get(f"{a_}.layer_norm.bias").f32, G.get(f"{a_}.fc.bias").f32, pre=self._fft_drop(tp, pf)[0])
        do = tp.act(f"{pf}.do", B, T, d, f32=False)
        g.conv_dgrad(dz1, P.get(f"{a_}.fc.weight"), None, do.hi, do.lo)
        with be.side():
            g.conv_wgrad(dz1, o, G.get(f"{a_}.fc.weight").f32)
        # attention
        dq_h, dq_l = tp.bf(f"{pf}.dqkv", (R, 3 * d))
        if self._fused_attn(dk):
            # P is recomputed from q, k and the saved log-sum-exp; the two key-tile kernels (dK, dV) and the query-tile kernel (dQ)
            # run side by side.  dP / dS are only written out for tapes a Hessian-vector pass will re-read.
            emit = self._attn_emit(tp, pf)
            dP = tp.f32(f"{pf}.dP", (B, H, T, Tp)) if emit else None
            ds_h, ds_l = tp.bf(f"{pf}.dS", (B, H, T, Tp)) if emit else (None, None)
            args = (qkv_h, qkv_l, lens, B, H, T, dk, o.hi, o.lo, self._attn_vec(tp, f"{pf}.lse", B, H, T), do.hi, do.lo,
                    self._attn_vec(tp, f"{pf}.dvec", B, H, T), dq_h, dq_l)
            sp = g.split_of("bmm")
            be.attn_bwd(L.ATTN_PREP, *args)
            with be.branch("att", local=True):
                be.attn_bwd(L.ATTN_DK, *args, split=sp)
            with be.branch("att2", local=True):
                be.attn_bwd(L.ATTN_DV, *args, split=sp)
            be.attn_bwd(L.ATTN_DQ, *args, dP, ds_h, ds_l, Tp, split=sp)
            be.join("att", local=True)
            be.join("att2", local=True)
        else:
            p_h, p_l = tp.bf(f"{pf}.P", (B, H, T, Tp))
            dP = tp.f32(f"{pf}.dP", (B, H, T, Tp))
            with be.branch("att", local=True):                  # dV needs only dO and P: beside dP -> softmax-bwd -> dQ
                g.bmm(pm(p_h, p_l), True, om(do.hi, do.lo), True, qm(dq_h, dq_l, 2), B, H)                  # dV = P^T dO
            g.bmm(om(do.hi, do.lo), False, qm(qkv_h, qkv_l, 2), False, pm(None, None, dP), B, H)
            ds_h, ds_l = tp.bf(f"{pf}.dS", (B, H, T, Tp))
            be.softmax(1, dP, None, p_h, p_l, None, None, lens, B * H, H, T, T, Tp, ds_h, ds_l)
            sc = 1.0 / math.sqrt(dk)
            g.bmm(pm(ds_h, ds_l), False, qm(qkv_h, qkv_l, 1), True, qm(dq_h, dq_l, 0), B, H, alpha=sc)       # dQ = dS K
            with be.branch("att", local=True):                  # dK runs beside dQ (disjoint column blocks of dqkv)
                g.bmm(pm(ds_h, ds_l), True, qm(qkv_h, qkv_l, 0), True, qm(dq_h, dq_l, 1), B, H, alpha=sc)    # dK = dS^T Q
            be.join("att", local=True)
        dqkv = Act(None, dq_h, dq_l, B, T, 3 * d)
        with be.side():
            g.conv_wgrad(dqkv, x, gwqkv.f32)
            be.colsum(None, dq_h, dq_l, 1, R, 3 * d, gbqkv.f32)
        g.conv_dgrad(dqkv, wqkv, dx_out, None, None, add_c=True)

    def _lin_t(self, x: Act, xd: Optional[Act], w: Wt, wd: Optional[Wt], bd, out_f32, out_hi, out_lo, relu_gate=None):
        """Tangent of y = conv(x, W) + b:  yd = conv(xd, W) + conv(x, Wd) + bd (then optional ReLU gate),
        ONE launch (two product terms accumulate in the same TMEM tile)."""
        assert xd is not None or wd is not None, "tangent of a linear op with zero input and weight tangents"
        if xd is not None and wd is not None:
            self.g.conv_fwd(xd, w, bd, out_f32, out_hi, out_lo, gate=relu_gate, x2=x, w2=wd)
        elif xd is not None:
            self.g.conv_fwd(xd, w, bd, out_f32, out_hi, out_lo, gate=relu_gate)
        else:
            self.g.conv_fwd(x, wd, bd, out_f32, out_hi, out_lo, gate=relu_gate)

    def _dgrad_t(self, dy: Act, ddy: Act, w: Wt, wd: Optional[Wt], out_f32, out_hi, out_lo, gate=None, add_c=False):
        """Tangent of dx = dgrad(dy, W):  ddx = dgrad(ddy, W) + dgrad(dy, Wd)  (+ existing out_f32 if add_c)."""
        if wd is None:
            self.g.conv_dgrad(ddy, w, out_f32, out_hi, out_lo, gate=gate, add_c=add_c)
        else:
            self.g.conv_dgrad(ddy, w, out_f32, out_hi, out_lo, gate=gate, add_c=add_c, dy2=dy, w2=wd)

    def _wgrad_t(self, dy: Act, ddy: Act, x: Act, xd: Optional[Act], hv_w: torch.Tensor):
        """Tangent of dW = wgrad(dy, x):  ddW += wgrad(ddy, x) + wgrad(dy, xd)."""
        if xd is None:
            self.g.conv_wgrad(ddy, x, hv_w)
        else:
            self.g.conv_wgrad(ddy, x, hv_w, dy2=dy, x2=xd)

    def fft_tfwd(self, P: ParamSet, Pd: ParamSet, pf: str, tp: Tape, tt: Tape, x: Act, xd: Optional[Act], lens, H: int) -> Act:
        """Tangent forward; xd = input tangent (None = zero), Pd = parameter tangents."""
        be, g, scr = self.be, self.g, self.scr
        B, T, d = x.B, x.T, self.d
        R = B * T
        a_, f_ = self._fft_names(pf)
        dk, Tp, qm, pm, om = self._attn_geom(B, T, H)
        wqkv, _ = self._qkv(P, pf)
        wdqkv, bdqkv = self._qkv(Pd, pf)
        gd = lambda n: (Pd.get(n).f32 if Pd.has(n) else None)  # noqa: E731
        wdt = lambda n: (Pd.get(n) if Pd.has(n) else None)  # noqa: E731
        qkv_h, qkv_l = tp.bf(f"{pf}.qkv", (R, 3 * d))
        p_h, p_l = tp.bf(f"{pf}.P", (B, H, T, Tp))
        o = tp.act(f"{pf}.o", B, T, d, f32=False)
        y1 = tp.act(f"{pf}.y1", B, T, d)
        h = tp.act(f"{pf}.h", B, T, self.d_inner, f32=False)
        # qkv tangent
        qd_h, qd_l = tt.bf(f"{pf}.qkvd", (R, 3 * d))
        qd_f = scr.scratch("qkv_f", (R, 3 * d))
        self._lin_t(x, xd, wqkv, wdqkv, bdqkv.f32 if bdqkv is not None else None, qd_f, qd_h, qd_l)
        # Sdot = scale (Qd K^T + Q Kd^T)
        Sd = scr.scratch("S", (B, H, T, Tp))
        sc = 1.0 / math.sqrt(dk)
        g.bmm(qm(qd_h, qd_l, 0), False, qm(qkv_h, qkv_l, 1), False, pm(None, None, Sd), B, H, alpha=sc,
              A2=qm(qkv_h, qkv_l, 0), B2=qm(qd_h, qd_l, 1))
        pd_h, pd_l = tt.bf(f"{pf}.Pd", (B, H, T, Tp))
        be.softmax(1, Sd, None, p_h, p_l, None, None, lens, B * H, H, T, T, Tp, pd_h, pd_l)
        # Od = Pd V + P Vd
        od = tt.act(f"{pf}.od", B, T, d, f32=False)
        g.bmm(pm(pd_h, pd_l), False, qm(qkv_h, qkv_l, 2), True, om(od.hi, od.lo), B, H, A2=pm(p_h, p_l), B2=qm(qd_h, qd_l, 2))
        # fc + LN1
        y0d = scr.scratch("y0", (B, T, d))
        self._lin_t(o, od, P.get(f"{a_}.fc.weight"), wdt(f"{a_}.fc.weight"), gd(f"{a_}.fc.bias"), y0d, None, None)
        y1d = tt.act(f"{pf}.y1d", B, T, d)
        be.ln_tfwd(y0d, xd.f32 if xd is not None else None, tp.f32(f"{pf}.z1", (B, T, d)), tp.f32(f"{pf}.st1", (R, 2)),
                   P.get(f"{a_}.layer_norm.weight").f32, gd(f"{a_}.layer_norm.weight"), gd(f"{a_}.layer_norm.bias"), lens, T,
                   R, d, tt.f32(f"{pf}.z1d", (B, T, d)), y1d.f32, y1d.hi, y1d.lo, pre=self._fft_drop(tp, pf)[0])
        # conv9 + relu gate, conv1, LN2
        hd = tt.act(f"{pf}.hd", B, T, self.d_inner, f32=False)
        hd_f = scr.scratch("h_f", (B, T, self.d_inner))
        self._lin_t(y1, y1d, P.get(f"{f_}.w_1.weight"), wdt(f"{f_}.w_1.weight"), gd(f"{f_}.w_1.bias"), hd_f, hd.hi, hd.lo,
                    relu_gate=h.hi)
        y2d = scr.scratch("y0", (B, T, d))
        self._lin_t(h, hd, P.get(f"{f_}.w_2.weight"), wdt(f"{f_}.w_2.weight"), gd(f"{f_}.w_2.bias"), y2d, None, None)
        outd = tt.act(f"{pf}.outd", B, T, d)
        be.ln_tfwd(y2d, y1d.f32, tp.f32(f"{pf}.z2", (B, T, d)), tp.f32(f"{pf}.st2", (R, 2)),
                   P.get(f"{f_}.layer_norm.weight").f32, gd(f"{f_}.layer_norm.weight"), gd(f"{f_}.layer_norm.bias"), lens, T,
                   R, d, tt.f32(f"{pf}.z2d", (B, T, d)), outd.f32, outd.hi, outd.lo, pre=self._fft_drop(tp, pf)[1])
        return outd

    def fft_tbwd(self, P: ParamSet, Pd: ParamSet, HV: ParamSet, pf: str, tp: Tape, tt: Tape, x: Act, xd: Optional[Act],
                 lens, H: int, dout: torch.Tensor, ddout: torch.Tensor, ddx_out: torch.Tensor):
        """Tangent backward: given the primal tape (forward + backward signals), the tangent tape and
        ddout = tangent of dL/d(out), accumulate (H v) into HV and write the tangent of dL/dx."""
        be, g, scr = self.be, self.g, self.scr
        B, T, d = x.B, x.T, self.d
        R = B * T
        a_, f_ = self._fft_names(pf)
        dk, Tp, qm, pm, om = self._attn_geom(B, T, H)
        wqkv, _ = self._qkv(P, pf)
        wdqkv, _ = self._qkv(Pd, pf)
        hvwqkv, hvbqkv = self._qkv(HV, pf)
        gd = lambda n: (Pd.get(n).f32 if Pd.has(n) else None)  # noqa: E731
        wdt = lambda n: (Pd.get(n) if Pd.has(n) else None)  # noqa: E731
        hv = lambda n: HV.get(n).f32  # noqa: E731
        qkv_h, qkv_l = tp.bf(f"{pf}.qkv", (R, 3 * d))
        p_h, p_l = tp.bf(f"{pf}.P", (B, H, T, Tp))
        o = tp.act(f"{pf}.o", B, T, d, f32=False)
        y1 = tp.act(f"{pf}.y1", B, T, d)
        h = tp.act(f"{pf}.h", B, T, self.d_inner, f32=False)
        dz2 = tp.act(f"{pf}.dz2", B, T, d)             # .f32 holds dL/dy1 (total), hi/lo hold dz2
        dh = tp.act(f"{pf}.dh", B, T, self.d_inner, f32=False)
        dz1 = Act(None, *tp.bf(f"{pf}.dz1", (B, T, d)), B, T, d)
        do = tp.act(f"{pf}.do", B, T, d, f32=False)
        dP = tp.f32(f"{pf}.dP", (B, H, T, Tp))
        ds_h, ds_l = tp.bf(f"{pf}.dS", (B, H, T, Tp))
        dq_h, dq_l = tp.bf(f"{pf}.dqkv", (R, 3 * d))
        qd_h, qd_l = tt.bf(f"{pf}.qkvd", (R, 3 * d))
        pd_h, pd_l = tt.bf(f"{pf}.Pd", (B, H, T, Tp))
        od = tt.act(f"{pf}.od", B, T, d, f32=False)
        y1d = tt.act(f"{pf}.y1d", B, T, d)
        hd = tt.act(f"{pf}.hd", B, T, self.d_inner, f32=False)
        # ---- LN2 ----
        ddz2 = tt.act(f"{pf}.ddz2", B, T, d)
        be.ln_tbwd(dout, ddout, tp.f32(f"{pf}.z2", (B, T, d)), tt.f32(f"{pf}.z2d", (B, T, d)), tp.f32(f"{pf}.st2", (R, 2)),
                   P.get(f"{f_}.layer_norm.weight").f32, gd(f"{f_}.layer_norm.weight"), lens, T, R, d, 0, ddz2.f32, ddz2.hi,
                   ddz2.lo, hv(f"{f_}.layer_norm.weight"), hv(f"{f_}.layer_norm.bias"), hv(f"{f_}.w_2.bias"),
                   pre=self._fft_drop(tp, pf)[1])
        # ---- w_2 (k=1) with ReLU gate ----
        ddh = tt.act(f"{pf}.ddh", B, T, self.d_inner, f32=False)
        ddh_f = scr.scratch("h_f", (B, T, self.d_inner))
        self._dgrad_t(dz2, ddz2, P.get(f"{f_}.w_2.weight"), wdt(f"{f_}.w_2.weight"), ddh_f, ddh.hi, ddh.lo, gate=h.hi)
        with be.side():
            self._wgrad_t(dz2, ddz2, h, hd, hv(f"{f_}.w_2.weight"))
            be.colsum(None, ddh.hi, ddh.lo, 1, R, self.d_inner, hv(f"{f_}.w_1.bias"))
            self._wgrad_t(dh, ddh, y1, y1d, hv(f"{f_}.w_1.weight"))
        # ---- w_1 (k=9): ddy1 = dgrad(ddh, W1) + dgrad(dh, W1d) + ddz2 (residual) ----
        self._dgrad_t(dh, ddh, P.get(f"{f_}.w_1.weight"), wdt(f"{f_}.w_1.weight"), ddz2.f32, None, None, add_c=True)
        # ---- LN1 ----
        ddz1 = Act(ddx_out, *tt.bf(f"{pf}.ddz1", (B, T, d)), B, T, d)
        be.ln_tbwd(dz2.f32, ddz2.f32, tp.f32(f"{pf}.z1", (B, T, d)), tt.f32(f"{pf}.z1d", (B, T, d)),
                   tp.f32(f"{pf}.st1", (R, 2)), P.get(f"{a_}.layer_norm.weight").f32, gd(f"{a_}.layer_norm.weight"), lens, T,
                   R, d, 0, ddz1.f32, ddz1.hi, ddz1.lo, hv(f"{a_}.layer_norm.weight"), hv(f"{a_}.layer_norm.bias"),
                   hv(f"{a_}.fc.bias"), pre=self._fft_drop(tp, pf)[0])
        # ---- fc ----
        ddo = tt.act(f"{pf}.ddo", B, T, d, f32=False)
        ddo_f = scr.scratch("o_f", (B, T, d))
        self._dgrad_t(dz1, ddz1, P.get(f"{a_}.fc.weight"), wdt(f"{a_}.fc.weight"), ddo_f, ddo.hi, ddo.lo)
        with be.side():
            self._wgrad_t(dz1, ddz1, o, od, hv(f"{a_}.fc.weight"))
        # ---- attention ----
        sc = 1.0 / math.sqrt(dk)
        ddP = scr.scratch("S", (B, H, T, Tp))
        g.bmm(om(ddo.hi, ddo.lo), False, qm(qkv_h, qkv_l, 2), False, pm(None, None, ddP), B, H,             # ddO V^T + dO Vd^T
              A2=om(do.hi, do.lo), B2=qm(qd_h, qd_l, 2))
        dds_h, dds_l = tt.bf(f"{pf}.ddS", (B, H, T, Tp))
        be.softmax(2, dP, ddP, p_h, p_l, pd_h, pd_l, lens, B * H, H, T, T, Tp, dds_h, dds_l)
        ddq_h, ddq_l = tt.bf(f"{pf}.ddqkv", (R, 3 * d))
        with be.branch("att", local=True):
            # ddV = Pd^T dO + P^T ddO
            g.bmm(pm(pd_h, pd_l), True, om(do.hi, do.lo), True, qm(ddq_h, ddq_l, 2), B, H, A2=pm(p_h, p_l), B2=om(ddo.hi, ddo.lo))
            # ddK = scale (ddS^T Q + dS^T Qd)
            g.bmm(pm(dds_h, dds_l), True, qm(qkv_h, qkv_l, 0), True, qm(ddq_h, ddq_l, 1), B, H, alpha=sc,
                  A2=pm(ds_h, ds_l), B2=qm(qd_h, qd_l, 0))
        # ddQ = scale (ddS K + dS Kd)
        g.bmm(pm(dds_h, dds_l), False, qm(qkv_h, qkv_l, 1), True, qm(ddq_h, ddq_l, 0), B, H, alpha=sc,
              A2=pm(ds_h, ds_l), B2=qm(qd_h, qd_l, 1))
        be.join("att", local=True)
        dqkv = Act(None, dq_h, dq_l, B, T, 3 * d)
        ddqkv = Act(None, ddq_h, ddq_l, B, T, 3 * d)
        with be.side():
            self._wgrad_t(dqkv, ddqkv, x, xd, hvwqkv.f32)
            be.colsum(None, ddq_h, ddq_l, 1, R, 3 * d, hvbqkv.f32)
        self._dgrad_t(dqkv, ddqkv, wqkv, wdqkv, ddx_out, None, None, add_c=True)

    # ---------------------------------------------------------------------------------------------
    # VariancePredictor (modules.py:197-250)
    # ---------------------------------------------------------------------------------------------
    @_strict_precision
    def vp_fwd(self, P, pf, tp, x: Act, lens, out: torch.Tensor):
        be, g = self.be, self.g
        B, Lq, d = x.B, x.T, self.d
        R = B * Lq
        c = f"{pf}.conv_layer"
        h1 = tp.f32(f"{pf}.h1", (B, Lq, d))
        g.conv_fwd(x, P.get(f"{c}.conv1d_1.conv.weight"), P.get(f"{c}.conv1d_1.conv.bias").f32, h1, None, None, relu=True)
        a1 = tp.act(f"{pf}.a1", B, Lq, d)
        be.ln_fwd(h1, None, P.get(f"{c}.layer_norm_1.weight").f32, P.get(f"{c}.layer_norm_1.bias").f32, None, Lq, R, d, None,
                  tp.f32(f"{pf}.st1", (R, 2)), a1.f32, a1.hi, a1.lo, post=self._vp_drop(tp, pf)[0])
        h2 = tp.f32(f"{pf}.h2", (B, Lq, d))
        g.conv_fwd(a1, P.get(f"{c}.conv1d_2.conv.weight"), P.get(f"{c}.conv1d_2.conv.bias").f32, h2, None, None, relu=True)
        a2 = tp.f32(f"{pf}.a2", (B, Lq, d))
        be.ln_fwd(h2, None, P.get(f"{c}.layer_norm_2.weight").f32, P.get(f"{c}.layer_norm_2.bias").f32, None, Lq, R, d, None,
                  tp.f32(f"{pf}.st2", (R, 2)), a2, None, None, post=self._vp_drop(tp, pf)[1])
        be.rowdot_fwd(a2, None, P.get(f"{pf}.linear_layer.weight").f32, None, P.get(f"{pf}.linear_layer.bias").f32, None, lens,
                      Lq, R, d, out)

    @_strict_precision
    def vp_bwd(self, P, G, pf, tp, x: Act, lens, dpred: torch.Tensor, dx_acc: torch.Tensor):
        """dx_acc += dL/dx (fp32 [B,L,d])."""
        be, g = self.be, self.g
        B, Lq, d = x.B, x.T, self.d
        R = B * Lq
        c = f"{pf}.conv_layer"
        a1 = tp.act(f"{pf}.a1", B, Lq, d)
        da2 = tp.f32(f"{pf}.da2", (B, Lq, d))
        be.rowdot_bwd(dpred, None, tp.f32(f"{pf}.a2", (B, Lq, d)), None, P.get(f"{pf}.linear_layer.weight").f32, None, lens,
                      Lq, R, d, da2, G.get(f"{pf}.linear_layer.weight").f32, G.get(f"{pf}.linear_layer.bias").f32)
        dc2 = tp.act(f"{pf}.dc2", B, Lq, d, f32=False)
        be.ln_bwd(da2, tp.f32(f"{pf}.h2", (B, Lq, d)), tp.f32(f"{pf}.st2", (R, 2)), P.get(f"{c}.layer_norm_2.weight").f32,
                  None, Lq, R, d, 1, None, dc2.hi, dc2.lo, G.get(f"{c}.layer_norm_2.weight").f32,
                  G.get(f"{c}.layer_norm_2.bias").f32, G.get(f"{c}.conv1d_2.conv.bias").f32, post=self._vp_drop(tp, pf)[1])
        with be.side():
            g.conv_wgrad(dc2, a1, G.get(f"{c}.conv1d_2.conv.weight").f32)
        da1 = tp.f32(f"{pf}.da1", (B, Lq, d))
        g.conv_dgrad(dc2, P.get(f"{c}.conv1d_2.conv.weight"), da1, None, None)
        dc1 = tp.act(f"{pf}.dc1", B, Lq, d, f32=False)
        be.ln_bwd(da1, tp.f32(f"{pf}.h1", (B, Lq, d)), tp.f32(f"{pf}.st1", (R, 2)), P.get(f"{c}.layer_norm_1.weight").f32,
                  None, Lq, R, d, 1, None, dc1.hi, dc1.lo, G.get(f"{c}.layer_norm_1.weight").f32,
                  G.get(f"{c}.layer_norm_1.bias").f32, G.get(f"{c}.conv1d_1.conv.bias").f32, post=self._vp_drop(tp, pf)[0])
        with be.side():
            g.conv_wgrad(dc1, x, G.get(f"{c}.conv1d_1.conv.weight").f32)
        g.conv_dgrad(dc1, P.get(f"{c}.conv1d_1.conv.weight"), dx_acc, None, None, add_c=True)

    @_strict_precision
    def vp_tfwd(self, P, Pd, pf, tp, tt, x: Act, xd: Optional[Act], lens, outd: torch.Tensor):
        be = self.be
        B, Lq, d = x.B, x.T, self.d
        R = B * Lq
        c = f"{pf}.conv_layer"
        gd = lambda n: (Pd.get(n).f32 if Pd.has(n) else None)  # noqa: E731
        wdt = lambda n: (Pd.get(n) if Pd.has(n) else None)  # noqa: E731
        h1 = tp.f32(f"{pf}.h1", (B, Lq, d))
        h2 = tp.f32(f"{pf}.h2", (B, Lq, d))
        a1 = tp.act(f"{pf}.a1", B, Lq, d)
        # relu gate needs a bf16 view of h1/h2: gate on the fp32 relu output via a hi copy kept in the tape
        h1g, _ = tp.bf(f"{pf}.h1g", (B, Lq, d))
        h2g, _ = tp.bf(f"{pf}.h2g", (B, Lq, d))
        be.split_(h1, h1g, None)
        be.split_(h2, h2g, None)
        h1d = tt.f32(f"{pf}.h1d", (B, Lq, d))
        self._lin_t(x, xd, P.get(f"{c}.conv1d_1.conv.weight"), wdt(f"{c}.conv1d_1.conv.weight"), gd(f"{c}.conv1d_1.conv.bias"),
                    h1d, None, None, relu_gate=h1g)
        a1d = tt.act(f"{pf}.a1d", B, Lq, d)
        be.ln_tfwd(h1d, None, h1, tp.f32(f"{pf}.st1", (R, 2)), P.get(f"{c}.layer_norm_1.weight").f32,
                   gd(f"{c}.layer_norm_1.weight"), gd(f"{c}.layer_norm_1.bias"), None, Lq, R, d, None, a1d.f32, a1d.hi, a1d.lo,
                   post=self._vp_drop(tp, pf)[0])
        h2d = tt.f32(f"{pf}.h2d", (B, Lq, d))
        self._lin_t(a1, a1d, P.get(f"{c}.conv1d_2.conv.weight"), wdt(f"{c}.conv1d_2.conv.weight"), gd(f"{c}.conv1d_2.conv.bias"),
                    h2d, None, None, relu_gate=h2g)
        a2d = tt.f32(f"{pf}.a2d", (B, Lq, d))
        be.ln_tfwd(h2d, None, h2, tp.f32(f"{pf}.st2", (R, 2)), P.get(f"{c}.layer_norm_2.weight").f32,
                   gd(f"{c}.layer_norm_2.weight"), gd(f"{c}.layer_norm_2.bias"), None, Lq, R, d, None, a2d, None, None,
                   post=self._vp_drop(tp, pf)[1])
        be.rowdot_fwd(tp.f32(f"{pf}.a2", (B, Lq, d)), a2d, P.get(f"{pf}.linear_layer.weight").f32, gd(f"{pf}.linear_layer.weight"),
                      None, gd(f"{pf}.linear_layer.bias"), lens, Lq, R, d, outd)

    @_strict_precision
    def vp_tbwd(self, P, Pd, HV, pf, tp, tt, x: Act, xd: Optional[Act], lens, dpred, ddpred, ddx_acc: torch.Tensor):
        be, g = self.be, self.g
        B, Lq, d = x.B, x.T, self.d
        R = B * Lq
        c = f"{pf}.conv_layer"
        gd = lambda n: (Pd.get(n).f32 if Pd.has(n) else None)  # noqa: E731
        wdt = lambda n: (Pd.get(n) if Pd.has(n) else None)  # noqa: E731
        hv = lambda n: HV.get(n).f32  # noqa: E731
        a1 = tp.act(f"{pf}.a1", B, Lq, d)
        a1d = tt.act(f"{pf}.a1d", B, Lq, d)
        dc2 = tp.act(f"{pf}.dc2", B, Lq, d, f32=False)
        dc1 = tp.act(f"{pf}.dc1", B, Lq, d, f32=False)
        dda2 = tt.f32(f"{pf}.dda2", (B, Lq, d))
        be.rowdot_bwd(dpred, ddpred, tp.f32(f"{pf}.a2", (B, Lq, d)), tt.f32(f"{pf}.a2d", (B, Lq, d)),
                      P.get(f"{pf}.linear_layer.weight").f32, gd(f"{pf}.linear_layer.weight"), lens, Lq, R, d, dda2,
                      hv(f"{pf}.linear_layer.weight"), hv(f"{pf}.linear_layer.bias"))
        ddc2 = tt.act(f"{pf}.ddc2", B, Lq, d, f32=False)
        be.ln_tbwd(tp.f32(f"{pf}.da2", (B, Lq, d)), dda2, tp.f32(f"{pf}.h2", (B, Lq, d)), tt.f32(f"{pf}.h2d", (B, Lq, d)),
                   tp.f32(f"{pf}.st2", (R, 2)), P.get(f"{c}.layer_norm_2.weight").f32, gd(f"{c}.layer_norm_2.weight"), None, Lq,
                   R, d, 1, None, ddc2.hi, ddc2.lo, hv(f"{c}.layer_norm_2.weight"), hv(f"{c}.layer_norm_2.bias"),
                   hv(f"{c}.conv1d_2.conv.bias"), post=self._vp_drop(tp, pf)[1])
        with be.side():
            self._wgrad_t(dc2, ddc2, a1, a1d, hv(f"{c}.conv1d_2.conv.weight"))
        dda1 = tt.f32(f"{pf}.dda1", (B, Lq, d))
        self._dgrad_t(dc2, ddc2, P.get(f"{c}.conv1d_2.conv.weight"), wdt(f"{c}.conv1d_2.conv.weight"), dda1, None, None)
        ddc1 = tt.act(f"{pf}.ddc1", B, Lq, d, f32=False)
        be.ln_tbwd(tp.f32(f"{pf}.da1", (B, Lq, d)), dda1, tp.f32(f"{pf}.h1", (B, Lq, d)), tt.f32(f"{pf}.h1d", (B, Lq, d)),
                   tp.f32(f"{pf}.st1", (R, 2)), P.get(f"{c}.layer_norm_1.weight").f32, gd(f"{c}.layer_norm_1.weight"), None, Lq,
                   R, d, 1, None, ddc1.hi, ddc1.lo, hv(f"{c}.layer_norm_1.weight"), hv(f"{c}.layer_norm_1.bias"),
                   hv(f"{c}.conv1d_1.conv.bias"), post=self._vp_drop(tp, pf)[0])
        with be.side():
            self._wgrad_t(dc1, ddc1, x, xd, hv(f"{c}.conv1d_1.conv.weight"))
        self._dgrad_t(dc1, ddc1, P.get(f"{c}.conv1d_1.conv.weight"), wdt(f"{c}.conv1d_1.conv.weight"), ddx_acc, None, None,
                      add_c=True)

    # ---------------------------------------------------------------------------------------------
    # whole model
    # ---------------------------------------------------------------------------------------------
    class _BranchScratch:
        """While issuing work for an auxiliary stream, bind the branch's own scratch buffers."""

        def __init__(self, eng):
            self.eng = eng

        def __enter__(self):
            self.eng.scr = self.eng._scr_branch

        def __exit__(self, *a):
            self.eng.scr = self.eng._scr_main
            return False

    def encoder_early(self, P: ParamSet, bt: Batch, tp: Tape, drop_pass: Optional[int] = None) -> Act:
        """The encoder of a pass whose other inputs are not ready yet (the query pass: the encoder is not adapted, so it
        does not depend on the inner loop).  Issued on the 'enc' branch; `forward(..., enc=...)` joins it."""
        tp.drop_pass = drop_pass
        with self.be.branch("enc"), FS2Engine._BranchScratch(self):
            return self.encoder_fwd(P, bt.texts, bt.src_lens, bt.B, bt.L, tp)

    # ---- forward-type passes of an FFT stack as concurrent chains over groups of utterances (see SPLIT_FWD) ----
    def _groups(self, B: int, T: int):
        if not SPLIT_FWD or B < 2 or B * T > SPLIT_FWD_MAX_ROWS:
            return [(0, B)]
        return [(0, B // 2), (B // 2, B)]

    class _ChainScratch:
        """Scratch buffers of one chain: chains run concurrently, so they must not share `scr` (keyed by the enclosing branch too:
        the query encoder's chains run beside the main stream's)."""

        def __init__(self, eng, h):
            self.eng, self.h = eng, h

        def __enter__(self):
            e = self.eng
            key = (getattr(e.be, "_cur_branch", None), self.h)
            pool = e.__dict__.setdefault("_chain_scr", {})
            if key not in pool:
                pool[key] = Tape(e.be, e.split)
            self.saved = e.scr
            e.scr = pool[key]

        def __exit__(self, *a):
            self.eng.scr = self.saved

    def _stack_fwd(self, P: ParamSet, stack: str, n_layers: int, tp: Tape, x: Act, lens, H: int) -> Act:
        groups = self._groups(x.B, x.T)
        if len(groups) == 1:
            for i in range(n_layers):
                x = self.fft_fwd(P, f"{stack}.layer_stack.{i}", tp, x, lens, H)
            return x
        be = self.be
        for h, (b0, b1) in enumerate(groups):
            with be.branch(f"grp{h}", local=True), FS2Engine._ChainScratch(self, h):
                tv = TapeView(tp, b0, b1, x.B, x.T)
                xh = _act_rows(x, b0, b1)
                for i in range(n_layers):
                    xh = self.fft_fwd(P, f"{stack}.layer_stack.{i}", tv, xh, lens[b0:b1], H)
        for h in range(len(groups)):
            be.join(f"grp{h}", local=True)
        return tp.act(f"{stack}.layer_stack.{n_layers - 1}.out", x.B, x.T, self.d)

    def _stack_tfwd(self, P: ParamSet, Pd: ParamSet, stack: str, n_layers: int, tp: Tape, tt: Tape, y: Act, yd: Optional[Act], lens, H: int):
        """Tangent forward through a stack; y = the stack's primal input (layer i's input is layer i-1's saved output)."""
        B, T, d = y.B, y.T, self.d
        be = self.be
        groups = self._groups(B, T)
        for h, (b0, b1) in enumerate(groups):
            one = len(groups) == 1
            tpv = tp if one else TapeView(tp, b0, b1, B, T)
            ttv = tt if one else TapeView(tt, b0, b1, B, T)

            def chain():
                yh, ydh = (y, yd) if one else (_act_rows(y, b0, b1), _act_rows(yd, b0, b1))
                for i in range(n_layers):
                    pf = f"{stack}.layer_stack.{i}"
                    ynext = tpv.act(f"{pf}.out", b1 - b0, T, d)
                    ydh = self.fft_tfwd(P, Pd, pf, tpv, ttv, yh, ydh, lens if one else lens[b0:b1], H)
                    yh = ynext
                return ydh

            if one:
                return chain()
            with be.branch(f"grp{h}", local=True), FS2Engine._ChainScratch(self, h):
                chain()
        for h in range(len(groups)):
            be.join(f"grp{h}", local=True)
        return tt.act(f"{stack}.layer_stack.{n_layers - 1}.outd", B, T, d)

    def encoder_fwd(self, P: ParamSet, texts, src_lens, B: int, Lq: int, tp: Tape) -> Act:
        """Encoder.forward (Models.py:73-100): embedding + position_enc, then the FFT blocks."""
        d = self.d
        x = tp.act("enc.x0", B, Lq, d)
        self.be.embed_fwd(texts, P.get("encoder.src_word_emb.weight").f32, None, self.consts["encoder.position_enc"], Lq,
                          B * Lq, d, x.f32, x.hi, x.lo)
        return self._stack_fwd(P, "encoder", self.n_enc, tp, x, src_lens, self.h_enc)

    def _position_table(self, name: str, T: int, eval_mode: bool):
        """Models.py:82-91 / 148-160: the stored table covers max_seq_len + 1 positions; under model.eval() a longer
        sequence gets a freshly computed sinusoid table (float64 on the host -> fp32, as get_sinusoid_encoding_table),
        in train mode the reference truncates the sequence instead (not supported here: refuse loudly)."""
        tab = self.consts[name]
        if T <= self.cfg["max_seq_len"]:
            return tab
        if not eval_mode:
            raise ValueError(f"sequence of {T} positions > max_seq_len={self.cfg['max_seq_len']} in train mode: the "
                             "reference truncates the decoder input (Models.py:161-166); not supported")
        key = (name, T)
        cache = self.__dict__.setdefault("_long_tables", {})
        if key not in cache:
            pos = torch.arange(T, dtype=torch.float64)[:, None]
            j = torch.arange(self.d, dtype=torch.float64)[None, :]
            ang = pos / torch.pow(torch.tensor(10000.0, dtype=torch.float64), 2 * torch.div(j, 2, rounding_mode="floor") / self.d)
            t = torch.where((torch.arange(self.d) % 2 == 0)[None, :], torch.sin(ang), torch.cos(ang)).to(torch.float32)
            cache[key] = t.contiguous().to(tab.device)
        return cache[key]

    def decoder_fwd(self, P: ParamSet, xin_f32, spk, mel_lens, B: int, T: int, tp: Tape, eval_mode: bool = False) -> Act:
        """Decoder.forward (Models.py:139-171) on (x + spk_emb): + position_enc, then the FFT blocks.
        `spk` may be None (plain Decoder module)."""
        d = self.d
        y = tp.act("dec.x0", B, T, d)
        self.be.add_rowvec(xin_f32, spk, d, self._position_table("decoder.position_enc", T, eval_mode), B, T, d, y.f32, y.hi, y.lo)
        return self._stack_fwd(P, "decoder", self.n_dec, tp, y, mel_lens, self.h_dec)

    def postnet_fwd(self, P: ParamSet, mel: Act, tp: Tape, update_bn: bool = True, eval_mode: bool = False) -> Act:
        """PostNet.forward (Layers.py:129-137): 4 x tanh(BN(conv5)) + BN(conv5); batch statistics in train mode,
        running statistics (no update, no dropout) under model.eval()."""
        be, g, scr = self.be, self.g, self.scr
        B, T = mel.B, mel.T
        R = B * T
        xin = mel
        for i in range(5):
            pre = f"postnet.convolutions.{i}"
            co = POSTNET_CH[i + 1]
            c = tp.f32(f"post.{i}.c", (B, T, co))
            g.conv_fwd(xin, P.get(f"{pre}.0.conv.weight"), P.get(f"{pre}.0.conv.bias").f32, c, None, None)
            o = tp.act(f"post.{i}.o", B, T, co, bf=(i < 4))
            if eval_mode:
                be.bn_eval(c, P.get(f"{pre}.1.weight").f32, P.get(f"{pre}.1.bias").f32, self.consts[f"{pre}.1.running_mean"],
                           self.consts[f"{pre}.1.running_var"], R, co, i < 4, o.f32, o.hi, o.lo)
                xin = o
                continue
            rm = self.consts[f"{pre}.1.running_mean"] if update_bn else None
            rv = self.consts[f"{pre}.1.running_var"] if update_bn else None
            be.bn_fwd(c, P.get(f"{pre}.1.weight").f32, P.get(f"{pre}.1.bias").f32, R, co, i < 4, rm, rv,
                      scr.scratch("bn.ws", (4 * 512,)), tp.f32(f"post.{i}.st", (2 * co,)), o.f32, o.hi, o.lo,
                      drop=self._site(tp, f"postnet.{i}", 0.5))
            xin = o
        return xin

    def forward(self, P: ParamSet, bt: Batch, tp: Tape, update_bn: bool = True, drop_pass: Optional[int] = None,
                enc: Optional[Act] = None, eval_mode: bool = False):
        """Teacher-forced forward + loss.  Returns dict with the reference's prediction tensors.
        drop_pass: None = dropout off (eval / parity-with-identity); an int = train-mode dropout, pass index mixed
        into every site seed (backward / tangent passes over `tp` reuse it).
        eval_mode: model.eval() semantics (PostNet BatchNorm uses running statistics; drop_pass must be None)."""
        be, g, scr, d = self.be, self.g, self.scr, self.d
        assert not (eval_mode and drop_pass is not None), "eval mode has no dropout"
        tp.drop_pass = drop_pass
        B, Lq, T = bt.B, bt.L, bt.T
        assert Lq <= self.cfg["max_seq_len"], "phoneme sequence longer than max_seq_len"
        # ---- encoder (Models.py:73-100) ----
        if enc is None:
            x = self.encoder_fwd(P, bt.texts, bt.src_lens, B, Lq, tp)
        else:
            be.join("enc")                                    # computed ahead of time by encoder_early (same tape)
            x = enc
        # ---- speaker embedding (base_adaptor.py:64-70) ----
        spk = tp.f32("spk", (B, d))
        be.spk_embed(bt.spk_ids, P.get("speaker_emb.model.weight").f32, bt.spk_ids.numel(), d, bt.average_spk, B, spk)
        x0 = tp.act("va.x0", B, Lq, d)
        be.add_rowvec(x.f32, spk, d, None, B, Lq, d, x0.f32, x0.hi, x0.lo)
        # ---- variance adaptor (modules.py:102-158) ----
        va = "variance_adaptor"
        logd = tp.f32("logd", (B, Lq))
        ppred = tp.f32("ppred", (B, Lq))
        epred = tp.f32("epred", (B, Lq))
        # The three predictor chains only feed the loss: they run on the 'vp' branch while this stream goes on through
        # the length regulator, decoder and postnet (15 small launches off the critical path); joined before the loss.
        with be.branch("vp"):
            self.vp_fwd(P, f"{va}.duration_predictor", tp, x0, bt.src_lens, logd)
            self.vp_fwd(P, f"{va}.pitch_predictor", tp, x0, bt.src_lens, ppred)
        idx_p = tp.buf("va.idx_p", (B, Lq), torch.int64)
        idx_e = tp.buf("va.idx_e", (B, Lq), torch.int64)
        be.bucketize(bt.pitches, self.consts[f"{va}.pitch_bins"], self.nbins - 1, B * Lq, idx_p)
        be.bucketize(bt.energies, self.consts[f"{va}.energy_bins"], self.nbins - 1, B * Lq, idx_e)
        x1 = tp.act("va.x1", B, Lq, d)
        be.embed_fwd(idx_p, P.get(f"{va}.pitch_embedding.weight").f32, x0.f32, None, Lq, B * Lq, d, x1.f32, x1.hi, x1.lo)
        with be.branch("vp"):
            self.vp_fwd(P, f"{va}.energy_predictor", tp, x1, bt.src_lens, epred)
        x2 = scr.scratch("va.x2", (B, Lq, d))
        be.embed_fwd(idx_e, P.get(f"{va}.energy_embedding.weight").f32, x1.f32, None, Lq, B * Lq, d, x2, None, None)
        lr_idx = tp.buf("lr.idx", (B, T), torch.int32)
        lr_len = tp.buf("lr.mel_len", (B,), torch.int64)
        be.lr_index(bt.durations, T, lr_idx, lr_len)
        xr = scr.scratch("lr.out", (B, T, d))
        be.lr_fwd(x2, lr_idx, xr)
        # ---- decoder (Models.py:139-171) ----
        y = self.decoder_fwd(P, xr, spk, bt.mel_lens, B, T, tp, eval_mode)
        # ---- mel_linear + postnet (fastspeech2.py:97-99, Layers.py:129-137) ----
        mel = tp.act("mel", B, T, N_MEL)
        g.conv_fwd(y, P.get("mel_linear.weight"), P.get("mel_linear.bias").f32, mel.f32, mel.hi, mel.lo)
        xin = self.postnet_fwd(P, mel, tp, update_bn, eval_mode)
        post = xin.f32
        be.axpby(1.0, mel.f32, 1.0, post)                    # postnet(output) + output
        loss6 = tp.f32("loss6", (6,))
        be.join("vp")
        be.loss_fwd(mel.f32, post, bt.mels, bt.mel_lens, ppred, bt.pitches, epred, bt.energies, logd, bt.durations,
                    bt.src_lens, B, T, Lq, N_MEL, scr.scratch("loss.ws", (8,)), loss6, tp.f32("loss.counts", (2,)))
        return {"mel": mel.f32, "postnet": post, "pitch": ppred, "energy": epred, "logd": logd, "loss6": loss6,
                "mel_len": lr_len}

    def synthesize(self, P: ParamSet, bt: Batch, tp: Tape, p_control: float = 1.0, e_control: float = 1.0,
                   d_control: float = 1.0, update_bn: bool = True, drop_pass: Optional[int] = None, eval_mode: bool = False):
        """Free-running forward: `forward_learner(learner, spk, texts, src_lens, max_src_len)` with every target None
        (base_adaptor.py:160-162,183-185 -> modules.py:85-99,132-139): pitch / energy embeddings come from the
        (control-scaled) PREDICTIONS, durations are clamp(round(exp(log_d) - 1) * d_control, 0), and the output length is
        data dependent.  ONE host synchronisation reads the B*L rounded durations (the reference syncs B*L times,
        modules.py:186); everything else is the training path's kernels.  `bt` needs spk_ids / texts / src_lens only.
        Returns the prediction dict (mel / postnet: [B, T, 80] with T = max(mel_len); d_rounded float as the reference)."""
        be, g, scr, d = self.be, self.g, self.scr, self.d
        assert not (eval_mode and drop_pass is not None), "eval mode has no dropout"
        tp.drop_pass = drop_pass
        B, Lq = bt.B, bt.L
        va = "variance_adaptor"
        x = self.encoder_fwd(P, bt.texts, bt.src_lens, B, Lq, tp)
        spk = tp.f32("spk", (B, d))
        be.spk_embed(bt.spk_ids, P.get("speaker_emb.model.weight").f32, bt.spk_ids.numel(), d, bt.average_spk, B, spk)
        x0 = tp.act("va.x0", B, Lq, d)
        be.add_rowvec(x.f32, spk, d, None, B, Lq, d, x0.f32, x0.hi, x0.lo)
        logd, ppred, epred = tp.f32("logd", (B, Lq)), tp.f32("ppred", (B, Lq)), tp.f32("epred", (B, Lq))
        self.vp_fwd(P, f"{va}.duration_predictor", tp, x0, bt.src_lens, logd)
        self.vp_fwd(P, f"{va}.pitch_predictor", tp, x0, bt.src_lens, ppred)
        if p_control != 1.0:
            be.unary(L.UN_SCALE, ppred, p_control, 0.0, ppred)              # prediction * control (modules.py:86)
        idx_p = tp.buf("va.idx_p", (B, Lq), torch.int64)
        be.bucketize(ppred, self.consts[f"{va}.pitch_bins"], self.nbins - 1, B * Lq, idx_p)
        x1 = tp.act("va.x1", B, Lq, d)
        be.embed_fwd(idx_p, P.get(f"{va}.pitch_embedding.weight").f32, x0.f32, None, Lq, B * Lq, d, x1.f32, x1.hi, x1.lo)
        self.vp_fwd(P, f"{va}.energy_predictor", tp, x1, bt.src_lens, epred)
        if e_control != 1.0:
            be.unary(L.UN_SCALE, epred, e_control, 0.0, epred)
        idx_e = tp.buf("va.idx_e", (B, Lq), torch.int64)
        be.bucketize(epred, self.consts[f"{va}.energy_bins"], self.nbins - 1, B * Lq, idx_e)
        x2 = scr.scratch("va.x2", (B, Lq, d))
        be.embed_fwd(idx_e, P.get(f"{va}.energy_embedding.weight").f32, x1.f32, None, Lq, B * Lq, d, x2, None, None)
        drnd = tp.f32("d_rounded", (B, Lq))
        be.duration_round(logd, d_control, drnd)
        # the output length is data dependent: one D2H of B*L floats (plumbing); int() truncation as modules.py:186-187
        T = int(drnd.detach().to("cpu").to(torch.int64).clamp_(min=0).sum(dim=1).max())
        if T <= 0:
            raise ValueError("free-running synthesis predicted zero frames for every utterance")
        if T > (1 << 20):
            raise ValueError(f"free-running synthesis predicted {T} frames (diverged duration predictor?)")
        if not eval_mode:
            T = min(T, self.cfg["max_seq_len"])      # train mode: the decoder keeps the first max_seq_len frames (Models.py:161-166);
                                                     # mel_len stays the full sum of durations, as the reference returns it
        lr_idx = tp.buf("lr.idx", (B, T), torch.int32)
        lr_len = tp.buf("lr.mel_len", (B,), torch.int64)
        be.lr_index(drnd, T, lr_idx, lr_len)
        xr = scr.scratch("lr.out", (B, T, d))
        be.lr_fwd(x2, lr_idx, xr)
        y = self.decoder_fwd(P, xr, spk, lr_len, B, T, tp, eval_mode)
        mel = tp.act("mel", B, T, N_MEL)
        g.conv_fwd(y, P.get("mel_linear.weight"), P.get("mel_linear.bias").f32, mel.f32, mel.hi, mel.lo)
        xin = self.postnet_fwd(P, mel, tp, update_bn, eval_mode)
        post = xin.f32
        be.axpby(1.0, mel.f32, 1.0, post)
        return {"mel": mel.f32, "postnet": post, "pitch": ppred, "energy": epred, "logd": logd, "d_rounded": drnd,
                "mel_len": lr_len, "T": T}

    def backward(self, P: ParamSet, G: ParamSet, bt: Batch, tp: Tape, loss_scale: float = 1.0, into_encoder: bool = True,
                 enc_G: Optional[ParamSet] = None):
        """dL*loss_scale/dparams accumulated into G (G must be zeroed by the caller when needed)."""
        be, g, scr, d = self.be, self.g, self.scr, self.d
        B, Lq, T = bt.B, bt.L, bt.T
        R = B * T
        va = "variance_adaptor"
        mel = tp.act("mel", B, T, N_MEL)
        post = tp.f32("post.4.o", (B, T, N_MEL))
        dmel = tp.act("dmel", B, T, N_MEL)                   # total dL/dmel (L1 + residual + postnet path)
        dpost = tp.f32("post.4.dout", (B, T, N_MEL))
        dp, de, dlogd = tp.f32("dp", (B, Lq)), tp.f32("de", (B, Lq)), tp.f32("dlogd", (B, Lq))
        be.loss_bwd(mel.f32, post, bt.mels, bt.mel_lens, tp.f32("ppred", (B, Lq)), bt.pitches, tp.f32("epred", (B, Lq)),
                    bt.energies, tp.f32("logd", (B, Lq)), bt.durations, bt.src_lens, B, T, Lq, N_MEL,
                    tp.f32("loss.counts", (2,)), loss_scale, 0, dmel.f32, dpost, dp, de, dlogd)
        be.axpby(1.0, dpost, 1.0, dmel.f32)                  # residual: postnet_output = postnet(mel) + mel
        # ---- variance predictors: they need only the loss gradients -> 'vp' branch, concurrently with the postnet /
        #      decoder backward; their input gradients land in separate buffers that are folded into dx after the join ----
        x0 = tp.act("va.x0", B, Lq, d)
        x1 = tp.act("va.x1", B, Lq, d)
        dxe, dx0 = tp.f32("va.dxe", (B, Lq, d)), tp.f32("va.dx0", (B, Lq, d))
        with be.branch("vp"):
            be.zero_(dxe)
            be.zero_(dx0)
            self.vp_bwd(P, G, f"{va}.energy_predictor", tp, x1, bt.src_lens, de, dxe)
            self.vp_bwd(P, G, f"{va}.pitch_predictor", tp, x0, bt.src_lens, dp, dx0)
            self.vp_bwd(P, G, f"{va}.duration_predictor", tp, x0, bt.src_lens, dlogd, dx0)
        # ---- postnet ----
        for i in range(4, -1, -1):
            pre = f"postnet.convolutions.{i}"
            ci, co = POSTNET_CH[i], POSTNET_CH[i + 1]
            xin = mel if i == 0 else tp.act(f"post.{i - 1}.o", B, T, ci)
            dout = tp.f32(f"post.{i}.dout", (B, T, co))
            dc = tp.act(f"post.{i}.dc", B, T, co)
            be.bn_bwd(dout, tp.f32(f"post.{i}.o", (B, T, co)) if i < 4 else None, tp.f32(f"post.{i}.c", (B, T, co)),
                      tp.f32(f"post.{i}.st", (2 * co,)), P.get(f"{pre}.1.weight").f32, R, co, i < 4, scr.scratch("bn.ws", (4 * 512,)),
                      dc.f32, dc.hi, dc.lo, G.get(f"{pre}.1.weight").f32, G.get(f"{pre}.1.bias").f32,
                      beta=P.get(f"{pre}.1.bias").f32, drop=self._site(tp, f"postnet.{i}", 0.5))
            with be.side():
                be.colsum(dc.f32, None, None, 1, R, co, G.get(f"{pre}.0.conv.bias").f32)
                g.conv_wgrad(dc, xin, G.get(f"{pre}.0.conv.weight").f32)
            if i > 0:
                g.conv_dgrad(dc, P.get(f"{pre}.0.conv.weight"), tp.f32(f"post.{i - 1}.dout", (B, T, ci)), None, None)
            else:
                g.conv_dgrad(dc, P.get(f"{pre}.0.conv.weight"), dmel.f32, None, None, add_c=True)
        # ---- mel_linear ----
        be.split_(dmel.f32, dmel.hi, dmel.lo)
        ylast = tp.act(f"decoder.layer_stack.{self.n_dec - 1}.out", B, T, d)
        with be.side():
            be.colsum(dmel.f32, None, None, 1, R, N_MEL, G.get("mel_linear.bias").f32)
            g.conv_wgrad(dmel, ylast, G.get("mel_linear.weight").f32)
        dcur = tp.f32(f"decoder.layer_stack.{self.n_dec - 1}.dout", (B, T, d))
        g.conv_dgrad(dmel, P.get("mel_linear.weight"), dcur, None, None)
        # ---- decoder ----
        for i in range(self.n_dec - 1, -1, -1):
            pf = f"decoder.layer_stack.{i}"
            xin = tp.act(f"decoder.layer_stack.{i - 1}.out", B, T, d) if i > 0 else tp.act("dec.x0", B, T, d)
            dnext = tp.f32(f"decoder.layer_stack.{i - 1}.dout", (B, T, d)) if i > 0 else tp.f32("dec.din", (B, T, d))
            self.fft_bwd(P, G, pf, tp, xin, bt.mel_lens, self.h_dec, dcur, dnext)
            dcur = dnext
        # ---- speaker add / length regulator ----
        dspk = tp.f32("dspk", (B, d))
        be.zero_(dspk)
        be.colsum(dcur, None, None, B, T, d, dspk)
        dx = tp.f32("va.dx", (B, Lq, d))
        be.lr_bwd(dcur, bt.durations, Lq, dx)
        # ---- variance adaptor ----
        be.embed_bwd(tp.buf("va.idx_e", (B, Lq), torch.int64), dx, B * Lq, d, -1, 1.0, G.get(f"{va}.energy_embedding.weight").f32)
        be.join("vp")
        be.axpby(1.0, dxe, 1.0, dx)                          # + energy predictor (input x1 = x0 + pitch embedding)
        be.embed_bwd(tp.buf("va.idx_p", (B, Lq), torch.int64), dx, B * Lq, d, -1, 1.0, G.get(f"{va}.pitch_embedding.weight").f32)
        be.axpby(1.0, dx0, 1.0, dx)                          # + pitch and duration predictors (input x0)
        be.colsum(dx, None, None, B, Lq, d, dspk)
        be.spk_embed_bwd(bt.spk_ids, dspk, bt.spk_ids.numel(), d, bt.average_spk, B, 1.0, G.get("speaker_emb.model.weight").f32)
        # ---- encoder ----
        if into_encoder and enc_G is not None:
            # nothing downstream on this stream needs the encoder gradient soon (the Hessian-vector passes follow):
            # run the encoder backward on the 'enc' branch into its own accumulation arena; the caller joins
            with be.branch("enc"), FS2Engine._BranchScratch(self):
                self._encoder_bwd(P, enc_G, bt, tp, dx)
        elif into_encoder:
            self._encoder_bwd(P, G, bt, tp, dx)
        be.join_side()                                      # weight-gradient branch joins before anyone reads G

    def _encoder_bwd(self, P, G, bt: Batch, tp: Tape, d_encout: torch.Tensor):
        """Plain backward through the encoder from dL/d(encoder output)."""
        be, d = self.be, self.d
        B, Lq = bt.B, bt.L
        dcur = d_encout
        for i in range(self.n_enc - 1, -1, -1):
            pf = f"encoder.layer_stack.{i}"
            xin = tp.act(f"encoder.layer_stack.{i - 1}.out", B, Lq, d) if i > 0 else tp.act("enc.x0", B, Lq, d)
            dnext = tp.f32(f"encoder.layer_stack.{i}.din", (B, Lq, d))
            self.fft_bwd(P, G, pf, tp, xin, bt.src_lens, self.h_enc, dcur, dnext)
            dcur = dnext
        be.embed_bwd(bt.texts, dcur, B * Lq, d, 0, 1.0, G.get("encoder.src_word_emb.weight").f32)

    # ---------------------------------------------------------------------------------------------
    # Hessian-vector product at the tape's parameters:  HV += d/d(eps) grad L(P + eps*Pd)
    # (Pd is non-zero on adapted parameters only; encoder = non-adapted => zero forward tangent.)
    # ---------------------------------------------------------------------------------------------
    def hvp(self, P: ParamSet, Pd: ParamSet, HV: ParamSet, bt: Batch, tp: Tape, tt: Tape, loss_scale: float = 1.0, before_encoder=None):
        """before_encoder: called when the ADAPTED region of HV is final (every adapted module's tangent backward has been issued),
        just before the pass walks back through the non-adapted encoder — the data-parallel step starts reducing the adapted 2/3 of
        the outer gradient there."""
        assert getattr(tp, "attn_emit", True), "this tape was recorded without the attention probabilities (attn_emit = False)"
        self.g.split_override = self.hvp_split if self.hvp_split != self.split else None
        self.g.tangent = True
        try:
            self._hvp(P, Pd, HV, bt, tp, tt, loss_scale, before_encoder)
        finally:
            self.g.split_override = None
            self.g.tangent = False

    def _hvp(self, P: ParamSet, Pd: ParamSet, HV: ParamSet, bt: Batch, tp: Tape, tt: Tape, loss_scale: float = 1.0, before_encoder=None):
        be, g, scr, d = self.be, self.g, self.scr, self.d
        B, Lq, T = bt.B, bt.L, bt.T
        R = B * T
        va = "variance_adaptor"
        lay = self.layout
        assert not lay.is_adapted_module("encoder"), "HVP with an adapted encoder is not implemented"
        gd = lambda n: (Pd.get(n).f32 if Pd.has(n) else None)  # noqa: E731
        wdt = lambda n: (Pd.get(n) if Pd.has(n) else None)  # noqa: E731
        hv = lambda n: HV.get(n).f32  # noqa: E731
        # ================= tangent forward =================
        x0 = tp.act("va.x0", B, Lq, d)
        x1 = tp.act("va.x1", B, Lq, d)
        spkd = None
        x0d = None
        if Pd.has("speaker_emb.model.weight"):
            spkd = tt.f32("spkd", (B, d))
            be.spk_embed(bt.spk_ids, Pd.get("speaker_emb.model.weight").f32, bt.spk_ids.numel(), d, bt.average_spk, B, spkd)
            x0d = tt.act("va.x0d", B, Lq, d)
            zl = scr.scratch("zeros.L", (B, Lq, d))
            be.zero_(zl)
            be.add_rowvec(zl, spkd, d, None, B, Lq, d, x0d.f32, x0d.hi, x0d.lo)
        va_adapted = lay.is_adapted_module(va)
        assert va_adapted and spkd is not None, "HVP expects speaker_emb and variance_adaptor in adapt.modules"
        logdd, ppd, epd = tt.f32("logdd", (B, Lq)), tt.f32("ppredd", (B, Lq)), tt.f32("epredd", (B, Lq))
        with be.branch("vp"):
            self.vp_tfwd(P, Pd, f"{va}.duration_predictor", tp, tt, x0, x0d, bt.src_lens, logdd)
            self.vp_tfwd(P, Pd, f"{va}.pitch_predictor", tp, tt, x0, x0d, bt.src_lens, ppd)
        idx_p = tp.buf("va.idx_p", (B, Lq), torch.int64)
        idx_e = tp.buf("va.idx_e", (B, Lq), torch.int64)
        x1d = tt.act("va.x1d", B, Lq, d)
        be.embed_fwd(idx_p, Pd.get(f"{va}.pitch_embedding.weight").f32, x0d.f32, None, Lq, B * Lq, d, x1d.f32, x1d.hi, x1d.lo)
        with be.branch("vp"):
            self.vp_tfwd(P, Pd, f"{va}.energy_predictor", tp, tt, x1, x1d, bt.src_lens, epd)
        x2d = scr.scratch("va.x2", (B, Lq, d))
        be.embed_fwd(idx_e, Pd.get(f"{va}.energy_embedding.weight").f32, x1d.f32, None, Lq, B * Lq, d, x2d, None, None)
        xrd = scr.scratch("lr.out", (B, T, d))
        be.lr_fwd(x2d, tp.buf("lr.idx", (B, T), torch.int32), xrd)
        yd = tt.act("dec.x0d", B, T, d)
        be.add_rowvec(xrd, spkd, d, None, B, T, d, yd.f32, yd.hi, yd.lo)
        y = tp.act("dec.x0", B, T, d)
        yd = self._stack_tfwd(P, Pd, "decoder", self.n_dec, tp, tt, y, yd, bt.mel_lens, self.h_dec)
        y = tp.act(f"decoder.layer_stack.{self.n_dec - 1}.out", B, T, d)
        mel = tp.act("mel", B, T, N_MEL)
        meld = tt.act("meld", B, T, N_MEL)
        self._lin_t(y, yd, P.get("mel_linear.weight"), wdt("mel_linear.weight"), gd("mel_linear.bias"), meld.f32, meld.hi, meld.lo)
        xin, xind = mel, meld
        for i in range(5):
            pre = f"postnet.convolutions.{i}"
            co = POSTNET_CH[i + 1]
            cd = tt.f32(f"post.{i}.cd", (B, T, co))
            self._lin_t(xin, xind, P.get(f"{pre}.0.conv.weight"), wdt(f"{pre}.0.conv.weight"), gd(f"{pre}.0.conv.bias"), cd, None, None)
            od = tt.act(f"post.{i}.od", B, T, co, bf=(i < 4))
            be.bn_tfwd(cd, tp.f32(f"post.{i}.c", (B, T, co)), tp.f32(f"post.{i}.st", (2 * co,)), P.get(f"{pre}.1.weight").f32,
                       gd(f"{pre}.1.weight"), gd(f"{pre}.1.bias"), tp.f32(f"post.{i}.o", (B, T, co)) if i < 4 else None, R, co,
                       i < 4, scr.scratch("bn.ws", (4 * 512,)), tt.f32(f"post.{i}.ts", (2 * co,)), od.f32, od.hi, od.lo,
                       beta=P.get(f"{pre}.1.bias").f32, drop=self._site(tp, f"postnet.{i}", 0.5))
            xin = tp.act(f"post.{i}.o", B, T, co, bf=(i < 4))
            xind = od
        # (postnet_output tangent = od4 + meld, but the L1 losses have zero curvature: not needed)
        # ================= tangent backward =================
        ddmel = tt.act("ddmel", B, T, N_MEL)
        ddpost = tt.f32("post.4.ddout", (B, T, N_MEL))
        ddp, dde, ddlogd = tt.f32("ddp", (B, Lq)), tt.f32("dde", (B, Lq)), tt.f32("ddlogd", (B, Lq))
        be.join("vp")
        be.loss_bwd(None, None, None, bt.mel_lens, ppd, None, epd, None, logdd, None, bt.src_lens, B, T, Lq, N_MEL,
                    tp.f32("loss.counts", (2,)), loss_scale, 1, ddmel.f32, ddpost, ddp, dde, ddlogd)
        ddxe, ddx0 = tt.f32("va.ddxe", (B, Lq, d)), tt.f32("va.ddx0", (B, Lq, d))
        with be.branch("vp"):                                # predictor tangent-backward chains, folded into ddx below
            be.zero_(ddxe)
            be.zero_(ddx0)
            self.vp_tbwd(P, Pd, HV, f"{va}.energy_predictor", tp, tt, x1, x1d, bt.src_lens, tp.f32("de", (B, Lq)), dde, ddxe)
            self.vp_tbwd(P, Pd, HV, f"{va}.pitch_predictor", tp, tt, x0, x0d, bt.src_lens, tp.f32("dp", (B, Lq)), ddp, ddx0)
            self.vp_tbwd(P, Pd, HV, f"{va}.duration_predictor", tp, tt, x0, x0d, bt.src_lens, tp.f32("dlogd", (B, Lq)), ddlogd, ddx0)
        # ddmel = 0 + ddpost(=0) so far; postnet chain
        for i in range(4, -1, -1):
            pre = f"postnet.convolutions.{i}"
            ci, co = POSTNET_CH[i], POSTNET_CH[i + 1]
            xin = mel if i == 0 else tp.act(f"post.{i - 1}.o", B, T, ci)
            xind = meld if i == 0 else tt.act(f"post.{i - 1}.od", B, T, ci)
            dc = tp.act(f"post.{i}.dc", B, T, co)
            ddc = tt.act(f"post.{i}.ddc", B, T, co)
            be.bn_tbwd(tp.f32(f"post.{i}.dout", (B, T, co)), tt.f32(f"post.{i}.ddout", (B, T, co)),
                       tp.f32(f"post.{i}.o", (B, T, co)) if i < 4 else None, tt.f32(f"post.{i}.od", (B, T, co)) if i < 4 else None,
                       tp.f32(f"post.{i}.c", (B, T, co)), tt.f32(f"post.{i}.cd", (B, T, co)), tp.f32(f"post.{i}.st", (2 * co,)),
                       tt.f32(f"post.{i}.ts", (2 * co,)), P.get(f"{pre}.1.weight").f32, gd(f"{pre}.1.weight"), R, co, i < 4,
                       scr.scratch("bn.ws", (4 * 512,)), ddc.f32, ddc.hi, ddc.lo, hv(f"{pre}.1.weight"), hv(f"{pre}.1.bias"),
                       beta=P.get(f"{pre}.1.bias").f32, bdot=gd(f"{pre}.1.bias"), drop=self._site(tp, f"postnet.{i}", 0.5))
            with be.side():
                be.colsum(ddc.f32, None, None, 1, R, co, hv(f"{pre}.0.conv.bias"))
                self._wgrad_t(dc, ddc, xin, xind, hv(f"{pre}.0.conv.weight"))
            if i > 0:
                self._dgrad_t(dc, ddc, P.get(f"{pre}.0.conv.weight"), wdt(f"{pre}.0.conv.weight"),
                              tt.f32(f"post.{i - 1}.ddout", (B, T, ci)), None, None)
            else:
                self._dgrad_t(dc, ddc, P.get(f"{pre}.0.conv.weight"), wdt(f"{pre}.0.conv.weight"), ddmel.f32, None, None,
                              add_c=True)
        # mel_linear
        dmel = tp.act("dmel", B, T, N_MEL)
        be.split_(ddmel.f32, ddmel.hi, ddmel.lo)
        ylast = tp.act(f"decoder.layer_stack.{self.n_dec - 1}.out", B, T, d)
        ylastd = tt.act(f"decoder.layer_stack.{self.n_dec - 1}.outd", B, T, d)
        with be.side():
            be.colsum(ddmel.f32, None, None, 1, R, N_MEL, hv("mel_linear.bias"))
            self._wgrad_t(dmel, ddmel, ylast, ylastd, hv("mel_linear.weight"))
        ddcur = tt.f32(f"decoder.layer_stack.{self.n_dec - 1}.ddout", (B, T, d))
        self._dgrad_t(dmel, ddmel, P.get("mel_linear.weight"), wdt("mel_linear.weight"), ddcur, None, None)
        # decoder
        for i in range(self.n_dec - 1, -1, -1):
            pf = f"decoder.layer_stack.{i}"
            xin = tp.act(f"decoder.layer_stack.{i - 1}.out", B, T, d) if i > 0 else tp.act("dec.x0", B, T, d)
            xind = tt.act(f"decoder.layer_stack.{i - 1}.outd", B, T, d) if i > 0 else tt.act("dec.x0d", B, T, d)
            ddnext = tt.f32(f"decoder.layer_stack.{i - 1}.ddout", (B, T, d)) if i > 0 else tt.f32("dec.ddin", (B, T, d))
            self.fft_tbwd(P, Pd, HV, pf, tp, tt, xin, xind, bt.mel_lens, self.h_dec, tp.f32(f"{pf}.dout", (B, T, d)), ddcur, ddnext)
            ddcur = ddnext
        ddspk = tt.f32("ddspk", (B, d))
        be.zero_(ddspk)
        be.colsum(ddcur, None, None, B, T, d, ddspk)
        ddx = tt.f32("va.ddx", (B, Lq, d))
        be.lr_bwd(ddcur, bt.durations, Lq, ddx)
        be.embed_bwd(idx_e, ddx, B * Lq, d, -1, 1.0, hv(f"{va}.energy_embedding.weight"))
        be.join("vp")
        be.axpby(1.0, ddxe, 1.0, ddx)
        be.embed_bwd(idx_p, ddx, B * Lq, d, -1, 1.0, hv(f"{va}.pitch_embedding.weight"))
        be.axpby(1.0, ddx0, 1.0, ddx)
        be.colsum(ddx, None, None, B, Lq, d, ddspk)
        be.spk_embed_bwd(bt.spk_ids, ddspk, bt.spk_ids.numel(), d, bt.average_spk, B, 1.0, hv("speaker_emb.model.weight"))
        if before_encoder is not None:
            be.join_side()                       # the adapted modules' weight-gradient products (side stream) are in HV
            before_encoder()
        # encoder: zero forward tangent => the tangent backward is a plain backward of ddx (mixed partials)
        self._encoder_bwd(P, HV, bt, tp, ddx)
        be.join_side()
