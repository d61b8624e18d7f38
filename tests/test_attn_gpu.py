"""Fused attention kernels (csrc/mtts_attn.cu: mtts_attn_fwd / mtts_attn_bwd) vs the CPU restatement of
ScaledDotProductAttention (Modules.py:14-25) and its autograd in oracle/ops_reference.RefOps, and vs torch autograd in fp64.
Covers ragged key lengths, T that is not a multiple of any tile size, one-tile sequences, the maximum training length (1000
frames), bf16x3 and single-pass bf16, and the emitted P / dP / dS tensors the Hessian-vector passes re-read.
Tolerance: 2e-5 of the tensor's max for bf16x3 (fp32-grade), 2e-2 for single-pass bf16."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

from meta_tts_b200 import lib as L  # noqa: E402
from meta_tts_b200.ops import CudaOps, split_bf16  # noqa: E402
from oracle.ops_reference import RefOps  # noqa: E402

DK = 128


def _inputs(B, H, T, seed, lens):
    g = torch.Generator().manual_seed(seed)
    qkv = torch.randn(B * T, 3 * H * DK, generator=g) * 0.9
    qkv[:, :H * DK] *= 1.7                               # sharper rows: scores of a few units
    do = torch.randn(B * T, H * DK, generator=g)
    klens = torch.tensor(lens, dtype=torch.int64) if lens is not None else None
    return qkv, do, klens


def _err(a, b):
    a, b = a.double(), b.double()
    assert torch.isfinite(a).all(), "non-finite values"
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def _run(dev, B, H, T, lens, split, emit, seed=0, tol=None):
    qkv, do, klens = _inputs(B, H, T, seed, lens)
    Tp, Tl = (T + 7) // 8 * 8, (T + 127) // 128 * 128
    qh, ql = split_bf16(qkv)
    dh, dl = split_bf16(do)
    out = {}
    for name, be, to in (("ref", RefOps(split=split), lambda t: t.clone()), ("cuda", CudaOps(split=split), lambda t: t.to(dev))):
        z = lambda *s, dt=torch.float32: to(torch.zeros(*s, dtype=dt))  # noqa: E731
        bz = lambda *s: z(*s, dt=torch.bfloat16)  # noqa: E731
        a = dict(qh=to(qh), ql=to(ql), kl=None if klens is None else to(klens), o_h=bz(B * T, H * DK), o_l=bz(B * T, H * DK),
                 lse=z(B, H, Tl), p_h=bz(B, H, T, Tp), p_l=bz(B, H, T, Tp), dh=to(dh), dl=to(dl), dvec=z(B, H, Tl),
                 dq_h=bz(B * T, 3 * H * DK), dq_l=bz(B * T, 3 * H * DK), dp=z(B, H, T, Tp), ds_h=bz(B, H, T, Tp), ds_l=bz(B, H, T, Tp))
        lo = (lambda t: t) if split == 3 else (lambda t: None)
        be.attn_fwd(a["qh"], lo(a["ql"]), a["kl"], B, H, T, DK, a["o_h"], lo(a["o_l"]), a["lse"],
                    a["p_h"] if emit else None, lo(a["p_l"]) if emit else None, Tp)
        be.attn_bwd(L.ATTN_PREP | L.ATTN_DQ | L.ATTN_DKV, a["qh"], lo(a["ql"]), a["kl"], B, H, T, DK, a["o_h"], lo(a["o_l"]), a["lse"],
                    a["dh"], lo(a["dl"]), a["dvec"], a["dq_h"], lo(a["dq_l"]), a["dp"] if emit else None,
                    a["ds_h"] if emit else None, lo(a["ds_l"]) if emit else None, Tp)
        if name == "cuda":
            torch.cuda.synchronize()
        out[name] = {k: (v.cpu() if torch.is_tensor(v) else v) for k, v in a.items()}
    r, c = out["ref"], out["cuda"]
    val = (lambda d, k: d[k + "_h"].double() + d[k + "_l"].double()) if split == 3 else (lambda d, k: d[k + "_h"].double())
    tol = tol or (2e-5 if split == 3 else 2e-2)
    errs = {"o": _err(val(c, "o"), val(r, "o")), "lse": _err(c["lse"][..., :T], r["lse"][..., :T]),
            "dvec": _err(c["dvec"][..., :T], r["dvec"][..., :T])}
    dq_c, dq_r = val(c, "dq").reshape(B * T, 3, H * DK), val(r, "dq").reshape(B * T, 3, H * DK)
    for i, n in enumerate(("dq", "dk", "dv")):
        errs[n] = _err(dq_c[:, i], dq_r[:, i])
    if emit:
        errs["P"] = _err(val(c, "p"), val(r, "p"))
        errs["dS"] = _err(val(c, "ds"), val(r, "ds"))
        errs["dP"] = _err(c["dp"][..., :T], r["dp"][..., :T])
    print(f"attn B{B} H{H} T{T} split{split} emit{emit}: " + " ".join(f"{k} {v:.1e}" for k, v in errs.items()))
    # split=1: dvec is computed from the single-bf16 O of each side, which differ by ~1e-2 relative
    for k, v in errs.items():
        assert v < tol, f"{k}: {v:.3e}"
    return out


@pytest.mark.parametrize("B,H,T,lens", [
    (2, 2, 24, [24, 17]),             # one partial tile
    (2, 2, 40, [1, 40]),              # a sequence with ONE valid key (softmax over a single element)
    (1, 2, 8, [5]),                   # shortest aligned sequence
    (1, 2, 128, None),                # exactly one resident tile, no mask (encoder-sized)
    (2, 2, 200, [200, 131]),          # T % 32 != 0, ragged
    (4, 2, 864, [864, 700, 515, 300]),  # configs[1] decoder shape
    (1, 2, 1000, [1000]),             # max_seq_len
])
def test_attn_fused_matches_reference(cuda_device, B, H, T, lens):
    # one valid key: P = 1 and dS = P (dP - D) is the difference of two equal numbers, each rounded on its own (the reference gets an
    # exact 0): the cancellation noise, ~1e-6 of |dP|, times |q| shows up in dK at 3e-5 of the tensor's maximum
    _run(cuda_device, B, H, T, lens, split=3, emit=False, tol=5e-5 if (lens and min(lens) == 1) else None)


@pytest.mark.parametrize("T,lens", [(72, [72, 40]), (864, [864, 333])])
def test_attn_fused_emits_tape_tensors(cuda_device, T, lens):
    _run(cuda_device, 2, 2, T, lens, split=3, emit=True, seed=3)


def test_attn_fused_single_pass_bf16(cuda_device):
    _run(cuda_device, 2, 2, 200, [200, 150], split=1, emit=True, seed=5)


def test_attn_fused_vs_autograd_fp64(cuda_device):
    """Independent check of the whole op against torch autograd in float64 (not the hand-derived backward of RefOps)."""
    B, H, T = 2, 2, 160
    qkv, do, klens = _inputs(B, H, T, 11, [160, 99])
    out = _run(cuda_device, B, H, T, [160, 99], split=3, emit=False, seed=11)["cuda"]
    qh, ql = split_bf16(qkv)
    dh, dl = split_bf16(do)
    x = (qh.double() + ql.double()).reshape(B, T, 3, H, DK).requires_grad_(True)
    q, k, v = (x[:, :, i].permute(0, 2, 1, 3) for i in range(3))
    s = q @ k.transpose(-1, -2) / math.sqrt(DK)
    km = (torch.arange(T)[None, :] < klens[:, None])[:, None, None, :]
    o = (torch.softmax(s.masked_fill(~km, -math.inf), -1) @ v).permute(0, 2, 1, 3).reshape(B * T, H * DK)
    (o * (dh.double() + dl.double())).sum().backward()
    got_o = out["o_h"].double() + out["o_l"].double()
    got_g = (out["dq_h"].double() + out["dq_l"].double()).reshape(B, T, 3, H, DK)
    assert _err(got_o, o.detach()) < 2e-5
    # dvec was formed from the kernel's own (hi+lo rounded) O: the gradients agree to that rounding
    for i in range(3):
        assert _err(got_g[:, :, i], x.grad[:, :, i]) < 5e-5


def test_attn_bwd_parts_are_independent(cuda_device):
    """PREP, DQ and DKV may be issued as separate calls (the engine runs DQ and DKV on concurrent streams)."""
    B, H, T = 2, 2, 96
    qkv, do, klens = _inputs(B, H, T, 7, [96, 50])
    dev = cuda_device
    be = CudaOps(split=3)
    qh, ql = (t.to(dev) for t in split_bf16(qkv))
    dh, dl = (t.to(dev) for t in split_bf16(do))
    kl = klens.to(dev)
    Tl = 128
    bz = lambda *s: torch.zeros(*s, dtype=torch.bfloat16, device=dev)  # noqa: E731
    o_h, o_l, lse, dvec = bz(B * T, H * DK), bz(B * T, H * DK), torch.zeros(B, H, Tl, device=dev), torch.zeros(B, H, Tl, device=dev)
    be.attn_fwd(qh, ql, kl, B, H, T, DK, o_h, o_l, lse)
    res = []
    for split_calls in (False, True):
        g_h, g_l = bz(B * T, 3 * H * DK), bz(B * T, 3 * H * DK)
        args = (qh, ql, kl, B, H, T, DK, o_h, o_l, lse, dh, dl, dvec, g_h, g_l)
        if split_calls:
            be.attn_bwd(L.ATTN_PREP, *args)
            be.attn_bwd(L.ATTN_DKV, *args)
            be.attn_bwd(L.ATTN_DQ, *args)
        else:
            be.attn_bwd(L.ATTN_PREP | L.ATTN_DQ | L.ATTN_DKV, *args)
        torch.cuda.synchronize()
        res.append((g_h.clone(), g_l.clone()))
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])     # deterministic: no atomics anywhere
