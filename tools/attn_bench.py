"""Fused attention kernels vs the three-launch chain they replace, at the decoder shape of BASELINE configs[1]
(B = 4 utterances, 2 heads, 864 frames, d_k = 128): CUDA-event timing of graph replays, cold-ish L2 (rotating buffer sets).

    python tools/attn_bench.py [B H T]
"""
import math
import sys

import torch

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from meta_tts_b200 import lib as L  # noqa: E402
from meta_tts_b200.engine import BMat, Gemm  # noqa: E402
from meta_tts_b200.ops import CudaOps, split_bf16  # noqa: E402

B, H, T = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (4, 2, 864)
DK = 128
d = H * DK
dev = torch.device("cuda:0")
be = CudaOps(split=3)
g = Gemm(be)
Tp, Tl = (T + 7) // 8 * 8, (T + 127) // 128 * 128
NSETS = 6


def mk():
    qkv = torch.randn(B * T, 3 * d, device=dev)
    do = torch.randn(B * T, d, device=dev)
    bz = lambda *s: torch.zeros(*s, dtype=torch.bfloat16, device=dev)  # noqa: E731
    z = lambda *s: torch.zeros(*s, device=dev)  # noqa: E731
    qh, ql = split_bf16(qkv)
    dh, dl = split_bf16(do)
    return dict(qh=qh, ql=ql, dh=dh, dl=dl, o_h=bz(B * T, d), o_l=bz(B * T, d), lse=z(B, H, Tl), dvec=z(B, H, Tl), S=z(B, H, T, Tp),
                p_h=bz(B, H, T, Tp), p_l=bz(B, H, T, Tp), dP=z(B, H, T, Tp), ds_h=bz(B, H, T, Tp), ds_l=bz(B, H, T, Tp),
                g_h=bz(B * T, 3 * d), g_l=bz(B * T, 3 * d))


sets = [mk() for _ in range(NSETS)]
klens = torch.full((B,), T, dtype=torch.int64, device=dev)
row = 3 * d
qm = lambda hi, lo, w: BMat(hi, lo, w * d, T * row, DK, row, T, DK)  # noqa: E731
pm = lambda hi, lo, f=None: BMat(hi, lo, 0, H * T * Tp, T * Tp, Tp, T, T, f)  # noqa: E731
om = lambda hi, lo: BMat(hi, lo, 0, T * d, DK, d, T, DK)  # noqa: E731
sc = 1.0 / math.sqrt(DK)


def fused_fwd(a, emit):
    be.attn_fwd(a["qh"], a["ql"], klens, B, H, T, DK, a["o_h"], a["o_l"], a["lse"], a["p_h"] if emit else None, a["p_l"] if emit else None, Tp)


def chain_fwd(a):
    g.bmm(qm(a["qh"], a["ql"], 0), False, qm(a["qh"], a["ql"], 1), False, pm(None, None, a["S"]), B, H, alpha=sc)
    be.softmax(0, a["S"], None, None, None, None, None, klens, B * H, H, T, T, Tp, a["p_h"], a["p_l"])
    g.bmm(pm(a["p_h"], a["p_l"]), False, qm(a["qh"], a["ql"], 2), True, om(a["o_h"], a["o_l"]), B, H)


def fused_bwd(a, parts, emit):
    be.attn_bwd(parts, a["qh"], a["ql"], klens, B, H, T, DK, a["o_h"], a["o_l"], a["lse"], a["dh"], a["dl"], a["dvec"], a["g_h"], a["g_l"],
                a["dP"] if emit else None, a["ds_h"] if emit else None, a["ds_l"] if emit else None, Tp)


def chain_bwd(a):
    g.bmm(pm(a["p_h"], a["p_l"]), True, om(a["dh"], a["dl"]), True, qm(a["g_h"], a["g_l"], 2), B, H)
    g.bmm(om(a["dh"], a["dl"]), False, qm(a["qh"], a["ql"], 2), False, pm(None, None, a["dP"]), B, H)
    be.softmax(1, a["dP"], None, a["p_h"], a["p_l"], None, None, klens, B * H, H, T, T, Tp, a["ds_h"], a["ds_l"])
    g.bmm(pm(a["ds_h"], a["ds_l"]), False, qm(a["qh"], a["ql"], 1), True, qm(a["g_h"], a["g_l"], 0), B, H, alpha=sc)
    g.bmm(pm(a["ds_h"], a["ds_l"]), True, qm(a["qh"], a["ql"], 0), True, qm(a["g_h"], a["g_l"], 1), B, H, alpha=sc)


def timeit(name, fn, flops):
    for a in sets:
        fn(a)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for a in sets:
            fn(a)
    for _ in range(3):
        gr.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for _ in range(reps):
        gr.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (reps * NSETS)
    print(f"{name:44s} {us:8.1f} us   {flops / us / 1e6:7.1f} algorithmic TFLOP/s")
    return us


f_fwd = 4.0 * B * H * T * T * DK
print(f"B {B} H {H} T {T} dk {DK}  (bf16x3)")
timeit("chain  fwd: scores + softmax + PV", chain_fwd, f_fwd)
timeit("fused  fwd", lambda a: fused_fwd(a, False), f_fwd)
timeit("fused  fwd + emit P", lambda a: fused_fwd(a, True), f_fwd)
timeit("chain  bwd: dV, dP, softmax', dQ, dK (serial)", chain_bwd, 2 * f_fwd)
timeit("fused  bwd: prep + dQ + dK + dV (serial)", lambda a: fused_bwd(a, 15, False), 2 * f_fwd)
timeit("fused  bwd: dQ only", lambda a: fused_bwd(a, L.ATTN_DQ, False), 0.75 * f_fwd)
timeit("fused  bwd: dQ only + emit dP, dS", lambda a: fused_bwd(a, L.ATTN_DQ, True), 0.75 * f_fwd)
timeit("fused  bwd: dK only", lambda a: fused_bwd(a, L.ATTN_DK, False), 0.75 * f_fwd)
timeit("fused  bwd: dV only", lambda a: fused_bwd(a, L.ATTN_DV, False), 0.5 * f_fwd)
timeit("fused  bwd: prep only", lambda a: fused_bwd(a, L.ATTN_PREP, False), 0.0)
