"""Micro-benchmark of mtts_gemm on the hot shapes of BASELINE config 2 (B=4, T=864): CUDA-event time of
back-to-back launches (L2-warm for weights; activations 14-28 MB), per (block_n, pair, ksplit) config."""
import math
import sys

import torch

sys.path.insert(0, ".")
from meta_tts_b200 import lib as L  # noqa: E402
from meta_tts_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
PEAK = 1444.4


def bf(*s):
    return torch.randn(*s, device=dev).to(torch.bfloat16)


def timeit(fn, iters=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3   # us


def report(name, flops, us, split):
    tf = flops / us / 1e6
    print(f"{name:58s} {us:8.1f} us  {tf:7.1f} TF/s alg  ({100 * tf * (3 if split == 3 else 1) / PEAK:5.1f}% of peak as issued MMA)")


def conv_fwd(split, B=4, T=864, cin=256, cout=1024, k=9):
    p = (k - 1) // 2
    x, xl = bf(B, T, cin), bf(B, T, cin)
    w, wl = bf(k, cout, cin), bf(k, cout, cin)
    oh, ol = torch.empty(B, T, cout, device=dev, dtype=torch.bfloat16), torch.empty(B, T, cout, device=dev, dtype=torch.bfloat16)
    flops = 2.0 * B * T * cout * cin * k
    for bn, pair in ((128, False), (256, False), (128, True), (256, True)):
        def fn():
            ops.gemm(ops.Opnd(x, xl if split == 3 else None, L.MAJOR_K, (cin, T, B), (1, cin, T * cin), src2=L.SRC_Z0,
                              shift_src=L.SRC_TAP, shift_base=-p, shift_step=1),
                     ops.Opnd(w, wl if split == 3 else None, L.MAJOR_K, (cin, cout, k), (1, cin, cout * cin), src2=L.SRC_TAP),
                     T, cout, cin, c_hi=oh, c_lo=ol if split == 3 else None, ldc=cout, c_sz0=T * cout, ntaps=k, nz0=B,
                     split=split, block_n=bn, pair=pair, flags=L.EPI_RELU)
        report(f"conv{k} fwd {cin}->{cout} split={split} bn={bn} pair={pair}", flops, timeit(fn), split)


def conv_dgrad(split, B=4, T=864, cin=256, cout=1024, k=9):
    p = (k - 1) // 2
    dy, dyl = bf(B, T, cout), bf(B, T, cout)
    w, wl = bf(k, cout, cin), bf(k, cout, cin)
    out = torch.zeros(B, T, cin, device=dev)
    flops = 2.0 * B * T * cout * cin * k
    for bn, pair, ks in ((128, False, 1), (64, False, 1), (256, True, 1), (256, True, 4), (128, True, 2), (256, False, 4)):
        def fn():
            ops.gemm(ops.Opnd(dy, dyl if split == 3 else None, L.MAJOR_K, (cout, T, B), (1, cout, T * cout), src2=L.SRC_Z0,
                              shift_src=L.SRC_TAP, shift_base=p, shift_step=-1),
                     ops.Opnd(w, wl if split == 3 else None, L.MAJOR_MN, (cin, cout, k), (1, cin, cout * cin), src2=L.SRC_TAP),
                     T, cin, cout, c_f32=out, ldc=cin, c_sz0=T * cin, ntaps=k, nz0=B, split=split, block_n=bn, pair=pair,
                     ksplit=ks, flags=L.EPI_ACCUM if ks > 1 else L.EPI_ADD_C)
        report(f"conv{k} dgrad {cout}->{cin} split={split} bn={bn} pair={pair} ks={ks}", flops, timeit(fn), split)


def conv_wgrad(split, B=4, T=864, cin=256, cout=1024, k=9):
    p = (k - 1) // 2
    dy, dyl = bf(B, T, cout), bf(B, T, cout)
    x, xl = bf(B, T, cin), bf(B, T, cin)
    dw = torch.zeros(k, cout, cin, device=dev)
    flops = 2.0 * B * T * cout * cin * k
    for bn, pair, ks in ((128, False, 1), (128, True, 1), (256, True, 1), (256, True, 2)):
        def fn():
            ops.gemm(ops.Opnd(dy, dyl if split == 3 else None, L.MAJOR_MN, (cout, T, B), (1, cout, T * cout), src2=L.SRC_KB),
                     ops.Opnd(x, xl if split == 3 else None, L.MAJOR_MN, (cin, T, B), (1, cin, T * cin), src2=L.SRC_KB,
                              shift_src=L.SRC_Z0, shift_base=-p, shift_step=1),
                     cout, cin, T, c_f32=dw, ldc=cin, c_sz0=cout * cin, nkb=B, nz0=k, split=split, flags=L.EPI_ACCUM,
                     ksplit=ks, block_n=bn, pair=pair)
        report(f"conv{k} wgrad split={split} bn={bn} pair={pair} ks={ks}", flops, timeit(fn), split)


def attn(split, B=4, H=2, T=864, dk=128):
    row = 3 * H * dk
    q, ql = bf(B * T, row), bf(B * T, row)
    S = torch.empty(B, H, T, T, device=dev)
    P, Pl = bf(B, H, T, T), bf(B, H, T, T)
    oh, ol = torch.empty(B * T, H * dk, device=dev, dtype=torch.bfloat16), torch.empty(B * T, H * dk, device=dev, dtype=torch.bfloat16)
    flops = 2.0 * B * H * T * T * dk
    lo = (lambda t: t) if split == 3 else (lambda t: None)
    for bn, pair in ((128, False), (256, False), (256, True)):
        def fn():
            ops.gemm(ops.Opnd(q, lo(ql), L.MAJOR_K, (dk, T, H, B), (1, row, dk, T * row), src2=L.SRC_Z0, src3=L.SRC_Z1),
                     ops.Opnd(q, lo(ql), L.MAJOR_K, (dk, T, H, B), (1, row, dk, T * row), src2=L.SRC_Z0, src3=L.SRC_Z1, offset=H * dk),
                     T, T, dk, c_f32=S, ldc=T, c_sz0=T * T, c_sz1=H * T * T, nz0=H, nz1=B, split=split, block_n=bn, pair=pair)
        report(f"attn QK^T split={split} bn={bn} pair={pair}", flops, timeit(fn), split)
    for bn, pair in ((64, False), (128, False), (128, True)):
        def fn():
            ops.gemm(ops.Opnd(P, lo(Pl), L.MAJOR_K, (T, T, H, B), (1, T, T * T, H * T * T), src2=L.SRC_Z0, src3=L.SRC_Z1),
                     ops.Opnd(q, lo(ql), L.MAJOR_MN, (dk, T, H, B), (1, row, dk, T * row), src2=L.SRC_Z0, src3=L.SRC_Z1, offset=2 * H * dk),
                     T, dk, T, c_hi=oh, c_lo=lo(ol), ldc=H * dk, c_sz0=dk, c_sz1=T * H * dk, nz0=H, nz1=B, split=split, block_n=bn,
                     pair=pair)
        report(f"attn PV split={split} bn={bn} pair={pair}", flops, timeit(fn), split)


def linear(split, M=3456, N=768, K=256):
    x, xl = bf(M, K), bf(M, K)
    w, wl = bf(N, K), bf(N, K)
    oh, ol = torch.empty(M, N, device=dev, dtype=torch.bfloat16), torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    lo = (lambda t: t) if split == 3 else (lambda t: None)
    flops = 2.0 * M * N * K
    for bn, pair in ((64, False), (128, False), (128, True), (256, True)):
        if bn > N:
            continue
        def fn():
            ops.gemm(ops.Opnd(x, lo(xl), L.MAJOR_K, (K, M), (1, K)), ops.Opnd(w, lo(wl), L.MAJOR_K, (K, N), (1, K)), M, N, K,
                     c_hi=oh, c_lo=lo(ol), ldc=N, split=split, block_n=bn, pair=pair)
        report(f"linear {M}x{N}x{K} split={split} bn={bn} pair={pair}", flops, timeit(fn), split)


if __name__ == "__main__":
    for split in (3, 1):
        conv_fwd(split)
        conv_dgrad(split)
        conv_wgrad(split)
        attn(split)
        linear(split, N=768, K=256)
        linear(split, N=256, K=256)
        linear(split, N=256, K=1024)
        conv_fwd(split, cin=512, cout=512, k=5)
    # empty-kernel launch floor
    t = torch.empty(256, device=dev)
    print("axpby floor us", timeit(lambda: ops.CudaOps(3).axpby(1.0, t, 1.0, t), 50))
