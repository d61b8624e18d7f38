"""GPU parity tests of the generic tcgen05 GEMM (mtts_gemm) through the C ABI.

Checker: float64 torch matmul / conv restated in-line on the same inputs (for split=1 the inputs
are the bf16-rounded values, so the only difference is fp32 accumulation order; for split=3 the
inputs are fp32 and the kernel must be fp32-grade).  Tolerances are written per case.
"""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

from meta_tts_b200 import lib as L  # noqa: E402
from meta_tts_b200 import ops  # noqa: E402


def _rel(a, b):
    a = a.double()
    b = b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _prep(x, split):
    """fp32 -> (hi, lo, value the kernel effectively sees)."""
    hi, lo = ops.split_bf16(x)
    if split == 1:
        return hi, None, hi.double()
    return hi, lo, x.double()


TOL = {1: 2e-5, 3: 5e-5}   # relative Frobenius error vs float64 reference (bf16x3: ~2^-17 per operand, grows ~sqrt(K))


def _check(name, got, ref, split):
    r = _rel(got, ref)
    print(f"[gemm] {name:32s} split={split} rel_err={r:.3e}")
    assert math.isfinite(r) and r < TOL[split], f"{name}: rel err {r}"


@pytest.mark.parametrize("split", [1, 3])
@pytest.mark.parametrize("block_n", [64, 128, 256])
def test_linear_fwd(cuda_device, split, block_n):
    torch.manual_seed(0)
    M, N, K = 300, 256, 320
    X = torch.randn(M, K, device=cuda_device)
    W = torch.randn(N, K, device=cuda_device) / math.sqrt(K)
    bias = torch.randn(N, device=cuda_device)
    xh, xl, xr = _prep(X, split)
    wh, wl, wr = _prep(W, split)
    out = torch.empty(M, N, device=cuda_device)
    oh = torch.empty(M, N, device=cuda_device, dtype=torch.bfloat16)
    ol = torch.empty(M, N, device=cuda_device, dtype=torch.bfloat16)
    ops.gemm(ops.Opnd(xh, xl, L.MAJOR_K, (K, M), (1, K)), ops.Opnd(wh, wl, L.MAJOR_K, (K, N), (1, K)),
             M, N, K, c_f32=out, c_hi=oh, c_lo=ol, ldc=N, bias=bias, flags=L.EPI_RELU, split=split,
             block_n=block_n)
    ref = torch.relu(xr @ wr.t() + bias.double())
    _check(f"linear_fwd bn={block_n}", out, ref, split)
    _check(f"linear_fwd hi+lo bn={block_n}", oh.float() + ol.float(), out, 3)
    assert _rel(oh.float(), out) < 4e-3


@pytest.mark.parametrize("split", [1, 3])
def test_linear_dgrad_mn_major_b(cuda_device, split):
    torch.manual_seed(1)
    M, N, K = 260, 256, 192          # dX[M,K] = dY[M,N] W[N,K]
    dY = torch.randn(M, N, device=cuda_device)
    W = torch.randn(N, K, device=cuda_device) / math.sqrt(N)
    yh, yl, yr = _prep(dY, split)
    wh, wl, wr = _prep(W, split)
    out = torch.empty(M, K, device=cuda_device)
    ops.gemm(ops.Opnd(yh, yl, L.MAJOR_K, (N, M), (1, N)), ops.Opnd(wh, wl, L.MAJOR_MN, (K, N), (1, K)),
             M, K, N, c_f32=out, ldc=K, split=split)
    _check("linear_dgrad (B MN-major)", out, yr @ wr, split)


@pytest.mark.parametrize("split", [1, 3])
@pytest.mark.parametrize("ksplit", [1, 4])
def test_linear_wgrad_mn_major_ab(cuda_device, split, ksplit):
    torch.manual_seed(2)
    T, N, K = 700, 256, 192          # dW[N,K] = dY[T,N]^T X[T,K]
    dY = torch.randn(T, N, device=cuda_device)
    X = torch.randn(T, K, device=cuda_device)
    yh, yl, yr = _prep(dY, split)
    xh, xl, xr = _prep(X, split)
    out = torch.zeros(N, K, device=cuda_device)
    ops.gemm(ops.Opnd(yh, yl, L.MAJOR_MN, (N, T), (1, N)), ops.Opnd(xh, xl, L.MAJOR_MN, (K, T), (1, K)),
             N, K, T, c_f32=out, ldc=K, split=split, ksplit=ksplit, flags=L.EPI_ACCUM, alpha=0.5)
    _check(f"linear_wgrad ksplit={ksplit}", out, 0.5 * (yr.t() @ xr), split)


def _conv_ref(x, w_kio, bias, pad):
    # x [B,T,Cin], w [k, Cout, Cin]
    w = w_kio.permute(1, 2, 0).contiguous()   # [Cout, Cin, k]
    y = torch.nn.functional.conv1d(x.transpose(1, 2), w, bias, padding=pad)
    return y.transpose(1, 2).contiguous()


@pytest.mark.parametrize("split", [1, 3])
@pytest.mark.parametrize("k,cin,cout,T", [(9, 256, 256, 200), (5, 80, 512, 150), (3, 256, 256, 37)])
def test_conv_fwd_dgrad_wgrad(cuda_device, split, k, cin, cout, T):
    torch.manual_seed(3)
    B, p = 3, (k - 1) // 2
    X = torch.randn(B, T, cin, device=cuda_device)
    W = torch.randn(k, cout, cin, device=cuda_device) / math.sqrt(cin * k)
    bias = torch.randn(cout, device=cuda_device)
    xh, xl, xr = _prep(X, split)
    wh, wl, wr = _prep(W, split)
    # forward
    Y = torch.empty(B, T, cout, device=cuda_device)
    ops.gemm(ops.Opnd(xh, xl, L.MAJOR_K, (cin, T, B), (1, cin, T * cin), src2=L.SRC_Z0,
                      shift_src=L.SRC_TAP, shift_base=-p, shift_step=1),
             ops.Opnd(wh, wl, L.MAJOR_K, (cin, cout, k), (1, cin, cout * cin), src2=L.SRC_TAP),
             T, cout, cin, c_f32=Y, ldc=cout, c_sz0=T * cout, bias=bias, ntaps=k, nz0=B, split=split)
    ref = _conv_ref(xr, wr, bias.double(), p)
    _check(f"conv_fwd k={k} {cin}->{cout}", Y, ref, split)
    # dgrad: dX[b,t] = sum_j dY[b,t-j+p] W_j
    dY = torch.randn(B, T, cout, device=cuda_device)
    yh, yl, yr = _prep(dY, split)
    dX = torch.empty(B, T, cin, device=cuda_device)
    ops.gemm(ops.Opnd(yh, yl, L.MAJOR_K, (cout, T, B), (1, cout, T * cout), src2=L.SRC_Z0,
                      shift_src=L.SRC_TAP, shift_base=p, shift_step=-1),
             ops.Opnd(wh, wl, L.MAJOR_MN, (cin, cout, k), (1, cin, cout * cin), src2=L.SRC_TAP),
             T, cin, cout, c_f32=dX, ldc=cin, c_sz0=T * cin, ntaps=k, nz0=B, split=split)
    xr_ = xr.clone().requires_grad_(True)
    wr_ = wr.clone().requires_grad_(True)
    _conv_ref(xr_, wr_, bias.double(), p).backward(yr)
    _check(f"conv_dgrad k={k}", dX, xr_.grad, split)
    # wgrad: dW[j] = sum_{b,t} dY[b,t]^T X[b,t+j-p]
    dW = torch.zeros(k, cout, cin, device=cuda_device)
    ops.gemm(ops.Opnd(yh, yl, L.MAJOR_MN, (cout, T, B), (1, cout, T * cout), src2=L.SRC_KB),
             ops.Opnd(xh, xl, L.MAJOR_MN, (cin, T, B), (1, cin, T * cin), src2=L.SRC_KB,
                      shift_src=L.SRC_Z0, shift_base=-p, shift_step=1),
             cout, cin, T, c_f32=dW, ldc=cin, c_sz0=cout * cin, nkb=B, nz0=k, split=split,
             flags=L.EPI_ACCUM, ksplit=2)
    _check(f"conv_wgrad k={k}", dW, wr_.grad, split)


@pytest.mark.parametrize("split", [1, 3])
@pytest.mark.parametrize("Lq", [128, 200])
def test_attention_products(cuda_device, split, Lq):
    torch.manual_seed(4)
    B, H, dk = 2, 2, 128
    row = 3 * H * dk
    QKV = torch.randn(B * Lq, row, device=cuda_device)
    qh, ql, qr = _prep(QKV, split)
    Lp = (Lq + 7) // 8 * 8
    S = torch.zeros(B, H, Lq, Lp, device=cuda_device)
    scale = 1.0 / math.sqrt(dk)
    ops.gemm(ops.Opnd(qh, ql, L.MAJOR_K, (dk, Lq, H, B), (1, row, dk, Lq * row), src2=L.SRC_Z0, src3=L.SRC_Z1),
             ops.Opnd(qh, ql, L.MAJOR_K, (dk, Lq, H, B), (1, row, dk, Lq * row), src2=L.SRC_Z0, src3=L.SRC_Z1,
                      offset=H * dk),
             Lq, Lq, dk, c_f32=S, ldc=Lp, c_sz0=Lq * Lp, c_sz1=H * Lq * Lp, alpha=scale, nz0=H, nz1=B,
             split=split)
    q4 = qr.view(B, Lq, 3, H, dk)
    Sref = torch.einsum("blhd,bmhd->bhlm", q4[:, :, 0], q4[:, :, 1]) * scale
    _check(f"attn QK^T L={Lq}", S[..., :Lq], Sref, split)
    # P V with V MN-major
    P = torch.softmax(Sref.float(), dim=-1)
    Pp = torch.zeros(B, H, Lq, Lp, device=cuda_device)
    Pp[..., :Lq] = P
    ph, pl, pr = _prep(Pp, split)
    O = torch.empty(B * Lq, H * dk, device=cuda_device)
    ops.gemm(ops.Opnd(ph, pl, L.MAJOR_K, (Lq, Lq, H, B), (1, Lp, Lq * Lp, H * Lq * Lp), src2=L.SRC_Z0, src3=L.SRC_Z1),
             ops.Opnd(qh, ql, L.MAJOR_MN, (dk, Lq, H, B), (1, row, dk, Lq * row), src2=L.SRC_Z0, src3=L.SRC_Z1,
                      offset=2 * H * dk),
             Lq, dk, Lq, c_f32=O, ldc=H * dk, c_sz0=dk, c_sz1=Lq * H * dk, nz0=H, nz1=B, split=split)
    Oref = torch.einsum("bhlm,bmhd->blhd", pr[..., :Lq], q4[:, :, 2]).reshape(B * Lq, H * dk)
    _check(f"attn PV L={Lq}", O, Oref, split)
    # dV = P^T dO  (A MN-major from P, B MN-major from dO) -> [B, L, H, dk] layout
    dO = torch.randn(B * Lq, H * dk, device=cuda_device)
    dh, dl, dr = _prep(dO, split)
    dV = torch.empty(B * Lq, H * dk, device=cuda_device)
    ops.gemm(ops.Opnd(ph, pl, L.MAJOR_MN, (Lq, Lq, H, B), (1, Lp, Lq * Lp, H * Lq * Lp), src2=L.SRC_Z0, src3=L.SRC_Z1),
             ops.Opnd(dh, dl, L.MAJOR_MN, (dk, Lq, H, B), (1, H * dk, dk, Lq * H * dk), src2=L.SRC_Z0, src3=L.SRC_Z1),
             Lq, dk, Lq, c_f32=dV, ldc=H * dk, c_sz0=dk, c_sz1=Lq * H * dk, nz0=H, nz1=B, split=split)
    dVref = torch.einsum("bhlm,blhd->bmhd", pr[..., :Lq], dr.view(B, Lq, H, dk)).reshape(B * Lq, H * dk)
    _check(f"attn dV L={Lq}", dV, dVref, split)


@pytest.mark.parametrize("split", [1, 3])
def test_odd_sizes_and_gate(cuda_device, split):
    torch.manual_seed(5)
    M, N, K = 40, 80, 80
    X = torch.randn(M, K, device=cuda_device)
    W = torch.randn(N, K, device=cuda_device)
    G = torch.randn(M, N, device=cuda_device)
    xh, xl, xr = _prep(X, split)
    wh, wl, wr = _prep(W, split)
    out = torch.empty(M, N, device=cuda_device)
    ops.gemm(ops.Opnd(xh, xl, L.MAJOR_K, (K, M), (1, K)), ops.Opnd(wh, wl, L.MAJOR_K, (K, N), (1, K)),
             M, N, K, c_f32=out, ldc=N, split=split, gate=G.to(torch.bfloat16), flags=L.EPI_GATE)
    ref = (xr @ wr.t()) * (G.to(torch.bfloat16).double() > 0)
    _check("odd sizes + gate", out, ref, split)


def test_large_decoder_shapes(cuda_device):
    """BASELINE config-2 decoder conv k=9 shape (B*T = 4*864 tokens, 256 -> 1024)."""
    torch.manual_seed(6)
    B, T, cin, cout, k, p = 4, 864, 256, 1024, 9, 4
    X = torch.randn(B, T, cin, device=cuda_device)
    W = torch.randn(k, cout, cin, device=cuda_device) / math.sqrt(cin * k)
    bias = torch.randn(cout, device=cuda_device)
    for split in (1, 3):
        xh, xl, xr = _prep(X, split)
        wh, wl, wr = _prep(W, split)
        Hh = torch.empty(B, T, cout, device=cuda_device, dtype=torch.bfloat16)
        Hl = torch.empty(B, T, cout, device=cuda_device, dtype=torch.bfloat16)
        ops.gemm(ops.Opnd(xh, xl, L.MAJOR_K, (cin, T, B), (1, cin, T * cin), src2=L.SRC_Z0,
                          shift_src=L.SRC_TAP, shift_base=-p, shift_step=1),
                 ops.Opnd(wh, wl, L.MAJOR_K, (cin, cout, k), (1, cin, cout * cin), src2=L.SRC_TAP),
                 T, cout, cin, c_hi=Hh, c_lo=Hl, ldc=cout, c_sz0=T * cout, bias=bias, ntaps=k, nz0=B,
                 split=split, flags=L.EPI_RELU)
        ref = torch.relu(_conv_ref(xr, wr, bias.double(), p))
        _check("decoder conv9 (hi+lo out)", Hh.float() + Hl.float(), ref, split)


@pytest.mark.parametrize("split", [1, 3])
@pytest.mark.parametrize("block_n", [64, 128, 256])
def test_two_term_conv_and_bn256(cuda_device, split, block_n):
    """C = conv(x1, W1) + conv(x2, W2) + bias in ONE launch (tangent passes), all tile widths incl. BN=256 in bf16x3."""
    torch.manual_seed(7)
    B, T, cin, cout, k, p = 2, 300, 256, 512, 3, 1
    X1, X2 = torch.randn(B, T, cin, device=cuda_device), torch.randn(B, T, cin, device=cuda_device)
    W1, W2 = (torch.randn(k, cout, cin, device=cuda_device) / 30 for _ in range(2))
    bias = torch.randn(cout, device=cuda_device)
    (x1h, x1l, x1r), (x2h, x2l, x2r) = _prep(X1, split), _prep(X2, split)
    (w1h, w1l, w1r), (w2h, w2l, w2r) = _prep(W1, split), _prep(W2, split)
    Y = torch.empty(B, T, cout, device=cuda_device)
    ops.gemm(ops.Opnd(x1h, x1l, L.MAJOR_K, (cin, T, B), (1, cin, T * cin), src2=L.SRC_Z0, shift_src=L.SRC_TAP,
                      shift_base=-p, shift_step=1, hi2=x2h, lo2=x2l),
             ops.Opnd(w1h, w1l, L.MAJOR_K, (cin, cout, k), (1, cin, cout * cin), src2=L.SRC_TAP, hi2=w2h, lo2=w2l),
             T, cout, cin, c_f32=Y, ldc=cout, c_sz0=T * cout, bias=bias, ntaps=k, nz0=B, split=split, block_n=block_n)
    ref = _conv_ref(x1r, w1r, bias.double(), p) + _conv_ref(x2r, w2r, None, p)
    _check(f"two-term conv bn={block_n}", Y, ref, split)


# ---------------------------------------------------------------------------------------------
# 2-CTA kernel (tcgen05 cta_group::2, one 256 x BN tile per CTA pair)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("split", [1, 3])
@pytest.mark.parametrize("block_n", [128, 256])
@pytest.mark.parametrize("M", [300, 128, 700])          # odd / even m-tile counts (dummy partner CTA for odd tails)
def test_pair_linear_all_majors(cuda_device, split, block_n, M):
    torch.manual_seed(11)
    N, K = 384, 320
    X = torch.randn(M, K, device=cuda_device)
    W = torch.randn(N, K, device=cuda_device) / math.sqrt(K)
    bias = torch.randn(N, device=cuda_device)
    xh, xl, xr = _prep(X, split)
    wh, wl, wr = _prep(W, split)
    out = torch.empty(M, N, device=cuda_device)
    oh = torch.empty(M, N, device=cuda_device, dtype=torch.bfloat16)
    ol = torch.empty(M, N, device=cuda_device, dtype=torch.bfloat16)
    ops.gemm(ops.Opnd(xh, xl, L.MAJOR_K, (K, M), (1, K)), ops.Opnd(wh, wl, L.MAJOR_K, (K, N), (1, K)),
             M, N, K, c_f32=out, c_hi=oh, c_lo=ol, ldc=N, bias=bias, flags=L.EPI_RELU, split=split, block_n=block_n, pair=True)
    _check(f"pair linear_fwd bn={block_n} M={M}", out, torch.relu(xr @ wr.t() + bias.double()), split)
    _check("pair hi+lo", oh.float() + ol.float(), out, 3)
    # dgrad (B MN-major): dX[M,K] = dY[M,N] W[N,K]
    dY = torch.randn(M, N, device=cuda_device)
    yh, yl, yr = _prep(dY, split)
    dX = torch.empty(M, K, device=cuda_device)
    ops.gemm(ops.Opnd(yh, yl, L.MAJOR_K, (N, M), (1, N)), ops.Opnd(wh, wl, L.MAJOR_MN, (K, N), (1, K)),
             M, K, N, c_f32=dX, ldc=K, split=split, block_n=block_n, pair=True)
    _check(f"pair dgrad (B MN-major) bn={block_n}", dX, yr @ wr, split)
    # wgrad (A and B MN-major, split-K accumulate): dW[N,K] = dY^T X
    dW = torch.zeros(N, K, device=cuda_device)
    ops.gemm(ops.Opnd(yh, yl, L.MAJOR_MN, (N, M), (1, N)), ops.Opnd(xh, xl, L.MAJOR_MN, (K, M), (1, K)),
             N, K, M, c_f32=dW, ldc=K, split=split, ksplit=2, flags=L.EPI_ACCUM, block_n=block_n, pair=True)
    _check(f"pair wgrad (A,B MN-major) bn={block_n}", dW, yr.t() @ xr, split)


@pytest.mark.parametrize("split", [1, 3])
def test_pair_conv_taps_and_two_terms(cuda_device, split):
    torch.manual_seed(12)
    B, T, cin, cout, k, p = 3, 864, 256, 1024, 9, 4
    X1, X2 = torch.randn(B, T, cin, device=cuda_device), torch.randn(B, T, cin, device=cuda_device)
    W1, W2 = (torch.randn(k, cout, cin, device=cuda_device) / 40 for _ in range(2))
    bias = torch.randn(cout, device=cuda_device)
    (x1h, x1l, x1r), (x2h, x2l, x2r) = _prep(X1, split), _prep(X2, split)
    (w1h, w1l, w1r), (w2h, w2l, w2r) = _prep(W1, split), _prep(W2, split)
    Y = torch.empty(B, T, cout, device=cuda_device)
    ops.gemm(ops.Opnd(x1h, x1l, L.MAJOR_K, (cin, T, B), (1, cin, T * cin), src2=L.SRC_Z0, shift_src=L.SRC_TAP,
                      shift_base=-p, shift_step=1, hi2=x2h, lo2=x2l),
             ops.Opnd(w1h, w1l, L.MAJOR_K, (cin, cout, k), (1, cin, cout * cin), src2=L.SRC_TAP, hi2=w2h, lo2=w2l),
             T, cout, cin, c_f32=Y, ldc=cout, c_sz0=T * cout, bias=bias, ntaps=k, nz0=B, split=split, block_n=256, pair=True)
    ref = _conv_ref(x1r, w1r, bias.double(), p) + _conv_ref(x2r, w2r, None, p)
    _check("pair two-term conv9 864x1024", Y, ref, split)
    # conv dgrad through MN-major weights, N_gemm = cin
    dY = torch.randn(B, T, cout, device=cuda_device)
    yh, yl, yr = _prep(dY, split)
    dX = torch.empty(B, T, cin, device=cuda_device)
    ops.gemm(ops.Opnd(yh, yl, L.MAJOR_K, (cout, T, B), (1, cout, T * cout), src2=L.SRC_Z0, shift_src=L.SRC_TAP,
                      shift_base=p, shift_step=-1),
             ops.Opnd(w1h, w1l, L.MAJOR_MN, (cin, cout, k), (1, cin, cout * cin), src2=L.SRC_TAP),
             T, cin, cout, c_f32=dX, ldc=cin, c_sz0=T * cin, ntaps=k, nz0=B, split=split, block_n=256, pair=True)
    xr_ = x1r.clone().requires_grad_(True)
    _conv_ref(xr_, w1r, None, p).backward(yr)
    _check("pair conv9 dgrad", dX, xr_.grad, split)


@pytest.mark.parametrize("split", [1, 3])
def test_pair_attention_products(cuda_device, split):
    torch.manual_seed(13)
    B, H, dk, Lq = 2, 2, 128, 200
    row = 3 * H * dk
    QKV = torch.randn(B * Lq, row, device=cuda_device)
    qh, ql, qr = _prep(QKV, split)
    Lp = (Lq + 7) // 8 * 8
    S = torch.zeros(B, H, Lq, Lp, device=cuda_device)
    ops.gemm(ops.Opnd(qh, ql, L.MAJOR_K, (dk, Lq, H, B), (1, row, dk, Lq * row), src2=L.SRC_Z0, src3=L.SRC_Z1),
             ops.Opnd(qh, ql, L.MAJOR_K, (dk, Lq, H, B), (1, row, dk, Lq * row), src2=L.SRC_Z0, src3=L.SRC_Z1, offset=H * dk),
             Lq, Lq, dk, c_f32=S, ldc=Lp, c_sz0=Lq * Lp, c_sz1=H * Lq * Lp, nz0=H, nz1=B, split=split, block_n=256, pair=True)
    q4 = qr.view(B, Lq, 3, H, dk)
    Sref = torch.einsum("blhd,bmhd->bhlm", q4[:, :, 0], q4[:, :, 1])
    _check("pair attn QK^T", S[..., :Lq], Sref, split)
    P = torch.softmax(Sref.float() / 11.3, dim=-1)
    Pp = torch.zeros(B, H, Lq, Lp, device=cuda_device)
    Pp[..., :Lq] = P
    ph, pl, pr = _prep(Pp, split)
    O = torch.empty(B * Lq, H * dk, device=cuda_device)
    ops.gemm(ops.Opnd(ph, pl, L.MAJOR_K, (Lq, Lq, H, B), (1, Lp, Lq * Lp, H * Lq * Lp), src2=L.SRC_Z0, src3=L.SRC_Z1),
             ops.Opnd(qh, ql, L.MAJOR_MN, (dk, Lq, H, B), (1, row, dk, Lq * row), src2=L.SRC_Z0, src3=L.SRC_Z1, offset=2 * H * dk),
             Lq, dk, Lq, c_f32=O, ldc=H * dk, c_sz0=dk, c_sz1=Lq * H * dk, nz0=H, nz1=B, split=split, block_n=128, pair=True)
    Oref = torch.einsum("bhlm,bmhd->blhd", pr[..., :Lq], q4[:, :, 2]).reshape(B * Lq, H * dk)
    _check("pair attn PV (B MN-major)", O, Oref, split)
    dO = torch.randn(B * Lq, H * dk, device=cuda_device)
    dh, dl, dr = _prep(dO, split)
    dV = torch.empty(B * Lq, H * dk, device=cuda_device)
    ops.gemm(ops.Opnd(ph, pl, L.MAJOR_MN, (Lq, Lq, H, B), (1, Lp, Lq * Lp, H * Lq * Lp), src2=L.SRC_Z0, src3=L.SRC_Z1),
             ops.Opnd(dh, dl, L.MAJOR_MN, (dk, Lq, H, B), (1, H * dk, dk, Lq * H * dk), src2=L.SRC_Z0, src3=L.SRC_Z1),
             Lq, dk, Lq, c_f32=dV, ldc=H * dk, c_sz0=dk, c_sz1=Lq * H * dk, nz0=H, nz1=B, split=split, block_n=128, pair=True)
    dVref = torch.einsum("bhlm,blhd->bmhd", pr[..., :Lq], dr.view(B, Lq, H, dk)).reshape(B * Lq, H * dk)
    _check("pair attn dV (A,B MN-major)", dV, dVref, split)


@pytest.mark.parametrize("split", [1, 3])
@pytest.mark.parametrize("M,K,T,lens,drop,with_res", [
    (3456, 256, 864, [864, 700, 515, 300], True, True),     # out-projection of the configs[1] decoder (27 clusters of 4 CTAs)
    (512, 1024, 128, [128, 90, 128, 7], False, True),       # conv k=1 (w_2) of the encoder
    (200, 320, 50, None, True, False),                      # partial row tile, K tail, no mask, no residual
])
def test_gemm_ln_epilogue(cuda_device, split, M, K, T, lens, drop, with_res):
    """mtts_gemm_ln: GEMM + bias -> dropout -> + residual -> LayerNorm -> pad-row zeroing in one launch (row statistics exchanged
    between the 4 CTAs of a cluster) against RefOps' GEMM emulator followed by its LayerNorm (float64)."""
    from oracle.ops_reference import RefOps

    N = 256
    g = torch.Generator().manual_seed(7)
    X = torch.randn(M, K, generator=g)
    W = torch.randn(N, K, generator=g) / math.sqrt(K)
    bias = torch.randn(N, generator=g)
    res = torch.randn(M, N, generator=g) + 0.5 if with_res else None        # non-zero row means
    gamma, beta = 1 + 0.1 * torch.randn(N, generator=g), 0.1 * torch.randn(N, generator=g)
    klens = torch.tensor(lens, dtype=torch.int64) if lens is not None else None
    xh, xl = ops.split_bf16(X)
    wh, wl = ops.split_bf16(W)
    pre = (int(0.2 * (1 << 24)), 0xC0FFEE11, 1.0 / 0.8) if drop else (0, 0, 1.0)
    salt = torch.tensor([0x1234567], dtype=torch.int32)
    outs = {}
    for name, be, to in (("ref", RefOps(split=split), lambda t: t.clone()), ("cuda", ops.CudaOps(split=split), lambda t: t.to(cuda_device))):
        be.drop_salt = to(salt)
        o = dict(z=to(torch.zeros(M, N)), st=to(torch.zeros(M, 2)), out=to(torch.zeros(M, N)),
                 oh=to(torch.zeros(M, N, dtype=torch.bfloat16)), ol=to(torch.zeros(M, N, dtype=torch.bfloat16)))
        lo = (lambda t: to(t)) if split == 3 else (lambda t: None)
        be.gemm(ops.Opnd(to(xh), lo(xl), L.MAJOR_K, (K, M), (1, K)), ops.Opnd(to(wh), lo(wl), L.MAJOR_K, (K, N), (1, K)), M, N, K,
                c_f32=o["out"], c_hi=o["oh"], c_lo=o["ol"], ldc=N, bias=to(bias), block_n=64,
                ln=dict(res=None if res is None else to(res), gamma=to(gamma), beta=to(beta), lens=None if klens is None else to(klens),
                        T=T, z=o["z"], stats=o["st"], pre=pre))
        if name == "cuda":
            torch.cuda.synchronize()
        outs[name] = {k: v.cpu() for k, v in o.items()}
    r, c = outs["ref"], outs["cuda"]
    errs = {}
    for k in ("z", "st", "out"):
        assert torch.isfinite(c[k]).all(), k
        errs[k] = ((c[k].double() - r[k].double()).abs().max() / r[k].double().abs().max()).item()
    errs["hi+lo"] = ((c["oh"].double() + c["ol"].double() - r["out"].double()).abs().max() / r["out"].double().abs().max()).item()
    print(f"[gemm_ln] M{M} K{K} split{split} drop{drop}: " + " ".join(f"{k} {v:.1e}" for k, v in errs.items()))
    for k, v in errs.items():
        assert v < 1e-4, f"{k}: {v:.3e}"
    if klens is not None:                      # pad rows are exactly zero
        rows = torch.arange(M)
        pad = (rows % T) >= klens[rows // T]
        assert pad.any() and (c["out"][pad] == 0).all() and (c["oh"][pad] == 0).all()
    if drop:                                    # the epilogue's hash drops exactly the elements the host hash drops
        zz = (c["z"] - (res if res is not None else 0)).abs()
        assert abs((zz < 1e-12).float().mean().item() - 0.2) < 0.02
