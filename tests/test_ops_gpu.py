"""Every libmtts row / reduction / elementwise kernel vs the independent CPU restatement
(oracle/ops_reference.RefOps; tangent forms there come from torch.func jvp/vjp).  Tolerance: 1e-5
relative (fp32 kernels; fast-math exp/rsqrt), bit-exact for integer paths."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from meta_tts_b200.ops import CudaOps  # noqa: E402
from oracle.ops_reference import RefOps  # noqa: E402

TOL = 2e-5


SALT = 0x9ABCDEF1          # > 2^31: exercises the uint32 wrap of the host-side hash


def _both(cuda_device, name, args, kwargs=None, tol=TOL, split=3, skip=(), salt=None):
    """Run op `name` on RefOps (CPU) and CudaOps (GPU); compare every tensor argument afterwards."""
    kwargs = kwargs or {}
    ref, cu = RefOps(split=split), CudaOps(split=split)
    if salt is not None:
        ref.drop_salt = torch.tensor([salt - (1 << 32) if salt >= (1 << 31) else salt], dtype=torch.int32)
        cu.drop_salt = ref.drop_salt.to(cuda_device)
    cargs = [a.clone() if torch.is_tensor(a) else a for a in args]
    gargs = [a.to(cuda_device) if torch.is_tensor(a) else a for a in args]
    ckw = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in kwargs.items()}
    gkw = {k: (v.to(cuda_device) if torch.is_tensor(v) else v) for k, v in kwargs.items()}
    getattr(ref, name)(*cargs, **ckw)
    getattr(cu, name)(*gargs, **gkw)
    torch.cuda.synchronize()
    i = 0
    while i < len(cargs):
        c, g = cargs[i], gargs[i]
        if not torch.is_tensor(c) or i in skip:
            i += 1
            continue
        g = g.cpu()
        if c.dtype in (torch.int64, torch.int32):
            assert torch.equal(c, g), f"{name}: int arg {i} differs"
        elif c.dtype == torch.bfloat16:
            # (hi, lo) pairs: the halves may round differently; their SUM is the fp32-grade value
            nxt = cargs[i + 1] if i + 1 < len(cargs) else None
            if torch.is_tensor(nxt) and nxt.dtype == torch.bfloat16:
                cs, gs = c.double() + nxt.double(), g.double() + gargs[i + 1].cpu().double()
                i += 1
                lim = tol
            else:
                cs, gs, lim = c.double(), g.double(), 1e-2
            assert torch.isfinite(gs).all(), f"{name}: arg {i} has non-finite values"
            err = (cs - gs).abs().max() / cs.abs().max().clamp_min(1e-30)
            assert err < lim, f"{name}: bf16 arg {i} max err {err:.3e}"
        else:
            c64, g64 = c.double(), g.double()
            assert torch.isfinite(g64).all(), f"{name}: arg {i} has non-finite values"
            scale = c64.abs().max().clamp_min(1e-30)
            err = (c64 - g64).abs().max() / scale
            assert err < tol, f"{name}: arg {i} max err {err:.3e} (scale {scale:.3e})"
        i += 1
    return cargs, [g.cpu() if torch.is_tensor(g) else g for g in gargs]


def _hl_check(c_hi, c_lo, g_hi, g_lo, tol=TOL):
    c = c_hi.double() + c_lo.double()
    g = g_hi.double() + g_lo.double()
    assert ((c - g).abs().max() / c.abs().max().clamp_min(1e-30)) < tol


def R_(*s, seed=0):
    return torch.randn(*s, generator=torch.Generator().manual_seed(seed))


def bf(*s):
    return torch.zeros(*s, dtype=torch.bfloat16)


def _site(p, seed):
    return (int(p * (1 << 24)), seed, 1.0 / (1.0 - p))


@pytest.mark.parametrize("drop", [False, True])
@pytest.mark.parametrize("with_res,with_lens", [(True, True), (False, False)])
def test_layernorm_family(cuda_device, with_res, with_lens, drop):
    """drop=True: both dropout sites active (pre 0.2 on the branch, post 0.5 on the output); the kernels' counter hash
    must select exactly the elements the host hash selects, else the errors are O(1)."""
    B, T, C = 3, 37, 256
    R = B * T
    kw = {"pre": _site(0.2, 0xDEADBEEF), "post": _site(0.5, 12345)} if drop else {}
    salt = SALT if drop else None
    lens = torch.tensor([37, 20, 5]) if with_lens else None
    y, res = R_(R, C, seed=1), (R_(R, C, seed=2) if with_res else None)
    gamma, beta = 1 + 0.1 * R_(C, seed=3), 0.1 * R_(C, seed=4)
    z, st, out, oh, ol = torch.zeros(R, C), torch.zeros(R, 2), torch.zeros(R, C), bf(R, C), bf(R, C)
    c, g = _both(cuda_device, "ln_fwd", [y, res, gamma, beta, lens, T, R, C, z, st, out, oh, ol], kw, salt=salt)
    if drop:
        keep = (c[10][:lens[0] if with_lens else T] != 0).float().mean().item()
        assert 0.45 < keep < 0.55                       # post site, p = 0.5
    _hl_check(c[11], c[12], g[11], g[12])
    z, st = c[8], c[9]
    dy = R_(R, C, seed=5)
    for gate in (0, 1):
        dz, dh, dl = torch.zeros(R, C), bf(R, C), bf(R, C)
        dg, db, dbias = torch.zeros(C), torch.zeros(C), torch.zeros(C)
        c2, g2 = _both(cuda_device, "ln_bwd", [dy, z, st, gamma, lens, T, R, C, gate, dz, dh, dl, dg, db, dbias], kw, salt=salt)
        _hl_check(c2[10], c2[11], g2[10], g2[11])
    ydot, resdot = R_(R, C, seed=6), (R_(R, C, seed=7) if with_res else None)
    gdot, bdot = R_(C, seed=8), R_(C, seed=9)
    zd, od, odh, odl = torch.zeros(R, C), torch.zeros(R, C), bf(R, C), bf(R, C)
    c3, _ = _both(cuda_device, "ln_tfwd", [ydot, resdot, z, st, gamma, gdot, bdot, lens, T, R, C, zd, od, odh, odl], kw, salt=salt)
    zd = c3[11]
    ddy = R_(R, C, seed=10)
    for gate in (0, 1):
        ddz, dh, dl = torch.zeros(R, C), bf(R, C), bf(R, C)
        dg, db, dbias = torch.zeros(C), torch.zeros(C), torch.zeros(C)
        _both(cuda_device, "ln_tbwd", [dy, ddy, z, zd, st, gamma, gdot, lens, T, R, C, gate, ddz, dh, dl, dg, db, dbias],
              kw, tol=5e-5, salt=salt)


def test_rowdot(cuda_device):
    B, T, C = 2, 19, 256
    R = B * T
    lens = torch.tensor([19, 7])
    h, hd, w, wd, b, bd = R_(R, C, seed=1), R_(R, C, seed=2), R_(C, seed=3), R_(C, seed=4), R_(1, seed=5), R_(1, seed=6)
    _both(cuda_device, "rowdot_fwd", [h, None, w, None, b, None, lens, T, R, C, torch.zeros(R)])
    _both(cuda_device, "rowdot_fwd", [h, hd, w, wd, None, bd, lens, T, R, C, torch.zeros(R)])
    dout, ddout = R_(R, seed=7), R_(R, seed=8)
    _both(cuda_device, "rowdot_bwd", [dout, None, h, None, w, None, lens, T, R, C, torch.zeros(R, C), torch.zeros(C), torch.zeros(1)])
    _both(cuda_device, "rowdot_bwd", [dout, ddout, h, hd, w, wd, lens, T, R, C, torch.zeros(R, C), torch.zeros(C), torch.zeros(1)])


@pytest.mark.parametrize("Lq,ld", [(24, 24), (50, 56), (50, 51), (150, 160), (330, 352), (864, 896)])
def test_softmax_family(cuda_device, Lq, ld):
    """ld % 4 == 0 -> the 128-bit vector kernels (1 / 2 / 4 / 7 key groups per lane), else the scalar kernel.
    Key lengths that straddle a 4-key group, and NaN in the padded columns [Lq, ld) the GEMMs never write."""
    B, H = 2, 2
    nz = B * H
    kl = torch.tensor([Lq, Lq // 2 + 1])
    S = 3 * R_(nz, Lq, ld, seed=1)
    S[..., Lq:] = float("nan")
    ph, pl = bf(nz, Lq, ld), bf(nz, Lq, ld)
    c, g = _both(cuda_device, "softmax", [0, S, None, None, None, None, None, kl, nz, H, Lq, Lq, ld, ph, pl], skip=(1, 2))
    _hl_check(c[13], c[14], g[13], g[14])
    ph, pl = c[13], c[14]
    dP = R_(nz, Lq, ld, seed=2)
    dP[..., Lq:] = float("nan")
    oh, ol = bf(nz, Lq, ld), bf(nz, Lq, ld)
    c1, g1 = _both(cuda_device, "softmax", [1, dP, None, ph, pl, None, None, kl, nz, H, Lq, Lq, ld, oh, ol], skip=(1, 2))
    _hl_check(c1[13], c1[14], g1[13], g1[14])
    pdh, pdl = c1[13], c1[14]
    ddP = R_(nz, Lq, ld, seed=3)
    ddP[..., Lq:] = float("nan")
    oh, ol = bf(nz, Lq, ld), bf(nz, Lq, ld)
    c2, g2 = _both(cuda_device, "softmax", [2, dP, ddP, ph, pl, pdh, pdl, kl, nz, H, Lq, Lq, ld, oh, ol], skip=(1, 2))
    _hl_check(c2[13], c2[14], g2[13], g2[14])


def test_softmax_long_rows_forward_only(cuda_device):
    """Rows longer than 1024 keys (eval-mode sequences beyond max_seq_len, Models.py:148-156): forward works, the backward /
    tangent forms refuse (train mode truncates to max_seq_len, so they cannot occur)."""
    B, H, Lq, ld = 1, 2, 1573, 1576
    nz = B * H
    kl = torch.tensor([1500])
    S = 3 * R_(nz, Lq, ld, seed=1)
    S[..., Lq:] = float("nan")
    c, g = _both(cuda_device, "softmax", [0, S, None, None, None, None, None, kl, nz, H, Lq, Lq, ld, bf(nz, Lq, ld), bf(nz, Lq, ld)], skip=(1, 2))
    _hl_check(c[13], c[14], g[13], g[14])
    from meta_tts_b200.lib import MttsError
    with pytest.raises(MttsError):
        CudaOps(split=3).softmax(1, S.to(cuda_device), None, g[13].to(cuda_device), g[14].to(cuda_device), None, None, kl.to(cuda_device), nz, H,
                                 Lq, Lq, ld, bf(nz, Lq, ld).to(cuda_device), bf(nz, Lq, ld).to(cuda_device))


def test_gathers_and_sums(cuda_device):
    B, T, C, V = 3, 11, 256, 40
    R = B * T
    g = torch.Generator().manual_seed(0)
    idx = torch.randint(0, V, (R,), generator=g)
    table, base, pos = R_(V, C, seed=1), R_(R, C, seed=2), R_(T + 3, C, seed=3)
    _both(cuda_device, "embed_fwd", [idx, table, base, pos, T, R, C, torch.zeros(R, C), bf(R, C), bf(R, C)])
    _both(cuda_device, "embed_fwd", [idx, table, None, None, T, R, C, torch.zeros(R, C), None, None])
    _both(cuda_device, "embed_bwd", [idx, R_(R, C, seed=4), R, C, 0, 0.5, torch.zeros(V, C)], tol=5e-5)
    bins = torch.linspace(-2.9, 10.2, 255)
    v = torch.cat([3 * R_(R - 3, seed=5), bins[[0, 100, 254]]])          # exact boundary hits included
    _both(cuda_device, "bucketize", [v, bins, 255, R, torch.zeros(R, dtype=torch.int64)])
    vec = R_(B, C, seed=6)
    _both(cuda_device, "add_rowvec", [R_(B, T, C, seed=7), vec, C, pos, B, T, C, torch.zeros(B, T, C), bf(B, T, C), bf(B, T, C)])
    _both(cuda_device, "add_rowvec", [R_(B, T, C, seed=7), vec, C, None, B, T, C, torch.zeros(B, T, C), None, None])
    ids = torch.tensor([3, 3, 7])
    for avg, n_out in ((False, 3), (True, 5)):
        _both(cuda_device, "spk_embed", [ids, table, 3, C, avg, n_out, torch.zeros(n_out, C)])
        _both(cuda_device, "spk_embed_bwd", [ids, R_(n_out, C, seed=8), 3, C, avg, n_out, 1.0, torch.zeros(V, C)])
    x = R_(B, T, C, seed=9)
    _both(cuda_device, "colsum", [x, None, None, B, T, C, torch.zeros(B, C)], tol=5e-5)
    xh = x.to(torch.bfloat16)
    xl = (x - xh.float()).to(torch.bfloat16)
    _both(cuda_device, "colsum", [None, xh, xl, 1, R, C, torch.zeros(C)], tol=5e-5)


@pytest.mark.parametrize("drop", [False, True])
@pytest.mark.parametrize("C,tanh", [(512, True), (80, False)])
def test_batchnorm_family(cuda_device, C, tanh, drop):
    R = 4 * 53
    kw = {"drop": _site(0.5, 0xC0FFEE)} if drop else {}
    salt = SALT if drop else None
    x, xd = 2 * R_(R, C, seed=1) + 0.5, R_(R, C, seed=2)
    gamma, beta, gd, bd = 1 + 0.1 * R_(C, seed=3), 0.1 * R_(C, seed=4), R_(C, seed=5), R_(C, seed=6)
    rm, rv = torch.zeros(C), torch.ones(C)
    ws, st, out, oh, ol = torch.zeros(4 * 512), torch.zeros(2 * C), torch.zeros(R, C), bf(R, C), bf(R, C)
    c, g = _both(cuda_device, "bn_fwd", [x, gamma, beta, R, C, tanh, rm, rv, ws, st, out, oh, ol], kw, tol=5e-5, skip=(8,), salt=salt)
    st, o = c[9], c[10]
    dout, ddout = R_(R, C, seed=7), R_(R, C, seed=8)
    _both(cuda_device, "bn_bwd", [dout, o if tanh else None, x, st, gamma, R, C, tanh, ws, torch.zeros(R, C), bf(R, C), bf(R, C),
                                  torch.zeros(C), torch.zeros(C)], {"beta": beta, **kw}, tol=1e-4, skip=(8,), salt=salt)
    ts, od = torch.zeros(2 * C), torch.zeros(R, C)
    c2, _ = _both(cuda_device, "bn_tfwd", [xd, x, st, gamma, gd, bd, o if tanh else None, R, C, tanh, ws, ts, od, bf(R, C), bf(R, C)],
                  {"beta": beta, **kw}, tol=1e-4, skip=(10,), salt=salt)
    ts, od = c2[11], c2[12]
    _both(cuda_device, "bn_tbwd", [dout, ddout, o if tanh else None, od if tanh else None, x, xd, st, ts, gamma, gd, R, C, tanh, ws,
                                   torch.zeros(R, C), bf(R, C), bf(R, C), torch.zeros(C), torch.zeros(C)],
          {"beta": beta, "bdot": bd, **kw}, tol=2e-4, skip=(13,), salt=salt)


def test_loss_family(cuda_device):
    B, T, Lq, NM = 3, 20, 7, 80
    mel_lens, src_lens = torch.tensor([20, 11, 3]), torch.tensor([7, 4, 2])
    mel, post, tgt = R_(B, T, NM, seed=1), R_(B, T, NM, seed=2), R_(B, T, NM, seed=3)
    p, pt, e, et, logd = R_(B, Lq, seed=4), R_(B, Lq, seed=5), R_(B, Lq, seed=6), R_(B, Lq, seed=7), R_(B, Lq, seed=8)
    dur = torch.randint(0, 9, (B, Lq), generator=torch.Generator().manual_seed(9))
    c, _ = _both(cuda_device, "loss_fwd", [mel, post, tgt, mel_lens, p, pt, e, et, logd, dur, src_lens, B, T, Lq, NM, torch.zeros(8),
                                           torch.zeros(6), torch.zeros(2)], skip=(15,))
    counts = c[17]
    outs = [torch.zeros(B, T, NM), torch.zeros(B, T, NM), torch.zeros(B, Lq), torch.zeros(B, Lq), torch.zeros(B, Lq)]
    _both(cuda_device, "loss_bwd", [mel, post, tgt, mel_lens, p, pt, e, et, logd, dur, src_lens, B, T, Lq, NM, counts, 0.7, 0, *outs])
    _both(cuda_device, "loss_bwd", [None, None, None, mel_lens, p, None, e, None, logd, None, src_lens, B, T, Lq, NM, counts, 0.7, 1,
                                    *[torch.ones_like(o) for o in outs]])


def test_elementwise(cuda_device):
    n = 4096 + 64
    x, g = R_(n, seed=1), R_(n, seed=2)
    c, gg = _both(cuda_device, "split_", [x, bf(n), bf(n)])
    _hl_check(c[1], c[2], gg[1], gg[2])
    _both(cuda_device, "sgd_split", [x, g, 0.001, torch.zeros(n), bf(n), bf(n)])
    _both(cuda_device, "axpby", [-0.3, x, 1.5, g.clone()])
    cu = CudaOps(split=3)
    ws = torch.zeros(2048, device=cuda_device)
    cu.sumsq(x.to(cuda_device), ws)
    ref = (x.double() ** 2).sum().item()
    assert abs(float(ws[0]) - ref) < 1e-5 * ref
    ws2 = torch.full((2048,), 7.0, device=cuda_device)       # stale workspace contents must not matter; same bits run to run
    cu.sumsq(x.to(cuda_device), ws2)
    assert float(ws2[0]) == float(ws[0])
    hyper = torch.tensor([0.01, 1 - 0.9, 1 - 0.98, 0.0])
    ss = (g.double() ** 2).sum().float().reshape(1)
    _both(cuda_device, "adam_clip", [x.clone(), g, 0.1 * R_(n, seed=3), R_(n, seed=4).abs(), ss, 0.5, 1.0, hyper, 0.9, 0.98, 1e-9,
                                     bf(n), bf(n)], tol=5e-5)


# ---- free-running synthesis / vocoder-side elementwise kernels (csrc/mtts_audio.cu) ----------------------------------
def test_duration_round_exact(cuda_device):
    """Integer path: clamp(round(exp(log_d) - 1) * d_control, 0) with round-half-to-even, incl. exact .5 cases."""
    g = torch.Generator().manual_seed(0)
    logd = torch.cat([torch.randn(4000, generator=g) * 1.5 + 1.0, torch.log(torch.tensor([1.5, 2.5, 3.5, 4.5, 1.0, 0.5])),
                      torch.tensor([-3.0, 0.0, 6.0])])
    for dc in (1.0, 1.3, 0.5):
        ref, got = _both(cuda_device, "duration_round", [logd, dc, torch.zeros_like(logd)], skip=(2,))
        # exp() differs by an ulp between libm and the device: only inputs whose exp(log_d) - 1 sits within 1e-5 (relative)
        # of a rounding boundary may legitimately differ; everything else must be identical
        v = torch.exp(logd.double()) - 1
        safe = ((v % 1.0) - 0.5).abs() > 1e-5 * v.abs().clamp_min(1.0)
        assert int(safe.sum()) > 3900 and torch.equal(ref[2][safe], got[2][safe])
        assert float(got[2].min()) >= 0.0 and (dc != 1.0 or torch.equal(got[2], got[2].round()))


def test_bn_eval_and_unary(cuda_device):
    g = torch.Generator().manual_seed(1)
    R, C = 333, 80
    x = torch.randn(R, C, generator=g) * 2 + 0.3
    gamma, beta = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.1
    rm, rv = torch.randn(C, generator=g) * 0.2, torch.rand(C, generator=g) + 0.2
    bf = lambda: torch.zeros(R, C, dtype=torch.bfloat16)  # noqa: E731
    for tanh in (True, False):
        _both(cuda_device, "bn_eval", [x, gamma, beta, rm, rv, R, C, tanh, torch.zeros(R, C), bf(), bf()])
    v = torch.rand(1000, generator=g) * 3
    _both(cuda_device, "unary", [0, v - 0.5, 1e-5, 1.0, torch.zeros(1000)])                 # log(clamp)
    _both(cuda_device, "unary", [1, v - 1.5, 0.25, 0.0, torch.zeros(1000), torch.zeros(1000, dtype=torch.bfloat16),
                                 torch.zeros(1000, dtype=torch.bfloat16)])                    # exp * a, + operand split
    _both(cuda_device, "unary", [2, v, 1.7, 0.0, torch.zeros(1000)])


def test_stft_elementwise_kernels(cuda_device):
    g = torch.Generator().manual_seed(2)
    B, N, pad, ld = 3, 700, 512, 1792
    x = torch.randn(B, N, generator=g)
    bf = lambda *s: torch.zeros(*s, dtype=torch.bfloat16)  # noqa: E731
    _both(cuda_device, "reflect_pad", [x, B, N, pad, ld, torch.zeros(B, ld), bf(B, ld), bf(B, ld)])
    _both(cuda_device, "reflect_pad", [x, B, N, pad, 1500, torch.zeros(B, 1500), bf(B, 1500), bf(B, 1500)])   # ld cuts the tail
    R, nb, io = 77, 513, 520
    ri = torch.randn(R, 2 * io, generator=g)
    mag, ph, en = torch.zeros(R, io), torch.zeros(R, io), torch.zeros(R)
    ref, got = _both(cuda_device, "stft_polar", [ri, R, nb, 2 * io, io, io, mag, ph, en, bf(R, io), bf(R, io)], tol=1e-5)
    assert float(got[6][:, nb:].abs().max()) == 0.0 and float(got[7][:, nb:].abs().max()) == 0.0      # pad columns zeroed
    _both(cuda_device, "stft_recombine", [ref[6], ref[7], None, R, nb, 2 * io, io, io, bf(R, 2 * io), bf(R, 2 * io)], tol=1e-5)
    _both(cuda_device, "stft_recombine", [ref[6], None, ri, R, nb, 2 * io, io, io, bf(R, 2 * io), bf(R, 2 * io)], tol=1e-5)
    n, trim = 2048, 512
    ola = torch.randn(2, n, generator=g)
    ws = torch.rand(n, generator=g)
    ws[:7] = 0.0                                                         # below `tiny`: left un-normalised
    _both(cuda_device, "istft_finish", [ola, ws, 1.1754944e-38, 4.0, 2, n, trim, torch.zeros(2, n - 2 * trim)])
    ws2 = ws.clone()
    ws2[600:620] = 0.0
    _both(cuda_device, "istft_finish", [ola, ws2, 1.1754944e-38, 4.0, 2, n, trim, torch.zeros(2, n - 2 * trim)])
