"""LengthRegulator CUDA kernels vs the oracle restatement of modules.py:167-194 (bit-exact)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from meta_tts_b200 import ops  # noqa: E402
from oracle.fs2_oracle import length_regulator_ref  # noqa: E402


@pytest.mark.parametrize("B,L,T,C", [(4, 128, 864, 256), (3, 17, 50, 256), (2, 5, 40, 12), (1, 64, 512, 256)])
@pytest.mark.parametrize("kind", ["int64", "float32"])
def test_length_regulator_bit_exact(cuda_device, B, L, T, C, kind):
    g = torch.Generator().manual_seed(B * 1000 + L)
    d = torch.randint(0, 2 * T // L + 2, (B, L), generator=g)
    d[0, : L // 2] = 0                      # zero durations (skipped phonemes)
    if L > 3:
        d[-1, 3] = -2                       # negative -> clamped to 0 (modules.py:187)
    # keep total <= T for some rows, exceed for none (reference pad() would fail on overflow)
    for b in range(B):
        while int(d[b].clamp_min(0).sum()) > T:
            d[b] = d[b] // 2
    x = torch.randn(B, L, C, generator=g)
    dur = d if kind == "int64" else d.float()
    ref_out, ref_len = length_regulator_ref(x, dur, T)
    idx, mel_len = ops.length_regulate_index(dur.to(cuda_device), T)
    out = ops.length_regulate_fwd(x.to(cuda_device), idx)
    assert torch.equal(mel_len.cpu(), ref_len)
    assert torch.equal(out.cpu(), ref_out)           # bit-exact payload copy
    # backward: segment sum == autograd of the reference gather
    dy = torch.randn(B, T, C, generator=g)
    xr = x.clone().requires_grad_(True)
    ro, _ = length_regulator_ref(xr, dur, T)
    ro.backward(dy)
    dx = ops.length_regulate_bwd(dy.to(cuda_device), dur.to(cuda_device), L)
    torch.testing.assert_close(dx.cpu(), xr.grad, rtol=1e-5, atol=1e-5)
