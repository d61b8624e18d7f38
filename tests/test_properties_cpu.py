"""Size-independent properties of the path's index / linear pieces, checked on seeded random inputs (oracle and product host
logic through the CPU op restatement): LengthRegulator mass conservation and monotone indices, STFT linearity and the
transform -> inverse round trip, collate padding invariants."""
import numpy as np
import pytest
import torch

from meta_tts_b200 import audio as PA
from meta_tts_b200 import collate as B
from oracle import collate_oracle as C
from oracle import fs2_oracle as O
from oracle.ops_reference import RefOps


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_length_regulator_properties(seed):
    g = torch.Generator().manual_seed(seed)
    Bn, L, Cc = 3, 11, 6
    d = torch.randint(-2, 7, (Bn, L), generator=g)
    x = torch.randn(Bn, L, Cc, generator=g)
    T = int(d.clamp_min(0).sum(1).max()) + 3
    out, mel_len = O.length_regulator_ref(x, d, T)
    be = RefOps()
    idx = torch.zeros(Bn, T, dtype=torch.int32)
    ml = torch.zeros(Bn, dtype=torch.int64)
    be.lr_index(d, T, idx, ml)
    got = torch.zeros(Bn, T, Cc)
    be.lr_fwd(x, idx, got)
    assert torch.equal(ml, mel_len) and torch.equal(ml, d.clamp_min(0).sum(1)) and torch.equal(got, out)
    # mass conservation: sum_t out[b, t] = sum_j max(d_j, 0) * x[b, j];   frames beyond mel_len are zero
    assert torch.allclose(out.sum(1), (d.clamp_min(0)[..., None] * x).sum(1), atol=1e-5)
    for b in range(Bn):
        assert float(out[b, int(ml[b]):].abs().max() if int(ml[b]) < T else 0.0) == 0.0
        valid = idx[b, :int(ml[b])]
        assert bool((valid[1:] >= valid[:-1]).all())                       # indices never go back
        assert torch.equal(torch.bincount(valid.long(), minlength=L), d[b].clamp_min(0))   # phoneme j appears exactly d_j times
    # backward is the adjoint: <LR(x), dy> == <x, LR^T(dy)>
    dy = torch.randn(Bn, T, Cc, generator=g)
    dx = torch.zeros(Bn, L, Cc)
    be.lr_bwd(dy, d, L, dx)
    assert torch.allclose((out * dy).sum(), (x * dx).sum(), rtol=1e-4, atol=1e-5)


def test_stft_linearity_and_round_trip():
    st = PA.STFT(1024, 256, 1024, backend=RefOps(split=3))
    g = torch.Generator().manual_seed(5)
    x, y = 0.3 * torch.randn(2, 256 * 11, generator=g), 0.3 * torch.randn(2, 256 * 11, generator=g)
    rx, ry, rz = st.transform_fm(x), st.transform_fm(y), st.transform_fm(0.7 * x - 1.9 * y)
    assert ((rz - (0.7 * rx - 1.9 * ry)).norm() / rz.norm()).item() < 2e-5          # the transform is linear before |.|
    assert float(rz[:, :, 513:520].abs().max()) == 0.0 and float(rz[:, :, 1033:].abs().max()) == 0.0   # pad columns stay zero
    rec = st.forward(x)[:, 0]
    assert ((rec[:, 1024:-1024] - x[:, 1024:rec.shape[1] - 1024]).norm() / x[:, 1024:-1024].norm()).item() < 1e-4
    # Parseval-type check on the un-windowed basis: energy of a frame's one-sided spectrum matches the time-domain energy
    st0 = PA.STFT(1024, 256, 1024, window=None, backend=RefOps(split=3))
    r0 = st0.transform_fm(x)[:, 3]                                            # one interior frame
    re, im = r0[:, :513], r0[:, 520:1033]
    e_freq = (re[:, 0] ** 2 + re[:, 512] ** 2 + 2 * (re[:, 1:512] ** 2 + im[:, 1:512] ** 2).sum(1)) / 1024
    frame = torch.nn.functional.pad(x[:, None], (512, 512), mode="reflect")[:, 0, 3 * 256:3 * 256 + 1024]
    assert torch.allclose(e_freq, (frame ** 2).sum(1), rtol=1e-4)


@pytest.mark.parametrize("seed", [0, 4, 9])
def test_collate_padding_invariants(seed):
    data = C.synth_dataset(n=7, seed=seed, lmin=1, lmax=23)
    t12 = B.reprocess(data, np.arange(len(data)))
    texts, tl, mels, ml, pit, ene, dur = t12[3], t12[4], t12[6], t12[7], t12[9], t12[10], t12[11]
    assert int(t12[5]) == int(tl.max()) and int(t12[8]) == int(ml.max())
    for i, dd in enumerate(data):
        L, T = int(tl[i]), int(ml[i])
        assert L == len(dd["text"]) and T == dd["mel"].shape[0] == int(dd["duration"].sum())
        assert np.array_equal(texts[i, :L].numpy(), dd["text"]) and int(texts[i, L:].abs().sum()) == 0
        assert np.array_equal(mels[i, :T].numpy(), dd["mel"]) and float(mels[i, T:].abs().sum()) == 0.0
        assert int(dur[i, L:].abs().sum()) == 0 and float(pit[i, L:].abs().sum()) == 0.0 and float(ene[i, L:].abs().sum()) == 0.0
    # sorted collate: a permutation of the same rows, longest text first
    s12 = B.get_single_collate(sort=True)(data)
    assert sorted(s12[0]) == sorted(t12[0]) and bool((s12[4][:-1] >= s12[4][1:]).all())
