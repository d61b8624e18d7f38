"""Generate tests/golden/audio_golden.npz from the REAL reference audio modules (container only; /root/reference).

    python -m oracle.make_golden_audio

Imports the unmodified `audio/stft.py`, `audio/audio_processing.py`, `audio/tools.py`.  librosa is absent in this image:
`librosa.util.{pad_center,tiny,normalize}` and `librosa.filters.mel` are stubbed with the restatements of
oracle/audio_oracle.py (so the mel filterbank matrix is NOT pinned by this script; everything downstream of it is).
`STFT.transform` hard-codes `.cuda()` (stft.py:67-72): `torch.Tensor.cuda` is patched to the identity for the run.
Griffin-Lim's unseeded `np.random.rand` start is made reproducible by seeding numpy right before the call (the oracle /
CUDA path take the same angles as an explicit argument).
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import audio_oracle as A  # noqa: E402
from oracle import refstub  # noqa: E402

CFG = dict(filter_length=1024, hop_length=256, win_length=1024, n_mel_channels=80, sampling_rate=22050, mel_fmin=0, mel_fmax=8000)
# preprocess/LibriTTS.yaml:25-35


def synth_wave(n: int, seed: int) -> np.ndarray:
    """A few decaying harmonics + noise in [-1, 1] (speech-shaped enough for mel / Griffin-Lim tests)."""
    rng = np.random.RandomState(seed)
    t = np.arange(n) / 22050.0
    f0 = 110.0 + 40.0 * np.sin(2 * np.pi * 1.5 * t + rng.rand())
    ph = 2 * np.pi * np.cumsum(f0) / 22050.0
    y = sum(np.sin(k * ph) / k for k in range(1, 9)) * (0.6 + 0.4 * np.sin(2 * np.pi * 3.0 * t))
    y = y + 0.05 * rng.randn(n)
    return (0.5 * y / np.abs(y).max()).astype(np.float32)


def install_librosa_stub():
    lib = types.ModuleType("librosa")
    util = types.ModuleType("librosa.util")
    util.pad_center = A.pad_center
    util.tiny = A.tiny
    util.normalize = lambda x, norm=None: x
    filt = types.ModuleType("librosa.filters")
    filt.mel = A.mel_filterbank
    lib.util, lib.filters = util, filt
    sys.modules.update({"librosa": lib, "librosa.util": util, "librosa.filters": filt})


def main():
    refstub.install_stubs()
    install_librosa_stub()
    sys.path.insert(0, refstub.REF)
    torch.Tensor.cuda = lambda self, *a, **k: self          # stft.py:67-72
    from audio import audio_processing as RP              # the real modules
    from audio import stft as RS
    from audio import tools as RT

    out = {}
    ref = RS.TacotronSTFT(**CFG)
    mine = A.TacotronSTFT(**CFG)
    assert torch.equal(ref.stft_fn.forward_basis, mine.stft_fn.forward_basis)
    assert torch.equal(ref.stft_fn.inverse_basis, mine.stft_fn.inverse_basis)
    # --- transform / inverse on a 2-utterance batch ---
    y = torch.from_numpy(np.stack([synth_wave(256 * 37, 1), synth_wave(256 * 37, 2)]))
    mag, ph = ref.stft_fn.transform(y)
    rec = ref.stft_fn.inverse(mag, ph)
    out["wave"] = y.numpy()
    out["mag"], out["phase"], out["recon"] = mag.numpy(), ph.numpy(), rec.numpy()
    m2, p2 = mine.stft_fn.transform(y)
    assert torch.equal(m2, mag) and torch.equal(p2, ph) and torch.equal(mine.stft_fn.inverse(m2, p2), rec)
    # --- mel_spectrogram / get_mel_from_wav ---
    mel, energy = RT.get_mel_from_wav(y[0].numpy(), ref)
    out["mel"], out["energy"] = mel, energy
    mel_m, en_m = A.get_mel_from_wav(y[0].numpy(), mine)
    assert np.array_equal(mel_m, mel) and np.array_equal(en_m, energy)
    # --- Griffin-Lim (tools.inv_mel_spec body; the shipped function dereferences a non-existent attribute, tools.py:28) ---
    mel_t = torch.from_numpy(mel)                      # [80, T]
    dec = ref.spectral_de_normalize(torch.stack([mel_t])).transpose(1, 2).data.cpu()
    spec = torch.mm(dec[0], ref.mel_basis).transpose(0, 1).unsqueeze(0) * 1000
    for iters in (0, 3, 30):
        np.random.seed(1234)
        audio = RP.griffin_lim(torch.autograd.Variable(spec[:, :, :-1]), ref.stft_fn, iters).squeeze().cpu().numpy()
        np.random.seed(1234)
        ang = np.angle(np.exp(2j * np.pi * np.random.rand(*spec[:, :, :-1].size()))).astype(np.float32)
        a2 = A.inv_mel_spec(mel_t, mine, iters, init_angles=ang)
        assert np.array_equal(a2, audio), iters
        out[f"gl_audio_{iters}"] = audio
    out["gl_init_angles"] = ang
    out["gl_spec"] = spec.numpy()
    out["window_sum"] = RP.window_sumsquare("hann", 36, hop_length=256, win_length=1024, n_fft=1024, dtype=np.float32)
    out["mel_basis"] = ref.mel_basis.numpy()             # (built by the stubbed librosa restatement: not a reference pin)
    path = os.path.join(ROOT, "tests", "golden", "audio_golden.npz")
    np.savez_compressed(path, **out)
    print("[golden] wrote", path, os.path.getsize(path) // 1024, "KiB; the restatement reproduces the real modules bit for bit")


if __name__ == "__main__":
    main()
