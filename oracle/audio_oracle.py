"""CPU oracle for the vocoder-side decode of BASELINE configs[4] — TEST INFRASTRUCTURE ONLY.

Restates, op for op in fp32 torch (the reference's own arithmetic), the reference's
  * `audio/stft.py:15-127`    STFT (conv1d with a windowed Fourier basis, conv_transpose1d with its pseudo-inverse,
                              window-sum-square normalisation) and `TacotronSTFT.mel_spectrogram` (:139-178),
  * `audio/audio_processing.py:7-100`  window_sumsquare, griffin_lim, dynamic_range_(de)compression,
  * `audio/tools.py:9-34`     get_mel_from_wav, inv_mel_spec (up to the wav file write).
Third-party arithmetic that is NOT under /root/reference: librosa (`requirements.txt`, unpinned; absent in this image):
`librosa.util.pad_center`, `librosa.util.tiny`, `librosa.util.normalize(norm=None)` (identity) and
`librosa.filters.mel` (Slaney scale, Slaney area normalisation — the positional call `librosa_mel_fn(sr, n_fft, n_mels,
fmin, fmax)` at stft.py:145-147 implies librosa < 0.10).  Their published algorithms are restated below.

Pinned: `oracle/make_golden_audio.py` imports the REAL `audio/stft.py`, `audio/audio_processing.py` and `audio/tools.py`
(librosa stubbed with the restatements here, `.cuda()` at stft.py:67-72 patched to identity) and stores
transform / inverse / Griffin-Lim / mel outputs in `tests/golden/audio_golden.npz`; `tests/test_audio_cpu.py` checks this
file against them.  The mel filterbank itself has no reference-side pin (librosa absent): "parity unpinned" for that
matrix only — it is a constant both sides build once.

The reference's `inv_mel_spec` is dead code as shipped (`_stft._stft_fn` does not exist, tools.py:28; the attribute is
`stft_fn`, stft.py:142): behaviour is defined here with the evident intent (`_stft.stft_fn`).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F
from scipy.signal import get_window


# ---- librosa restatements ------------------------------------------------------------------------------------------
def pad_center(data: np.ndarray, size: int) -> np.ndarray:
    """librosa.util.pad_center (1-D): centre `data` in a zero vector of length `size`."""
    n = data.shape[-1]
    lpad = int((size - n) // 2)
    assert lpad >= 0
    return np.pad(data, (lpad, int(size - n - lpad)), mode="constant")


def tiny(x) -> float:
    """librosa.util.tiny: smallest positive normal number of x's dtype."""
    x = np.asarray(x)
    dt = x.dtype if np.issubdtype(x.dtype, np.floating) else np.dtype(np.float32)
    return float(np.finfo(dt).tiny)


def _hz_to_mel(f):
    f = np.asanyarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep, mels)


def _mel_to_hz(m):
    m = np.asanyarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    freqs = f_sp * m
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), freqs)


def mel_filterbank(sr: int, n_fft: int, n_mels: int, fmin: float, fmax) -> np.ndarray:
    """librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax) with its defaults (htk=False, norm='slaney') -> [n_mels, 1+n_fft/2] f32."""
    if fmax is None:
        fmax = sr / 2.0
    fftfreqs = np.linspace(0, sr / 2.0, int(1 + n_fft // 2), endpoint=True)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(fmin), _hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    w = np.zeros((n_mels, len(fftfreqs)))
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        w[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    w *= enorm[:, None]
    return w.astype(np.float32)


# ---- audio/stft.py ---------------------------------------------------------------------------------------------------
def stft_bases(filter_length: int, hop_length: int, win_length: int, window: str = "hann"):
    """stft.py:18-52 -> (forward_basis [2*cutoff, 1, n_fft], inverse_basis [2*cutoff, 1, n_fft]) float32 tensors."""
    scale = filter_length / hop_length
    fourier_basis = np.fft.fft(np.eye(filter_length))
    cutoff = int(filter_length / 2 + 1)
    fourier_basis = np.vstack([np.real(fourier_basis[:cutoff, :]), np.imag(fourier_basis[:cutoff, :])])
    forward_basis = torch.FloatTensor(fourier_basis[:, None, :])
    inverse_basis = torch.FloatTensor(np.linalg.pinv(scale * fourier_basis).T[:, None, :])
    if window is not None:
        assert filter_length >= win_length
        fft_window = get_window(window, win_length, fftbins=True)
        fft_window = torch.from_numpy(pad_center(fft_window, filter_length)).float()
        forward_basis *= fft_window
        inverse_basis *= fft_window
    return forward_basis.float(), inverse_basis.float()


def window_sumsquare(window, n_frames, hop_length, win_length, n_fft, dtype=np.float32):
    """audio_processing.py:7-60 (librosa 0.6 window_sumsquare; normalize(norm=None) is the identity)."""
    if win_length is None:
        win_length = n_fft
    n = n_fft + hop_length * (n_frames - 1)
    x = np.zeros(n, dtype=dtype)
    win_sq = get_window(window, win_length, fftbins=True) ** 2
    win_sq = pad_center(win_sq, n_fft)
    for i in range(n_frames):
        sample = i * hop_length
        x[sample:min(n, sample + n_fft)] += win_sq[:max(0, min(n_fft, n - sample))]
    return x


class STFT:
    """stft.py:15-127 on the CPU (the reference's `.cuda()` calls removed)."""

    def __init__(self, filter_length, hop_length, win_length, window="hann"):
        self.filter_length, self.hop_length, self.win_length, self.window = filter_length, hop_length, win_length, window
        self.forward_basis, self.inverse_basis = stft_bases(filter_length, hop_length, win_length, window)

    def transform(self, input_data: torch.Tensor):
        nb, ns = input_data.shape
        x = input_data.view(nb, 1, ns)
        p = int(self.filter_length / 2)
        x = F.pad(x.unsqueeze(1), (p, p, 0, 0), mode="reflect").squeeze(1)
        ft = F.conv1d(x, self.forward_basis, stride=self.hop_length, padding=0)
        cutoff = int(self.filter_length / 2 + 1)
        re, im = ft[:, :cutoff, :], ft[:, cutoff:, :]
        return torch.sqrt(re ** 2 + im ** 2), torch.atan2(im, re)

    def inverse(self, magnitude: torch.Tensor, phase: torch.Tensor):
        rec = torch.cat([magnitude * torch.cos(phase), magnitude * torch.sin(phase)], dim=1)
        inv = F.conv_transpose1d(rec, self.inverse_basis, stride=self.hop_length, padding=0)
        if self.window is not None:
            ws = window_sumsquare(self.window, magnitude.size(-1), hop_length=self.hop_length, win_length=self.win_length,
                                  n_fft=self.filter_length, dtype=np.float32)
            nz = torch.from_numpy(np.where(ws > tiny(ws))[0])
            ws = torch.from_numpy(ws)
            inv[:, :, nz] /= ws[nz]
            inv *= float(self.filter_length) / self.hop_length
        inv = inv[:, :, int(self.filter_length / 2):]
        inv = inv[:, :, :-int(self.filter_length / 2):]
        return inv


def dynamic_range_compression(x, C=1, clip_val=1e-5):
    return torch.log(torch.clamp(x, min=clip_val) * C)


def dynamic_range_decompression(x, C=1):
    return torch.exp(x) / C


class TacotronSTFT:
    """stft.py:130-178."""

    def __init__(self, filter_length, hop_length, win_length, n_mel_channels, sampling_rate, mel_fmin, mel_fmax):
        self.n_mel_channels, self.sampling_rate = n_mel_channels, sampling_rate
        self.stft_fn = STFT(filter_length, hop_length, win_length)
        self.mel_basis = torch.from_numpy(mel_filterbank(sampling_rate, filter_length, n_mel_channels, mel_fmin, mel_fmax)).float()

    def mel_spectrogram(self, y: torch.Tensor):
        assert torch.min(y) >= -1 and torch.max(y) <= 1
        mag, _ = self.stft_fn.transform(y)
        mel = dynamic_range_compression(torch.matmul(self.mel_basis, mag))
        return mel, torch.norm(mag, dim=1)


def griffin_lim(magnitudes: torch.Tensor, stft_fn: STFT, n_iters: int = 30, init_angles=None):
    """audio_processing.py:63-82.  init_angles replaces the reference's unseeded np.random draw
    (`np.angle(np.exp(2j*pi*rand))`, float32) so both sides start from the same phases."""
    if init_angles is None:
        init_angles = np.angle(np.exp(2j * np.pi * np.random.rand(*magnitudes.size()))).astype(np.float32)
    angles = torch.as_tensor(init_angles, dtype=torch.float32)
    signal = stft_fn.inverse(magnitudes, angles).squeeze(1)
    for _ in range(n_iters):
        _, angles = stft_fn.transform(signal)
        signal = stft_fn.inverse(magnitudes, angles).squeeze(1)
    return signal


def get_mel_from_wav(audio, _stft: TacotronSTFT):
    """tools.py:9-16."""
    audio = torch.clip(torch.FloatTensor(audio).unsqueeze(0), -1, 1)
    mel, energy = _stft.mel_spectrogram(audio)
    return torch.squeeze(mel, 0).numpy().astype(np.float32), torch.squeeze(energy, 0).numpy().astype(np.float32)


def inv_mel_spec(mel: torch.Tensor, _stft: TacotronSTFT, griffin_iters: int = 60, init_angles=None) -> np.ndarray:
    """tools.py:18-34 without the file write: mel [n_mel, T] (log-compressed) -> waveform [hop*(T-2)] float32."""
    mel = torch.stack([mel])
    dec = dynamic_range_decompression(mel).transpose(1, 2)
    spec = torch.mm(dec[0], _stft.mel_basis).transpose(0, 1).unsqueeze(0) * 1000
    audio = griffin_lim(spec[:, :, :-1], _stft.stft_fn, griffin_iters, init_angles)
    return audio.squeeze().numpy()
