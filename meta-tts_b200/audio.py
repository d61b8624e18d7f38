"""Vocoder-side decode of BASELINE configs[4] on the B200 — drop-in for the reference's `audio/` package
(`audio/stft.py: STFT, TacotronSTFT`; `audio/audio_processing.py: griffin_lim, dynamic_range_*`;
`audio/tools.py: get_mel_from_wav, inv_mel_spec`): same class / function names, arguments and return layouts.

B200-first design (nothing here is a translation of the conv1d / conv_transpose1d calls):
  * `STFT.transform` (stft.py:54-81, a stride-`hop` conv1d with a [2*cutoff, n_fft] windowed Fourier basis) is ONE
    tcgen05 GEMM launch with n_fft/hop taps over the hop-reshaped reflect-padded signal `[B, rows, hop]`: tap j reads the
    signal rows shifted by j — the frames (4x overlapping) are never materialised, the taps accumulate in TMEM.
  * `STFT.inverse` (stft.py:83-122, conv_transpose1d = per-frame synthesis + overlap-add) is the matching "dgrad" form:
    output row r (hop samples) = sum_j X[r - j] * inv_basis_j; the overlap-add happens in the accumulator, rows outside
    the frame range read zero through TMA out-of-bounds fill.
  * operands are bf16 hi/lo pairs (bf16x3 => fp32-grade products, like the training path); magnitude / phase /
    recombination / window-sum normalisation / reflect padding are HBM-bound elementwise kernels (csrc/mtts_audio.cu).
  * internal layout is FRAME-major (`[B, n_frames, ...]`, rows 16-byte aligned: real parts in columns [0, cutoff),
    imaginary parts in [im_off, im_off + cutoff)); the reference's bins-major `[B, cutoff, n_frames]` tensors exist only
    as transposed views at the public boundary.
The Fourier bases, the window-sum envelope and the mel filterbank are constants built once on the host exactly as the
reference builds them (numpy float64 -> float32); `librosa` (absent here, unpinned in the reference) is replaced by its
published formulas (Slaney mel scale / area normalisation, pad_center, tiny).
"""
from __future__ import annotations

from typing import Dict, Optional

import math

import numpy as np
import torch
from scipy.signal import get_window

from . import lib as L
from .engine import _pick_cfg
from .ops import Opnd, _stream


def _rup(x: int, m: int) -> int:
    return (x + m - 1) // m * m


def _pad_center(data: np.ndarray, size: int) -> np.ndarray:
    """librosa.util.pad_center for 1-D input (stft.py:43, audio_processing.py:54)."""
    n = data.shape[-1]
    lpad = int((size - n) // 2)
    return np.pad(data, (lpad, int(size - n - lpad)), mode="constant")


def _mel_filterbank(sr: int, n_fft: int, n_mels: int, fmin: float, fmax) -> np.ndarray:
    """librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax), defaults htk=False / norm='slaney' (stft.py:145-147)."""
    f_sp, min_log_hz = 200.0 / 3, 1000.0
    min_log_mel, logstep = min_log_hz / f_sp, np.log(6.4) / 27.0
    hz2mel = lambda f: np.where(np.asarray(f, np.float64) >= min_log_hz,  # noqa: E731
                                min_log_mel + np.log(np.maximum(np.asarray(f, np.float64), 1e-30) / min_log_hz) / logstep,
                                np.asarray(f, np.float64) / f_sp)
    mel2hz = lambda m: np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)  # noqa: E731
    fmax = sr / 2.0 if fmax is None else fmax
    freqs = np.linspace(0, sr / 2.0, 1 + n_fft // 2)
    mel_f = mel2hz(np.linspace(hz2mel(fmin), hz2mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, freqs)
    w = np.zeros((n_mels, len(freqs)))
    for i in range(n_mels):
        w[i] = np.maximum(0, np.minimum(-ramps[i] / fdiff[i], ramps[i + 2] / fdiff[i + 1]))
    w *= (2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels]))[:, None]
    return w.astype(np.float32)


def _default_backend(device, split):
    from .ops import CudaOps           # raises without a B200: there is no CPU fallback

    return CudaOps(split=split, device=device)


class STFT:
    """audio/stft.py:15-127.  `transform(input_data[B, N]) -> (magnitude[B, cutoff, F], phase[B, cutoff, F])`,
    `inverse(magnitude, phase) -> [B, 1, hop*(F-1)]`, `forward(x) = inverse(*transform(x))`."""

    def __init__(self, filter_length, hop_length, win_length, window="hann", *, backend=None, device="cuda:0", split: int = 3):
        assert filter_length % hop_length == 0, "the tap formulation needs hop | n_fft (reference configs: 1024 / 256)"
        self.filter_length, self.hop_length, self.win_length, self.window = filter_length, hop_length, win_length, window
        self.be = backend if backend is not None else _default_backend(device, split)
        self.device = self.be.device
        self.taps = filter_length // hop_length
        self.cutoff = int(filter_length / 2 + 1)
        self.im_off = _rup(self.cutoff, 8)            # 16-byte aligned start of the imaginary half
        self.NP = 2 * self.im_off                     # padded [re | im] row
        # --- constants, as stft.py:25-52 ---
        scale = filter_length / hop_length
        fourier = np.fft.fft(np.eye(filter_length))
        fourier = np.vstack([np.real(fourier[:self.cutoff, :]), np.imag(fourier[:self.cutoff, :])])
        fwd = torch.FloatTensor(fourier)                                            # [2*cutoff, n_fft]
        inv = torch.FloatTensor(np.linalg.pinv(scale * fourier).T)                  # [2*cutoff, n_fft]
        if window is not None:
            assert filter_length >= win_length
            w = torch.from_numpy(_pad_center(get_window(window, win_length, fftbins=True), filter_length)).float()
            fwd, inv = fwd * w, inv * w
        self.forward_basis = fwd[:, None, :].float()        # reference buffer layouts (state_dict compatibility)
        self.inverse_basis = inv[:, None, :].float()
        T, H, NP, c, io = self.taps, hop_length, self.NP, self.cutoff, self.im_off
        # per-tap GEMM operands: Fw[j][n][k] = fwd[n, j*H + k] (rows padded to NP);  Iw[j][k][n] = inv[n, j*H + k]
        Fw = torch.zeros(T, NP, H)
        Iw = torch.zeros(T, H, NP)
        for j in range(T):
            Fw[j, :c], Fw[j, io:io + c] = fwd[:c, j * H:(j + 1) * H], fwd[c:, j * H:(j + 1) * H]
            Iw[j, :, :c], Iw[j, :, io:io + c] = inv[:c, j * H:(j + 1) * H].t(), inv[c:, j * H:(j + 1) * H].t()
        self._Fw = self._operand(Fw)
        self._Iw = self._operand(Iw)
        self._wsum: Dict[int, torch.Tensor] = {}
        self.num_samples = None
        self.launches = 0

    # ---- plumbing ----
    def _operand(self, t: torch.Tensor):
        t = t.contiguous().to(self.device)
        hi = t.to(torch.bfloat16)
        lo = (t - hi.float()).to(torch.bfloat16) if self.be.split == 3 else None     # init-time constant preparation
        return hi, lo

    def _bf(self, shape):
        return (self.be.empty(shape, torch.bfloat16), self.be.empty(shape, torch.bfloat16) if self.be.split == 3 else None)

    def n_frames(self, num_samples: int) -> int:
        return num_samples // self.hop_length + 1           # conv1d(stride=hop) over N + n_fft samples (stft.py:67-72)

    def window_sum(self, n_frames: int) -> torch.Tensor:
        """audio_processing.py:7-60 (window_sumsquare), cached per frame count; zeros when window is None."""
        ws = self._wsum.get(n_frames)
        if ws is None:
            n = self.filter_length + self.hop_length * (n_frames - 1)
            x = np.zeros(n, dtype=np.float32)
            if self.window is not None:
                win_sq = _pad_center(get_window(self.window, self.win_length, fftbins=True) ** 2, self.filter_length)
                for i in range(n_frames):
                    s = i * self.hop_length
                    x[s:min(n, s + self.filter_length)] += win_sq[:max(0, min(self.filter_length, n - s))]
            ws = self._wsum[n_frames] = torch.from_numpy(x).to(self.device)
        return ws

    # ---- frame-major core ----
    def transform_fm(self, x: torch.Tensor) -> torch.Tensor:
        """x [B, N] f32 (device) -> ri [B, F, NP] f32: reflect pad + split, then the 4-tap Fourier-basis GEMM."""
        be, H, NP = self.be, self.hop_length, self.NP
        B, N = x.shape
        F_ = self.n_frames(N)
        rows = F_ + self.taps - 1
        sig_hi, sig_lo = self._bf((B, rows * H))
        be.reflect_pad(x, B, N, self.filter_length // 2, rows * H, None, sig_hi, sig_lo)
        ri = be.empty((B, F_, NP))
        a = Opnd(sig_hi, sig_lo, L.MAJOR_K, (H, rows, B), (1, H, rows * H), src2=L.SRC_Z0, shift_src=L.SRC_TAP, shift_base=0,
                 shift_step=1)
        b = Opnd(self._Fw[0], self._Fw[1], L.MAJOR_K, (H, NP, self.taps), (1, H, NP * H), src2=L.SRC_TAP)
        bn, pair, _ = _pick_cfg((F_ + 127) // 128, B, NP, 0, False)
        be.gemm(a, b, F_, NP, H, c_f32=ri, ldc=NP, c_sz0=F_ * NP, ntaps=self.taps, nz0=B, block_n=bn, pair=pair)
        self.launches += 2
        return ri

    def inverse_fm(self, X_hi: torch.Tensor, X_lo: Optional[torch.Tensor], B: int, F_: int) -> torch.Tensor:
        """X [B, F, NP] (bf16 hi/lo of [mag cos | mag sin]) -> signal [B, hop*(F-1)] f32."""
        be, H, NP = self.be, self.hop_length, self.NP
        rows = F_ + self.taps - 1
        ola = be.empty((B, rows * H))
        a = Opnd(X_hi, X_lo, L.MAJOR_K, (NP, F_, B), (1, NP, F_ * NP), src2=L.SRC_Z0, shift_src=L.SRC_TAP, shift_base=0,
                 shift_step=-1)
        b = Opnd(self._Iw[0], self._Iw[1], L.MAJOR_K, (NP, H, self.taps), (1, NP, H * NP), src2=L.SRC_TAP)
        # few output tiles (rows/128 x B), long contraction (taps * NP): split it across CTAs, partial sums meet in fp32 reds
        bn, pair, ks = _pick_cfg((rows + 127) // 128, B, H, self.taps * ((NP + 63) // 64), True)
        if ks > 1:
            be.zero_(ola)
        be.gemm(a, b, rows, H, NP, c_f32=ola, ldc=H, c_sz0=rows * H, ntaps=self.taps, nz0=B, block_n=bn, pair=pair, ksplit=ks,
                flags=L.EPI_ACCUM if ks > 1 else 0)
        n = rows * H                                   # = n_fft + hop*(F-1)  (conv_transpose1d output length)
        out = be.empty((B, n - self.filter_length))
        if self.window is not None:
            be.istft_finish(ola, self.window_sum(F_), float(np.finfo(np.float32).tiny), float(self.filter_length) / H, B, n,
                            self.filter_length // 2, out)
        else:
            be.istft_finish(ola, self.window_sum(F_), 0.0, 1.0, B, n, self.filter_length // 2, out)
        self.launches += 2
        return out

    def polar_fm(self, ri: torch.Tensor, want_phase=True, want_energy=False, want_split=False):
        """ri [B, F, NP] -> (mag [B, F, im_off], phase or None, energy [B, F] or None, (mag_hi, mag_lo) or None)."""
        be = self.be
        B, F_, _ = ri.shape
        mag = be.empty((B, F_, self.im_off))
        phase = be.empty((B, F_, self.im_off)) if want_phase else None
        energy = be.empty((B, F_)) if want_energy else None
        mh, ml = self._bf((B, F_, self.im_off)) if want_split else (None, None)
        be.stft_polar(ri, B * F_, self.cutoff, self.NP, self.im_off, self.im_off, mag, phase, energy, mh, ml)
        self.launches += 1
        return mag, phase, energy, (mh, ml)

    def recombine_fm(self, mag: torch.Tensor, phase: Optional[torch.Tensor], ri: Optional[torch.Tensor]):
        B, F_, _ = mag.shape
        X_hi, X_lo = self._bf((B, F_, self.NP))
        self.be.stft_recombine(mag, phase, ri, B * F_, self.cutoff, self.NP, self.im_off, self.im_off, X_hi, X_lo)
        self.launches += 1
        return X_hi, X_lo

    def _to_fm(self, t: torch.Tensor) -> torch.Tensor:
        """Boundary plumbing: reference layout [B, cutoff, F] -> frame-major padded [B, F, im_off] on the device."""
        t = torch.as_tensor(t, dtype=torch.float32)
        B, c, F_ = t.shape
        out = torch.zeros((B, F_, self.im_off), dtype=torch.float32, device=self.device)
        out[:, :, :c] = t.to(self.device).transpose(1, 2)
        return out

    # ---- reference API ----
    def transform(self, input_data):
        x = torch.as_tensor(input_data, dtype=torch.float32).to(self.device).contiguous()
        self.num_samples = x.size(1)
        mag, phase, _, _ = self.polar_fm(self.transform_fm(x))
        c = self.cutoff
        return mag[:, :, :c].transpose(1, 2), phase[:, :, :c].transpose(1, 2)

    def inverse(self, magnitude, phase):
        mag, ph = self._to_fm(magnitude), self._to_fm(phase)
        B, F_, _ = mag.shape
        X_hi, X_lo = self.recombine_fm(mag, ph, None)
        return self.inverse_fm(X_hi, X_lo, B, F_)[:, None, :]

    def forward(self, input_data):
        self.magnitude, self.phase = self.transform(input_data)
        return self.inverse(self.magnitude, self.phase)

    __call__ = forward


def dynamic_range_compression(x, C=1, clip_val=1e-5):
    """audio_processing.py:85-91: log(clamp(x, min=clip_val) * C).  CUDA tensors go through mtts_unary (no torch kernels on the
    device path); host tensors (the reference calls it on numpy-backed tensors during preprocessing) use torch on the CPU."""
    if x.is_cuda:
        xin = x.contiguous().float()
        out = torch.empty_like(xin)
        L.call("mtts_unary", L.UN_LOGCLAMP, xin.data_ptr(), xin.numel(), float(clip_val), float(C), out.data_ptr(), None, None, _stream())
        return out
    return torch.log(torch.clamp(x, min=clip_val) * C)


def dynamic_range_decompression(x, C=1):
    """audio_processing.py:94-100: exp(x) / C (CUDA tensors: mtts_unary)."""
    if x.is_cuda:
        xin = x.contiguous().float()
        out = torch.empty_like(xin)
        L.call("mtts_unary", L.UN_EXP, xin.data_ptr(), xin.numel(), 1.0 / float(C), 0.0, out.data_ptr(), None, None, _stream())
        return out
    return torch.exp(x) / C


class TacotronSTFT:
    """audio/stft.py:130-178: `mel_spectrogram(y[B, N]) -> (mel[B, n_mel, F], energy[B, F])`."""

    def __init__(self, filter_length, hop_length, win_length, n_mel_channels, sampling_rate, mel_fmin, mel_fmax, *,
                 backend=None, device="cuda:0", split: int = 3):
        self.n_mel_channels, self.sampling_rate = n_mel_channels, sampling_rate
        self.stft_fn = STFT(filter_length, hop_length, win_length, backend=backend, device=device, split=split)
        st = self.stft_fn
        mb = torch.from_numpy(_mel_filterbank(sampling_rate, filter_length, n_mel_channels, mel_fmin, mel_fmax)).float()
        self.mel_basis = mb                                           # [n_mel, cutoff], the reference buffer
        self.NM = _rup(n_mel_channels, 8)
        fwd = torch.zeros(self.NM, st.im_off)                          # B operand of mag -> mel   ([N = mel, K = bin])
        fwd[:n_mel_channels, :st.cutoff] = mb
        invb = torch.zeros(st.im_off, self.NM)                         # B operand of mel -> mag   ([N = bin, K = mel])
        invb[:st.cutoff, :n_mel_channels] = mb.t()
        self._mb_fwd, self._mb_inv = st._operand(fwd), st._operand(invb)

    def spectral_normalize(self, magnitudes):
        return dynamic_range_compression(magnitudes)

    def spectral_de_normalize(self, magnitudes):
        return dynamic_range_decompression(magnitudes)

    def mel_fm(self, x: torch.Tensor):
        """x [B, N] -> (mel [B, F, NM] f32 log-compressed, energy [B, F])."""
        st, be = self.stft_fn, self.stft_fn.be
        ri = st.transform_fm(x)
        B, F_, _ = ri.shape
        _, _, energy, (mh, ml) = st.polar_fm(ri, want_phase=False, want_energy=True, want_split=True)
        R, K, NM = B * F_, st.im_off, self.NM
        lin = be.empty((B, F_, NM))
        a = Opnd(mh, ml, L.MAJOR_K, (K, R), (1, K))
        b = Opnd(self._mb_fwd[0], self._mb_fwd[1], L.MAJOR_K, (K, NM), (1, K))
        bn, pair, _ = _pick_cfg((R + 127) // 128, 1, NM, 0, False)
        be.gemm(a, b, R, NM, K, c_f32=lin, ldc=NM, block_n=bn, pair=pair)
        mel = be.empty((B, F_, NM))
        be.unary(L.UN_LOGCLAMP, lin, 1e-5, 1.0, mel)                  # dynamic_range_compression(C=1, clip_val=1e-5)
        st.launches += 2
        return mel, energy

    def mel_spectrogram(self, y):
        y = torch.as_tensor(y, dtype=torch.float32)
        assert torch.min(y) >= -1 and torch.max(y) <= 1              # stft.py:166-167
        mel, energy = self.mel_fm(y.to(self.stft_fn.device).contiguous())
        return mel[:, :, :self.n_mel_channels].transpose(1, 2), energy

    def spec_from_mel_fm(self, mel: torch.Tensor, scaling: float = 1000.0) -> torch.Tensor:
        """mel [B, T, n_mel] f32 (log-compressed, device) -> linear magnitudes [B, T, im_off] = (exp(mel) @ mel_basis)*scaling
        (tools.py:20-25)."""
        st, be = self.stft_fn, self.stft_fn.be
        B, T, nm = mel.shape
        assert nm == self.NM, "mel rows must be padded to a multiple of 8 channels"
        dh, dl = st._bf((B, T, nm))
        be.unary(L.UN_EXP, mel, 1.0, 1.0, None, dh, dl)               # dynamic_range_decompression(C=1)
        R, K = B * T, st.im_off
        spec = be.empty((B, T, K))
        a = Opnd(dh, dl, L.MAJOR_K, (nm, R), (1, nm))
        b = Opnd(self._mb_inv[0], self._mb_inv[1], L.MAJOR_K, (nm, K), (1, nm))
        bn, pair, _ = _pick_cfg((R + 127) // 128, 1, K, 0, False)
        be.gemm(a, b, R, K, nm, c_f32=spec, ldc=K, alpha=scaling, block_n=bn, pair=pair)
        st.launches += 2
        return spec


def griffin_lim_fm(mag: torch.Tensor, stft_fn: STFT, n_iters: int, init_angles: torch.Tensor, use_graph: Optional[bool] = None) -> torch.Tensor:
    """Frame-major Griffin-Lim (audio_processing.py:63-82): mag, init_angles [B, F, im_off] on the device -> signal
    [B, hop*(F-1)].  Per iteration: reflect-pad+split, Fourier GEMM, angle+recombine, inverse GEMM, normalise = 5 kernels
    (+ a memset when the inverse GEMM is split along its contraction).  The iteration is a fixed launch sequence on fixed
    shapes, so it CAN be captured in a CUDA graph (`use_graph=True`: signal buffer updated in place, one capture + n-1
    replays).  Measured on a B200 at 931 frames x 60 iterations (tools/config5_bench.py): eager 7.2 ms (120 us per
    iteration: the host keeps up), so eager is the default; a capture per call does not pay for a single utterance.
    (Round 1 recorded 23.9 ms for `inv_mel_spec` and blamed the capture: it was the reference-style HOST draw of the
    initial angles — since moved to the device, `inv_mel_spec` now takes 7.4 ms.)"""
    B, F_, _ = mag.shape
    X_hi, X_lo = stft_fn.recombine_fm(mag, init_angles, None)
    signal = stft_fn.inverse_fm(X_hi, X_lo, B, F_)

    def iteration(sig):
        ri = stft_fn.transform_fm(sig)
        xh, xl = stft_fn.recombine_fm(mag, None, ri)                  # keeps only the angles of the new transform
        return stft_fn.inverse_fm(xh, xl, B, F_)

    if not use_graph:
        for _ in range(n_iters):
            signal = iteration(signal)
        return signal
    signal.copy_(iteration(signal))                                   # iteration 1, eager: warms allocator and kernel attributes
    stft_fn.window_sum(F_)                                            # (cached constant: must exist before capture)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        signal.copy_(iteration(signal))
    for _ in range(n_iters - 1):
        graph.replay()
    return signal


def griffin_lim(magnitudes, stft_fn: STFT, n_iters=30, init_angles=None):
    """audio/audio_processing.py:63-82: magnitudes [B, cutoff, F] -> signal [B, hop*(F-1)].
    `init_angles` (same layout) replaces the reference's unseeded host draw when given."""
    magnitudes = torch.as_tensor(magnitudes, dtype=torch.float32)
    if init_angles is None:
        init_angles = np.angle(np.exp(2j * np.pi * np.random.rand(*magnitudes.size()))).astype(np.float32)
    return griffin_lim_fm(stft_fn._to_fm(magnitudes), stft_fn, n_iters, stft_fn._to_fm(torch.as_tensor(init_angles)))


def get_mel_from_wav(audio, _stft: TacotronSTFT):
    """audio/tools.py:9-16 -> (mel [n_mel, F] float32 ndarray, energy [F] float32 ndarray)."""
    audio = torch.clip(torch.FloatTensor(audio).unsqueeze(0), -1, 1)
    melspec, energy = _stft.mel_spectrogram(audio)
    return (torch.squeeze(melspec, 0).cpu().numpy().astype(np.float32),
            torch.squeeze(energy, 0).cpu().numpy().astype(np.float32))


def inv_mel_spec(mel, out_filename, _stft: TacotronSTFT, griffin_iters=60, init_angles=None):
    """audio/tools.py:18-34: mel [n_mel, T] (log-compressed) -> Griffin-Lim waveform, written to `out_filename` as the
    reference does (scipy.io.wavfile.write) unless it is None; returns the waveform (float32 ndarray).
    The reference dereferences `_stft._stft_fn` (tools.py:28), which does not exist; `_stft.stft_fn` is used."""
    st = _stft.stft_fn
    mel = torch.as_tensor(mel, dtype=torch.float32)
    T = mel.shape[1]
    mel_fm = torch.zeros((1, T, _stft.NM), dtype=torch.float32, device=st.device)
    mel_fm[0, :, :mel.shape[0]] = mel.to(st.device).t()
    spec = _stft.spec_from_mel_fm(mel_fm)[:, :T - 1].contiguous()                  # spec_from_mel[:, :, :-1]
    if init_angles is None:
        # the reference draws np.angle(np.exp(2j*pi*U)), U ~ uniform[0,1) unseeded, on the host (audio_processing.py:70-72): the same
        # distribution — 2*pi*U wrapped into (-pi, pi] — drawn on the device (the host draw + H2D was 17 of inv_mel_spec's 25 ms)
        ang = 2.0 * math.pi * torch.rand((1, T - 1, st.cutoff), dtype=torch.float32, device=st.device)
        ang = torch.where(ang > math.pi, ang - 2.0 * math.pi, ang)
        angles_fm = torch.zeros((1, T - 1, st.im_off), dtype=torch.float32, device=st.device)
        angles_fm[:, :, :st.cutoff] = ang
    else:
        angles_fm = st._to_fm(torch.as_tensor(init_angles))
    audio = griffin_lim_fm(spec, st, griffin_iters, angles_fm)
    audio = audio.squeeze().cpu().numpy()
    if out_filename is not None:
        from scipy.io.wavfile import write

        write(out_filename, _stft.sampling_rate, audio)
    return audio
