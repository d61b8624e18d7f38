"""SURVEY §8 row f3 on the B200 (through the C ABI): STFT / iSTFT / mel / Griffin-Lim against the goldens generated from
the REAL reference audio modules and against the CPU oracle at BASELINE configs[4] size (864 frames, 60 iterations).
Tolerance: fp32 relative 1e-3 (north_star); the single-transform products agree to ~1e-6 (bf16x3)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from meta_tts_b200 import audio as PA  # noqa: E402
from meta_tts_b200 import ops as _ops  # noqa: E402
from oracle import audio_oracle as A  # noqa: E402

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "audio_golden.npz"), allow_pickle=False)
CFG = dict(filter_length=1024, hop_length=256, win_length=1024, n_mel_channels=80, sampling_rate=22050, mel_fmin=0, mel_fmax=8000)


def _rel(a, b):
    a = torch.as_tensor(np.asarray(a.cpu() if torch.is_tensor(a) else a)).double()
    b = torch.as_tensor(np.asarray(b)).double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def prod(cuda_device):
    return PA.TacotronSTFT(**CFG, device="cuda:0")


def test_transform_inverse_vs_reference_goldens(prod):
    y = torch.from_numpy(G["wave"])
    n0 = _ops.launch_count
    mag, ph = prod.stft_fn.transform(y)
    torch.cuda.synchronize()
    assert _ops.launch_count - n0 == 3, "reflect-pad+split, Fourier GEMM, polar"
    assert tuple(mag.shape) == G["mag"].shape
    r_mag = _rel(mag, G["mag"])
    z = (mag * torch.exp(1j * ph)).cpu()
    zg = torch.from_numpy(G["mag"]) * torch.exp(1j * torch.from_numpy(G["phase"]))
    r_z = ((z - zg).abs().norm() / zg.abs().norm()).item()
    rec = prod.stft_fn.inverse(torch.from_numpy(G["mag"]), torch.from_numpy(G["phase"]))
    r_rec = _rel(rec, G["recon"])
    print(f"[audio] transform mag rel {r_mag:.2e} complex rel {r_z:.2e}; inverse rel {r_rec:.2e}")
    assert r_mag < 2e-5 and r_z < 2e-5 and r_rec < 2e-5
    # round trip property on the device path itself
    back = prod.stft_fn.forward(y)[:, 0].cpu()
    assert _rel(back[:, 1024:-1024], y[:, 1024:back.shape[1] - 1024]) < 1e-4


def test_ragged_batch_and_single_frame_edges(prod):
    rng = np.random.RandomState(0)
    t = A.STFT(1024, 256, 1024)
    for B, N in ((3, 256 * 9 + 77), (1, 513), (5, 256 * 130)):            # N > n_fft/2 is the reflect-pad limit
        y = torch.from_numpy((0.3 * rng.randn(B, N)).astype(np.float32))
        mag_o, _ = t.transform(y)
        mag, _ = prod.stft_fn.transform(y)
        assert tuple(mag.shape) == tuple(mag_o.shape) and _rel(mag, mag_o.numpy()) < 2e-5, (B, N)
    with pytest.raises(Exception):
        prod.stft_fn.transform(torch.zeros(1, 512))                      # torch's reflect pad rejects pad >= N too


def test_mel_spectrogram_vs_reference_goldens(prod):
    mel, en = PA.get_mel_from_wav(G["wave"][0], prod)
    print(f"[audio] mel max abs err {np.abs(mel - G['mel']).max():.2e}, energy rel {_rel(en, G['energy']):.2e}")
    assert mel.shape == G["mel"].shape and np.abs(mel - G["mel"]).max() < 2e-3 and _rel(en, G["energy"]) < 2e-5


def test_griffin_lim_vs_reference_goldens(prod):
    for iters, tol in ((0, 2e-5), (3, 1e-3), (30, 1e-3)):
        a = PA.inv_mel_spec(torch.from_numpy(G["mel"]), None, prod, iters, init_angles=G["gl_init_angles"])
        r = _rel(a, G[f"gl_audio_{iters}"])
        print(f"[audio] griffin-lim {iters} iterations: rel {r:.2e}")
        assert a.shape == G[f"gl_audio_{iters}"].shape and r < tol
    # the CUDA-graph form of the loop (one captured iteration replayed) gives the same waveform as the eager loop
    st = prod.stft_fn
    spec = st._to_fm(torch.from_numpy(G["gl_spec"])[:, :, :-1])
    ang = st._to_fm(torch.from_numpy(G["gl_init_angles"]))
    eager = PA.griffin_lim_fm(spec, st, 8, ang, use_graph=False).clone()
    graph = PA.griffin_lim_fm(spec, st, 8, ang, use_graph=True)
    assert _rel(graph, eager.cpu().numpy()) < 2e-4      # (not bit-equal: the split-K inverse GEMM reduces with fp32 atomics)


def test_config5_size_griffin_lim_60_iterations(prod):
    """BASELINE configs[4] decode size: one 864-frame mel, 60 Griffin-Lim iterations (tools.py:18: griffin_iters=60)."""
    rng = np.random.RandomState(5)
    t = A.TacotronSTFT(**CFG)
    from oracle.make_golden_audio import synth_wave
    wav = synth_wave(256 * 863, 7)
    mel, _ = A.get_mel_from_wav(wav, t)                                    # [80, 864]
    assert mel.shape == (80, 864)
    ang = np.angle(np.exp(2j * np.pi * rng.rand(1, 513, 863))).astype(np.float32)
    ref = A.inv_mel_spec(torch.from_numpy(mel), t, 60, init_angles=ang)
    n0 = _ops.launch_count
    got = PA.inv_mel_spec(torch.from_numpy(mel), None, prod, 60, init_angles=ang)
    torch.cuda.synchronize()
    r = _rel(got, ref)
    print(f"[audio] config-5 size Griffin-Lim x60: rel {r:.2e}, {_ops.launch_count - n0} launches")
    assert got.shape == ref.shape == (256 * 862,) and r < 1e-3
    # size-independent property: spectral convergence — the STFT magnitude of the result approaches the target
    spec = prod.spec_from_mel_fm(torch.from_numpy(mel).t()[None].contiguous().to("cuda:0"))
    tgt = spec[0, :863, :513].cpu()

    def sc(x):
        m, _ = prod.stft_fn.transform(torch.from_numpy(x)[None])
        m = m[0].t().cpu()
        return ((m - tgt).norm() / tgt.norm()).item()
    e0 = sc(PA.inv_mel_spec(torch.from_numpy(mel), None, prod, 0, init_angles=ang))
    e60 = sc(got)
    print(f"[audio] spectral convergence: {e0:.3f} (0 iterations) -> {e60:.3f} (60 iterations)")
    assert e60 < 0.5 * e0          # CPU oracle: 0.609 -> 0.180
