"""SURVEY §8 row f3 on the CPU: (1) the audio oracle reproduces the goldens generated from the REAL reference modules
(`oracle/make_golden_audio.py`: audio/stft.py, audio_processing.py, tools.py), and (2) the product module
`meta_tts_b200.audio` — driven through the CPU restatement of the op set, which emulates the GEMM's TMA coordinate / tap /
out-of-bounds semantics — matches the oracle: validates the tap formulation of STFT / iSTFT, the padded frame-major
layouts and the Griffin-Lim loop without a GPU."""
import os

import numpy as np
import pytest
import torch

from meta_tts_b200 import audio as PA
from oracle import audio_oracle as A
from oracle.ops_reference import RefOps

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "audio_golden.npz"), allow_pickle=False)
CFG = dict(filter_length=1024, hop_length=256, win_length=1024, n_mel_channels=80, sampling_rate=22050, mel_fmin=0, mel_fmax=8000)


def _rel(a, b):
    a, b = torch.as_tensor(np.asarray(a)).double(), torch.as_tensor(np.asarray(b)).double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def test_oracle_matches_reference_goldens():
    t = A.TacotronSTFT(**CFG)
    y = torch.from_numpy(G["wave"])
    mag, ph = t.stft_fn.transform(y)
    assert np.array_equal(mag.numpy(), G["mag"]) and np.array_equal(ph.numpy(), G["phase"])
    assert np.array_equal(t.stft_fn.inverse(mag, ph).numpy(), G["recon"])
    mel, en = A.get_mel_from_wav(G["wave"][0], t)
    assert np.array_equal(mel, G["mel"]) and np.array_equal(en, G["energy"])
    assert np.array_equal(A.window_sumsquare("hann", 36, 256, 1024, 1024), G["window_sum"])
    for iters in (0, 3):
        a = A.inv_mel_spec(torch.from_numpy(G["mel"]), t, iters, init_angles=G["gl_init_angles"])
        assert np.array_equal(a, G[f"gl_audio_{iters}"])


def test_stft_reconstructs_signal():
    """Size-independent property: inverse(transform(x)) == x away from the edges (the basis pair is a pseudo-inverse)."""
    t = A.STFT(1024, 256, 1024)
    y = torch.from_numpy(G["wave"])
    rec = t.inverse(*t.transform(y))[:, 0]
    assert _rel(rec[:, 1024:-1024], y[:, 1024:rec.shape[1] - 1024]) < 1e-4


@pytest.fixture(scope="module")
def prod():
    return PA.TacotronSTFT(**CFG, backend=RefOps(split=3))


def test_product_constants_match_oracle(prod):
    t = A.TacotronSTFT(**CFG)
    assert torch.equal(prod.stft_fn.forward_basis, t.stft_fn.forward_basis)
    assert torch.equal(prod.stft_fn.inverse_basis, t.stft_fn.inverse_basis)
    assert torch.equal(prod.mel_basis, t.mel_basis) and np.array_equal(prod.mel_basis.numpy(), G["mel_basis"])
    assert np.array_equal(prod.stft_fn.window_sum(36).numpy(), G["window_sum"])


def test_product_transform_inverse(prod):
    y = torch.from_numpy(G["wave"])
    mag, ph = prod.stft_fn.transform(y)
    assert mag.shape == G["mag"].shape and ph.shape == G["phase"].shape
    assert _rel(mag, G["mag"]) < 2e-5
    # phases of near-silent bins are ill-conditioned: compare the complex spectrum instead of raw angles
    z, zg = mag * torch.exp(1j * ph), torch.from_numpy(G["mag"]) * torch.exp(1j * torch.from_numpy(G["phase"]))
    assert ((z - zg).abs().norm() / zg.abs().norm()).item() < 2e-5
    rec = prod.stft_fn.inverse(torch.from_numpy(G["mag"]), torch.from_numpy(G["phase"]))
    assert rec.shape == G["recon"].shape and _rel(rec, G["recon"]) < 2e-5


def test_product_ragged_length_and_batch(prod):
    """num_samples not a multiple of hop (tail samples beyond the last full frame are ignored, as conv1d does)."""
    rng = np.random.RandomState(0)
    y = torch.from_numpy((0.3 * rng.randn(3, 256 * 9 + 77)).astype(np.float32))
    t = A.STFT(1024, 256, 1024)
    mag_o, _ = t.transform(y)
    mag, _ = prod.stft_fn.transform(y)
    assert mag.shape == mag_o.shape and _rel(mag, mag_o) < 2e-5


def test_product_mel_spectrogram(prod):
    mel, en = PA.get_mel_from_wav(G["wave"][0], prod)
    assert mel.shape == G["mel"].shape and en.shape == G["energy"].shape
    assert np.abs(mel - G["mel"]).max() < 2e-3 and _rel(en, G["energy"]) < 2e-5     # log-compressed: absolute error


def test_product_griffin_lim(prod):
    for iters, tol in ((0, 2e-5), (3, 1e-3)):
        a = PA.inv_mel_spec(torch.from_numpy(G["mel"]), None, prod, iters, init_angles=G["gl_init_angles"])
        assert a.shape == G[f"gl_audio_{iters}"].shape
        assert _rel(a, G[f"gl_audio_{iters}"]) < tol, iters
    spec = torch.from_numpy(G["gl_spec"])[:, :, :-1]
    sig = PA.griffin_lim(spec, prod.stft_fn, 1, init_angles=G["gl_init_angles"])
    ref = A.griffin_lim(spec, A.STFT(1024, 256, 1024), 1, init_angles=G["gl_init_angles"])
    assert _rel(sig, ref) < 2e-4
