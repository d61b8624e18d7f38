"""Synthetic LibriTTS-shaped workloads for benchmarks and smoke runs (there is no corpus and no checkpoint
offline): the reference's 12-tuple wire format (lightning/collate.py:47-60) filled with seeded random content, and a
random-init `state_dict` of the reference architecture (PyTorch default initialisers, as the reference's constructors use).
Product-side so that `bench.py`'s measured arm never touches `oracle/`."""
from __future__ import annotations

import json
import os
import tempfile

import torch

N_MEL, N_SYMBOLS = 80, 360
DEFAULT_STATS = {"pitch": [-2.9, 10.2, 180.0, 50.0], "energy": [-1.4, 8.6, 30.0, 20.0]}      # LibriTTS-like normalised ranges


def synth_batch(n: int, L: int, T: int, seed: int, speaker: int = 0, ragged: bool = False, n_speaker: int = 16):
    """n utterances: texts ~ U{1..360}; positive integer durations summing to the utterance's mel length (1 + a random
    share of the T - L spare frames); mels, pitch, energy ~ N(0, 1); utterance 0 has the full (L, T), the others are
    shorter when `ragged`."""
    g = torch.Generator().manual_seed(seed)
    texts, durs = torch.zeros(n, L, dtype=torch.long), torch.zeros(n, L, dtype=torch.long)
    mels, pitch, energy = torch.zeros(n, T, N_MEL), torch.zeros(n, L), torch.zeros(n, L)
    src_lens, mel_lens = torch.zeros(n, dtype=torch.long), torch.zeros(n, dtype=torch.long)
    for i in range(n):
        Li = L if (not ragged or i == 0) else int(torch.randint(max(2, L // 2), L + 1, (1,), generator=g))
        Ti = T if (not ragged or i == 0) else int(torch.randint(max(Li, T // 2), T + 1, (1,), generator=g))
        w = torch.rand(Li, generator=g) + 0.2
        d = 1 + torch.floor((Ti - Li) * w / w.sum()).long()
        d[0] += Ti - int(d.sum())
        texts[i, :Li] = torch.randint(1, N_SYMBOLS + 1, (Li,), generator=g)
        durs[i, :Li] = d
        mels[i, :Ti] = torch.randn(Ti, N_MEL, generator=g)
        pitch[i, :Li] = torch.randn(Li, generator=g)
        energy[i, :Li] = torch.randn(Li, generator=g)
        src_lens[i], mel_lens[i] = Li, Ti
    ids = [f"synth-{seed}-{i}" for i in range(n)]
    return (ids, ids, torch.full((n,), speaker % n_speaker, dtype=torch.long), texts, src_lens, int(src_lens.max()), mels, mel_lens,
            int(mel_lens.max()), pitch, energy, durs)


def synth_task(task: int, shots: int, queries: int, L: int, T: int, rank: int = 0, ragged: bool = False):
    """(support 12-tuple, query 12-tuple) of one speaker task; generator seed = 1000*task + 10*is_query + rank."""
    return (synth_batch(shots, L, T, seed=1000 * task + rank, speaker=task, ragged=ragged),
            synth_batch(queries, L, T, seed=1000 * task + 10 + rank, speaker=task, ragged=ragged))


def init_state_dict(model_config, n_speaker: int = 16, seed: int = 0, stats=None):
    """Random-init weights of the reference architecture: the drop-in `modules.FastSpeech2` constructed under
    torch.manual_seed(seed) (nn.Linear / nn.Conv1d / nn.Embedding / LayerNorm / BatchNorm default initialisers, sinusoid
    position tables, linspace pitch / energy bins)."""
    from .modules import FastSpeech2

    d = tempfile.mkdtemp(prefix="mtts_synth_")
    with open(os.path.join(d, "stats.json"), "w") as f:
        json.dump(stats or DEFAULT_STATS, f)
    with open(os.path.join(d, "speakers.json"), "w") as f:
        json.dump({f"spk{i}": i for i in range(n_speaker)}, f)
    pre = {"path": {"preprocessed_path": d},
           "preprocessing": {"pitch": {"feature": "phoneme_level"}, "energy": {"feature": "phoneme_level"}, "mel": {"n_mel_channels": N_MEL}}}
    algo = {"adapt": {"speaker_emb": "table"}}
    with torch.random.fork_rng():
        torch.manual_seed(seed)
        model = FastSpeech2(pre, model_config, algo)
    return {k: v.detach().clone() for k, v in model.state_dict().items()}
