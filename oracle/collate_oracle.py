"""CPU restatement of the reference's batch producer — TEST INFRASTRUCTURE ONLY (see oracle/fs2_oracle.py's header).

Follows lightning/collate.py:9-60 (`reprocess`) and utils/tools.py:270-301 (`pad_1D`, `pad_2D`) literally (np.pad per
utterance, np.stack, torch.from_numpy with the reference's dtype conversions).  Pinned: oracle/make_golden.py imports the
REAL modules and stores their outputs on a seeded synthetic dataset in tests/golden/collate_golden.npz.
"""
import numpy as np
import torch


def pad_1D(inputs, PAD=0):                                   # utils/tools.py:270-281
    max_len = max(len(x) for x in inputs)
    return np.stack([np.pad(x, (0, max_len - x.shape[0]), mode="constant", constant_values=PAD) for x in inputs])


def pad_2D(inputs, maxlen=None):                              # utils/tools.py:284-301
    def pad(x, max_len):
        if np.shape(x)[0] > max_len:
            raise ValueError("not max_len")
        s = np.shape(x)[1]
        return np.pad(x, (0, max_len - np.shape(x)[0]), mode="constant", constant_values=0)[:, :s]
    max_len = maxlen if maxlen else max(np.shape(x)[0] for x in inputs)
    return np.stack([pad(x, max_len) for x in inputs])


def reprocess(data, idxs):                                    # lightning/collate.py:9-60 (table speaker embedding branch)
    ids = [data[i]["id"] for i in idxs]
    speakers = np.array([data[i]["speaker"] for i in idxs])
    texts = [data[i]["text"] for i in idxs]
    raw_texts = [data[i]["raw_text"] for i in idxs]
    mels = [data[i]["mel"] for i in idxs]
    text_lens = np.array([t.shape[0] for t in texts])
    mel_lens = np.array([m.shape[0] for m in mels])
    return (ids, raw_texts, torch.from_numpy(speakers).long(), torch.from_numpy(pad_1D(texts)).long(), torch.from_numpy(text_lens),
            max(text_lens), torch.from_numpy(pad_2D(mels)).float(), torch.from_numpy(mel_lens), max(mel_lens),
            torch.from_numpy(pad_1D([data[i]["pitch"] for i in idxs])).float(),
            torch.from_numpy(pad_1D([data[i]["energy"] for i in idxs])),
            torch.from_numpy(pad_1D([data[i]["duration"] for i in idxs])).long())


def synth_dataset(n=9, seed=0, n_mel=80, lmin=3, lmax=17):
    """Dataset items as lightning/dataset.py:47-70 yields them (id, speaker, text, raw_text, mel, pitch, energy, duration)."""
    rng = np.random.RandomState(seed)
    data = []
    for i in range(n):
        L = int(rng.randint(lmin, lmax + 1))
        dur = rng.randint(0, 6, size=L).astype(np.int64)          # zero durations included
        dur[0] = max(dur[0], 1)
        T = int(dur.sum())
        data.append({"id": f"utt{i}", "speaker": int(rng.randint(0, 16)), "text": rng.randint(1, 361, size=L).astype(np.int64),
                     "raw_text": f"raw text {i}", "mel": rng.randn(T, n_mel).astype(np.float32),
                     "pitch": rng.randn(L).astype(np.float32), "energy": rng.randn(L).astype(np.float32), "duration": dur})
    return data
