"""A tiny synthetic preprocessed corpus in the reference's on-disk layout (shared by oracle/make_golden_dataset.py and the tests)."""
import json
import os

import numpy as np

PHONES = ["HH", "AH0", "L", "OW1", "W", "ER1", "D", "sp", "T", "EH1", "S"]
UTTS = [("utt_a", "spk1", "HH AH0 L OW1 sp W ER1 L D", "hello world"), ("utt_b", "spk2", "T EH1 S T", "test"),
        ("utt_c", "spk1", "W ER1 D sp T EH1 S T S", "word tests")]


def make_corpus(root: str, seed: int = 0):
    rng = np.random.RandomState(seed)
    for k in ("mel", "pitch", "energy", "duration"):
        os.makedirs(os.path.join(root, k), exist_ok=True)
    lines = []
    for base, spk, phones, raw in UTTS:
        L = len(phones.split())
        dur = rng.randint(1, 6, size=L).astype(np.int64)
        T = int(dur.sum())
        np.save(os.path.join(root, "mel", f"{spk}-mel-{base}.npy"), rng.randn(T, 80).astype(np.float32))
        np.save(os.path.join(root, "pitch", f"{spk}-pitch-{base}.npy"), rng.randn(L).astype(np.float64))
        np.save(os.path.join(root, "energy", f"{spk}-energy-{base}.npy"), rng.randn(L).astype(np.float32))
        np.save(os.path.join(root, "duration", f"{spk}-duration-{base}.npy"), dur)
        lines.append(f"{base}|{spk}|{{{phones}}}|{raw}")
    with open(os.path.join(root, "train.txt"), "w", encoding="utf-8") as f:
        f.write("\n".join(lines) + "\n")
    with open(os.path.join(root, "speakers.json"), "w") as f:
        json.dump({"spk1": 0, "spk2": 1}, f)
    pre = {"dataset": "LibriTTS", "path": {"preprocessed_path": root}, "preprocessing": {"text": {"text_cleaners": ["english_cleaners"]}}}
    train = {"optimizer": {"batch_size": 2}}
    return pre, train
