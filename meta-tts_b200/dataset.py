"""On-disk sample format of the batch producer — drop-in for the reference's `dataset.py: TTSDataset`
(dataset.py:14-110): `{preprocessed_path}/{mel,pitch,energy,duration}/{speaker}-{kind}-{basename}.npy`, a metadata file with
`basename|speaker|{phoneme string}|raw text` lines and `speakers.json`.  `__getitem__` yields the dictionary
(`id, speaker, text, raw_text, mel, pitch, energy, duration`) that `collate.reprocess` / `SpeakerTaskCollate` consume.

The text front-end (`text.text_to_sequence`, cleaners, symbol table) is outside the hot path: it is injected as a callable —
pass the reference's own `text.text_to_sequence` when its tree is importable; the default splits a `{P1 P2 ...}` phoneme string
and looks the symbols up in a caller-supplied table.
"""
from __future__ import annotations

import json
import os
from typing import Callable, Dict, Optional, Sequence

import numpy as np
from torch.utils.data import Dataset

KINDS = ("mel", "pitch", "energy", "duration")


def phoneme_table_lookup(symbol_to_id: Dict[str, int]) -> Callable[[str, Sequence[str]], list]:
    """A minimal `text_to_sequence(text, cleaners)` for already-phonemised `{AA1 B ...}` strings (what the LibriTTS / MFA
    preprocessing of the reference writes into train.txt): curly-brace content is split on whitespace and mapped through
    `symbol_to_id` with the reference's `@` prefix for ARPAbet symbols (text/__init__.py) when the bare symbol is absent."""
    def text_to_sequence(text: str, cleaners=None):
        body = text.strip()
        if body.startswith("{") and body.endswith("}"):
            body = body[1:-1]
        out = []
        for sym in body.split():
            if sym in symbol_to_id:
                out.append(symbol_to_id[sym])
            elif "@" + sym in symbol_to_id:
                out.append(symbol_to_id["@" + sym])
            else:
                raise KeyError(f"phoneme symbol {sym!r} is not in the table")
        return out
    return text_to_sequence


class TTSDataset(Dataset):
    """dataset.py:14-110 (the `spk_refer_wav` d-vector side channel is outside the table-speaker hot path)."""

    def __init__(self, filename, preprocess_config, train_config, sort=False, drop_last=False, spk_refer_wav=False,
                 text_to_sequence: Optional[Callable] = None):
        assert not spk_refer_wav, "speaker reference mels (d-vector / GE2E encoders) are out of scope: table speaker ids only"
        self.dataset_name = preprocess_config["dataset"]
        self.preprocessed_path = preprocess_config["path"]["preprocessed_path"]
        self.cleaners = preprocess_config["preprocessing"]["text"]["text_cleaners"]
        self.batch_size = train_config["optimizer"]["batch_size"]
        self.spk_refer_wav = spk_refer_wav
        if text_to_sequence is None:
            from text import text_to_sequence                 # the reference's front-end, when its tree is on sys.path
        self.text_to_sequence = text_to_sequence
        self.basename, self.speaker, self.text, self.raw_text = self.process_meta(filename)
        with open(os.path.join(self.preprocessed_path, "speakers.json")) as f:
            self.speaker_map = json.load(f)
        self.sort, self.drop_last = sort, drop_last

    def __len__(self):
        return len(self.text)

    def path(self, kind: str, speaker: str, basename: str) -> str:
        return os.path.join(self.preprocessed_path, kind, f"{speaker}-{kind}-{basename}.npy")

    def __getitem__(self, idx):
        basename, speaker = self.basename[idx], self.speaker[idx]
        sample = {"id": basename, "speaker": self.speaker_map[speaker],
                  "text": np.array(self.text_to_sequence(self.text[idx], self.cleaners)), "raw_text": self.raw_text[idx]}
        for kind in KINDS:
            sample[kind] = np.load(self.path(kind, speaker, basename))
        return sample

    def process_meta(self, filename):
        name, speaker, text, raw_text = [], [], [], []
        with open(os.path.join(self.preprocessed_path, filename), "r", encoding="utf-8") as f:
            for line in f.readlines():
                n, s, t, r = line.strip("\n").split("|")
                name.append(n)
                speaker.append(s)
                text.append(t)
                raw_text.append(r)
        return name, speaker, text, raw_text
