"""SURVEY §8 row f2, on-disk side: the drop-in `TTSDataset` reads the reference's preprocessed layout
(`{kind}/{speaker}-{kind}-{basename}.npy`, `train.txt`, `speakers.json`, dataset.py:14-110) and yields the same samples as the REAL
reference class did on the same synthetic corpus (tests/golden/dataset_golden.npz, oracle/make_golden_dataset.py); the samples
flow through the drop-in collate into the 12-tuple wire format."""
import os
import tempfile

import numpy as np
import torch

from meta_tts_b200 import collate as B
from meta_tts_b200.dataset import TTSDataset, phoneme_table_lookup
from tests_helpers_dataset import make_corpus

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "dataset_golden.npz"), allow_pickle=False)


def _dataset():
    d = tempfile.mkdtemp(prefix="mtts_ds_")
    pre, train = make_corpus(d)
    table = {str(p): int(i) for p, i in zip(G["phones"], G["phone_ids"])}      # slice of the reference's symbol table
    return TTSDataset("train.txt", pre, train, text_to_sequence=phoneme_table_lookup(table))


def test_samples_match_the_real_reference_dataset():
    ds = _dataset()
    assert len(ds) == int(G["n"]) == 3
    for i in range(len(ds)):
        s = ds[i]
        assert s["id"] == str(G[f"{i}_id"]) and s["speaker"] == int(G[f"{i}_speaker"]) and s["raw_text"] == str(G[f"{i}_raw_text"])
        for k in ("text", "mel", "pitch", "energy", "duration"):
            assert s[k].dtype == G[f"{i}_{k}"].dtype and np.array_equal(s[k], G[f"{i}_{k}"]), (i, k)
    assert int(G["n_symbols"]) == 360          # = N_SYMBOLS of the phoneme embedding table (encoder.src_word_emb has 361 rows)


def test_samples_feed_the_collate_wire_format():
    ds = _dataset()
    data = [ds[i] for i in range(len(ds))]
    t12 = B.reprocess(data, np.arange(len(data)))
    assert len(t12) == 12 and t12[3].dtype == torch.int64 and t12[6].shape[2] == 80
    assert t12[4].tolist() == [len(d["text"]) for d in data] and t12[7].tolist() == [d["mel"].shape[0] for d in data]
    assert all(int(d["duration"].sum()) == d["mel"].shape[0] for d in data)
    for d in data:
        d["speaker"] = 0
    sup, qry = B.SpeakerTaskCollate().get_meta_collate(shots=2, queries=1)(data)
    assert sup[0][3].shape[0] == 2 and qry[0][3].shape[0] == 1
