"""Generate tests/golden/dataset_golden.npz with the REAL reference `dataset.py: TTSDataset` and `text.text_to_sequence`
(container only): a tiny synthetic preprocessed directory (written by `tests/tests_helpers_dataset.make_corpus`) is read by the
reference class; its samples (phoneme ids from the reference's own symbol table, arrays, speaker ids) are stored so that the
drop-in reader can be checked against them without the reference.

    python -m oracle.make_golden_dataset
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import refstub  # noqa: E402
from tests_helpers_dataset import PHONES, make_corpus  # noqa: E402


def main():
    refstub.load_reference()
    import types
    ra = types.ModuleType("resemblyzer.audio")            # dataset.py:7-8 imports two wav helpers it only uses for d-vector references
    ra.preprocess_wav = ra.wav_to_mel_spectrogram = lambda *a, **k: None
    sys.modules["resemblyzer"].__path__ = []
    sys.modules["resemblyzer"].audio = ra
    sys.modules["resemblyzer.audio"] = ra
    import dataset as RD                      # the real module
    from text import text_to_sequence         # the real front-end
    from text.symbols import symbols
    d = tempfile.mkdtemp(prefix="mtts_ds_")
    pre, train = make_corpus(d)
    ds = RD.TTSDataset("train.txt", pre, train)
    out = {"n": np.array(len(ds))}
    for i in range(len(ds)):
        s = ds[i]
        out[f"{i}_id"] = np.array(s["id"])
        out[f"{i}_speaker"] = np.array(s["speaker"])
        out[f"{i}_raw_text"] = np.array(s["raw_text"])
        for k in ("text", "mel", "pitch", "energy", "duration"):
            out[f"{i}_{k}"] = np.asarray(s[k])
    # the slice of the reference symbol table the synthetic corpus uses (for the table-lookup front-end of the drop-in)
    ids = {p: text_to_sequence("{" + p + "}", ["english_cleaners"])[0] for p in PHONES}
    out["phones"] = np.array(PHONES)
    out["phone_ids"] = np.array([ids[p] for p in PHONES])
    out["n_symbols"] = np.array(len(symbols))
    path = os.path.join(ROOT, "tests", "golden", "dataset_golden.npz")
    np.savez_compressed(path, **out)
    print("[golden] wrote", path, os.path.getsize(path) // 1024, "KiB;", len(ds), "samples,", len(symbols), "symbols in the reference table")


if __name__ == "__main__":
    main()
