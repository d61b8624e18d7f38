"""Generate tests/golden/collate_golden.npz from the REAL reference batch producer (container only; /root/reference).

    python -m oracle.make_golden_collate

Imports the unmodified `lightning/collate.py` (reprocess, SpeakerTaskCollate) and `utils/tools.py` (pad_1D, pad_2D) through
the stub loader, runs them on the seeded synthetic dataset of oracle/collate_oracle.synth_dataset, and stores every tensor
field of the resulting 12-tuples (+ dtypes), so tests can pin the oracle and the B200 collate without the reference.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import collate_oracle as C  # noqa: E402
from oracle import refstub  # noqa: E402

FIELDS = {2: "speakers", 3: "texts", 4: "text_lens", 6: "mels", 7: "mel_lens", 9: "pitches", 10: "energies", 11: "durations"}


def pack(prefix, t12, out):
    for i, name in FIELDS.items():
        v = t12[i]
        out[f"{prefix}_{name}"] = v.numpy()
        out[f"{prefix}_{name}_dtype"] = np.array(str(v.dtype))
    out[f"{prefix}_max_text_len"] = np.array(int(t12[5]))
    out[f"{prefix}_max_mel_len"] = np.array(int(t12[8]))
    out[f"{prefix}_ids"] = np.array(list(t12[0]))


def main():
    refstub.install_stubs()
    sys.path.insert(0, refstub.REF)
    from lightning import collate as RC            # the real module
    from utils import tools as RT
    data = C.synth_dataset(n=9, seed=0)
    out = {}
    pack("plain", RC.reprocess(data, np.arange(9)), out)
    pack("sorted", RC.get_single_collate(sort=True)(data), out)
    sup, qry = RC.SpeakerTaskCollate().get_meta_collate(shots=5, queries=4)(data)
    pack("sup", sup[0], out)
    pack("qry", qry[0], out)
    out["pad1d"] = RT.pad_1D([d["pitch"] for d in data])
    out["pad2d"] = RT.pad_2D([d["mel"] for d in data])
    out["pad2d_maxlen"] = RT.pad_2D([d["mel"] for d in data[:3]], maxlen=80)
    # the restatement reproduces the real module bit for bit
    for name, idx in (("plain", np.arange(9)),):
        mine = C.reprocess(data, idx)
        for i, f in FIELDS.items():
            assert mine[i].dtype == getattr(torch, str(out[f"{name}_{f}_dtype"]).replace("torch.", "")), f
            assert np.array_equal(mine[i].numpy(), out[f"{name}_{f}"]), f
    path = os.path.join(ROOT, "tests", "golden", "collate_golden.npz")
    np.savez_compressed(path, **out)
    print("[golden] wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
