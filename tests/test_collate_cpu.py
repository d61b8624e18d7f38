"""SURVEY.md §8 row f2 (batch producer): the oracle restatement and the B200 collate against goldens made from the REAL
`lightning/collate.py` / `utils/tools.py` (oracle/make_golden_collate.py) — bit-exact, dtypes included — plus the
staging-layout contract with systems._StaticBatch and the ragged -> padded pack semantics (CPU restatement of the kernel)."""
import os

import numpy as np
import torch

from meta_tts_b200 import collate as B
from meta_tts_b200.systems import _StaticBatch
from oracle import collate_oracle as C
from oracle.ops_reference import RefOps

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "collate_golden.npz"), allow_pickle=False)
FIELDS = {2: "speakers", 3: "texts", 4: "text_lens", 6: "mels", 7: "mel_lens", 9: "pitches", 10: "energies", 11: "durations"}
DATA = C.synth_dataset(n=9, seed=0)


def _check(prefix, t12):
    assert len(t12) == 12
    for i, name in FIELDS.items():
        assert str(t12[i].dtype) == str(G[f"{prefix}_{name}_dtype"]), (prefix, name)
        assert np.array_equal(t12[i].numpy(), G[f"{prefix}_{name}"]), (prefix, name)
    assert int(t12[5]) == int(G[f"{prefix}_max_text_len"]) and int(t12[8]) == int(G[f"{prefix}_max_mel_len"])
    assert list(t12[0]) == list(G[f"{prefix}_ids"])


def test_oracle_restatement_matches_reference_goldens():
    _check("plain", C.reprocess(DATA, np.arange(9)))
    assert np.array_equal(C.pad_1D([d["pitch"] for d in DATA]), G["pad1d"])
    assert np.array_equal(C.pad_2D([d["mel"] for d in DATA]), G["pad2d"])
    assert np.array_equal(C.pad_2D([d["mel"] for d in DATA[:3]], maxlen=80), G["pad2d_maxlen"])


def test_b200_collate_matches_reference_goldens():
    _check("plain", B.reprocess(DATA, np.arange(9), pin=False))
    _check("sorted", B.get_single_collate(sort=True)(DATA))
    sup, qry = B.SpeakerTaskCollate().get_meta_collate(shots=5, queries=4)(DATA)
    _check("sup", sup[0])
    _check("qry", qry[0])
    assert np.array_equal(B.pad_1D([d["pitch"] for d in DATA]), G["pad1d"])
    assert np.array_equal(B.pad_2D([d["mel"] for d in DATA]), G["pad2d"])
    assert np.array_equal(B.pad_2D([d["mel"] for d in DATA[:3]], maxlen=80), G["pad2d_maxlen"])
    try:
        B.pad_2D([d["mel"] for d in DATA], maxlen=2)
        assert False, "pad_2D must reject sequences longer than maxlen (tools.py:287-288)"
    except ValueError:
        pass


def test_staged_buffer_is_the_static_batch_layout():
    """reprocess() writes the batch in the byte layout of systems._StaticBatch: upload() is then ONE copy of that buffer."""
    t12 = B.reprocess(DATA, np.arange(5), pin=False)
    n, L, T = 5, int(t12[5]), int(t12[8])
    sb = _StaticBatch(torch.device("cpu"), n, L, T, n, False)
    assert t12.staged is not None and t12.staged.numel() == sb.nbytes
    sb.upload(t12, salt=0x89ABCDEF)                                 # fast path (staged buffer)
    ref = _StaticBatch(torch.device("cpu"), n, L, T, n, False)
    ref.upload(tuple(t12), salt=0x89ABCDEF)                         # generic path (per-field copies)
    assert torch.equal(sb.dev_buf, ref.dev_buf)
    assert torch.equal(sb.dev.texts, t12[3]) and torch.equal(sb.dev.mels, t12[6]) and torch.equal(sb.dev.durations, t12[11])
    assert int(sb.salt[0]) & 0xFFFFFFFF == 0x89ABCDEF


def test_ragged_pack_semantics():
    """pack_on_device (ragged H2D + mtts_pack_rows) == the reference's padded tensors; here through the CPU restatement."""
    idx = np.arange(9)
    rag = B.reprocess_ragged(DATA, idx, pin=False)
    assert int(rag["off_t"][-1]) == sum(d["mel"].shape[0] for d in DATA)
    out = B.pack_on_device(RefOps(), rag, "cpu")
    ref = C.reprocess(DATA, idx)
    for k, i in (("texts", 3), ("mels", 6), ("pitches", 9), ("energies", 10), ("durations", 11), ("src_lens", 4), ("mel_lens", 7)):
        assert torch.equal(out[k], ref[i]), k
