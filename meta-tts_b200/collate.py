"""Batch producer for the meta-training step — drop-in for `lightning/collate.py` (reference file:line cited per function).

Same call signatures and the same 12-tuple wire format as the reference (`reprocess`, `get_single_collate`,
`SpeakerTaskCollate.get_meta_collate`), built B200-first:
  * every padded field is written ONCE, straight into a single pinned staging buffer laid out exactly like the static
    device batch of `systems._StaticBatch` (one `cudaMemcpyAsync` per task, no per-field copies in the training loop:
    the 12-tuple's tensors are views of that buffer), instead of np.pad per utterance + np.stack + torch.from_numpy;
  * alternatively the utterances can stay RAGGED on the host (`ragged=True`): fields are concatenated back to back,
    copied in one H2D of sum(len) rows, and padded on the device by `mtts_pack_rows` (`pack_on_device`).
Results are bit-identical to the reference's collate (tests/test_collate_*.py against goldens made from the real module).
"""
from __future__ import annotations

from functools import partial
from typing import Dict, List, Sequence

import numpy as np
import torch
import torch.utils.data

N_MEL_DEFAULT = 80


def pad_1D(inputs: Sequence[np.ndarray], PAD=0) -> np.ndarray:
    """utils/tools.py:270-281: right-pad every 1-D array to the longest, stack."""
    max_len = max(len(x) for x in inputs)
    out = np.full((len(inputs), max_len), PAD, dtype=np.result_type(*[x.dtype for x in inputs]))
    for i, x in enumerate(inputs):
        out[i, :len(x)] = x
    return out


def pad_2D(inputs: Sequence[np.ndarray], maxlen=None) -> np.ndarray:
    """utils/tools.py:284-301: right-pad the first axis of every [T_i, C] array to max T (or maxlen), stack."""
    max_len = maxlen if maxlen else max(np.shape(x)[0] for x in inputs)
    for x in inputs:
        if np.shape(x)[0] > max_len:
            raise ValueError("not max_len")                    # tools.py:287-288
    c = np.shape(inputs[0])[1]
    out = np.zeros((len(inputs), max_len, c), dtype=np.result_type(*[x.dtype for x in inputs]))
    for i, x in enumerate(inputs):
        out[i, :np.shape(x)[0]] = x
    return out


class Batch12(tuple):
    """The reference's 12-tuple (collate.py:47-60) plus, as attributes, the pinned staging buffer its tensors live in
    (`staged`, in `_StaticBatch` layout) or the ragged buffers (`ragged`)."""
    staged = None
    ragged = None


def _field_layout(n: int, L: int, T: int, n_spk: int, n_mel: int):
    """Byte layout of systems._StaticBatch (kept in sync by tests/test_collate_cpu.py)."""
    from .systems import _StaticBatch
    shapes = {"spk_ids": (n_spk,), "texts": (n, L), "src_lens": (n,), "mels": (n, T, n_mel), "mel_lens": (n,),
              "pitches": (n, L), "energies": (n, L), "durations": (n, L), "salt": (1,)}
    offs, off = {}, 0
    for f, dt in _StaticBatch.FIELDS:
        nbytes = torch.tensor([], dtype=dt).element_size() * int(np.prod(shapes[f]))
        offs[f] = (off, nbytes, dt, shapes[f])
        off = (off + nbytes + 63) // 64 * 64
    return offs, off


def _default_pin() -> bool:
    """Pin the staging buffer only in the trainer process: inside a (forked) DataLoader worker `pin_memory()` would
    initialise CUDA in the child ('Cannot re-initialize CUDA in forked subprocess'); there the DataLoader's own
    `pin_memory=True` thread (or the consumer) pins instead."""
    if torch.utils.data.get_worker_info() is not None:
        return False
    return torch.cuda.is_available()


def reprocess(data, idxs, pin: bool = None):
    """collate.py:9-60.  Same inputs (list of dataset dicts, index array) and the same 12-tuple:
    (ids, raw_texts, speaker_args i64[B], texts i64[B,L], text_lens i64[B], max_text_len, mels f32[B,T,80],
     mel_lens i64[B], max_mel_len, pitches f32[B,L], energies [B,L] (dtype as stored), durations i64[B,L])."""
    idxs = list(idxs)
    items = [data[i] for i in idxs]
    ids = [d["id"] for d in items]
    raw_texts = [d["raw_text"] for d in items]
    assert "spk_ref_mel_slices" not in data[0], "reference-encoder speaker args are not on the hot path (table embedding only)"
    texts = [np.asarray(d["text"]) for d in items]
    mels = [np.asarray(d["mel"]) for d in items]
    text_lens = np.array([t.shape[0] for t in texts])
    mel_lens = np.array([m.shape[0] for m in mels])
    n, L, T, n_mel = len(items), int(text_lens.max()), int(mel_lens.max()), int(mels[0].shape[1])
    e_dtype = np.result_type(*[np.asarray(d["energy"]).dtype for d in items])
    fast = e_dtype == np.float32                  # the staging layout stores energies as f32 (the dtype the dataset writes)
    pin = _default_pin() if pin is None else pin
    offs, nbytes = _field_layout(n, L, T, n, n_mel)
    buf = torch.zeros(nbytes, dtype=torch.uint8)
    if pin:
        buf = buf.pin_memory()
    view = {f: buf[o:o + nb].view(dt).view(shape) for f, (o, nb, dt, shape) in offs.items()}
    view["spk_ids"].copy_(torch.from_numpy(np.array([d["speaker"] for d in items]).astype(np.int64)))
    view["src_lens"].copy_(torch.from_numpy(text_lens.astype(np.int64)))
    view["mel_lens"].copy_(torch.from_numpy(mel_lens.astype(np.int64)))
    tn = {f: view[f].numpy() for f in ("texts", "mels", "pitches", "energies", "durations")}
    for i, d in enumerate(items):                 # one pass: each value is written once, into its final (padded) place
        tn["texts"][i, :text_lens[i]] = texts[i]
        tn["mels"][i, :mel_lens[i]] = mels[i]
        p, e, du = np.asarray(d["pitch"]), np.asarray(d["energy"]), np.asarray(d["duration"])
        tn["pitches"][i, :p.shape[0]] = p
        if fast:
            tn["energies"][i, :e.shape[0]] = e
        tn["durations"][i, :du.shape[0]] = du
    energies = view["energies"] if fast else torch.from_numpy(pad_1D([np.asarray(d["energy"]) for d in items]))
    out = Batch12((ids, raw_texts, view["spk_ids"], view["texts"], torch.from_numpy(text_lens), text_lens.max(), view["mels"],
                   torch.from_numpy(mel_lens), mel_lens.max(), view["pitches"], energies, view["durations"]))
    out.staged = buf if fast else None
    return out


def reprocess_ragged(data, idxs, pin: bool = None) -> Dict[str, torch.Tensor]:
    """The same utterances, NOT padded: fields concatenated back to back + int64 row offsets (for `pack_on_device`)."""
    items = [data[i] for i in list(idxs)]
    pin = _default_pin() if pin is None else pin
    cat = lambda k, dt: torch.from_numpy(np.concatenate([np.asarray(d[k]) for d in items]).astype(dt))  # noqa: E731
    out = {"texts": cat("text", np.int64), "durations": cat("duration", np.int64), "pitches": cat("pitch", np.float32),
           "energies": cat("energy", np.float32), "mels": cat("mel", np.float32),
           "spk_ids": torch.from_numpy(np.array([d["speaker"] for d in items]).astype(np.int64))}
    tl = np.array([np.asarray(d["text"]).shape[0] for d in items], dtype=np.int64)
    ml = np.array([np.asarray(d["mel"]).shape[0] for d in items], dtype=np.int64)
    out["src_lens"], out["mel_lens"] = torch.from_numpy(tl), torch.from_numpy(ml)
    out["off_l"] = torch.from_numpy(np.concatenate([[0], np.cumsum(tl)]).astype(np.int64))
    out["off_t"] = torch.from_numpy(np.concatenate([[0], np.cumsum(ml)]).astype(np.int64))
    if pin:
        out = {k: v.pin_memory() for k, v in out.items()}
    return out


def pack_on_device(be, ragged: Dict[str, torch.Tensor], device, L: int = None, T: int = None) -> Dict[str, torch.Tensor]:
    """Ragged host fields -> padded device tensors: one H2D per field of sum(len) rows, padding by `mtts_pack_rows`."""
    B = ragged["src_lens"].numel()
    L = int(ragged["src_lens"].max()) if L is None else L
    T = int(ragged["mel_lens"].max()) if T is None else T
    dev = {k: v.to(device, non_blocking=True) for k, v in ragged.items()}
    out = {"spk_ids": dev["spk_ids"], "src_lens": dev["src_lens"], "mel_lens": dev["mel_lens"]}
    for k, off, Lm in (("texts", "off_l", L), ("durations", "off_l", L), ("pitches", "off_l", L), ("energies", "off_l", L),
                       ("mels", "off_t", T)):
        src = dev[k]
        row = src[0].numel() * src.element_size() if src.dim() > 1 else src.element_size()
        dst = torch.empty((B, Lm) + tuple(src.shape[1:]), dtype=src.dtype, device=device)
        be.pack_rows(src, dev[off], B, Lm, row, dst)
        out[k] = dst
    return out


def get_single_collate(sort=True):
    """collate.py:128-143"""
    def collate_fn(data):
        if sort:
            idx_arr = np.argsort(-np.array([d["text"].shape[0] for d in data]))
        else:
            idx_arr = np.arange(len(data))
        return reprocess(data, idx_arr)
    return collate_fn


class SpeakerTaskCollate:
    """collate.py:146-196: 1 way (speaker), K shots, Q queries -> ([sup 12-tuple], [qry 12-tuple])."""

    def get_meta_collate(self, shots, queries, sort=False, split=True):
        return partial(self.meta_collate_fn, shots=shots, queries=queries, sort=sort, split=split)

    def meta_collate_fn(self, data, shots, queries, sort=False, split=True):
        batch_size = shots + queries
        assert len(data) == batch_size, "n_batch=1 for speaker adaptation"
        if sort:
            idx_arr = np.argsort(-np.array([d["text"].shape[0] for d in data]))
        else:
            idx_arr = np.arange(len(data))
        idx_arr = idx_arr.reshape((-1, batch_size))
        if split:
            sup_idx = np.zeros(batch_size, dtype=bool)
            sup_idx[np.arange(shots)] = True
            qry_idx = ~sup_idx
            return ([reprocess(data, idx) for idx in idx_arr[:, sup_idx]], [reprocess(data, idx) for idx in idx_arr[:, qry_idx]])
        return [reprocess(data, idx) for idx in idx_arr]


def split_reprocess(batch, idxs):
    """lightning/collate.py:63-125 (table speaker ids): the rows `idxs` of a collated 12-tuple, re-trimmed to their own
    maximum lengths (used by the 1-shot test protocol, base_adaptor.py:144-151)."""
    (ids, raw_texts, speaker_args, texts, text_lens, max_text_lens, mels, mel_lens, max_mel_lens, pitches, energies,
     durations) = batch
    idxs = np.asarray(idxs)
    sub_text_lens, sub_mel_lens = text_lens[idxs], mel_lens[idxs]
    Ls, Ts = sub_text_lens.max(), sub_mel_lens.max()
    cut = lambda t: t[idxs][:, :Ls] if t.shape[1] == max_text_lens else t[idxs][:, :Ts]  # noqa: E731
    return ([ids[i] for i in idxs], [raw_texts[i] for i in idxs], speaker_args[idxs], texts[idxs][:, :Ls], sub_text_lens, Ls,
            mels[idxs][:, :Ts], sub_mel_lens, Ts, cut(pitches), cut(energies), durations[idxs][:, :Ls])
