"""Drop-in host API for the MAML step: same names / arguments / return values as the reference's
`lightning.systems` classes for the hot path, minus Lightning (which only orchestrates).

    reference                                         here
    ------------------------------------------------  -------------------------------------------
    MAML(module, lr).clone()/adapt_()                 MAML (compatibility shim over MamlEngine)
      lightning/systems/utils.py:17-77
    BaseAdaptorSystem.adapt(batch, steps, ...)        MetaSystem.adapt        (base_adaptor.py:98-112)
    BaseAdaptorSystem.meta_learn(batch, idx, train)   MetaSystem.meta_learn   (base_adaptor.py:114-124)
    MetaSystem.training_step(batch, idx)              MetaSystem.training_step (meta.py:68-80)
    Lightning: backward, DDP allreduce, clip, Adam    MetaSystem.optimizer_step (main.py:57-64,
      + LambdaLR                                        optimizer.py:6-16, scheduler.py:6-29)

`batch` keeps the reference layout `[ ( [sup 12-tuple], [qry 12-tuple] ) ]` (base_adaptor.py:126-131).
The autograd-free engine cannot accept a bare loss tensor in `MAML.adapt_(loss)`; the fused path is
exposed at adapt()/meta_learn() level with the reference signatures.

The whole task step (K inner steps + query + outer backward [+ HVP recursion]) is captured once per
input shape in a CUDA graph and replayed; per step the host copies the batch into static device
buffers (pinned H2D) and reads back the 6 query losses.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Optional, Tuple

import os
import time

import torch

from . import ops as _ops
from .engine import Batch, N_MEL
from .maml import MamlEngine, batch_from_tuple
from .ops import CudaOps


def _assert_meta_batch(batch) -> None:
    """base_adaptor.py:126-131"""
    assert len(batch) == 1, "meta_batch_per_gpu"
    assert len(batch[0]) == 2, "sup + qry"
    assert len(batch[0][0]) == 1, "n_batch == 1"
    assert len(batch[0][0][0]) == 12, "data with 12 elements"


class _StaticBatch:
    """Static device buffers for one 12-tuple shape + a ring of pinned host staging buffers.
    All tensor fields live in ONE contiguous byte buffer: one cudaMemcpyAsync per batch; the host may run
    one step ahead of the device (ring of 2, guarded by events)."""
    FIELDS = (("spk_ids", torch.int64), ("texts", torch.int64), ("src_lens", torch.int64), ("mels", torch.float32),
              ("mel_lens", torch.int64), ("pitches", torch.float32), ("energies", torch.float32), ("durations", torch.int64),
              ("salt", torch.int32))        # per-step dropout salt (uint32 bits): rides in the same H2D copy
    RING = 4        # the host may enqueue up to RING - 1 steps ahead of the device

    def __init__(self, device, n: int, L: int, T: int, n_spk_ids: int, average_spk: bool, max_T: Optional[int] = None):
        # train-mode truncation (Models.py:161-166, loss.py:42-43): keep the first max_seq_len frames of the mel targets
        if max_T is not None and T > max_T:
            T = max_T
        self.T = T
        shapes = {"spk_ids": (n_spk_ids,), "texts": (n, L), "src_lens": (n,), "mels": (n, T, N_MEL), "mel_lens": (n,),
                  "pitches": (n, L), "energies": (n, L), "durations": (n, L), "salt": (1,)}
        offs, off = {}, 0
        for f, dt in self.FIELDS:
            nbytes = int(torch.tensor([], dtype=dt).element_size()) * int(torch.Size(shapes[f]).numel())
            offs[f] = (off, nbytes)
            off = (off + nbytes + 63) // 64 * 64
        self.nbytes = off
        self._salt_off = offs["salt"][0]
        self._keep = []
        self.dev_buf = torch.zeros(off, dtype=torch.uint8, device=device)
        view = lambda buf, f, dt: buf[offs[f][0]:offs[f][0] + offs[f][1]].view(dt).view(shapes[f])  # noqa: E731
        d = {f: view(self.dev_buf, f, dt) for f, dt in self.FIELDS}
        self.salt = d.pop("salt")
        self.dev = Batch(average_spk=average_spk, B=n, L=L, T=T, **d)
        pin = torch.cuda.is_available() and self.dev_buf.is_cuda
        self.host_bufs = [torch.zeros(off, dtype=torch.uint8).pin_memory() if pin else torch.zeros(off, dtype=torch.uint8)
                          for _ in range(self.RING)]
        self.host = [{f: view(hb, f, dt) for f, dt in self.FIELDS} for hb in self.host_bufs]
        self.events = [None] * self.RING
        self.slot = 0
        self.h2d_bytes = off

    def upload(self, b12, spk_ids=None, salt: int = 0) -> None:
        k = self.slot
        self.slot = (k + 1) % self.RING
        if self.events[k] is not None:
            self.events[k].synchronize()                 # the H2D that last used this staging slot has completed
        staged = getattr(b12, "staged", None)
        if staged is not None and spk_ids is None and staged.numel() == self.nbytes and int(b12[8]) == self.T:
            # the batch producer (meta_tts_b200.collate.reprocess) already wrote this batch in the static layout into
            # pinned memory: no per-field host copies, one H2D straight from the producer's buffer
            salt &= 0xFFFFFFFF
            staged[self._salt_off:self._salt_off + 4].view(torch.int32)[0] = salt - (1 << 32) if salt >= (1 << 31) else salt
            self.dev_buf.copy_(staged, non_blocking=True)
            if self.dev_buf.is_cuda:
                ev = self.events[k] or torch.cuda.Event()
                ev.record()
                self.events[k] = ev
            self._keep = (self._keep + [staged])[-self.RING:]        # keep the producer's buffers alive until their copies ran
            return
        h = self.host[k]
        # 12-tuple positions (lightning/collate.py:47-60): 2 speaker, 3 texts, 4 text_lens, 6 mels, 7 mel_lens, 9 pitch,
        # 10 energy, 11 durations
        for f, i in (("texts", 3), ("src_lens", 4), ("mels", 6), ("mel_lens", 7), ("pitches", 9), ("energies", 10), ("durations", 11)):
            src = b12[i]
            dst = h[f]
            src = src if torch.is_tensor(src) else torch.as_tensor(src)
            if f == "mels" and src.shape[1] > self.T:
                src = src[:, :self.T]                    # train-mode truncation to max_seq_len frames
            dst.copy_(src)                               # dtype conversion + gather into pinned staging
        spk = b12[2] if spk_ids is None else spk_ids
        h["spk_ids"].copy_(spk if torch.is_tensor(spk) else torch.as_tensor(spk))
        salt &= 0xFFFFFFFF
        h["salt"][0] = salt - (1 << 32) if salt >= (1 << 31) else salt
        self.dev_buf.copy_(self.host_bufs[k], non_blocking=True)       # ONE H2D per batch
        if self.dev_buf.is_cuda:
            ev = self.events[k] or torch.cuda.Event()
            ev.record()
            self.events[k] = ev


class _Predictions(tuple):
    """The reference's 10-tuple of predictions (base_adaptor.py:91-95).  Device tensors; the two boolean masks
    (entries 6, 7; True = padding) are built on first access so that a training loop that never reads them
    launches nothing for them."""

    def __new__(cls, out, dev, Lq, T):
        self = super().__new__(cls, (out["mel"], out["postnet"], out["pitch"], out["energy"], out["logd"], dev.durations,
                                     None, None, dev.src_lens, out["mel_len"]))
        self._dev, self._Lq, self._T, self._masks = dev, Lq, T, None
        return self

    def _get_masks(self):
        if self._masks is None:
            d = self._dev
            ar_l = torch.arange(self._Lq, device=d.src_lens.device)[None, :]
            ar_t = torch.arange(self._T, device=d.src_lens.device)[None, :]
            self._masks = (ar_l >= d.src_lens[:, None], ar_t >= d.mel_lens[:, None])
        return self._masks

    def __getitem__(self, i):
        if isinstance(i, int) and i in (6, 7, -4, -3):
            return self._get_masks()[0 if i in (6, -4) else 1]
        if isinstance(i, slice):
            return tuple(self[j] for j in range(*i.indices(10)))
        return super().__getitem__(i)

    def __iter__(self):
        return iter([self[j] for j in range(10)])


class LazyLosses:
    """The step's 6 query losses, copied device -> pinned host asynchronously right after the step was
    enqueued (the reference returns CUDA tensors from training_step and only syncs when it logs).  Reading
    a value waits for that copy.  Behaves like the reference's 6-tuple."""

    def __init__(self, host: torch.Tensor, event):
        self._host, self._event = host, event

    def wait(self):
        if self._event is not None:
            self._event.synchronize()
            self._event = None
        return self._host

    def __len__(self):
        return 6

    def __getitem__(self, i):
        return self.wait()[i]

    def __iter__(self):
        h = self.wait()
        return iter([h[i] for i in range(6)])


class LazyScalar:
    """One entry of a LazyLosses 6-tuple that has not been waited for yet: `float(x)` / `x.item()` / formatting /
    arithmetic wait for the step's asynchronous device-to-host copy.  (The reference returns a CUDA tensor from
    training_step and only synchronises when it logs — meta.py:76-80; returning a Python float here would force a
    full device sync inside every training_step and forbid the host from running ahead.)"""

    def __init__(self, losses: "LazyLosses", i: int):
        self._l, self._i = losses, i

    def item(self) -> float:
        return float(self._l.wait()[self._i])

    __float__ = item

    def __repr__(self):
        return f"{self.item():.6f}"

    def __format__(self, spec):
        return format(self.item(), spec)

    def __add__(self, o):
        return self.item() + float(o)

    __radd__ = __add__

    def __mul__(self, o):
        return self.item() * float(o)

    __rmul__ = __mul__

    def __truediv__(self, o):
        return self.item() / float(o)

    def __lt__(self, o):
        return self.item() < float(o)

    def __gt__(self, o):
        return self.item() > float(o)


DEFAULT_MODEL_CONFIG = {
    "transformer": {"encoder_layer": 4, "encoder_head": 2, "encoder_hidden": 256, "decoder_layer": 6, "decoder_head": 2,
                    "decoder_hidden": 256, "conv_filter_size": 1024, "conv_kernel_size": [9, 1],
                    "encoder_dropout": 0.2, "decoder_dropout": 0.2},
    "variance_predictor": {"filter_size": 256, "kernel_size": 3, "dropout": 0.5},
    "variance_embedding": {"pitch_quantization": "linear", "energy_quantization": "linear", "n_bins": 256},
    "multi_speaker": True, "max_seq_len": 1000,
}                                                                      # config/model/base.yaml
DEFAULT_ALGORITHM_CONFIG = {
    "name": "meta_emb_vad", "type": "meta",
    "adapt": {"type": "spk", "speaker_emb": "table",
              "modules": ["speaker_emb", "variance_adaptor", "decoder", "mel_linear", "postnet"],
              "task": {"ways": 1, "shots": 5, "queries": 5, "lr": 0.001},
              "train": {"steps": 5, "meta_batch_size": 8}, "test": {"steps": 100}},
}                                                                      # config/algorithm/meta_emb_vad.yaml
DEFAULT_TRAIN_CONFIG = {
    "optimizer": {"betas": [0.9, 0.98], "eps": 1e-9, "weight_decay": 0.0, "grad_clip_thresh": 1.0, "grad_acc_step": 1,
                  "warm_up_step": 4000, "anneal_steps": [300000, 400000, 500000], "anneal_rate": 0.3},
}                                                                      # config/train/base.yaml


def _metasystem_init(self, preprocess_config=None, model_config=None, train_config=None, algorithm_config=None,
                     log_dir=None, result_dir=None, *, n_speaker: int = 16, device: str = "cuda:0", split: int = 3,
                     use_cuda_graph: bool = True, second_order: bool = True, process_group=None, backend=None,
                     dropout: bool = True, seed: int = 0, overlap_allreduce: Optional[bool] = None):
    self.preprocess_config = preprocess_config
    self.model_config = model_config or DEFAULT_MODEL_CONFIG
    self.train_config = train_config or DEFAULT_TRAIN_CONFIG
    self.algorithm_config = algorithm_config or DEFAULT_ALGORITHM_CONFIG
    ad = self.algorithm_config["adapt"]
    self.adaptation_steps = ad["train"]["steps"]
    self.test_adaptation_steps = ad["test"]["steps"]
    assert self.test_adaptation_steps % self.adaptation_steps == 0        # base_adaptor.py:39
    assert ad.get("speaker_emb", "table") == "table", "only the table speaker embedding is on the hot path"
    self.device = torch.device(device)
    self.second_order = second_order       # reference: first_order = not train  (base_adaptor.py:107)
    # learner.train() (base_adaptor.py:103): dropout is active in every meta_learn forward.  Masks come from a counter
    # hash of (site, pass, element, salt); salt = seed + step * world + rank rides in the support batch's H2D copy.
    self.dropout = dropout
    self.seed = seed
    self._drop_steps = 0
    # `backend` exists for host-logic tests (tests inject the CPU restatement); the product path is CudaOps,
    # which raises without a B200 — there is no fallback.
    self.be = backend if backend is not None else CudaOps(split=split, device=device)
    self.maml = MamlEngine(self.be, self.model_config, n_speaker, ad["modules"], inner_lr=ad["task"]["lr"],
                           max_inner_steps=self.adaptation_steps)
    self.use_cuda_graph = use_cuda_graph
    self.process_group = process_group
    # Data-parallel exchange INSIDE the task step (opt-in: MTTS_OVERLAP_ALLREDUCE=1 or overlap_allreduce=True): the allreduce
    # of the outer gradient is issued by the step itself, the adapted 2/3 of the buffer under the last Hessian-vector pass' walk
    # through the encoder, captured in the CUDA graph with the rest of the step.  Measured on B200: 10.93 -> 10.80 ms/step at 2 GPUs,
    # no change at 8 GPUs (10.97 ms both ways) — so the default stays the single allreduce in optimizer_step(); call close()
    # before destroying the process group when it is on (graphs that contain the collective must go first).
    self.overlap_allreduce = (os.environ.get("MTTS_OVERLAP_ALLREDUCE", "0") == "1") if overlap_allreduce is None else bool(overlap_allreduce)
    self._reduced_in_step = False
    # CUDA-graph cache: one entry (static batches, graph, activation tapes) per input-shape signature.  Real corpora
    # produce a new (L, T) almost every step, so the cache is an LRU of `graph_cache_size` entries (evicted graphs / tapes /
    # pinned rings are freed) and a signature is only captured on its `graph_min_hits`-th sighting; before that the step
    # runs eagerly on throw-away tapes (same kernels, no capture, no synchronisation).
    self._graphs: "OrderedDict[Tuple, list]" = OrderedDict()
    self.graph_cache_size = int(os.environ.get("MTTS_GRAPH_CACHE", "8"))
    self.graph_min_hits = int(os.environ.get("MTTS_GRAPH_MIN_HITS", "2"))
    self._shape_hits: "OrderedDict[Tuple, int]" = OrderedDict()
    self._adapt_cache: "OrderedDict[Tuple, tuple]" = OrderedDict()      # stand-alone adapt(): (static batch, tapes) per shape
    self._last_qdev = None
    self.host_prof = {"upload": 0.0, "replay": 0.0, "optimizer": 0.0, "d2h": 0.0, "step": 0.0}      # host seconds spent enqueueing (diagnostics)
    self._pending_tasks = 0
    self.launches_per_task_step: Optional[int] = None
    self.h2d_bytes_per_step = 0
    self.d2h_bytes_per_step = 6 * 4




def _load_state_dict(self, state_dict, strict: bool = True):
    """Accepts the reference's FastSpeech2 state_dict (optionally with Lightning's 'model.' prefix)."""
    sd = {(k[6:] if k.startswith("model.") else k): v for k, v in state_dict.items()}
    self.maml.load_state_dict(sd)
    self._graphs.clear()


def _state_dict(self):
    return self.maml.state_dict()


def _task_key(self, sup12, qry12, steps: int, first_order: bool, accumulate_scale: Optional[float], reduce_now: bool = False):
    S, Q = sup12[3].shape[0], qry12[3].shape[0]
    Ls, Ts, Lq, Tq = int(sup12[5]), int(sup12[8]), int(qry12[5]), int(qry12[8])
    # accumulate_scale is baked into the captured `axpby(scale, g_task, 1, g_outer)` (absent when None: validation), so it is
    # part of the signature — a graph captured by a validation step must never serve a training step and vice versa.
    return (S, Ls, Ts, Q, Lq, Tq, steps, first_order, accumulate_scale, reduce_now)


def _get_task(self, key):
    """LRU lookup / creation of the graph-cache entry of `key` = [sup static batch, qry static batch, graph, result, tapes]."""
    ent = self._graphs.get(key)
    if ent is not None:
        self._graphs.move_to_end(key)
        return ent
    S, Ls, Ts, Q, Lq, Tq = key[:6]
    max_T = self.model_config["max_seq_len"]
    sb = _StaticBatch(self.device, S, Ls, Ts, S, False, max_T)
    qb = _StaticBatch(self.device, Q, Lq, Tq, S, True, max_T)
    ent = [sb, qb, None, None, self.maml.new_tapes()]
    self._graphs[key] = ent
    while len(self._graphs) > max(1, self.graph_cache_size):
        _, old = self._graphs.popitem(last=False)          # least recently used: drop its graph, tapes and pinned rings
        old.clear()
    return ent


def _run_task(self, sup12, qry12, steps: int, first_order: bool, accumulate_scale: Optional[float], reduce_now: bool = False):
    """reduce_now: this is the last accumulated task of a data-parallel step — the step issues the allreduce itself."""
    key = _task_key(self, sup12, qry12, steps, first_order, accumulate_scale, reduce_now)
    pg = self.process_group
    reduce = (lambda t: torch.distributed.all_reduce(t, group=pg)) if reduce_now else None
    drop_base = 0 if self.dropout else None
    max_T = self.model_config["max_seq_len"]
    if self.use_cuda_graph and key not in self._graphs:
        hits = self._shape_hits.pop(key, 0) + 1
        self._shape_hits[key] = hits
        while len(self._shape_hits) > 1024:
            self._shape_hits.popitem(last=False)
        if hits < self.graph_min_hits:
            # a signature seen for the first time: run it eagerly on throw-away buffers (no capture, no host sync)
            t0 = time.perf_counter()
            sup = batch_from_tuple(sup12, self.device, max_T=max_T)
            qry = batch_from_tuple(qry12, self.device, spk_ids=sup12[2], average_spk=True, max_T=max_T)
            _set_salt(self, self.next_salt() if self.dropout else 0)
            self.h2d_bytes_per_step = sum(t.numel() * t.element_size() for b in (sup, qry) for t in
                                          (b.spk_ids, b.texts, b.src_lens, b.mels, b.mel_lens, b.pitches, b.energies, b.durations))
            self.host_prof["upload"] += time.perf_counter() - t0
            self.maml.use_tapes(self.maml.new_tapes())
            n0 = _ops.launch_count
            result = self.maml.task_step(sup, qry, steps, first_order, accumulate_scale, drop_base, reduce=reduce)
            self.launches_per_task_step = _ops.launch_count - n0
            self._last_qdev = qry
            return result
    ent = _get_task(self, key)
    sb, qb, graph, result, tapes = ent
    self.h2d_bytes_per_step = sb.h2d_bytes + qb.h2d_bytes
    self._last_qdev = qb.dev
    self.maml.use_tapes(tapes)
    self.be.drop_salt = sb.salt
    t0 = time.perf_counter()
    sb.upload(sup12, salt=self.next_salt() if self.dropout else 0)
    qb.upload(qry12, spk_ids=sup12[2])            # query uses the SUPPORT speaker ids, averaged (base_adaptor.py:122)
    self.host_prof["upload"] += time.perf_counter() - t0
    if not self.use_cuda_graph:
        n0 = _ops.launch_count
        result = self.maml.task_step(sb.dev, qb.dev, steps, first_order, accumulate_scale, drop_base, reduce=reduce)
        self.launches_per_task_step = _ops.launch_count - n0
        return result
    if graph is None:
        # eager warm-up run: allocates every tape buffer and sets kernel attributes outside capture
        bn0 = self.maml.bn_batches
        saved = {k: v.clone() for k, v in self.maml.consts.items() if k.endswith(("running_mean", "running_var"))}
        g_outer_saved = self.maml.g_outer_full.clone()
        self.maml.task_step(sb.dev, qb.dev, steps, first_order, accumulate_scale, drop_base, reduce=reduce)     # (every rank warms up: the collective matches)
        torch.cuda.synchronize()
        for k, v in saved.items():
            self.maml.consts[k].copy_(v)           # the warm-up must not advance BatchNorm running statistics
        self.maml.g_outer_full.copy_(g_outer_saved)
        self.maml.bn_batches = bn0
        graph = torch.cuda.CUDAGraph()
        n0 = _ops.launch_count
        # thread_local: the NCCL watchdog thread's event queries must not invalidate a capture that contains the collective
        with torch.cuda.graph(graph, capture_error_mode="thread_local" if reduce_now else "global"):
            result = self.maml.task_step(sb.dev, qb.dev, steps, first_order, accumulate_scale, drop_base, reduce=reduce)
        self.launches_per_task_step = _ops.launch_count - n0
        self.maml.bn_batches = bn0
        ent[2], ent[3] = graph, result
    t0 = time.perf_counter()
    tr = getattr(self, "trace_events", None)
    if tr is not None:
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ea.record()
    graph.replay()
    if tr is not None:
        eb.record()
        tr.append((ea, eb))
    self.host_prof["replay"] += time.perf_counter() - t0
    self.maml.bn_batches += steps + 1
    return ent[3]


def next_salt(self) -> int:
    """Dropout salt of the next task step: distinct per step and per rank (uint32)."""
    dist = torch.distributed.is_initialized()
    world = torch.distributed.get_world_size(self.process_group) if dist else 1
    rank = torch.distributed.get_rank(self.process_group) if dist else 0
    salt = (self.seed * 0x9E3779B1 + self._drop_steps * world + rank) & 0xFFFFFFFF
    self.last_salt = salt
    self._drop_steps += 1
    return salt


def _scale(self) -> float:
    acc = self.train_config["optimizer"].get("grad_acc_step", 1)
    world = torch.distributed.get_world_size(self.process_group) if torch.distributed.is_initialized() else 1
    return 1.0 / (acc * world)


def adapt(self, batch, adaptation_steps: int = 5, learner=None, train: bool = True):
    """base_adaptor.py:98-112.  Returns the number of inner steps held in the fast-weight arenas
    (the reference returns the adapted learner object; here the engine owns the fast weights)."""
    _assert_meta_batch(batch)
    sup12 = batch[0][0][0]
    start = int(learner) if learner is not None else 0
    # static batch + activation tapes per input shape (a small LRU: the tapes' buffers are shape-specific)
    akey = (sup12[3].shape[0], int(sup12[5]), int(sup12[8]))
    ent = self._adapt_cache.pop(akey, None)
    if ent is None:
        ent = (_StaticBatch(self.device, akey[0], akey[1], akey[2], akey[0], False, self.model_config["max_seq_len"]),
               self.maml.new_tapes())
    self._adapt_cache[akey] = ent
    while len(self._adapt_cache) > 4:
        self._adapt_cache.popitem(last=False)
    sb = ent[0]
    self.maml.use_tapes(ent[1])
    self.be.drop_salt = sb.salt
    sb.upload(sup12, salt=self.next_salt() if self.dropout else 0)
    self.maml.adapt(sb.dev, adaptation_steps, start=start, drop_base=start if self.dropout else None,
                    second_order=False)          # no Hessian-vector pass follows a stand-alone adapt(): nothing re-reads P / dP / dS
    return start + adaptation_steps


def meta_learn(self, batch, batch_idx, train: bool = True):
    """base_adaptor.py:114-124 -> (6-tuple losses, 10-tuple predictions).  Also leaves this task's
    outer gradient (Lightning's later `loss.backward()`) accumulated, pre-scaled by 1/(acc*world)."""
    _assert_meta_batch(batch)
    sup12, qry12 = batch[0][0][0], batch[0][1][0]
    steps = min(self.adaptation_steps, self.test_adaptation_steps)
    first_order = (not train) or (not self.second_order)
    acc = self.train_config["optimizer"].get("grad_acc_step", 1)
    dist_on = torch.distributed.is_initialized() and torch.distributed.get_world_size(self.process_group) > 1
    reduce_now = bool(train and dist_on and self.overlap_allreduce and (self._pending_tasks + 1) % acc == 0)
    loss6, out = _run_task(self, sup12, qry12, steps, first_order, _scale(self) if train else None, reduce_now)
    self._pending_tasks += 1 if train else 0
    self._reduced_in_step = self._reduced_in_step or reduce_now
    # D2H read of the 6 losses (the step's result): asynchronous copy into a pinned ring, waited on access
    if not hasattr(self, "_loss_ring"):
        pin = loss6.is_cuda
        self._loss_ring = [torch.zeros(6).pin_memory() if pin else torch.zeros(6) for _ in range(8)]
        self._loss_events = [None] * len(self._loss_ring)
        self._loss_slot = 0
    k = self._loss_slot
    hbuf = self._loss_ring[k]
    self._loss_slot = (k + 1) % len(self._loss_ring)
    t0 = time.perf_counter()
    if self._loss_events[k] is not None:
        self._loss_events[k].synchronize()       # the D2H copy that last used this slot has landed (bounds the host's run-ahead)
    hbuf.copy_(loss6, non_blocking=True)
    ev = None
    if loss6.is_cuda:
        ev = torch.cuda.Event()
        ev.record()
        self._loss_events[k] = ev
    self.host_prof["d2h"] += time.perf_counter() - t0
    losses = LazyLosses(hbuf, ev)
    dev = self._last_qdev
    preds = _Predictions(out, dev, dev.L, dev.T)
    return losses, preds


def training_step(self, batch, batch_idx):
    """meta.py:68-80"""
    t0 = time.perf_counter()
    train_loss, predictions = self.meta_learn(batch, batch_idx, train=True)
    self.host_prof["step"] += time.perf_counter() - t0
    qry_batch = batch[0][1][0]
    return {"loss": LazyScalar(train_loss, 0), "losses": train_loss, "output": predictions, "_batch": qry_batch}


def validation_step(self, batch, batch_idx):
    """meta.py:85-97 (NB: the reference validates with train=True, i.e. second order; no optimizer step follows)."""
    val_loss, predictions = self.meta_learn(batch, batch_idx, train=False)
    return {"losses": val_loss, "output": predictions, "_batch": batch[0][1][0]}


# dropout pass indices of a test step (shared with oracle/fs2_oracle.py: test_time_adaptation)
TEST_RECON_PASS, TEST_SYNTH_PASS = 10000, 20000


def _pred10(out, bt: Batch, free_running: bool):
    """Engine prediction dict -> the reference's 10-tuple (base_adaptor.py:91-95); masks True = padding (tools.py:91-99)."""
    Lq, T = bt.L, int(out["mel"].shape[1])
    dev = bt.src_lens.device
    src_masks = torch.arange(Lq, device=dev)[None, :] >= bt.src_lens[:, None]
    mel_masks = torch.arange(T, device=dev)[None, :] >= out["mel_len"][:, None]
    d_rounded = out["d_rounded"] if free_running else bt.durations
    return (out["mel"], out["postnet"], out["pitch"], out["energy"], out["logd"], d_rounded, src_masks, mel_masks,
            bt.src_lens, out["mel_len"])


def _set_salt(self, salt: int) -> None:
    salt &= 0xFFFFFFFF
    self.be.drop_salt = torch.tensor([salt - (1 << 32) if salt >= (1 << 31) else salt], dtype=torch.int32, device=self.device)


def forward_learner(self, learner, speaker_args, texts, src_lens, max_src_len, mels=None, mel_lens=None, max_mel_len=None,
                    p_targets=None, e_targets=None, d_targets=None, p_control=1.0, e_control=1.0, d_control=1.0,
                    average_spk_emb=False):
    """base_adaptor.py:41-95: one model forward with the learner's (adapted) modules where it has them -> the reference's 10-tuple.
    `learner` is what `adapt()` returned — the number of inner steps held in the fast-weight arenas — or None / 0 / `self.learner`
    for the meta parameters.  Teacher forced when the targets are given, free running otherwise.  Mode follows the reference: an
    adapted learner is in train mode (`adapt()` calls `learner.train()`), the un-adapted one follows `self.training`
    (`.train()` / `.eval()`; Lightning's test loop puts the module in eval mode).  Dropout needs a step salt and is applied only
    inside `training_step` / `test_step`; here it is off."""
    k = 0 if (learner is None or learner is getattr(self, "learner", None)) else int(learner)
    free = d_targets is None
    assert (p_targets is None) == free and (e_targets is None) == free, "give all of (p, e, d) targets or none of them"
    b12 = (None, None, speaker_args, texts, src_lens, max_src_len, mels, mel_lens, max_mel_len, p_targets, e_targets, d_targets)
    bt = batch_from_tuple(b12, self.device, average_spk=average_spk_emb, targets=not free)
    eval_mode = (k == 0) and not getattr(self, "training", True)
    out = self.maml.predict(bt, k, free, eval_mode, None, p_control, e_control, d_control)
    return _pred10(out, bt, free)


def _train(self, mode: bool = True):
    self.training = bool(mode)
    return self


def _test_step(self, batch, batch_idx):
    """base_adaptor.py:153-189 — few-shot adaptation inference (BASELINE configs[4]): evaluate the un-adapted learner
    (eval mode), then `test.steps / train.steps` rounds of first-order adaptation on the support set, after each of which the
    query is reconstructed (teacher forced, with losses) and — at `saving_steps` — synthesised free-running.  The adapted
    learner is in train mode (the reference's adapt() calls `learner.train()` and nothing switches it back), so those
    forwards run with dropout and BatchNorm batch statistics."""
    _assert_meta_batch(batch)
    outputs = {}
    test_cfg = self.algorithm_config["adapt"]["test"]
    saving_steps = test_cfg.get("saving_steps", [5, 10, 20, 50, 100])
    sup12, qry12 = batch[0][0][0], batch[0][1][0]
    outputs["_batch"] = qry12
    m, dev = self.maml, self.device
    sup = batch_from_tuple(sup12, dev)
    qry_tf = batch_from_tuple(qry12, dev, spk_ids=sup12[2], average_spk=True)              # *qry_batch[3:]
    qry_fr = batch_from_tuple(qry12, dev, spk_ids=sup12[2], average_spk=True, targets=False)   # *qry_batch[3:6]
    drop = self.dropout
    if drop:
        _set_salt(self, self.next_salt())

    def recon(adapted, eval_mode, drop_pass):
        out = m.predict(qry_tf, adapted, False, eval_mode, drop_pass)
        loss6 = out["loss6"].clone()
        return {"losses": tuple(loss6[i] for i in range(6)), "output": _pred10(out, qry_tf, False)}

    def synth(adapted, eval_mode, drop_pass):
        return {"output": _pred10(m.predict(qry_fr, adapted, True, eval_mode, drop_pass), qry_fr, True)}

    # the initial model: Lightning's test loop runs the module in eval mode (dropout off, BN running statistics)
    outputs["step_0"] = {"recon": recon(False, True, None)}
    outputs["step_0"].update({"synth": synth(False, True, None)})
    tape = m.engine.new_tape()
    done = 0
    for ft_step in range(self.adaptation_steps, self.test_adaptation_steps + 1, self.adaptation_steps):
        m.adapt_rolling(sup, self.adaptation_steps, tape, fresh=(done == 0), drop_base=done if drop else None)
        done += self.adaptation_steps
        outputs[f"step_{ft_step}"] = {"recon": recon(True, False, TEST_RECON_PASS + ft_step if drop else None)}
        if ft_step in saving_steps:
            outputs[f"step_{ft_step}"].update({"synth": synth(True, False, TEST_SYNTH_PASS + ft_step if drop else None)})
    return outputs


def test_step(self, batch, batch_idx):
    """base_adaptor.py:139-151: one outputs dict per evaluation (one per support utterance in the "1-shot" protocol)."""
    _assert_meta_batch(batch)
    all_outputs = []
    qry12 = batch[0][1][0]
    if self.algorithm_config["adapt"]["test"].get("1-shot", False):
        from .collate import split_reprocess

        sup12 = batch[0][0][0]
        for i in range(len(sup12[0])):                       # Task(batch_size=1, shuffle=False)
            all_outputs.append(_test_step(self, [([split_reprocess(sup12, [i])], [qry12])], batch_idx))
    else:
        all_outputs.append(_test_step(self, batch, batch_idx))
    if torch.distributed.is_initialized():
        torch.distributed.barrier(group=self.process_group)
    return all_outputs


def close(self):
    """Drop every captured graph, static batch and tape.  Call before `torch.distributed.destroy_process_group()`: the graphs of a
    data-parallel run contain the in-step NCCL allreduce, and tearing the communicator down under live graphs can hang."""
    self._graphs.clear()
    self._adapt_cache.clear()
    if torch.cuda.is_available() and self.device.type == "cuda":
        torch.cuda.synchronize(self.device)


def optimizer_step(self):
    """What Lightning does after `accumulate_grad_batches` training_steps: DDP mean-allreduce of the
    outer gradient (one flat buffer, summed; the 1/(acc*world) scale was folded in at accumulation),
    clip_grad_norm_(grad_clip_thresh), Adam, LambdaLR; then zero the accumulation buffer."""
    m = self.maml
    t0 = time.perf_counter()
    if (torch.distributed.is_initialized() and torch.distributed.get_world_size(self.process_group) > 1
            and not self._reduced_in_step):          # (else the last task step of the accumulation window has already reduced it)
        torch.distributed.all_reduce(m.g_outer_full, group=self.process_group)
    # mean of the 6 query losses over every task of the step (all ranks x accumulated micro-steps): meta.py:77-79 sync_dist=True
    self.synced_losses = m.g_outer_full[m.layout.total:m.layout.total + 6].clone()
    opt = self.train_config["optimizer"]
    m.outer_update(1.0, float(opt.get("grad_clip_thresh", 1.0)), tuple(opt["betas"]), float(opt["eps"]),
                   warmup=int(opt.get("warm_up_step", 4000)), anneal_steps=tuple(opt.get("anneal_steps", ())),
                   anneal_rate=float(opt.get("anneal_rate", 0.3)))
    self.be.zero_(m.g_outer_full)
    self._pending_tasks = 0
    self._reduced_in_step = False
    self.host_prof["optimizer"] += time.perf_counter() - t0


def _prefixed_state_dict(self):
    return {("model." + k): v for k, v in self.maml.state_dict().items()}


def on_load_checkpoint(self, checkpoint: dict) -> None:
    """system.py:115-194: make a reference (Lightning) checkpoint loadable — rename the old speaker-table key, reconcile a
    speaker table of another size, drop unknown keys, report missing ones; a changed checkpoint loses its optimizer state."""
    from .checkpoint import adapt_checkpoint

    self.test_global_step = checkpoint.get("global_step", 0)
    self.checkpoint_changes = adapt_checkpoint(checkpoint, _prefixed_state_dict(self), self.preprocess_config,
                                               self.algorithm_config, verbose=getattr(self, "local_rank", 0) == 0 and
                                               getattr(self, "verbose_checkpoint", False))


def load_checkpoint(self, checkpoint: dict) -> None:
    """What Lightning's `Trainer(resume_from_checkpoint=...)` / `trainer.test(ckpt_path=...)` do with a checkpoint dict
    (`torch.load(path)`): on_load_checkpoint, load_state_dict (missing keys keep the current values), optimizer + scheduler
    state (torch.optim.Adam layout over `model.parameters()`), global step."""
    from .checkpoint import import_adam_state

    self.on_load_checkpoint(checkpoint)
    sd = _prefixed_state_dict(self)
    sd.update(checkpoint["state_dict"])
    self.load_state_dict(sd)
    opt = checkpoint.get("optimizer_states")
    if opt:
        import_adam_state(self.maml, opt[0])
    sch = checkpoint.get("lr_schedulers")
    if sch and not opt:
        self.maml.opt_step = int(sch[0].get("last_epoch", 0))
    self.global_step = int(checkpoint.get("global_step", self.maml.opt_step))


def save_checkpoint(self) -> dict:
    """A Lightning-format checkpoint dict the reference can load (`torch.save` it): `state_dict` with the `model.` prefix,
    `global_step`, `optimizer_states` (torch Adam layout), `lr_schedulers` (LambdaLR last_epoch)."""
    from .checkpoint import export_adam_state, export_scheduler_state

    opt = self.train_config["optimizer"]
    return {"global_step": self.maml.opt_step, "epoch": 0, "state_dict": _prefixed_state_dict(self),
            "optimizer_states": [export_adam_state(self.maml, tuple(opt["betas"]), float(opt["eps"]),
                                                   float(opt.get("weight_decay", 0.0)))],
            "lr_schedulers": [export_scheduler_state(self.maml)]}


def on_test_start(self) -> None:
    """system.py:197-212: with `adapt.test.avg_train_spk_emb` on LibriTTS the 39 test speakers' table rows are replaced by the
    mean of the 247 train-clean-100 rows before testing."""
    ad = self.algorithm_config["adapt"]
    if ad.get("speaker_emb", "table") != "table":
        return
    if (self.preprocess_config or {}).get("dataset") == "LibriTTS" and ad.get("test", {}).get("avg_train_spk_emb", False):
        sd = self.maml.state_dict()
        w = sd["speaker_emb.model.weight"]
        w[-39:] = w[:247].mean(dim=0)
        self.load_state_dict(sd)


class MetaSystem:
    """B200-native counterpart of `lightning.systems.meta.MetaSystem` (hot path only)."""

    __init__ = _metasystem_init
    load_state_dict = _load_state_dict
    state_dict = _state_dict
    adapt = adapt
    meta_learn = meta_learn
    training_step = training_step
    validation_step = validation_step
    test_step = test_step
    _test_step = _test_step
    forward_learner = forward_learner
    training = True
    learner = None                      # the reference's `self.learner` handle: the un-adapted meta parameters
    train = _train
    eval = lambda self: _train(self, False)  # noqa: E731
    on_load_checkpoint = on_load_checkpoint
    load_checkpoint = load_checkpoint
    save_checkpoint = save_checkpoint
    on_save_checkpoint = staticmethod(lambda checkpoint: checkpoint)       # system.py:111-113
    on_test_start = on_test_start
    optimizer_step = optimizer_step
    close = close
    next_salt = next_salt
    _on_meta_batch_start = staticmethod(_assert_meta_batch)


class MAML:
    """Compatibility shim for `lightning.systems.utils.MAML` (l2l.algorithms.MAML): exposes `.lr`,
    `.module` names and `clone()`; `adapt_` with a bare loss tensor is not meaningful without an
    autograd graph and raises with a pointer to MetaSystem.adapt / meta_learn."""

    def __init__(self, module, lr, first_order=False, allow_unused=None, allow_nograd=False):
        self.module, self.lr, self.first_order = module, lr, first_order
        self.allow_unused, self.allow_nograd = allow_unused, allow_nograd

    def clone(self, first_order=None, allow_unused=None, allow_nograd=None):
        return MAML(self.module, self.lr, self.first_order if first_order is None else first_order,
                    allow_unused, allow_nograd)

    def adapt_(self, loss, first_order=None, allow_unused=None, allow_nograd=None):
        raise RuntimeError("the B200 engine adapts without an autograd graph: call MetaSystem.adapt(batch, steps) / "
                           "meta_learn(batch, idx) (same signatures as lightning/systems/base_adaptor.py:98-124)")

    adapt = adapt_
