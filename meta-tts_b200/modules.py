"""Drop-in counterparts of the reference's model classes for the hot path — same constructor
arguments, forward signatures, return values and `state_dict` keys:

    transformer.Models.Encoder / Decoder            (transformer/Models.py:33-171)
    transformer.Layers.FFTBlock / PostNet / ConvNorm (transformer/Layers.py:11-137)
    transformer.SubLayers.MultiHeadAttention / PositionwiseFeedForward (parameter containers)
    lightning.model.modules.VarianceAdaptor / VariancePredictor / LengthRegulator / Conv
    lightning.model.fastspeech2.FastSpeech2, lightning.model.loss.FastSpeech2Loss

Each class builds the same torch.nn parameter containers in the same order as the reference (so
`torch.manual_seed(s)` gives bit-identical initial weights and `load_state_dict` works both ways),
but `forward` runs the hand-written sm_100a kernels through the engine (no torch ops on the compute
path, no autograd: outputs are plain tensors — training goes through `systems.MetaSystem`).
Masks follow the reference convention (True = padding, prefix-valid as produced by
`utils.tools.get_mask_from_lengths`); they are converted to lengths for the kernels.
"""
from __future__ import annotations

import json
import os
from collections import OrderedDict
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn as nn

from .engine import Act, Batch, FS2Engine, N_MEL, N_SYMBOLS, ParamLayout, ParamSet, Tape, const_names
from .ops import CudaOps

_ALL_MODULES = ("encoder", "variance_adaptor", "decoder", "mel_linear", "postnet", "speaker_emb")


def get_sinusoid_encoding_table(n_position, d_hid, padding_idx=None):
    """transformer/Models.py:10-30 (float64 numpy -> FloatTensor)."""
    pos = np.arange(n_position, dtype=np.float64)[:, None]
    j = np.arange(d_hid)[None, :]
    table = pos / np.power(10000, 2 * (j // 2) / d_hid)
    table[:, 0::2] = np.sin(table[:, 0::2])
    table[:, 1::2] = np.cos(table[:, 1::2])
    if padding_idx is not None:
        table[padding_idx] = 0.0
    return torch.FloatTensor(table)


def get_mask_from_lengths(lengths, max_len=None):
    """utils/tools.py:91-99, on the lengths' own device (the reference uses a module-global device)."""
    if max_len is None:
        max_len = int(torch.max(lengths).item())
    ids = torch.arange(0, int(max_len), device=lengths.device).unsqueeze(0).expand(lengths.shape[0], -1)
    return ids >= lengths.unsqueeze(1).expand(-1, int(max_len))


def _lens_from_mask(mask: torch.Tensor) -> torch.Tensor:
    return (~mask).sum(dim=1).to(torch.int64)


# ------------------------------------------------------------------------------------------------
# parameter containers (identical structure / init order to the reference modules)
# ------------------------------------------------------------------------------------------------
class MultiHeadAttention(nn.Module):
    def __init__(self, n_head, d_model, d_k, d_v, dropout=0.1):     # SubLayers.py:11-27
        super().__init__()
        self.n_head, self.d_k, self.d_v = n_head, d_k, d_v
        self.w_qs = nn.Linear(d_model, n_head * d_k)
        self.w_ks = nn.Linear(d_model, n_head * d_k)
        self.w_vs = nn.Linear(d_model, n_head * d_v)
        self.layer_norm = nn.LayerNorm(d_model)
        self.fc = nn.Linear(n_head * d_v, d_model)
        self.dropout_p = dropout


class PositionwiseFeedForward(nn.Module):
    def __init__(self, d_in, d_hid, kernel_size, dropout=0.1):      # SubLayers.py:63-83
        super().__init__()
        self.w_1 = nn.Conv1d(d_in, d_hid, kernel_size=kernel_size[0], padding=(kernel_size[0] - 1) // 2)
        self.w_2 = nn.Conv1d(d_hid, d_in, kernel_size=kernel_size[1], padding=(kernel_size[1] - 1) // 2)
        self.layer_norm = nn.LayerNorm(d_in)
        self.dropout_p = dropout


class ConvNorm(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=1, stride=1, padding=None, dilation=1, bias=True,
                 w_init_gain="linear"):                                   # Layers.py:33-64
        super().__init__()
        if padding is None:
            assert kernel_size % 2 == 1
            padding = int(dilation * (kernel_size - 1) / 2)
        self.conv = nn.Conv1d(in_channels, out_channels, kernel_size=kernel_size, stride=stride, padding=padding,
                              dilation=dilation, bias=bias)


class Conv(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=1, stride=1, padding=0, dilation=1, bias=True, w_init="linear"):
        super().__init__()                                                # modules.py:253-296
        self.conv = nn.Conv1d(in_channels, out_channels, kernel_size=kernel_size, stride=stride, padding=padding,
                              dilation=dilation, bias=bias)


# ------------------------------------------------------------------------------------------------
# engine plumbing shared by the drop-in modules
# ------------------------------------------------------------------------------------------------
class _Runtime:
    """Owns a backend, a flat parameter arena for ONE nn.Module tree (keys = reference names with a
    prefix), and tapes keyed by input shape.  Re-packs when parameters change (version counters)."""

    def __init__(self, cfg, n_speaker: int, device="cuda:0", split: int = 3, backend=None):
        self.cfg = cfg
        self.be = backend if backend is not None else CudaOps(split=split, device=device)
        self.layout = ParamLayout(cfg, n_speaker, adapt_modules=())
        dev = self.be.device
        self.theta = self.be.zeros((self.layout.total,))
        self.theta_hi = self.be.zeros((self.layout.total,), torch.bfloat16)
        self.theta_lo = self.be.zeros((self.layout.total,), torch.bfloat16) if self.be.split == 3 else None
        self.consts: Dict[str, torch.Tensor] = {}
        self.engine: Optional[FS2Engine] = None
        self.tapes: Dict[tuple, Tape] = {}
        self._versions = None
        self.device = dev

    def sync(self, named_tensors: Dict[str, torch.Tensor]) -> None:
        """named_tensors: reference-keyed parameters/buffers available in the calling module (a subset
        of the full model is fine: missing entries stay zero and are never touched by its forward)."""
        vers = tuple((k, v._version, v.data_ptr()) for k, v in named_tensors.items())
        if vers == self._versions:
            return
        lay = self.layout
        flat = torch.zeros(lay.total, dtype=torch.float32)
        for name, e in lay.entries.items():
            if name in named_tensors:
                t = named_tensors[name].detach().to("cpu", torch.float32)
                if len(e.sd_shape) == 3:
                    t = t.permute(2, 0, 1).contiguous()
                flat[e.offset:e.offset + e.numel].copy_(t.reshape(-1))
        self.theta.copy_(flat)
        self.be.split_(self.theta, self.theta_hi, self.theta_lo)
        for name in const_names(self.cfg):
            if name in named_tensors and not name.endswith("num_batches_tracked"):
                t = named_tensors[name].detach().to(torch.float32)
                self.consts[name] = (t[0] if name.endswith("position_enc") else t).contiguous().to(self.device)
        # modules that do not own a table still need the constants the kernels index
        d = self.cfg["transformer"]["encoder_hidden"]
        for nm in ("encoder.position_enc", "decoder.position_enc"):
            self.consts.setdefault(nm, get_sinusoid_encoding_table(self.cfg["max_seq_len"] + 1, d).to(self.device))
        if self.engine is None:
            self.engine = FS2Engine(self.be, self.cfg, self.layout, self.consts)
        self._versions = vers

    def params(self) -> ParamSet:
        return ParamSet(self.layout, self.theta, self.theta_hi, self.theta_lo)

    def tape(self, key) -> Tape:
        if key not in self.tapes:
            self.tapes[key] = self.engine.new_tape()
        return self.tapes[key]

    def dev(self, t, dtype):
        return torch.as_tensor(t).to(device=self.device, dtype=dtype).contiguous()


def _named(module: nn.Module, prefix: str) -> Dict[str, torch.Tensor]:
    out = {}
    for k, v in module.state_dict(keep_vars=True).items():
        out[prefix + k] = v
    return out


class _B200Module(nn.Module):
    """Base: lazily creates the runtime on first forward (so construction needs no GPU)."""
    _prefix = ""
    _backend = None          # tests may inject the CPU restatement; the product default is CudaOps (raises w/o B200)

    def _mirror_bn(self, rt: "_Runtime", postnet: nn.Module) -> None:
        """After a train-mode pass: the engine advanced the BatchNorm running statistics in its constants; copy them into the
        module's buffers (state_dict / checkpoints see them) without triggering a re-pack of the parameters."""
        with torch.no_grad():
            for i in range(5):
                bn = postnet.convolutions[i][1]
                bn.running_mean.copy_(rt.consts[f"postnet.convolutions.{i}.1.running_mean"])
                bn.running_var.copy_(rt.consts[f"postnet.convolutions.{i}.1.running_var"])
                bn.num_batches_tracked += 1
        rt._versions = tuple((k, v._version, v.data_ptr()) for k, v in _named(self, self._prefix).items())

    def _rt(self) -> _Runtime:
        rt = self.__dict__.get("_runtime")
        if rt is None:
            rt = _Runtime(self._cfg, getattr(self, "_n_speaker", 1), backend=self._backend)
            self.__dict__["_runtime"] = rt
        rt.sync(_named(self, self._prefix))
        return rt


def _default_model_config(d_model=256, n_head=2, d_inner=1024, kernel_size=(9, 1), n_layers=1, max_seq_len=1000):
    return {"transformer": {"encoder_layer": n_layers, "encoder_head": n_head, "encoder_hidden": d_model,
                            "decoder_layer": n_layers, "decoder_head": n_head, "decoder_hidden": d_model,
                            "conv_filter_size": d_inner, "conv_kernel_size": list(kernel_size),
                            "encoder_dropout": 0.2, "decoder_dropout": 0.2},
            "variance_predictor": {"filter_size": d_model, "kernel_size": 3, "dropout": 0.5},
            "variance_embedding": {"pitch_quantization": "linear", "energy_quantization": "linear", "n_bins": 256},
            "multi_speaker": True, "max_seq_len": max_seq_len}


# ------------------------------------------------------------------------------------------------
# transformer.*
# ------------------------------------------------------------------------------------------------
class FFTBlock(_B200Module):
    """transformer/Layers.py:11-30.  forward(enc_input, mask, slf_attn_mask) -> (enc_output, None): the
    attention matrix the reference returns is never consumed (Models.py:97-100) and is not materialised."""
    _prefix = "encoder.layer_stack.0."

    def __init__(self, d_model, n_head, d_k, d_v, d_inner, kernel_size, dropout=0.1):
        super().__init__()
        self.slf_attn = MultiHeadAttention(n_head, d_model, d_k, d_v, dropout=dropout)
        self.pos_ffn = PositionwiseFeedForward(d_model, d_inner, kernel_size, dropout=dropout)
        self._cfg = _default_model_config(d_model, n_head, d_inner, kernel_size, 1)
        self._n_head = n_head

    def forward(self, enc_input, mask=None, slf_attn_mask=None):
        rt = self._rt()
        B, T, d = enc_input.shape
        tp = rt.tape(("fft", B, T))
        x = tp.act("x", B, T, d)
        xin = rt.dev(enc_input, torch.float32)
        rt.be.add_rowvec(xin, None, 0, None, B, T, d, x.f32, x.hi, x.lo)
        lens = _lens_from_mask(rt.dev(mask, torch.bool)) if mask is not None else None
        out = rt.engine.fft_fwd(rt.params(), "encoder.layer_stack.0", tp, x, lens, self._n_head)
        return out.f32, None


class Encoder(_B200Module):
    """transformer/Models.py:33-100.  forward(src_seq, mask, return_attns=False) -> [B, L, d]."""
    _prefix = "encoder."

    def __init__(self, config):
        super().__init__()
        tr = config["transformer"]
        n_position = config["max_seq_len"] + 1
        d = tr["encoder_hidden"]
        self.max_seq_len, self.d_model = config["max_seq_len"], d
        self.src_word_emb = nn.Embedding(N_SYMBOLS + 1, d, padding_idx=0)
        self.position_enc = nn.Parameter(get_sinusoid_encoding_table(n_position, d).unsqueeze(0), requires_grad=False)
        self.layer_stack = nn.ModuleList([
            _FFTParams(d, tr["encoder_head"], d // tr["encoder_head"], d // tr["encoder_head"], tr["conv_filter_size"],
                       tr["conv_kernel_size"], dropout=tr["encoder_dropout"]) for _ in range(tr["encoder_layer"])])
        self._cfg = config

    def forward(self, src_seq, mask, return_attns=False):
        assert src_seq.shape[1] <= self.max_seq_len, "sequences beyond max_seq_len use the eval-only table path (not on the hot path)"
        rt = self._rt()
        B, Lq = src_seq.shape
        tp = rt.tape(("enc", B, Lq))
        lens = _lens_from_mask(rt.dev(mask, torch.bool))
        out = rt.engine.encoder_fwd(rt.params(), rt.dev(src_seq, torch.int64), lens, B, Lq, tp)
        return out.f32


class Decoder(_B200Module):
    """transformer/Models.py:103-171.  forward(enc_seq, mask, return_attns=False) -> (dec_output, mask)."""
    _prefix = "decoder."

    def __init__(self, config):
        super().__init__()
        tr = config["transformer"]
        n_position = config["max_seq_len"] + 1
        d = tr["decoder_hidden"]
        self.max_seq_len, self.d_model = config["max_seq_len"], d
        self.position_enc = nn.Parameter(get_sinusoid_encoding_table(n_position, d).unsqueeze(0), requires_grad=False)
        self.layer_stack = nn.ModuleList([
            _FFTParams(d, tr["decoder_head"], d // tr["decoder_head"], d // tr["decoder_head"], tr["conv_filter_size"],
                       tr["conv_kernel_size"], dropout=tr["decoder_dropout"]) for _ in range(tr["decoder_layer"])])
        self._cfg = config

    def forward(self, enc_seq, mask, return_attns=False):
        rt = self._rt()
        long_eval = (not self.training) and enc_seq.shape[1] > self.max_seq_len    # Models.py:148-156: keep the length, fresh table
        max_len = enc_seq.shape[1] if long_eval else min(enc_seq.shape[1], self.max_seq_len)      # Models.py:161-166: truncate
        x = rt.dev(enc_seq[:, :max_len, :], torch.float32)
        mask = mask[:, :max_len]
        B, T, _ = x.shape
        tp = rt.tape(("dec", B, T))
        lens = _lens_from_mask(rt.dev(mask, torch.bool))
        out = rt.engine.decoder_fwd(rt.params(), x, None, lens, B, T, tp, eval_mode=long_eval)
        return out.f32, mask


class _FFTParams(nn.Module):
    """Parameter container with FFTBlock's structure, used inside Encoder / Decoder layer stacks."""

    def __init__(self, d_model, n_head, d_k, d_v, d_inner, kernel_size, dropout=0.1):
        super().__init__()
        self.slf_attn = MultiHeadAttention(n_head, d_model, d_k, d_v, dropout=dropout)
        self.pos_ffn = PositionwiseFeedForward(d_model, d_inner, kernel_size, dropout=dropout)


class PostNet(_B200Module):
    """transformer/Layers.py:67-137.  forward(x [B,T,80]) -> [B,T,80]; BatchNorm uses batch statistics in train mode (and advances
    the running ones), the running statistics under eval()."""
    _prefix = "postnet."

    def __init__(self, n_mel_channels=80, postnet_embedding_dim=512, postnet_kernel_size=5, postnet_n_convolutions=5):
        super().__init__()
        assert (n_mel_channels, postnet_embedding_dim, postnet_kernel_size, postnet_n_convolutions) == (80, 512, 5, 5)
        self.convolutions = nn.ModuleList()
        chans = [80, 512, 512, 512, 512, 80]
        for i in range(5):
            self.convolutions.append(nn.Sequential(
                ConvNorm(chans[i], chans[i + 1], kernel_size=5, stride=1, padding=2, dilation=1,
                         w_init_gain="tanh" if i < 4 else "linear"),
                nn.BatchNorm1d(chans[i + 1])))
        self._cfg = _default_model_config()

    def forward(self, x):
        rt = self._rt()
        B, T, _ = x.shape
        tp = rt.tape(("post", B, T))
        mel = tp.act("mel", B, T, N_MEL)
        rt.be.add_rowvec(rt.dev(x, torch.float32), None, 0, None, B, T, N_MEL, mel.f32, mel.hi, mel.lo)
        out = rt.engine.postnet_fwd(rt.params(), mel, tp, update_bn=self.training, eval_mode=not self.training)   # eval: running statistics
        if self.training:                                                    # running stats live in the runtime: mirror back
            self._mirror_bn(rt, self)
        return out.f32


# ------------------------------------------------------------------------------------------------
# lightning.model.modules.*
# ------------------------------------------------------------------------------------------------
class LengthRegulator(nn.Module):
    """lightning/model/modules.py:161-194.  forward(x, duration, max_len) -> (output, mel_len)."""

    def __init__(self):
        super().__init__()

    def forward(self, x, duration, max_len):
        from . import ops
        dev = x.device if x.is_cuda else torch.device("cuda:0")
        xx = x.to(dev, torch.float32).contiguous()
        dur = duration.to(dev)
        dur = dur.contiguous() if dur.dtype == torch.int64 else dur.float().contiguous()
        if max_len is None:                                                     # pad() without a max: data-dependent
            max_len = int(dur.clamp_min(0).long().sum(1).max().item())
        idx, mel_len = ops.length_regulate_index(dur, int(max_len))
        return ops.length_regulate_fwd(xx, idx), mel_len


class VariancePredictor(_B200Module):
    """lightning/model/modules.py:197-250.  forward(encoder_output, mask) -> [B, L]."""
    _prefix = "variance_adaptor.duration_predictor."

    def __init__(self, model_config):
        super().__init__()
        self.input_size = model_config["transformer"]["encoder_hidden"]
        self.filter_size = model_config["variance_predictor"]["filter_size"]
        self.kernel = model_config["variance_predictor"]["kernel_size"]
        self.conv_output_size = self.filter_size
        self.dropout = model_config["variance_predictor"]["dropout"]
        self.conv_layer = nn.Sequential(OrderedDict([
            ("conv1d_1", Conv(self.input_size, self.filter_size, kernel_size=self.kernel, padding=(self.kernel - 1) // 2)),
            ("relu_1", nn.ReLU()), ("layer_norm_1", nn.LayerNorm(self.filter_size)), ("dropout_1", nn.Dropout(self.dropout)),
            ("conv1d_2", Conv(self.filter_size, self.filter_size, kernel_size=self.kernel, padding=1)),
            ("relu_2", nn.ReLU()), ("layer_norm_2", nn.LayerNorm(self.filter_size)), ("dropout_2", nn.Dropout(self.dropout))]))
        self.linear_layer = nn.Linear(self.conv_output_size, 1)
        self._cfg = model_config

    def forward(self, encoder_output, mask):
        rt = self._rt()
        B, Lq, d = encoder_output.shape
        tp = rt.tape(("vp", B, Lq))
        x = tp.act("x", B, Lq, d)
        rt.be.add_rowvec(rt.dev(encoder_output, torch.float32), None, 0, None, B, Lq, d, x.f32, x.hi, x.lo)
        lens = _lens_from_mask(rt.dev(mask, torch.bool)) if mask is not None else None
        out = tp.f32("out", (B, Lq))
        rt.engine.vp_fwd(rt.params(), "variance_adaptor.duration_predictor", tp, x, lens, out)
        return out


class FastSpeech2Loss(nn.Module):
    """lightning/model/loss.py:5-92 (phoneme-level pitch / energy).  forward(inputs, predictions) -> 6 scalars."""

    def __init__(self, preprocess_config=None, model_config=None):
        super().__init__()
        if preprocess_config is not None:
            assert preprocess_config["preprocessing"]["pitch"]["feature"] == "phoneme_level"
            assert preprocess_config["preprocessing"]["energy"]["feature"] == "phoneme_level"

    def forward(self, inputs, predictions):
        be = getattr(self, "_be", None)
        if be is None:
            be = CudaOps(split=3)
            self.__dict__["_be"] = be
        dev = be.device
        to = lambda t, dt: torch.as_tensor(t).to(device=dev, dtype=dt).contiguous()  # noqa: E731
        mel_t, _, _, p_t, e_t, d_t = inputs[6:]
        mel, post, p, e, logd, _, src_masks, mel_masks, _, _ = predictions
        B, T, NM = mel.shape
        Lq = p.shape[1]
        out6, counts, ws = be.zeros((6,)), be.zeros((2,)), be.zeros((8,))
        be.loss_fwd(to(mel, torch.float32), to(post, torch.float32), to(mel_t[:, :T, :], torch.float32),
                    _lens_from_mask(to(mel_masks, torch.bool)), to(p, torch.float32), to(p_t, torch.float32),
                    to(e, torch.float32), to(e_t, torch.float32), to(logd, torch.float32), to(d_t, torch.int64),
                    _lens_from_mask(to(src_masks, torch.bool)), B, T, Lq, NM, ws, out6, counts)
        return tuple(out6[i] for i in range(6))


class FastSpeech2(_B200Module):
    """lightning/model/fastspeech2.py:16-112: same constructor and teacher-forced forward (10-tuple).
    Sub-modules carry the reference's names, so `state_dict()` keys match; the whole forward runs as
    one engine pass."""
    _prefix = ""

    def __init__(self, preprocess_config, model_config, algorithm_config):
        super().__init__()
        self.model_config = model_config
        self.encoder = Encoder(model_config)
        self.variance_adaptor = _VarianceAdaptorParams(preprocess_config, model_config)
        self.decoder = Decoder(model_config)
        self.mel_linear = nn.Linear(model_config["transformer"]["decoder_hidden"],
                                    preprocess_config["preprocessing"]["mel"]["n_mel_channels"])
        self.postnet = PostNet()
        assert algorithm_config["adapt"]["speaker_emb"] == "table", "only the table speaker embedding is on the hot path"
        with open(os.path.join(preprocess_config["path"]["preprocessed_path"], "speakers.json")) as f:
            n_speaker = len(json.load(f))
        self.speaker_emb = _SpeakerTable(n_speaker, model_config["transformer"]["encoder_hidden"])
        self._cfg = model_config
        self._n_speaker = n_speaker

    def forward(self, speaker_args, texts, src_lens, max_src_len, mels=None, mel_lens=None, max_mel_len=None,
                p_targets=None, e_targets=None, d_targets=None, p_control=1.0, e_control=1.0, d_control=1.0):
        rt = self._rt()
        B, Lq = texts.shape[0], int(max_src_len)
        free = d_targets is None
        assert (p_targets is None) == free and (e_targets is None) == free, "give all of (p, e, d) targets or none of them"
        if free:
            # free-running synthesis (fastspeech2.py:73-92 with targets None -> modules.py:85-99,132-139): predicted pitch /
            # energy / durations; model.eval() -> BatchNorm running statistics, train mode -> batch statistics
            bt = Batch(spk_ids=rt.dev(speaker_args, torch.int64), average_spk=False, texts=rt.dev(texts, torch.int64),
                       src_lens=rt.dev(src_lens, torch.int64), mels=None, mel_lens=None, pitches=None, energies=None, durations=None,
                       B=B, L=Lq, T=0)
            out = rt.engine.synthesize(rt.params(), bt, rt.engine.new_tape(), p_control, e_control, d_control,
                                       update_bn=self.training, eval_mode=not self.training)
            if self.training:
                self._mirror_bn(rt, self.postnet)
            src_masks = get_mask_from_lengths(bt.src_lens, Lq)
            mel_masks = get_mask_from_lengths(out["mel_len"], int(out["mel"].shape[1]))
            return (out["mel"], out["postnet"], out["pitch"], out["energy"], out["logd"], out["d_rounded"], src_masks, mel_masks,
                    bt.src_lens, out["mel_len"])
        T = int(max_mel_len)
        bt = Batch(spk_ids=rt.dev(speaker_args, torch.int64), average_spk=False, texts=rt.dev(texts, torch.int64),
                   src_lens=rt.dev(src_lens, torch.int64), mels=rt.dev(mels, torch.float32),
                   mel_lens=rt.dev(mel_lens, torch.int64), pitches=rt.dev(p_targets, torch.float32),
                   energies=rt.dev(e_targets, torch.float32), durations=rt.dev(d_targets, torch.int64), B=B, L=Lq, T=T)
        out = rt.engine.forward(rt.params(), bt, rt.tape(("fs2", B, Lq, T)), update_bn=self.training, eval_mode=not self.training)
        if self.training:
            self._mirror_bn(rt, self.postnet)
        src_masks = get_mask_from_lengths(bt.src_lens, Lq)
        mel_masks = get_mask_from_lengths(bt.mel_lens, T)
        return (out["mel"], out["postnet"], out["pitch"], out["energy"], out["logd"], bt.durations, src_masks, mel_masks,
                bt.src_lens, out["mel_len"])


class _SpeakerTable(nn.Module):
    def __init__(self, n_speaker, d):                                     # speaker_encoder.py:47-51 ("table")
        super().__init__()
        self.model = nn.Embedding(n_speaker, d)


class _VarianceAdaptorParams(_B200Module):
    """lightning/model/modules.py:17-158 `VarianceAdaptor`: same constructor, parameter names and
    `forward(x, src_mask, mel_mask, max_len, pitch_target, energy_target, duration_target, p_control, e_control, d_control)`
    -> `(x [B,T,d], pitch_prediction, energy_prediction, log_duration_prediction, duration_rounded, mel_len, mel_mask)`
    (phoneme-level pitch / energy, linear quantisation: preprocess/LibriTTS.yaml:37,40).  Inside FastSpeech2 the same steps run
    as part of one engine pass; this standalone forward exists for drop-in use of the module alone.  Dropout of the predictors
    (modules.py:223,235) is not applied here: the standalone module has no step salt (use MetaSystem / FastSpeech2 for training)."""
    _prefix = "variance_adaptor."

    def __init__(self, preprocess_config, model_config):
        super().__init__()
        self.duration_predictor = VariancePredictor(model_config)
        self.length_regulator = LengthRegulator()
        self.pitch_predictor = VariancePredictor(model_config)
        self.energy_predictor = VariancePredictor(model_config)
        assert preprocess_config["preprocessing"]["pitch"]["feature"] == "phoneme_level"
        assert preprocess_config["preprocessing"]["energy"]["feature"] == "phoneme_level"
        assert model_config["variance_embedding"]["pitch_quantization"] == "linear"
        assert model_config["variance_embedding"]["energy_quantization"] == "linear"
        n_bins = model_config["variance_embedding"]["n_bins"]
        with open(os.path.join(preprocess_config["path"]["preprocessed_path"], "stats.json")) as f:
            stats = json.load(f)
        pmin, pmax = stats["pitch"][:2]
        emin, emax = stats["energy"][:2]
        self.pitch_bins = nn.Parameter(torch.linspace(pmin, pmax, n_bins - 1), requires_grad=False)
        self.energy_bins = nn.Parameter(torch.linspace(emin, emax, n_bins - 1), requires_grad=False)
        d = model_config["transformer"]["encoder_hidden"]
        self.pitch_embedding = nn.Embedding(n_bins, d)
        self.energy_embedding = nn.Embedding(n_bins, d)
        self._cfg = model_config

    def forward(self, x, src_mask, mel_mask=None, max_len=None, pitch_target=None, energy_target=None, duration_target=None,
                p_control=1.0, e_control=1.0, d_control=1.0):
        from . import lib as L
        rt = self._rt()
        eng, be, P = rt.engine, rt.be, rt.params()
        va = "variance_adaptor"
        B, Lq, d = x.shape
        tp = eng.new_tape()
        tp.drop_pass = None
        x0 = tp.act("va.x0", B, Lq, d)
        be.add_rowvec(rt.dev(x, torch.float32), None, 0, None, B, Lq, d, x0.f32, x0.hi, x0.lo)
        lens = _lens_from_mask(rt.dev(src_mask, torch.bool))
        logd, ppred, epred = tp.f32("logd", (B, Lq)), tp.f32("ppred", (B, Lq)), tp.f32("epred", (B, Lq))
        eng.vp_fwd(P, f"{va}.duration_predictor", tp, x0, lens, logd)
        eng.vp_fwd(P, f"{va}.pitch_predictor", tp, x0, lens, ppred)

        def embed(pred, target, control, bins, table, base, out: Act):
            if target is not None:
                src = rt.dev(target, torch.float32)
            else:
                if control != 1.0:
                    be.unary(L.UN_SCALE, pred, control, 0.0, pred)          # prediction * control (modules.py:86,97)
                src = pred
            idx = tp.buf(f"idx.{bins}", (B, Lq), torch.int64)
            be.bucketize(src, rt.consts[f"{va}.{bins}"], eng.nbins - 1, B * Lq, idx)
            be.embed_fwd(idx, P.get(f"{va}.{table}.weight").f32, base, None, Lq, B * Lq, d, out.f32, out.hi, out.lo)

        x1 = tp.act("va.x1", B, Lq, d)
        embed(ppred, pitch_target, p_control, "pitch_bins", "pitch_embedding", x0.f32, x1)
        eng.vp_fwd(P, f"{va}.energy_predictor", tp, x1, lens, epred)
        x2 = tp.act("va.x2", B, Lq, d, bf=False)
        embed(epred, energy_target, e_control, "energy_bins", "energy_embedding", x1.f32, x2)
        if duration_target is not None:
            dur = rt.dev(duration_target, torch.int64 if not torch.as_tensor(duration_target).is_floating_point() else torch.float32)
            d_rounded = duration_target
        else:
            dur = tp.f32("d_rounded", (B, Lq))
            be.duration_round(logd, d_control, dur)
            d_rounded = dur
        if max_len is None:                                                   # pad() without a max: data dependent, one sync
            max_len = int(dur.detach().to("cpu").to(torch.int64).clamp_(min=0).sum(dim=1).max())
        T = int(max_len)
        idx = tp.buf("lr.idx", (B, T), torch.int32)
        mel_len = tp.buf("lr.mel_len", (B,), torch.int64)
        be.lr_index(dur, T, idx, mel_len)
        out = tp.f32("lr.out", (B, T, d))
        be.lr_fwd(x2.f32, idx, out)
        if duration_target is None:
            mel_mask = get_mask_from_lengths(mel_len)                          # modules.py:139
        return out, ppred, epred, logd, d_rounded, mel_len, mel_mask


VarianceAdaptor = _VarianceAdaptorParams
