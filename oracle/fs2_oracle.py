"""CPU oracle for the Meta-TTS meta-training hot path — TEST INFRASTRUCTURE ONLY.

A plain-PyTorch (fp32, autograd) restatement of the reference's algorithm for the path
`MetaSystem.training_step -> meta_learn -> adapt -> forward_learner -> FastSpeech2 -> loss`.
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` leg may
import this module; the product (`meta-tts_b200/`) never does and has no CPU fallback.

Pinned how: `oracle/make_golden.py` imports the REAL reference modules from /root/reference (with
import stubs) in the build container and stores (a) their outputs on seeded inputs under
`tests/golden/` and (b) proves this restatement reproduces them (tests/test_oracle_golden.py).
The reference ships no tests / golden vectors of its own, and the MAML update rule
lives in learn2learn (requirements.txt:2, unpinned, absent): its published algorithm is restated in
`maml_task_step` below (clone_module / maml_update / update_module), anchored on the reference's
call sites lightning/systems/base_adaptor.py:98-124 and lightning/systems/utils.py:17-77.

Every function cites the reference file:line it follows (paths relative to /root/reference).
Dropout: `learner.train()` (base_adaptor.py:103) keeps every nn.Dropout / F.dropout active.  torch's CPU
Philox stream cannot be matched by a CUDA kernel, so the masks are defined by a counter
hash instead (`drop_keep`: murmur3 finaliser of element index + site/pass seed + per-step salt; spec in
include/mtts.h) which this file evaluates with exact integer arithmetic and the kernels evaluate on the
device — same sites, same rates, same 1/(1-p) scaling as the reference, and parity holds WITH dropout.
`drop_seed=None` gives identity dropout (used for the goldens from the real reference modules, which were
generated with the Dropout modules neutralised); `drop_seed="torch"` uses torch's own dropout (timing only).
BatchNorm keeps train-mode batch statistics.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]

# ------------------------------------------------------------------------------------------------
# configuration (config/model/base.yaml, config/algorithm/meta_emb_vad.yaml, preprocess/LibriTTS.yaml)
# ------------------------------------------------------------------------------------------------
N_SYMBOLS = 360          # len(text.symbols.symbols), text/symbols.py:21-29  => vocab 361

BASE_MODEL_CONFIG = {
    "transformer": {
        "encoder_layer": 4, "encoder_head": 2, "encoder_hidden": 256,
        "decoder_layer": 6, "decoder_head": 2, "decoder_hidden": 256,
        "conv_filter_size": 1024, "conv_kernel_size": [9, 1],
        "encoder_dropout": 0.2, "decoder_dropout": 0.2,
    },
    "variance_predictor": {"filter_size": 256, "kernel_size": 3, "dropout": 0.5},
    "variance_embedding": {"pitch_quantization": "linear", "energy_quantization": "linear", "n_bins": 256},
    "multi_speaker": True,
    "max_seq_len": 1000,
}
DEFAULT_STATS = {"pitch": [-2.9, 10.2, 180.0, 50.0], "energy": [-1.4, 8.6, 30.0, 20.0]}
ADAPT_MODULES = ["speaker_emb", "variance_adaptor", "decoder", "mel_linear", "postnet"]   # meta_emb_vad.yaml:14-19
N_MEL = 80


def small_model_config(enc_layers=1, dec_layers=1):
    import copy
    c = copy.deepcopy(BASE_MODEL_CONFIG)
    c["transformer"]["encoder_layer"] = enc_layers
    c["transformer"]["decoder_layer"] = dec_layers
    return c


# ------------------------------------------------------------------------------------------------
# parameters: same constructors, same order => same RNG consumption as the reference modules
# ------------------------------------------------------------------------------------------------
def sinusoid_table(n_position: int, d_hid: int) -> torch.Tensor:
    """transformer/Models.py:10-30 (float64 numpy, sin on even / cos on odd columns)."""
    pos = np.arange(n_position, dtype=np.float64)[:, None]
    j = np.arange(d_hid)[None, :]
    table = pos / np.power(10000, 2 * (j // 2) / d_hid)
    table[:, 0::2] = np.sin(table[:, 0::2])
    table[:, 1::2] = np.cos(table[:, 1::2])
    return torch.FloatTensor(table)


def _fft_block_params(P: Params, prefix: str, d_model, n_head, d_k, d_inner, ks):
    # MultiHeadAttention.__init__ SubLayers.py:11-27 ; PositionwiseFeedForward.__init__ :63-83
    for name in ("w_qs", "w_ks", "w_vs"):
        lin = nn.Linear(d_model, n_head * d_k)
        P[f"{prefix}.slf_attn.{name}.weight"], P[f"{prefix}.slf_attn.{name}.bias"] = lin.weight, lin.bias
    ln = nn.LayerNorm(d_model)
    P[f"{prefix}.slf_attn.layer_norm.weight"], P[f"{prefix}.slf_attn.layer_norm.bias"] = ln.weight, ln.bias
    fc = nn.Linear(n_head * d_k, d_model)
    P[f"{prefix}.slf_attn.fc.weight"], P[f"{prefix}.slf_attn.fc.bias"] = fc.weight, fc.bias
    w1 = nn.Conv1d(d_model, d_inner, kernel_size=ks[0], padding=(ks[0] - 1) // 2)
    w2 = nn.Conv1d(d_inner, d_model, kernel_size=ks[1], padding=(ks[1] - 1) // 2)
    P[f"{prefix}.pos_ffn.w_1.weight"], P[f"{prefix}.pos_ffn.w_1.bias"] = w1.weight, w1.bias
    P[f"{prefix}.pos_ffn.w_2.weight"], P[f"{prefix}.pos_ffn.w_2.bias"] = w2.weight, w2.bias
    ln2 = nn.LayerNorm(d_model)
    P[f"{prefix}.pos_ffn.layer_norm.weight"], P[f"{prefix}.pos_ffn.layer_norm.bias"] = ln2.weight, ln2.bias


def _variance_predictor_params(P: Params, prefix: str, cfg):
    # VariancePredictor.__init__ modules.py:200-240
    d = cfg["transformer"]["encoder_hidden"]
    f = cfg["variance_predictor"]["filter_size"]
    k = cfg["variance_predictor"]["kernel_size"]
    c1 = nn.Conv1d(d, f, kernel_size=k, padding=(k - 1) // 2)
    ln1 = nn.LayerNorm(f)
    c2 = nn.Conv1d(f, f, kernel_size=k, padding=1)
    ln2 = nn.LayerNorm(f)
    lin = nn.Linear(f, 1)
    P[f"{prefix}.conv_layer.conv1d_1.conv.weight"], P[f"{prefix}.conv_layer.conv1d_1.conv.bias"] = c1.weight, c1.bias
    P[f"{prefix}.conv_layer.layer_norm_1.weight"], P[f"{prefix}.conv_layer.layer_norm_1.bias"] = ln1.weight, ln1.bias
    P[f"{prefix}.conv_layer.conv1d_2.conv.weight"], P[f"{prefix}.conv_layer.conv1d_2.conv.bias"] = c2.weight, c2.bias
    P[f"{prefix}.conv_layer.layer_norm_2.weight"], P[f"{prefix}.conv_layer.layer_norm_2.bias"] = ln2.weight, ln2.bias
    P[f"{prefix}.linear_layer.weight"], P[f"{prefix}.linear_layer.bias"] = lin.weight, lin.bias


def init_params(seed: int = 0, model_config=None, n_speaker: int = 16, stats=None) -> Params:
    """Build the FastSpeech2 state_dict exactly as `FastSpeech2.__init__` would
    (lightning/model/fastspeech2.py:19-33) under torch.manual_seed(seed): identical module
    constructors in identical order, so the values equal the reference's bit for bit."""
    cfg = model_config or BASE_MODEL_CONFIG
    stats = stats or DEFAULT_STATS
    tr = cfg["transformer"]
    torch.manual_seed(seed)
    P: Params = OrderedDict()
    # Encoder.__init__ Models.py:36-71
    d = tr["encoder_hidden"]
    emb = nn.Embedding(N_SYMBOLS + 1, d, padding_idx=0)
    P["encoder.src_word_emb.weight"] = emb.weight
    P["encoder.position_enc"] = nn.Parameter(sinusoid_table(cfg["max_seq_len"] + 1, d).unsqueeze(0), requires_grad=False)
    for i in range(tr["encoder_layer"]):
        _fft_block_params(P, f"encoder.layer_stack.{i}", d, tr["encoder_head"], d // tr["encoder_head"],
                          tr["conv_filter_size"], tr["conv_kernel_size"])
    # VarianceAdaptor.__init__ modules.py:20-78
    _variance_predictor_params(P, "variance_adaptor.duration_predictor", cfg)
    _variance_predictor_params(P, "variance_adaptor.pitch_predictor", cfg)
    _variance_predictor_params(P, "variance_adaptor.energy_predictor", cfg)
    n_bins = cfg["variance_embedding"]["n_bins"]
    pmin, pmax = stats["pitch"][:2]
    emin, emax = stats["energy"][:2]
    P["variance_adaptor.pitch_bins"] = nn.Parameter(torch.linspace(pmin, pmax, n_bins - 1), requires_grad=False)
    P["variance_adaptor.energy_bins"] = nn.Parameter(torch.linspace(emin, emax, n_bins - 1), requires_grad=False)
    P["variance_adaptor.pitch_embedding.weight"] = nn.Embedding(n_bins, d).weight
    P["variance_adaptor.energy_embedding.weight"] = nn.Embedding(n_bins, d).weight
    # Decoder.__init__ Models.py:106-137
    dd = tr["decoder_hidden"]
    P["decoder.position_enc"] = nn.Parameter(sinusoid_table(cfg["max_seq_len"] + 1, dd).unsqueeze(0), requires_grad=False)
    for i in range(tr["decoder_layer"]):
        _fft_block_params(P, f"decoder.layer_stack.{i}", dd, tr["decoder_head"], dd // tr["decoder_head"],
                          tr["conv_filter_size"], tr["conv_kernel_size"])
    # mel_linear fastspeech2.py:26-29
    ml = nn.Linear(dd, N_MEL)
    P["mel_linear.weight"], P["mel_linear.bias"] = ml.weight, ml.bias
    # PostNet.__init__ Layers.py:72-127  (80 -> 512 x3 -> 80, k=5, each followed by BatchNorm1d)
    chans = [N_MEL, 512, 512, 512, 512, N_MEL]
    for i in range(5):
        conv = nn.Conv1d(chans[i], chans[i + 1], kernel_size=5, stride=1, padding=2, dilation=1, bias=True)
        bn = nn.BatchNorm1d(chans[i + 1])
        P[f"postnet.convolutions.{i}.0.conv.weight"], P[f"postnet.convolutions.{i}.0.conv.bias"] = conv.weight, conv.bias
        P[f"postnet.convolutions.{i}.1.weight"], P[f"postnet.convolutions.{i}.1.bias"] = bn.weight, bn.bias
        P[f"postnet.convolutions.{i}.1.running_mean"] = bn.running_mean
        P[f"postnet.convolutions.{i}.1.running_var"] = bn.running_var
        P[f"postnet.convolutions.{i}.1.num_batches_tracked"] = bn.num_batches_tracked
    # SpeakerEncoder "table" speaker_encoder.py:47-51
    P["speaker_emb.model.weight"] = nn.Embedding(n_speaker, d).weight
    return P


def is_trainable(name: str, t: torch.Tensor) -> bool:
    return t.is_floating_point() and not name.endswith(("position_enc", "_bins", "running_mean", "running_var"))


def trainable_names(P: Params) -> List[str]:
    return [k for k, v in P.items() if is_trainable(k, v)]


def adapted_names(P: Params, modules: Sequence[str] = ADAPT_MODULES) -> List[str]:
    """Parameters inside `learner = MAML(ModuleDict{adapt.modules})` that require grad
    (base_adaptor.py:31-35; l2l MAML.adapt(allow_nograd=True) differentiates only those)."""
    return [k for k in trainable_names(P) if k.split(".")[0] in modules]


# ------------------------------------------------------------------------------------------------
# dropout: the counter-based hash of include/mtts.h evaluated on the host (exact integer arithmetic), so the
# oracle and the kernels drop the SAME elements and parity holds in train mode with dropout active
# ------------------------------------------------------------------------------------------------
def _mul32(h: torch.Tensor, c: int) -> torch.Tensor:
    lo = (h * (c & 0xFFFF)) & 0xFFFFFFFF
    hi = ((h * (c >> 16)) & 0xFFFF) << 16
    return (lo + hi) & 0xFFFFFFFF


def drop_keep(thr: int, seed: int, numel: int) -> torch.Tensor:
    """bool [numel]: element i kept iff (mix32(i + seed*0x9E3779B9) >> 8) >= thr  (seed = the EFFECTIVE seed)."""
    c = ((int(seed) & 0xFFFFFFFF) * 0x9E3779B9) & 0xFFFFFFFF
    h = (torch.arange(numel, dtype=torch.int64) + c) & 0xFFFFFFFF
    h = h ^ (h >> 16)
    h = _mul32(h, 0x85EBCA6B)
    h = h ^ (h >> 13)
    h = _mul32(h, 0xC2B2AE35)
    h = h ^ (h >> 16)
    return (h >> 8) >= int(thr)


def drop_mask(p: float, seed: int, numel: int) -> torch.Tensor:
    """float32 [numel]: 1/(1-p) where kept, 0 where dropped; thr = floor(p*2^24)."""
    if p is None or p <= 0.0 or seed is None:
        return torch.ones(numel)
    return drop_keep(int(p * (1 << 24)), seed, numel).float() * (1.0 / (1.0 - p))


def site_seed(pass_seed, site: str) -> Optional[int]:
    """Effective seed of one dropout site of one forward pass.  pass_seed = (pass_index, salt) or None:
    the launch-time scalar is crc32(site) ^ (pass_index * 2654435761) and the device adds salt * 0x632BE5AB
    (include/mtts.h); an int pass_seed means salt 0."""
    if pass_seed is None:
        return None
    import zlib
    idx, salt = pass_seed if isinstance(pass_seed, tuple) else (pass_seed, 0)
    scalar = (zlib.crc32(site.encode()) ^ ((int(idx) * 2654435761) & 0xFFFFFFFF)) & 0xFFFFFFFF
    return (scalar + (int(salt) & 0xFFFFFFFF) * 0x632BE5AB) & 0xFFFFFFFF


def _drop(x: torch.Tensor, p: float, pass_seed: Optional[int], site: str) -> torch.Tensor:
    if pass_seed is None:
        return x
    if pass_seed == "torch":          # the reference's own nn.Dropout / F.dropout (Philox): used only when TIMING the CPU arm
        return F.dropout(x, p, True)
    return x * drop_mask(p, site_seed(pass_seed, site), x.numel()).reshape(x.shape).to(x.dtype)


# ------------------------------------------------------------------------------------------------
# model forward
# ------------------------------------------------------------------------------------------------
def get_mask_from_lengths(lengths: torch.Tensor, max_len: Optional[int] = None) -> torch.Tensor:
    """utils/tools.py:91-99 (True = padding)."""
    if max_len is None:
        max_len = int(lengths.max().item())
    ids = torch.arange(0, int(max_len), device=lengths.device).unsqueeze(0).expand(lengths.shape[0], -1)
    return ids >= lengths.unsqueeze(1).expand(-1, int(max_len))


def fft_block(P: Params, prefix: str, x, mask, n_head: int, p_drop: float = 0.0, drop_seed: Optional[int] = None):
    """transformer/Layers.py:21-30 -> SubLayers.py:29-57 -> Modules.py:14-25 ; SubLayers.py:85-93."""
    B, Lq, d_model = x.shape
    d_k = d_model // n_head
    residual = x
    q = F.linear(x, P[f"{prefix}.slf_attn.w_qs.weight"], P[f"{prefix}.slf_attn.w_qs.bias"]).view(B, Lq, n_head, d_k)
    k = F.linear(x, P[f"{prefix}.slf_attn.w_ks.weight"], P[f"{prefix}.slf_attn.w_ks.bias"]).view(B, Lq, n_head, d_k)
    v = F.linear(x, P[f"{prefix}.slf_attn.w_vs.weight"], P[f"{prefix}.slf_attn.w_vs.bias"]).view(B, Lq, n_head, d_k)
    q = q.permute(2, 0, 1, 3).contiguous().view(-1, Lq, d_k)
    k = k.permute(2, 0, 1, 3).contiguous().view(-1, Lq, d_k)
    v = v.permute(2, 0, 1, 3).contiguous().view(-1, Lq, d_k)
    slf_attn_mask = mask.unsqueeze(1).expand(-1, Lq, -1).repeat(n_head, 1, 1)
    attn = torch.bmm(q, k.transpose(1, 2)) / np.power(d_k, 0.5)
    attn = attn.masked_fill(slf_attn_mask, -np.inf)
    attn = torch.softmax(attn, dim=2)
    out = torch.bmm(attn, v)
    out = out.view(n_head, B, Lq, d_k).permute(1, 2, 0, 3).contiguous().view(B, Lq, -1)
    out = F.linear(out, P[f"{prefix}.slf_attn.fc.weight"], P[f"{prefix}.slf_attn.fc.bias"])
    out = _drop(out, p_drop, drop_seed, f"{prefix}.slf_attn")                                  # SubLayers.py:54
    out = F.layer_norm(out + residual, (d_model,), P[f"{prefix}.slf_attn.layer_norm.weight"],
                       P[f"{prefix}.slf_attn.layer_norm.bias"])
    out = out.masked_fill(mask.unsqueeze(-1), 0)
    residual = out
    w1, b1 = P[f"{prefix}.pos_ffn.w_1.weight"], P[f"{prefix}.pos_ffn.w_1.bias"]
    w2, b2 = P[f"{prefix}.pos_ffn.w_2.weight"], P[f"{prefix}.pos_ffn.w_2.bias"]
    h = F.conv1d(out.transpose(1, 2), w1, b1, padding=(w1.shape[2] - 1) // 2)
    h = F.conv1d(F.relu(h), w2, b2, padding=(w2.shape[2] - 1) // 2).transpose(1, 2)
    h = _drop(h.contiguous(), p_drop, drop_seed, f"{prefix}.pos_ffn")                          # SubLayers.py:90
    out = F.layer_norm(h + residual, (d_model,), P[f"{prefix}.pos_ffn.layer_norm.weight"],
                       P[f"{prefix}.pos_ffn.layer_norm.bias"])
    return out.masked_fill(mask.unsqueeze(-1), 0)


def encoder(P: Params, cfg, src_seq, mask, drop_seed: Optional[int] = None):
    """transformer/Models.py:73-100 (training branch)."""
    L = src_seq.shape[1]
    x = F.embedding(src_seq, P["encoder.src_word_emb.weight"], padding_idx=0) + P["encoder.position_enc"][:, :L, :]
    for i in range(cfg["transformer"]["encoder_layer"]):
        x = fft_block(P, f"encoder.layer_stack.{i}", x, mask, cfg["transformer"]["encoder_head"],
                      cfg["transformer"]["encoder_dropout"], drop_seed)
    return x


def decoder(P: Params, cfg, enc_seq, mask, drop_seed: Optional[int] = None, training: bool = True):
    """transformer/Models.py:139-171: train mode truncates to max_seq_len (:161-166); under model.eval() a longer sequence
    keeps its length and gets a freshly computed sinusoid table (:148-156)."""
    if not training and enc_seq.shape[1] > cfg["max_seq_len"]:
        max_len = enc_seq.shape[1]
        x = enc_seq + sinusoid_table(max_len, enc_seq.shape[2])[:max_len, :].unsqueeze(0).to(enc_seq.dtype)
    else:
        max_len = min(enc_seq.shape[1], cfg["max_seq_len"])
        x = enc_seq[:, :max_len, :] + P["decoder.position_enc"][:, :max_len, :]
        mask = mask[:, :max_len]
    for i in range(cfg["transformer"]["decoder_layer"]):
        x = fft_block(P, f"decoder.layer_stack.{i}", x, mask, cfg["transformer"]["decoder_head"],
                      cfg["transformer"]["decoder_dropout"], drop_seed)
    return x, mask


def variance_predictor(P: Params, prefix: str, x, mask, p_drop: float = 0.0, drop_seed: Optional[int] = None):
    """lightning/model/modules.py:242-250 (+ Conv 291-296)."""
    c = f"{prefix}.conv_layer"
    h = F.conv1d(x.transpose(1, 2), P[f"{c}.conv1d_1.conv.weight"], P[f"{c}.conv1d_1.conv.bias"], padding=1).transpose(1, 2)
    h = F.layer_norm(F.relu(h), (h.shape[-1],), P[f"{c}.layer_norm_1.weight"], P[f"{c}.layer_norm_1.bias"])
    h = _drop(h.contiguous(), p_drop, drop_seed, f"{prefix}.1")                                # modules.py:223
    h = F.conv1d(h.transpose(1, 2), P[f"{c}.conv1d_2.conv.weight"], P[f"{c}.conv1d_2.conv.bias"], padding=1).transpose(1, 2)
    h = F.layer_norm(F.relu(h), (h.shape[-1],), P[f"{c}.layer_norm_2.weight"], P[f"{c}.layer_norm_2.bias"])
    h = _drop(h.contiguous(), p_drop, drop_seed, f"{prefix}.2")                                # modules.py:235
    out = F.linear(h, P[f"{prefix}.linear_layer.weight"], P[f"{prefix}.linear_layer.bias"]).squeeze(-1)
    if mask is not None:
        out = out.masked_fill(mask, 0.0)
    return out


def length_regulator_ref(x, duration, max_len):
    """lightning/model/modules.py:167-194 + utils/tools.py:304-322, loop for loop."""
    outs, mel_len = [], []
    for batch, expand_target in zip(x, duration):
        rows = []
        for i, vec in enumerate(batch):
            rows.append(vec.expand(max(int(expand_target[i].item()), 0), -1))
        e = torch.cat(rows, 0)
        outs.append(e)
        mel_len.append(e.shape[0])
    if max_len is None:
        max_len = max(mel_len)
    padded = torch.stack([F.pad(o, (0, 0, 0, int(max_len) - o.size(0)), "constant", 0.0) for o in outs])
    return padded, torch.LongTensor(mel_len)


def variance_adaptor(P: Params, x, src_mask, mel_mask, max_len, pitch_target, energy_target, duration_target,
                     p_control=1.0, e_control=1.0, d_control=1.0, p_drop: float = 0.0, drop_seed: Optional[int] = None):
    """lightning/model/modules.py:102-158, phoneme-level pitch/energy (preprocess/LibriTTS.yaml:37,40)."""
    va = "variance_adaptor"
    log_d = variance_predictor(P, f"{va}.duration_predictor", x, src_mask, p_drop, drop_seed)
    p_pred = variance_predictor(P, f"{va}.pitch_predictor", x, src_mask, p_drop, drop_seed)
    if pitch_target is not None:
        p_emb = F.embedding(torch.bucketize(pitch_target, P[f"{va}.pitch_bins"]), P[f"{va}.pitch_embedding.weight"])
    else:
        p_pred = p_pred * p_control
        p_emb = F.embedding(torch.bucketize(p_pred, P[f"{va}.pitch_bins"]), P[f"{va}.pitch_embedding.weight"])
    x = x + p_emb
    e_pred = variance_predictor(P, f"{va}.energy_predictor", x, src_mask, p_drop, drop_seed)
    if energy_target is not None:
        e_emb = F.embedding(torch.bucketize(energy_target, P[f"{va}.energy_bins"]), P[f"{va}.energy_embedding.weight"])
    else:
        e_pred = e_pred * e_control
        e_emb = F.embedding(torch.bucketize(e_pred, P[f"{va}.energy_bins"]), P[f"{va}.energy_embedding.weight"])
    x = x + e_emb
    if duration_target is not None:
        x, mel_len = length_regulator_ref(x, duration_target, max_len)
        d_rounded = duration_target
    else:
        d_rounded = torch.clamp(torch.round(torch.exp(log_d) - 1) * d_control, min=0)
        x, mel_len = length_regulator_ref(x, d_rounded, max_len)
        mel_mask = get_mask_from_lengths(mel_len)
    return x, p_pred, e_pred, log_d, d_rounded, mel_len, mel_mask


def postnet(P: Params, x, training: bool = True, drop_seed: Optional[int] = None):
    """transformer/Layers.py:129-137: 4x tanh(BN(conv5)) + BN(conv5); dropout = id.
    BatchNorm1d in train mode: batch statistics over all B*T positions (padded frames included),
    running stats updated in place (momentum 0.1, unbiased variance)."""
    h = x.contiguous().transpose(1, 2)
    for i in range(5):
        pre = f"postnet.convolutions.{i}"
        h = F.conv1d(h, P[f"{pre}.0.conv.weight"], P[f"{pre}.0.conv.bias"], padding=2)
        if training:
            P[f"{pre}.1.num_batches_tracked"].add_(1)
        h = F.batch_norm(h, P[f"{pre}.1.running_mean"], P[f"{pre}.1.running_var"], P[f"{pre}.1.weight"],
                         P[f"{pre}.1.bias"], training, 0.1, 1e-5)
        if i < 4:
            h = torch.tanh(h)
        # F.dropout(..., 0.5, self.training) Layers.py:133-134; the kernels index elements token-major [B, T, C]
        h = _drop(h.transpose(1, 2).contiguous(), 0.5, drop_seed, f"postnet.{i}").transpose(1, 2)
    return h.contiguous().transpose(1, 2)


def fs2_forward(P: Params, cfg, speaker_args, texts, src_lens, max_src_len, mels=None, mel_lens=None,
                max_mel_len=None, p_targets=None, e_targets=None, d_targets=None,
                p_control=1.0, e_control=1.0, d_control=1.0, average_spk_emb=False, training=True,
                drop_seed: Optional[int] = None):
    """lightning/systems/base_adaptor.py:41-95 (`forward_learner`; maths identical to
    lightning/model/fastspeech2.py:40-112).  Returns the reference's 10-tuple."""
    max_src_len = int(max_src_len)
    src_masks = get_mask_from_lengths(src_lens, max_src_len)
    output = encoder(P, cfg, texts, src_masks, drop_seed)
    mel_masks = get_mask_from_lengths(mel_lens, int(max_mel_len)) if mel_lens is not None else None
    spk_emb = F.embedding(speaker_args, P["speaker_emb.model.weight"])            # speaker_encoder.py:62-65
    if average_spk_emb:
        spk_emb = spk_emb.mean(dim=0, keepdim=True).expand(output.shape[0], -1)   # base_adaptor.py:66-67
    output = output + spk_emb.unsqueeze(1).expand(-1, max_src_len, -1)
    (output, p_pred, e_pred, log_d_pred, d_rounded, mel_lens_out, mel_masks) = variance_adaptor(
        P, output, src_masks, mel_masks, max_mel_len, p_targets, e_targets, d_targets, p_control, e_control, d_control,
        cfg["variance_predictor"]["dropout"], drop_seed)
    output = output + spk_emb.unsqueeze(1).expand(-1, int(max(mel_lens_out)), -1)   # base_adaptor.py:80-84
    output, mel_masks = decoder(P, cfg, output, mel_masks, drop_seed, training)
    output = F.linear(output, P["mel_linear.weight"], P["mel_linear.bias"])
    postnet_output = postnet(P, output, training, drop_seed) + output
    return (output, postnet_output, p_pred, e_pred, log_d_pred, d_rounded, src_masks, mel_masks, src_lens, mel_lens_out)


def fs2_loss(inputs, predictions):
    """lightning/model/loss.py:19-92 (phoneme-level pitch/energy): masked L1 x2 + masked MSE x3."""
    mel_targets, _, _, pitch_targets, energy_targets, duration_targets = inputs[6:]
    (mel_pred, post_pred, pitch_pred, energy_pred, log_d_pred, _, src_masks, mel_masks, _, _) = predictions
    src_masks = ~src_masks
    mel_masks = ~mel_masks
    log_d_targets = torch.log(duration_targets.float() + 1)
    mel_targets = mel_targets[:, : mel_masks.shape[1], :]
    pitch_pred = pitch_pred.masked_select(src_masks)
    pitch_targets = pitch_targets.masked_select(src_masks)
    energy_pred = energy_pred.masked_select(src_masks)
    energy_targets = energy_targets.masked_select(src_masks)
    log_d_pred = log_d_pred.masked_select(src_masks)
    log_d_targets = log_d_targets.masked_select(src_masks)
    mel_pred = mel_pred.masked_select(mel_masks.unsqueeze(-1))
    post_pred = post_pred.masked_select(mel_masks.unsqueeze(-1))
    mel_targets = mel_targets.masked_select(mel_masks.unsqueeze(-1))
    mel_loss = F.l1_loss(mel_pred, mel_targets)
    post_loss = F.l1_loss(post_pred, mel_targets)
    pitch_loss = F.mse_loss(pitch_pred, pitch_targets)
    energy_loss = F.mse_loss(energy_pred, energy_targets)
    duration_loss = F.mse_loss(log_d_pred, log_d_targets)
    total = mel_loss + post_loss + duration_loss + pitch_loss + energy_loss
    return (total, mel_loss, post_loss, pitch_loss, energy_loss, duration_loss)


# ------------------------------------------------------------------------------------------------
# MAML task step (learn2learn restated; base_adaptor.py:98-124, systems/utils.py:17-77, meta.py:68-80)
# ------------------------------------------------------------------------------------------------
def maml_task_step(P: Params, cfg, sup_batch, qry_batch, adaptation_steps: int, lr: float = 0.001,
                   first_order: bool = False, adapt_modules: Sequence[str] = ADAPT_MODULES,
                   return_fast_weights: bool = False, drop_seed: Optional[int] = None):
    """One task of one outer meta-step.

    learner = self.learner.clone()          l2l clone_module: theta_0[k] = P[k].clone()  (differentiable)
    K x { preds = forward_learner(learner, *sup[2:]); loss = loss_func(sup, preds)[0]
          learner.adapt_(loss, first_order, allow_nograd=True) }
              l2l MAML.adapt: g = autograd.grad(loss, [p for p in params if p.requires_grad],
                                                retain_graph=so, create_graph=so)
              maml_update / update_module: p <- p + (-lr * g)
    predictions = forward_learner(learner, sup[2], *qry[3:], average_spk_emb=True)   base_adaptor.py:122
    valid_error = loss_func(qry, predictions)
    outer gradient = d valid_error[0] / d (all trainable parameters)  — what Lightning's backward
    produces for the DDP allreduce.
    Returns (loss 6-tuple (detached), predictions 10-tuple, {name: outer grad}).
    """
    names_tr = trainable_names(P)
    names_ad = [k for k in names_tr if k.split(".")[0] in adapt_modules]
    leaves = {k: P[k].detach().clone().requires_grad_(True) for k in names_tr}
    base: Params = dict(P)
    base.update(leaves)
    theta = {k: leaves[k].clone() for k in names_ad}            # clone_module
    second_order = not first_order

    def merged():
        m = dict(base)
        m.update(theta)
        return m

    # dropout: forward pass k of the task uses pass seed drop_seed + k (query pass: + adaptation_steps); None = identity
    if drop_seed is None or drop_seed == "torch":
        ds = lambda k: drop_seed
    else:
        d_base, d_salt = drop_seed if isinstance(drop_seed, tuple) else (drop_seed, 0)
        ds = lambda k: (d_base + k, d_salt)
    for step_i in range(adaptation_steps):
        preds = fs2_forward(merged(), cfg, *sup_batch[2:], drop_seed=ds(step_i))
        loss = fs2_loss(sup_batch, preds)[0]
        keys = list(theta.keys())
        grads = torch.autograd.grad(loss, [theta[k] for k in keys], retain_graph=second_order,
                                    create_graph=second_order, allow_unused=False)
        theta = {k: theta[k] + (-lr * g) for k, g in zip(keys, grads)}
    predictions = fs2_forward(merged(), cfg, sup_batch[2], *qry_batch[3:], average_spk_emb=True, drop_seed=ds(adaptation_steps))
    valid = fs2_loss(qry_batch, predictions)
    outer = torch.autograd.grad(valid[0], [leaves[k] for k in names_tr], allow_unused=True)
    grads = {k: (g if g is not None else torch.zeros_like(leaves[k])) for k, g in zip(names_tr, outer)}
    losses = tuple(v.detach() for v in valid)
    preds_d = tuple(t.detach() if torch.is_tensor(t) and t.is_floating_point() else t for t in predictions)
    if return_fast_weights:
        return losses, preds_d, grads, {k: v.detach() for k, v in theta.items()}
    return losses, preds_d, grads


# pass indices of the counter-hash dropout during a test step (shared with meta-tts_b200/systems.py: test_step):
# adaptation step s (0-based, counted across the adapt() rounds) -> s; the teacher-forced "recon" forward after ft_step
# inner steps -> TEST_RECON_PASS + ft_step; the free-running "synth" forward -> TEST_SYNTH_PASS + ft_step.
TEST_RECON_PASS, TEST_SYNTH_PASS = 10000, 20000


def test_time_adaptation(P: Params, cfg, sup_batch, qry_batch, adaptation_steps: int, test_adaptation_steps: int,
                         saving_steps=(5, 10, 20, 50, 100), lr: float = 0.001, adapt_modules: Sequence[str] = ADAPT_MODULES,
                         drop_seed=None):
    """`BaseAdaptorSystem._test_step` (lightning/systems/base_adaptor.py:160-189) — few-shot adaptation inference
    (BASELINE configs[4]).

    step_0: the un-adapted learner.  Under `trainer.test` Lightning puts the module in eval mode and `self.learner` wraps
    the model's own submodules, so these two forwards run with dropout off and BatchNorm running statistics:
        recon = forward_learner(self.learner, sup[2], *qry[3:], average_spk_emb=True) + loss    (teacher forced)
        synth = forward_learner(self.learner, sup[2], *qry[3:6], average_spk_emb=True)          (free running)
    then for ft_step = adaptation_steps, 2*adaptation_steps, ..., test_adaptation_steps:
        learner = self.adapt(batch, adaptation_steps, learner=learner, train=False)   first call: clone() + .train()
            -> FIRST-ORDER inner steps (first_order = not train, base_adaptor.py:107), continuing from the previous
               learner (base_adaptor.py:101-103): theta <- theta - lr * grad, no graph kept
        recon (+ synth when ft_step in saving_steps) with the adapted learner, which is in TRAIN mode: dropout active
        (drop_seed) and BatchNorm batch statistics; l2l's clone_module shares buffers that do not require grad, so every
        one of these forwards also advances the model's running statistics (P is updated in place, as the reference).
    Returns {f"step_{k}": {"recon": {"losses": 6-tuple, "output": 10-tuple}, "synth": {"output": 10-tuple}}}.
    """
    assert test_adaptation_steps % adaptation_steps == 0                         # base_adaptor.py:39
    names_ad = [k for k in trainable_names(P) if k.split(".")[0] in adapt_modules]
    if drop_seed is None:
        ds = lambda k: None  # noqa: E731
    else:
        d_base, d_salt = drop_seed if isinstance(drop_seed, tuple) else (drop_seed, 0)
        ds = lambda k: (d_base + k, d_salt)  # noqa: E731
    det = lambda tup: tuple(t.detach() if torch.is_tensor(t) and t.is_floating_point() else t for t in tup)  # noqa: E731
    out = {}
    with torch.no_grad():
        preds = fs2_forward(P, cfg, sup_batch[2], *qry_batch[3:], average_spk_emb=True, training=False)
        out["step_0"] = {"recon": {"losses": tuple(v.detach() for v in fs2_loss(qry_batch, preds)), "output": det(preds)}}
        preds = fs2_forward(P, cfg, sup_batch[2], *qry_batch[3:6], average_spk_emb=True, training=False)
        out["step_0"]["synth"] = {"output": det(preds)}
    theta = {k: P[k].detach().clone() for k in names_ad}                         # clone_module
    done = 0
    for ft_step in range(adaptation_steps, test_adaptation_steps + 1, adaptation_steps):
        for _ in range(adaptation_steps):
            leaves = {k: v.detach().requires_grad_(True) for k, v in theta.items()}
            m = dict(P)
            m.update(leaves)
            preds = fs2_forward(m, cfg, *sup_batch[2:], drop_seed=ds(done))
            loss = fs2_loss(sup_batch, preds)[0]
            keys = list(leaves.keys())
            grads = torch.autograd.grad(loss, [leaves[k] for k in keys])
            theta = {k: (leaves[k] + (-lr * g)).detach() for k, g in zip(keys, grads)}
            done += 1
        m = dict(P)
        m.update(theta)
        with torch.no_grad():
            preds = fs2_forward(m, cfg, sup_batch[2], *qry_batch[3:], average_spk_emb=True, drop_seed=ds(TEST_RECON_PASS + ft_step))
            out[f"step_{ft_step}"] = {"recon": {"losses": tuple(v.detach() for v in fs2_loss(qry_batch, preds)),
                                                "output": det(preds)}}
            if ft_step in saving_steps:
                preds = fs2_forward(m, cfg, sup_batch[2], *qry_batch[3:6], average_spk_emb=True,
                                    drop_seed=ds(TEST_SYNTH_PASS + ft_step))
                out[f"step_{ft_step}"]["synth"] = {"output": det(preds)}
    return out, theta


# ------------------------------------------------------------------------------------------------
# iMAML: lightning/systems/imaml.py:51-150, lightning/systems/utils.py:78-189,
# hypertorch/hypergrad/CG_torch.py:6-41
# ------------------------------------------------------------------------------------------------
def split_batch(batch, idxs):
    """lightning/collate.py:63-125 (`split_reprocess`, table speaker ids): rows `idxs` of a 12-tuple, re-trimmed."""
    import numpy as np
    (ids, raw, spk, texts, tl, mtl, mels, ml, mml, pit, ene, dur) = batch
    idxs = np.asarray(idxs)
    stl, sml = tl[idxs], ml[idxs]
    Ls, Ts = stl.max(), sml.max()
    cut = lambda t: t[idxs][:, :Ls] if t.shape[1] == mtl else t[idxs][:, :Ts]  # noqa: E731
    return ([ids[i] for i in idxs], [raw[i] for i in idxs], spk[idxs], texts[idxs][:, :Ls], stl, Ls, mels[idxs][:, :Ts], sml, Ts,
            cut(pit), cut(ene), dur[idxs][:, :Ls])


class SupportTask:
    """lightning/systems/utils.py:78-116 (`Task`): mini-batches of the support set drawn by
    BatchSampler(RandomSampler(range(n)), batch_size, drop_last=True); the iterator restarts (reshuffles) when exhausted."""

    def __init__(self, sup_data, batch_size, shuffle=True):
        from torch.utils.data import BatchSampler, RandomSampler
        self.sup = sup_data
        n = len(sup_data[0])
        self.sampler = BatchSampler(RandomSampler(range(n)) if shuffle else range(n), batch_size=batch_size, drop_last=True)
        self.it = iter(self.sampler)

    def reset_iterator(self):
        self.it = iter(self.sampler)

    def next_batch(self):
        try:
            idxs = next(self.it)
        except StopIteration:
            self.reset_iterator()
            idxs = next(self.it)
        return split_batch(self.sup, idxs)


def cg_solve(Ax, b, max_iter, epsilon):
    """hypertorch/hypergrad/CG_torch.py:6-41, statement for statement (NB: on early exit the last update of x is NOT
    applied — `x_last` is returned)."""
    cat = lambda ts: torch.cat([t.reshape(-1) for t in ts])  # noqa: E731
    x_last = [torch.zeros_like(bb) for bb in b]
    r_last = [bb.clone() for bb in b]
    p_last = [rr.clone() for rr in r_last]
    for _ in range(max_iter):
        Ap = Ax(p_last)
        rTr = torch.sum(cat(r_last) * cat(r_last))
        pAp = torch.sum(cat(p_last) * cat(Ap))
        alpha = rTr / pAp
        x = [xx + alpha * pp for xx, pp in zip(x_last, p_last)]
        r = [rr - alpha * pp for rr, pp in zip(r_last, Ap)]
        r_vec = cat(r)
        if float(torch.norm(r_vec)) < epsilon:
            break
        beta = torch.sum(r_vec * r_vec) / rTr
        p = [rr + beta * pp for rr, pp in zip(r, p_last)]
        x_last, p_last, r_last = x, p, r
    return x_last


def imaml_task_step(P: Params, cfg, sup_batch, qry_batch, adaptation_steps: int, lr: float, reg_param: float, cg_iters: int,
                    batch_size: int, stochastic: bool = True, adapt_modules: Sequence[str] = ADAPT_MODULES, drop_seed=None,
                    cg_eps: float = 1e-10):
    """One iMAML task (imaml.py:51-133 + systems/utils.py:120-189 `CG`), up to the hypergradient (clip / reduce / Adam excluded).

    adapt():   K first-order steps on support MINI-batches of  L(w) + 0.5 * reg * ||theta0 - w||^2   (imaml.py:68-75)
    CG():      b = d L_query(w) / d w;   A v = v - J_fp^T v with fp_map(w) = w - lr * grad[L_mb(w) + 0.5 reg ||theta0 - w||^2]
               (a fresh mini-batch per product when stochastic), i.e. A = lr * (H_mb + reg I);  v = cg(A, b, K iterations)
               hypergradient = (d fp_map / d theta0)^T v + d L_query / d theta0 = lr * reg * v      (theta0 enters fp_map only
               through the proximal term; the query loss is evaluated at w, so its direct term is zero)
    The reference then writes the gradients with `update_tensor_grads(hparams, grads)` where `hparams` are ALL trainable
    model parameters but `grads` only the adapted ones (imaml.py:140-141) — positionally misaligned unless every module is
    adapted; the evident intent (adapted parameters receive their hypergradient, the rest none) is what is returned here.
    Dropout pass indices: inner step s -> s, the query forward -> K, CG product j -> K + 1 + j.
    Returns (query losses, query predictions, {adapted name: hypergradient}, adapted weights w)."""
    names_ad = [k for k in trainable_names(P) if k.split(".")[0] in adapt_modules]
    if drop_seed is None:
        ds = lambda k: None  # noqa: E731
    else:
        d_base, d_salt = drop_seed if isinstance(drop_seed, tuple) else (drop_seed, 0)
        ds = lambda k: (d_base + k, d_salt)  # noqa: E731
    theta0 = {k: P[k].detach() for k in names_ad}
    w = {k: P[k].detach().clone() for k in names_ad}
    task = SupportTask(sup_batch, batch_size)

    def reg_loss(mb, leaves, seed):
        m = dict(P)
        m.update(leaves)
        loss = fs2_loss(mb, fs2_forward(m, cfg, *mb[2:], drop_seed=seed))[0]
        return loss + 0.5 * reg_param * sum(((theta0[k] - leaves[k]) ** 2).sum() for k in names_ad)

    for s in range(adaptation_steps):
        leaves = {k: v.detach().requires_grad_(True) for k, v in w.items()}
        g = torch.autograd.grad(reg_loss(task.next_batch(), leaves, ds(s)), [leaves[k] for k in names_ad])
        w = {k: (leaves[k] - lr * gi).detach() for k, gi in zip(names_ad, g)}
    task.reset_iterator()                                                       # imaml.py:116
    params = {k: v.detach().requires_grad_(True) for k, v in w.items()}
    m = dict(P)
    m.update(params)
    preds = fs2_forward(m, cfg, sup_batch[2], *qry_batch[3:], average_spk_emb=True, drop_seed=ds(adaptation_steps))
    valid = fs2_loss(qry_batch, preds)
    b = list(torch.autograd.grad(valid[0], [params[k] for k in names_ad]))
    counter = [0]

    def Ax(xs):
        mb = task.next_batch() if stochastic else sup_batch
        loss = reg_loss(mb, params, ds(adaptation_steps + 1 + counter[0]))
        counter[0] += 1
        g = torch.autograd.grad(loss, [params[k] for k in names_ad], create_graph=True)
        mapped = [params[k] - lr * gi for k, gi in zip(names_ad, g)]
        J = torch.autograd.grad(mapped, [params[k] for k in names_ad], grad_outputs=xs)
        return [v - j for v, j in zip(xs, J)]

    vs = cg_solve(Ax, b, cg_iters, cg_eps)
    grads = {k: (lr * reg_param * v).detach() for k, v in zip(names_ad, vs)}
    losses = tuple(v.detach() for v in valid)
    preds_d = tuple(t.detach() if torch.is_tensor(t) and t.is_floating_point() else t for t in preds)
    return losses, preds_d, grads, {k: v.detach() for k, v in w.items()}


# ------------------------------------------------------------------------------------------------
# synthetic LibriTTS-shaped tasks — shared by tests, smoke and bench
# ------------------------------------------------------------------------------------------------
def synth_batch(n: int, L: int, T: int, seed: int, speaker: int = 0, ragged: bool = False, n_speaker: int = 16):
    """The reference's 12-tuple (lightning/collate.py:47-60) with synthetic content:
    texts ~ U{1..360}; durations positive ints summing to the utterance's mel length;
    mels ~ N(0,1); pitch/energy ~ N(0,1)."""
    g = torch.Generator().manual_seed(seed)
    texts = torch.zeros(n, L, dtype=torch.long)
    durs = torch.zeros(n, L, dtype=torch.long)
    mels = torch.zeros(n, T, N_MEL)
    pitch = torch.zeros(n, L)
    energy = torch.zeros(n, L)
    src_lens = torch.zeros(n, dtype=torch.long)
    mel_lens = torch.zeros(n, dtype=torch.long)
    for i in range(n):
        Li = L if (not ragged or i == 0) else int(torch.randint(max(2, L // 2), L + 1, (1,), generator=g))
        Ti = T if (not ragged or i == 0) else int(torch.randint(max(Li, T // 2), T + 1, (1,), generator=g))
        w = torch.rand(Li, generator=g) + 0.2
        d = 1 + torch.floor((Ti - Li) * w / w.sum()).long()
        d[0] += Ti - int(d.sum())
        texts[i, :Li] = torch.randint(1, N_SYMBOLS + 1, (Li,), generator=g)
        durs[i, :Li] = d
        mels[i, :Ti] = torch.randn(Ti, N_MEL, generator=g)
        pitch[i, :Li] = torch.randn(Li, generator=g)
        energy[i, :Li] = torch.randn(Li, generator=g)
        src_lens[i], mel_lens[i] = Li, Ti
    spk = torch.full((n,), speaker % n_speaker, dtype=torch.long)
    ids = [f"synth-{seed}-{i}" for i in range(n)]
    return (ids, ids, spk, texts, src_lens, int(src_lens.max()), mels, mel_lens, int(mel_lens.max()),
            pitch, energy, durs)


def synth_task(task: int, shots: int, queries: int, L: int, T: int, rank: int = 0, ragged: bool = False):
    sup = synth_batch(shots, L, T, seed=1000 * task + rank, speaker=task, ragged=ragged)
    qry = synth_batch(queries, L, T, seed=1000 * task + 10 + rank, speaker=task, ragged=ragged)
    return sup, qry
