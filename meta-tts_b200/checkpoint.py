"""Checkpoint compatibility with the reference: Lightning-format checkpoints of
`lightning/systems/system.py` (`state_dict` with the `model.` prefix, `global_step`, torch-Adam `optimizer_states`,
`lr_schedulers`) map onto the flat arenas of `MamlEngine` and back, and `on_load_checkpoint` / `on_test_start` mirror the
reference's key renames and speaker-table handling (system.py:115-212).  Host-side dictionary plumbing only.
"""
from __future__ import annotations

from typing import Dict, List

import torch

from .engine import const_names, param_specs

BN_BUFFERS = ("running_mean", "running_var", "num_batches_tracked")


def reference_state_dict_keys(cfg) -> List[str]:
    """`FastSpeech2.state_dict()` key order of the reference (module registration order: fastspeech2.py:23-37,
    Models.py:36-71/106-137, SubLayers.py:12-27/69-83, modules.py:20-78/209-240, Layers.py:72-127) — checked against the
    key list of the real reference model in tests/test_checkpoint_cpu.py."""
    tr = cfg["transformer"]
    keys: List[str] = []

    def fft(pf):
        for n in ("w_qs", "w_ks", "w_vs"):
            keys.extend([f"{pf}.slf_attn.{n}.weight", f"{pf}.slf_attn.{n}.bias"])
        keys.extend([f"{pf}.slf_attn.layer_norm.weight", f"{pf}.slf_attn.layer_norm.bias", f"{pf}.slf_attn.fc.weight",
                     f"{pf}.slf_attn.fc.bias", f"{pf}.pos_ffn.w_1.weight", f"{pf}.pos_ffn.w_1.bias", f"{pf}.pos_ffn.w_2.weight",
                     f"{pf}.pos_ffn.w_2.bias", f"{pf}.pos_ffn.layer_norm.weight", f"{pf}.pos_ffn.layer_norm.bias"])

    keys.extend(["encoder.position_enc", "encoder.src_word_emb.weight"])
    for i in range(tr["encoder_layer"]):
        fft(f"encoder.layer_stack.{i}")
    keys.extend(["variance_adaptor.pitch_bins", "variance_adaptor.energy_bins"])
    for p in ("duration_predictor", "pitch_predictor", "energy_predictor"):
        c = f"variance_adaptor.{p}"
        keys.extend([f"{c}.conv_layer.conv1d_1.conv.weight", f"{c}.conv_layer.conv1d_1.conv.bias",
                     f"{c}.conv_layer.layer_norm_1.weight", f"{c}.conv_layer.layer_norm_1.bias",
                     f"{c}.conv_layer.conv1d_2.conv.weight", f"{c}.conv_layer.conv1d_2.conv.bias",
                     f"{c}.conv_layer.layer_norm_2.weight", f"{c}.conv_layer.layer_norm_2.bias",
                     f"{c}.linear_layer.weight", f"{c}.linear_layer.bias"])
    keys.extend(["variance_adaptor.pitch_embedding.weight", "variance_adaptor.energy_embedding.weight", "decoder.position_enc"])
    for i in range(tr["decoder_layer"]):
        fft(f"decoder.layer_stack.{i}")
    keys.extend(["mel_linear.weight", "mel_linear.bias"])
    for i in range(5):
        c = f"postnet.convolutions.{i}"
        keys.extend([f"{c}.0.conv.weight", f"{c}.0.conv.bias", f"{c}.1.weight", f"{c}.1.bias", f"{c}.1.running_mean",
                     f"{c}.1.running_var", f"{c}.1.num_batches_tracked"])
    keys.append("speaker_emb.model.weight")
    return keys


def reference_parameter_keys(cfg) -> List[str]:
    """`model.parameters()` order = state_dict order without buffers: the index space of torch.optim.Adam's state
    (lightning/optimizer.py:6-16 passes `model.parameters()`, frozen tables included)."""
    return [k for k in reference_state_dict_keys(cfg) if not k.endswith(BN_BUFFERS)]


def adapt_checkpoint(checkpoint: dict, model_state_dict: Dict[str, torch.Tensor], preprocess_config=None,
                     algorithm_config=None, verbose: bool = False) -> Dict[str, list]:
    """system.py:115-194 (`on_load_checkpoint`): edits `checkpoint["state_dict"]` in place so that it loads into a model whose
    state_dict is `model_state_dict` (keys WITH the `model.` prefix, as Lightning stores them):
      * old checkpoints' `model.speaker_emb.weight` is renamed to `model.speaker_emb.model.weight`;
      * a speaker table of another size: LibriTTS 326 -> 2390 rows copies the 247 training and the last 79 dev/test rows
        (system.py:139-151); another corpus keeps the model's rows, or the mean of the 247 training rows with
        `adapt.test.avg_train_spk_emb`; any other shape mismatch keeps the model's tensor ("skip");
      * keys the model does not have are dropped, keys the checkpoint lacks are reported ("miss");
      * if anything changed, the optimizer state is discarded (system.py:193-194).
    Returns the change log {"skip", "drop", "replace", "miss"}."""
    sd = checkpoint["state_dict"]
    pre = preprocess_config or {}
    algo = (algorithm_config or {}).get("adapt", {})
    changes = {"skip": [], "drop": [], "replace": [], "miss": []}
    changed = False
    if "model.speaker_emb.weight" in sd:
        assert "model.speaker_emb.model.weight" in model_state_dict and "model.speaker_emb.model.weight" not in sd
        sd["model.speaker_emb.model.weight"] = sd.pop("model.speaker_emb.weight")
        changes["replace"].append(["model.speaker_emb.weight", "model.speaker_emb.model.weight"])
        changed = True
    for k in list(sd.keys()):
        if k in model_state_dict:
            if tuple(sd[k].shape) != tuple(model_state_dict[k].shape):
                if k == "model.speaker_emb.model.weight":
                    assert algo.get("speaker_emb", "table") == "table"
                    if pre.get("dataset") == "LibriTTS":
                        assert sd[k].shape[0] == 326 and model_state_dict[k].shape[0] == 2390, \
                            f"state_dict: {tuple(sd[k].shape)}, model: {tuple(model_state_dict[k].shape)}"
                        model_state_dict[k][:247] = sd[k][:247]
                        model_state_dict[k][-79:] = sd[k][-79:]
                    else:
                        assert sd[k].shape[0] in (326, 2390)
                        if algo.get("test", {}).get("avg_train_spk_emb", False):
                            model_state_dict[k][:] = sd[k][:247].mean(dim=0)
                changes["skip"].append([k, tuple(model_state_dict[k].shape), tuple(sd[k].shape)])
                sd[k] = model_state_dict[k]
                changed = True
        else:
            changes["drop"].append(k)
            changed = True
    for k in model_state_dict:
        if k not in sd:
            changes["miss"].append(k)
            changed = True
    for k in changes["drop"]:
        del sd[k]
    if verbose:
        for a, b in changes["replace"]:
            print(f"Replace: {a}\n\t-> {b}")
        for k, need, have in changes["skip"]:
            print(f"Skip parameter: {k}, \n\trequired shape: {need}, loaded shape: {have}")
        for k in changes["drop"]:
            print(f"Dropping parameter: {k}")
        for k in changes["miss"]:
            print(f"Missing parameter: {k}")
    if changed:
        checkpoint.pop("optimizer_states", None)
    return changes


def export_adam_state(maml, betas, eps, weight_decay: float = 0.0) -> dict:
    """The flat Adam moments as a torch.optim.Adam state_dict over `model.parameters()` (what Lightning stores)."""
    keys = reference_parameter_keys(maml.cfg)
    m = maml.layout.unpack(maml.adam_m.detach().cpu())
    v = maml.layout.unpack(maml.adam_v.detach().cpu())
    state = {}
    if maml.opt_step > 0:
        for i, k in enumerate(keys):
            if k in m:
                state[i] = {"step": torch.tensor(float(maml.opt_step)), "exp_avg": m[k], "exp_avg_sq": v[k]}
    # after N optimizer + scheduler steps torch stores lr = initial_lr * lambda(last_epoch = N): the NEXT step's rate
    group = {"lr": maml.lr_schedule(maml.opt_step), "betas": tuple(betas), "eps": eps, "weight_decay": weight_decay,
             "amsgrad": False, "maximize": False, "foreach": None, "capturable": False, "differentiable": False, "fused": None,
             "initial_lr": maml.cfg["transformer"]["encoder_hidden"] ** -0.5, "params": list(range(len(keys)))}
    return {"state": state, "param_groups": [group]}


def export_scheduler_state(maml) -> dict:
    """torch.optim.lr_scheduler.LambdaLR.state_dict() after `opt_step` scheduler steps (lightning/scheduler.py:6-29): the lambda
    itself is not picklable state (`lr_lambdas: [None]`, which LambdaLR.load_state_dict pops and skips)."""
    base = maml.cfg["transformer"]["encoder_hidden"] ** -0.5
    return {"base_lrs": [base], "last_epoch": maml.opt_step, "_step_count": maml.opt_step + 1, "_is_initial": False,
            "_get_lr_called_within_step": False, "_last_lr": [maml.lr_schedule(maml.opt_step)], "lr_lambdas": [None]}


def import_adam_state(maml, opt_state: dict) -> None:
    keys = reference_parameter_keys(maml.cfg)
    st = opt_state["state"]
    zeros = {k: torch.zeros(e.sd_shape) for k, e in maml.layout.entries.items()}
    m, v, step = dict(zeros), {k: t.clone() for k, t in zeros.items()}, 0
    for i, k in enumerate(keys):
        s = st.get(i, st.get(str(i)))
        if s is None or k not in maml.layout.entries:
            continue
        m[k], v[k] = s["exp_avg"].float(), s["exp_avg_sq"].float()
        step = max(step, int(float(s["step"])))
    flat = torch.zeros(maml.layout.total)
    maml.layout.pack(m, flat)
    maml.adam_m.copy_(flat)
    maml.layout.pack(v, flat)
    maml.adam_v.copy_(flat)
    maml.opt_step = step


__all__ = ["reference_state_dict_keys", "reference_parameter_keys", "adapt_checkpoint", "export_adam_state",
           "export_scheduler_state", "import_adam_state", "const_names", "param_specs"]
