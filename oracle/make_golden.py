"""Generate tests/golden/*.npz from the REAL reference modules (container only; /root/reference).

    python -m oracle.make_golden

What is pinned (the reference has no tests/goldens of its own):
  1. state_dict of `lightning.model.FastSpeech2` built under torch.manual_seed(0): key list, shapes
     and a checksum per tensor  -> proves oracle.init_params reproduces the reference init bit for bit;
  2. `FastSpeech2.forward` (train mode, dropout neutralised) + `FastSpeech2Loss` on seeded ragged
     synthetic batches -> mel / postnet / pitch / energy / log-duration outputs and the 6 losses;
  3. one MAML task step (K inner steps, first- and second-order) driven through the reference
     modules with learn2learn's published clone_module / maml_update algorithm restated below
     (l2l is not installable here) -> query losses, outer-gradient norms and leading elements,
     and the adapted (fast) weights' norms;
  4. LengthRegulator index path on adversarial durations (zeros, negatives, floats).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import fs2_oracle as O  # noqa: E402
from oracle import refstub  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def build_reference_model(seed=0, model_config=None, n_speaker=16):
    FastSpeech2, FastSpeech2Loss, _ = refstub.load_reference()
    refstub.neutralise_dropout()
    cfg = model_config or O.BASE_MODEL_CONFIG
    pre_dir = refstub.make_preprocessed_dir(O.DEFAULT_STATS, n_speaker)
    pre = refstub.preprocess_config(pre_dir)
    torch.manual_seed(seed)
    model = FastSpeech2(pre, cfg, refstub.ALGORITHM_CONFIG)
    loss = FastSpeech2Loss(pre, cfg)
    model.train()
    return model, loss


# ---- learn2learn restated on nn.Modules (l2l/utils clone_module, update_module; algorithms/maml.py) ----
def l2l_clone_module(module, memo=None):
    """learn2learn.utils.clone_module: new module object whose _parameters are `p.clone()` (graph-
    connected to the originals); buffers cloned only if they require grad; recursion over children.
    The reference's own annotated copy of this traversal: lightning/systems/utils.py:192-293."""
    if memo is None:
        memo = {}
    clone = module.__new__(type(module))
    clone.__dict__ = module.__dict__.copy()
    clone._parameters = clone._parameters.copy()
    clone._buffers = clone._buffers.copy()
    clone._modules = clone._modules.copy()
    for k, p in module._parameters.items():
        if p is not None:
            ptr = p.data_ptr()
            if ptr in memo:
                clone._parameters[k] = memo[ptr]
            else:
                c = p.clone()
                clone._parameters[k] = c
                memo[ptr] = c
    for k, b in module._buffers.items():
        if b is not None and b.requires_grad:
            clone._buffers[k] = b.clone()
    for k, m in module._modules.items():
        clone._modules[k] = l2l_clone_module(m, memo)
    return clone


def l2l_adapt(module, loss, lr, first_order):
    """learn2learn.algorithms.MAML.adapt(allow_nograd=True) + maml_update + update_module."""
    second_order = not first_order
    params = [p for p in module.parameters() if p.requires_grad]
    grads = torch.autograd.grad(loss, params, retain_graph=second_order, create_graph=second_order, allow_unused=False)
    gmap = {id(p): g for p, g in zip(params, grads)}

    def update(m, memo):
        for k, p in m._parameters.items():
            if p is not None and id(p) in gmap:
                if id(p) in memo:
                    m._parameters[k] = memo[id(p)]
                else:
                    new = p + (-lr * gmap[id(p)])
                    memo[id(p)] = new
                    m._parameters[k] = new
        for c in m._modules.values():
            update(c, memo)

    update(module, {})


def reference_forward_learner(model, learner_modules, speaker_args, texts, src_lens, max_src_len, mels=None,
                              mel_lens=None, max_mel_len=None, p_targets=None, e_targets=None, d_targets=None,
                              average_spk_emb=False):
    """lightning/systems/base_adaptor.py:41-95 verbatim control flow over the reference's modules."""
    from utils.tools import get_mask_from_lengths  # reference

    get = lambda name: learner_modules[name] if name in learner_modules else getattr(model, name, None)  # noqa: E731
    encoder, va, dec = get("encoder"), get("variance_adaptor"), get("decoder")
    mel_linear, postnet, speaker_emb = get("mel_linear"), get("postnet"), get("speaker_emb")
    src_masks = get_mask_from_lengths(src_lens, max_src_len)
    output = encoder(texts, src_masks)
    mel_masks = get_mask_from_lengths(mel_lens, max_mel_len) if mel_lens is not None else None
    spk_emb = speaker_emb(speaker_args)
    if average_spk_emb:
        spk_emb = spk_emb.mean(dim=0, keepdim=True).expand(output.shape[0], -1)
    output += spk_emb.unsqueeze(1).expand(-1, max_src_len, -1)
    (output, p_pred, e_pred, log_d, d_rounded, mel_lens, mel_masks) = va(
        output, src_masks, mel_masks, max_mel_len, p_targets, e_targets, d_targets, 1.0, 1.0, 1.0)
    output += spk_emb.unsqueeze(1).expand(-1, max(mel_lens), -1)
    output, mel_masks = dec(output, mel_masks)
    output = mel_linear(output)
    postnet_output = postnet(output) + output
    return (output, postnet_output, p_pred, e_pred, log_d, d_rounded, src_masks, mel_masks, src_lens, mel_lens)


def reference_task_step(model, loss_fn, sup, qry, K, lr, first_order):
    mods = torch.nn.ModuleDict({k: getattr(model, k) for k in O.ADAPT_MODULES})
    learner = l2l_clone_module(mods)
    learner.train()
    lm = {k: learner[k] for k in O.ADAPT_MODULES}
    for _ in range(K):
        preds = reference_forward_learner(model, lm, *sup[2:])
        loss = loss_fn(sup, preds)[0]
        l2l_adapt(learner, loss, lr, first_order)
    preds = reference_forward_learner(model, lm, sup[2], *qry[3:], average_spk_emb=True)
    valid = loss_fn(qry, preds)
    names = [n for n, p in model.named_parameters() if p.requires_grad]
    params = [p for _, p in model.named_parameters() if p.requires_grad]
    grads = torch.autograd.grad(valid[0], params, allow_unused=True)
    gd = {n: (g if g is not None else torch.zeros_like(p)) for n, p, g in zip(names, params, grads)}
    fast = {f"{mk}.{n}": p.detach() for mk in O.ADAPT_MODULES for n, p in lm[mk].named_parameters() if p.requires_grad}
    return tuple(v.detach() for v in valid), preds, gd, fast


def tensor_fingerprint(t: torch.Tensor):
    t = t.detach().double().flatten()
    return np.array([t.sum().item(), t.abs().sum().item(), (t * t).sum().item()])


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    out = {}
    # ---------------- 1. init parity ----------------
    model, loss_fn = build_reference_model(seed=0)
    sd = model.state_dict()
    keys = list(sd.keys())
    out["sd_keys"] = np.array(keys)
    out["sd_fp"] = np.stack([tensor_fingerprint(sd[k].float()) for k in keys])
    P = O.init_params(seed=0)
    assert sorted(P.keys()) == sorted(keys), "oracle.init_params keys differ from the reference state_dict"
    for k in keys:
        assert torch.equal(P[k].detach(), sd[k]), f"init mismatch at {k}"
    print(f"[golden] init parity: {len(keys)} tensors bit-identical")

    # ---------------- 2. forward + loss ----------------
    cfg = O.BASE_MODEL_CONFIG
    for tag, (n, L, T, ragged) in {"fwd_small": (3, 12, 40, True), "fwd_mid": (2, 32, 150, True)}.items():
        batch = O.synth_batch(n, L, T, seed=7, speaker=3, ragged=ragged)
        with torch.no_grad():
            preds = model(*batch[2:])
            losses = loss_fn(batch, preds)
        out[f"{tag}_cfg"] = np.array([n, L, T, int(ragged), 7, 3])
        for name, t in zip(["mel", "postnet", "pitch", "energy", "logd"], preds[:5]):
            out[f"{tag}_{name}"] = t.numpy()
        out[f"{tag}_losses"] = np.array([v.item() for v in losses])
        print(f"[golden] {tag}: losses {out[f'{tag}_losses']}")

    # ---------------- 3. MAML task steps ----------------
    for tag, (K, fo, S, Q, L, T) in {"maml_so_k2": (2, False, 2, 2, 10, 36), "maml_fo_k2": (2, True, 2, 2, 10, 36),
                                     "maml_so_k1": (1, False, 3, 2, 16, 60)}.items():
        model, loss_fn = build_reference_model(seed=0)       # fresh BN running stats
        sup, qry = O.synth_task(task=5, shots=S, queries=Q, L=L, T=T, ragged=True)
        losses, preds, gd, fast = reference_task_step(model, loss_fn, sup, qry, K, 0.001, fo)
        names = sorted(gd.keys())
        out[f"{tag}_cfg"] = np.array([K, int(fo), S, Q, L, T, 5])
        out[f"{tag}_losses"] = np.array([v.item() for v in losses])
        out[f"{tag}_grad_names"] = np.array(names)
        out[f"{tag}_grad_norm"] = np.array([gd[k].double().norm().item() for k in names])
        out[f"{tag}_grad_head"] = np.stack([np.pad(gd[k].flatten()[:8].numpy(), (0, max(0, 8 - gd[k].numel()))) for k in names])
        fnames = sorted(fast.keys())
        out[f"{tag}_fast_names"] = np.array(fnames)
        out[f"{tag}_fast_norm"] = np.array([fast[k].double().norm().item() for k in fnames])
        out[f"{tag}_mel"] = preds[0].detach().numpy()
        out[f"{tag}_bn_running_mean0"] = model.postnet.convolutions[0][1].running_mean.numpy().copy()
        print(f"[golden] {tag}: query losses {out[f'{tag}_losses']}, |g| total "
              f"{np.sqrt((out[f'{tag}_grad_norm'] ** 2).sum()):.6f}")

    # ---------------- 4. LengthRegulator ----------------
    from lightning.model.modules import LengthRegulator  # reference
    import lightning.model.modules as ref_modules
    ref_modules.device = torch.device("cpu")
    lr = LengthRegulator()
    g = torch.Generator().manual_seed(11)
    x = torch.randn(3, 9, 4, generator=g)
    d = torch.tensor([[2, 0, 3, 1, 0, 0, 4, 1, 2], [0, 0, 0, 5, -3, 2, 1, 1, 1], [1, 1, 1, 1, 1, 1, 1, 1, 1]])
    o_i, l_i = lr(x, d, 16)
    o_f, l_f = lr(x, d.float() + 0.7, 16)
    out["lr_x"], out["lr_d"] = x.numpy(), d.numpy()
    out["lr_out_int"], out["lr_len_int"] = o_i.numpy(), l_i.numpy()
    out["lr_out_float"], out["lr_len_float"] = o_f.numpy(), l_f.numpy()

    path = os.path.join(GOLDEN_DIR, "fs2_golden.npz")
    np.savez_compressed(path, **out)
    print("[golden] wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
