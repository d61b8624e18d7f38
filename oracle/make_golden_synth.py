"""Generate tests/golden/synth_golden.npz from the REAL reference modules (container only; /root/reference): pins the oracle's
free-running synthesis (targets None: predicted pitch / energy / durations, modules.py:85-99,132-139), its eval-mode branches
(BatchNorm running statistics; Decoder beyond max_seq_len, Models.py:148-156), the train-mode truncation (Models.py:161-166),
and the few-shot adaptation protocol of `BaseAdaptorSystem._test_step` (base_adaptor.py:160-189) driven through the reference's
modules with learn2learn's clone_module / first-order maml_update restated (oracle/make_golden.py).

    python -m oracle.make_golden_synth
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import fs2_oracle as O  # noqa: E402
from oracle import make_golden as MG  # noqa: E402

BIAS = 1.3          # duration-predictor output bias: random init predicts zero-length utterances


def tweak(model):
    with torch.no_grad():
        model.variance_adaptor.duration_predictor.linear_layer.bias += BIAS


def ref_forward(model, lm, spk, batch, free, d_control=1.0):
    """forward_learner (base_adaptor.py:41-95) over the reference modules; `free` drops every target (qry_batch[3:6])."""
    from utils.tools import get_mask_from_lengths  # reference

    get = lambda name: lm[name] if name in lm else getattr(model, name, None)  # noqa: E731
    texts, src_lens, max_src_len = batch[3], batch[4], batch[5]
    mels, mel_lens, max_mel_len, p_t, e_t, d_t = (None,) * 6 if free else batch[6:12]
    src_masks = get_mask_from_lengths(src_lens, max_src_len)
    output = get("encoder")(texts, src_masks)
    mel_masks = get_mask_from_lengths(mel_lens, max_mel_len) if mel_lens is not None else None
    spk_emb = get("speaker_emb")(spk).mean(dim=0, keepdim=True).expand(output.shape[0], -1)
    output += spk_emb.unsqueeze(1).expand(-1, max_src_len, -1)
    (output, p_pred, e_pred, log_d, d_rounded, mel_lens, mel_masks) = get("variance_adaptor")(
        output, src_masks, mel_masks, max_mel_len, p_t, e_t, d_t, 1.0, 1.0, d_control)
    output += spk_emb.unsqueeze(1).expand(-1, max(mel_lens), -1)
    output, mel_masks = get("decoder")(output, mel_masks)
    output = get("mel_linear")(output)
    postnet_output = get("postnet")(output) + output
    return (output, postnet_output, p_pred, e_pred, log_d, d_rounded, src_masks, mel_masks, src_lens, mel_lens)


FRAME_STRIDE = 8      # mel / postnet are stored every 8th frame (keeps the fixture small; all phoneme-level outputs in full)


def pack(out, tag, preds):
    for name, t in zip(["mel", "postnet", "pitch", "energy", "logd", "d_rounded"], preds[:6]):
        t = t.detach().float()
        out[f"{tag}_{name}"] = (t[:, ::FRAME_STRIDE] if name in ("mel", "postnet") else t).numpy()
    out[f"{tag}_shape"] = np.array(preds[1].shape)
    out[f"{tag}_mel_len"] = preds[9].numpy()
    out[f"{tag}_mel_mask"] = preds[7].numpy()


def main():
    from oracle import refstub
    refstub.load_reference()                                   # puts /root/reference on sys.path behind the import stubs
    import lightning.model.modules as ref_modules
    import utils.tools as ref_tools
    out = {"bias": np.array(BIAS)}
    cfg = O.BASE_MODEL_CONFIG

    def oracle_params():
        P = O.init_params(seed=0)
        P["variance_adaptor.duration_predictor.linear_layer.bias"] = P["variance_adaptor.duration_predictor.linear_layer.bias"] + BIAS
        return P

    # ---------------- 1. free-running forwards: eval, train, beyond max_seq_len ----------------
    sup, qry = O.synth_task(task=6, shots=3, queries=2, L=30, T=110, ragged=True)
    for tag, train, dc in (("free_eval", False, 1.0), ("free_train", True, 1.0), ("free_eval_long", False, 12.0), ("free_train_long", True, 12.0)):
        model, _ = MG.build_reference_model(seed=0)
        ref_modules.device = ref_tools.device = torch.device("cpu")
        tweak(model)
        model.train(train)
        with torch.no_grad():
            preds = ref_forward(model, {}, sup[2], qry, free=True, d_control=dc)
        pack(out, tag, preds)
        P = oracle_params()
        with torch.no_grad():
            mine = O.fs2_forward(P, cfg, sup[2], *qry[3:6], d_control=dc, average_spk_emb=True, training=train)
        assert torch.equal(mine[5], preds[5]) and torch.equal(mine[9], preds[9]), tag
        for i in range(5):
            assert mine[i].shape == preds[i].shape and torch.allclose(mine[i], preds[i], rtol=1e-5, atol=1e-6), (tag, i)
        print(f"[golden] {tag}: T = {preds[1].shape[1]}, mel_len {preds[9].tolist()}")
    out["task_cfg"] = np.array([6, 3, 2, 30, 110])

    # ---------------- 2. _test_step: eval step_0, 2 rounds x 2 first-order steps, recon + synth ----------------
    model, loss_fn = MG.build_reference_model(seed=0)
    tweak(model)
    steps, total = 2, 4
    res = {}
    model.eval()                                               # trainer.test puts the LightningModule in eval mode
    mods = torch.nn.ModuleDict({k: getattr(model, k) for k in O.ADAPT_MODULES})
    with torch.no_grad():
        p = ref_forward(model, {}, sup[2], qry, free=False)
        res["step_0"] = {"recon": (loss_fn(qry, p), p), "synth": ref_forward(model, {}, sup[2], qry, free=True)}
    learner = None
    for ft in range(steps, total + 1, steps):
        if learner is None:
            learner = MG.l2l_clone_module(mods)
            learner.train()                                    # base_adaptor.py:101-103
        lm = {k: learner[k] for k in O.ADAPT_MODULES}
        for _ in range(steps):
            preds = MG.reference_forward_learner(model, lm, *sup[2:])
            MG.l2l_adapt(learner, loss_fn(sup, preds)[0], 0.001, first_order=True)
        with torch.no_grad():
            p = ref_forward(model, lm, sup[2], qry, free=False)
            res[f"step_{ft}"] = {"recon": (loss_fn(qry, p), p), "synth": ref_forward(model, lm, sup[2], qry, free=True)}
    for k, v in res.items():
        out[f"tta_{k}_losses"] = np.array([x.item() for x in v["recon"][0]])
        pack(out, f"tta_{k}_recon", v["recon"][1])
        pack(out, f"tta_{k}_synth", v["synth"])
    out["tta_running_mean0"] = model.postnet.convolutions[0][1].running_mean.numpy().copy()
    ref, _ = O.test_time_adaptation(oracle_params(), cfg, sup, qry, steps, total, saving_steps=(2, 4))
    for k in res:
        a, b = ref[k]["recon"]["output"], res[k]["recon"][1]
        assert torch.allclose(a[1], b[1], rtol=1e-4, atol=1e-5), k
        assert torch.equal(ref[k]["synth"]["output"][5], res[k]["synth"][5]), k
        assert torch.allclose(torch.stack(list(ref[k]["recon"]["losses"])), torch.stack([x.detach() for x in res[k]["recon"][0]]), rtol=1e-5), k
    print("[golden] test-time adaptation: query losses", {k: round(float(v["recon"][0][0]), 4) for k, v in res.items()})
    path = os.path.join(ROOT, "tests", "golden", "synth_golden.npz")
    np.savez_compressed(path, **out)
    print("[golden] wrote", path, os.path.getsize(path) // 1024, "KiB; the oracle reproduces the real modules")


if __name__ == "__main__":
    main()
