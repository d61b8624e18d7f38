"""The oracle's free-running synthesis, eval-mode branches and few-shot adaptation protocol against goldens produced by the REAL
reference modules (`oracle/make_golden_synth.py`: transformer/*, lightning/model/* with learn2learn's first-order update restated).
Re-checked here without the reference (which does not exist on the GPU box)."""
import os

import numpy as np
import torch

from oracle import fs2_oracle as O

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "synth_golden.npz"), allow_pickle=False)
STRIDE = 8


def params():
    P = O.init_params(seed=0)
    P["variance_adaptor.duration_predictor.linear_layer.bias"] = P["variance_adaptor.duration_predictor.linear_layer.bias"] + float(G["bias"])
    return P


def task():
    t, S, Q, L, T = [int(v) for v in G["task_cfg"]]
    return O.synth_task(task=t, shots=S, queries=Q, L=L, T=T, ragged=True)


def check(tag, preds, tol=1e-5):
    assert tuple(preds[1].shape) == tuple(int(v) for v in G[f"{tag}_shape"]), tag
    assert np.array_equal(preds[5].float().numpy(), G[f"{tag}_d_rounded"]), tag            # integer path: exact
    assert np.array_equal(preds[9].numpy(), G[f"{tag}_mel_len"]) and np.array_equal(preds[7].numpy(), G[f"{tag}_mel_mask"]), tag
    for i, name in enumerate(["mel", "postnet", "pitch", "energy", "logd"]):
        t = preds[i].detach().float()
        t = t[:, ::STRIDE] if name in ("mel", "postnet") else t
        ref = torch.from_numpy(G[f"{tag}_{name}"])
        assert t.shape == ref.shape and ((t - ref).norm() / ref.norm().clamp_min(1e-30)).item() < tol, (tag, name)


def test_free_running_forward_matches_the_real_modules():
    sup, qry = task()
    for tag, train, dc in (("free_eval", False, 1.0), ("free_train", True, 1.0), ("free_eval_long", False, 12.0), ("free_train_long", True, 12.0)):
        with torch.no_grad():
            preds = O.fs2_forward(params(), O.BASE_MODEL_CONFIG, sup[2], *qry[3:6], d_control=dc, average_spk_emb=True, training=train)
        check(tag, preds)
    assert int(G["free_eval_long_shape"][1]) > 1000 and int(G["free_train_long_shape"][1]) == 1000      # eval keeps, train truncates


def test_test_time_adaptation_matches_the_real_modules():
    sup, qry = task()
    P = params()
    ref, _ = O.test_time_adaptation(P, O.BASE_MODEL_CONFIG, sup, qry, 2, 4, saving_steps=(2, 4))
    for k in ("step_0", "step_2", "step_4"):
        check(f"tta_{k}_recon", ref[k]["recon"]["output"], tol=2e-5)
        check(f"tta_{k}_synth", ref[k]["synth"]["output"], tol=2e-5)
        losses = torch.stack(list(ref[k]["recon"]["losses"])).numpy()
        assert np.allclose(losses, G[f"tta_{k}_losses"], rtol=1e-5), k
    assert np.allclose(P["postnet.convolutions.0.1.running_mean"].numpy(), G["tta_running_mean0"], rtol=1e-4, atol=1e-7)
