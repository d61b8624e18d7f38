"""Host-logic test of the drop-in MetaSystem API (training_step / optimizer_step) with the CPU
restatement of the op set injected (no CUDA graph): same nested batch layout, same return dict, and the
accumulated + clipped + Adam-updated weights equal a hand-rolled torch computation from oracle gradients."""
import copy

import torch

from meta_tts_b200.systems import DEFAULT_ALGORITHM_CONFIG, DEFAULT_TRAIN_CONFIG, MetaSystem
from oracle import fs2_oracle as O
from oracle.ops_reference import RefOps

CFG = O.small_model_config(1, 1)


def test_training_step_and_optimizer_step_match_torch_adam():
    algo = copy.deepcopy(DEFAULT_ALGORITHM_CONFIG)
    algo["adapt"]["train"]["steps"] = 1
    algo["adapt"]["test"]["steps"] = 1
    train = copy.deepcopy(DEFAULT_TRAIN_CONFIG)
    train["optimizer"]["grad_acc_step"] = 2
    sysm = MetaSystem(None, CFG, train, algo, n_speaker=16, device="cpu", use_cuda_graph=False, backend=RefOps(split=3))
    P = O.init_params(seed=0, model_config=CFG)
    sysm.load_state_dict({("model." + k): v.detach().clone() for k, v in P.items()})     # Lightning prefix accepted
    names = O.trainable_names(P)
    acc = {k: torch.zeros_like(P[k]) for k in names}
    for t in range(2):                                     # two micro-steps (grad_acc_step = 2), then one optimizer step
        sup, qry = O.synth_task(task=t + 2, shots=2, queries=2, L=5, T=12, ragged=True)   # (task 1 sits on a ReLU kink with dropout on)
        out = sysm.training_step([([sup], [qry])], t)
        assert set(out) == {"loss", "losses", "output", "_batch"} and len(out["losses"]) == 6 and len(out["output"]) == 10
        Pc = {k: v.detach().clone() for k, v in P.items()}
        losses, preds, grads = O.maml_task_step(Pc, CFG, sup, qry, 1, 0.001, False, drop_seed=(0, sysm.last_salt))   # dropout ON
        assert abs(float(out["loss"]) - float(losses[0])) < 1e-4 * abs(float(losses[0]))
        assert torch.equal(out["output"][6], preds[6]) and torch.equal(out["output"][7], preds[7])   # masks
        for k in names:
            acc[k] += grads[k] / 2
    sysm.optimizer_step()
    # torch reference: clip_grad_norm_(1.0) + Adam(lr = 256^-0.5 * LambdaLR(step 0))
    params = [P[k].detach().clone().requires_grad_(True) for k in names]
    for p, k in zip(params, names):
        p.grad = acc[k].clone()
    torch.nn.utils.clip_grad_norm_(params, 1.0)
    lr = 256 ** -0.5 * min(1.0, 4000 ** -1.5 * 1)
    opt = torch.optim.Adam(params, lr=lr, betas=(0.9, 0.98), eps=1e-9, weight_decay=0.0)
    opt.step()
    new = sysm.state_dict()
    gtot = torch.sqrt(sum((acc[k].double() ** 2).sum() for k in names))
    for p, k in zip(params, names):
        if acc[k].norm() < 1e-5 * gtot:
            continue      # analytically-zero gradients (e.g. key bias): Adam (eps 1e-9) turns fp noise into +-lr steps
        sig = acc[k].abs() > 1e-3 * acc[k].abs().max()        # first Adam step = lr*sign(g): only compare where g is not noise
        delta_ref = (p.detach() - P[k].detach())[sig]
        delta = (new[k] - P[k].detach())[sig]
        assert (delta - delta_ref).abs().max() <= 0.02 * delta_ref.abs().max() + 6e-8, k   # lr(step 0) = 2.5e-7: fp32 ulp noise
    assert float(sysm.maml.g_outer.abs().max()) == 0.0     # accumulation buffer cleared


def _sys(algo_steps=1, max_seq_len=None):
    algo = copy.deepcopy(DEFAULT_ALGORITHM_CONFIG)
    algo["adapt"]["train"]["steps"] = algo_steps
    algo["adapt"]["test"]["steps"] = algo_steps
    cfg = copy.deepcopy(CFG)
    if max_seq_len is not None:
        cfg["max_seq_len"] = max_seq_len
    s = MetaSystem(None, cfg, DEFAULT_TRAIN_CONFIG, algo, n_speaker=16, device="cpu", use_cuda_graph=False, backend=RefOps(split=3),
                   dropout=False)
    P = O.init_params(seed=0, model_config=cfg)
    s.load_state_dict(P)
    return s, P, cfg


def test_graph_cache_key_separates_train_and_validation_and_is_bounded():
    """accumulate_scale is baked into the captured graph, so train (scale = 1/(acc*world)) and validation
    (scale None) steps of the same shape must not share a cache entry; (medium) the cache is an LRU of bounded size."""
    from meta_tts_b200 import systems as S

    s, _, _ = _sys()
    sup, qry = O.synth_task(task=2, shots=2, queries=2, L=5, T=12, ragged=True)
    k_train = S._task_key(s, sup, qry, 1, True, 1.0)
    k_val = S._task_key(s, sup, qry, 1, True, None)
    assert k_train != k_val and k_train[:-2] == k_val[:-2] and k_train[-1] == k_val[-1]      # (scale, reduce_now) are the last two fields
    s.graph_cache_size = 2
    for T in (10, 11, 12, 13):
        sup, qry = O.synth_task(task=2, shots=2, queries=2, L=5, T=T)
        S._get_task(s, S._task_key(s, sup, qry, 1, False, 1.0))
    assert len(s._graphs) == 2 and [k[2] for k in s._graphs] == [12, 13]
    # validation after training on the same shape leaves the accumulated outer gradient untouched
    s, _, _ = _sys()
    sup, qry = O.synth_task(task=2, shots=2, queries=2, L=5, T=12, ragged=True)
    s.training_step([([sup], [qry])], 0)
    g = s.maml.g_outer.clone()
    assert float(g.abs().max()) > 0
    s.validation_step([([sup], [qry])], 0)
    assert torch.equal(s.maml.g_outer, g)


def test_train_mode_truncates_beyond_max_seq_len():
    """utterances longer than max_seq_len train through (Models.py:161-166 keeps the first max_seq_len
    frames, loss.py:42-43 crops the targets) instead of raising; losses and the outer gradient equal the oracle's, which
    follows the reference's truncation."""
    s, P, cfg = _sys(max_seq_len=10)
    sup, qry = O.synth_task(task=3, shots=2, queries=2, L=5, T=14)          # 14 frames > max_seq_len = 10
    out = s.training_step([([sup], [qry])], 0)
    assert out["output"][0].shape[1] == 10 and out["output"][7].shape[1] == 10
    Pc = {k: v.detach().clone() for k, v in P.items()}
    losses, preds, grads = O.maml_task_step(Pc, cfg, sup, qry, 1, 0.001, False)
    assert preds[0].shape[1] == 10
    for i in range(6):
        assert abs(float(out["losses"][i]) - float(losses[i])) <= 2e-5 * abs(float(losses[i])) + 1e-7, i
    got = s.maml.task_grads()
    num = sum(((got[k] - grads[k]).double() ** 2).sum() for k in grads)
    den = sum((grads[k].double() ** 2).sum() for k in grads)
    assert float(num / den) ** 0.5 < 1e-3
    # stand-alone adapt() on two different shapes in a row (tapes are keyed by shape)
    s.adapt([([sup], [qry])], 1)
    sup2, qry2 = O.synth_task(task=4, shots=2, queries=2, L=6, T=9)
    s.adapt([([sup2], [qry2])], 1)
