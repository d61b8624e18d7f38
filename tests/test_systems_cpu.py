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
