"""The public drop-in API on the GPU: MetaSystem.training_step through the captured CUDA graph (side-stream
branch, PDL launches, static staged inputs) must give the same query losses / outer gradient as the oracle,
for the batch that was used for capture AND for a different batch replayed through the same graph."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu

from meta_tts_b200.systems import DEFAULT_ALGORITHM_CONFIG, DEFAULT_TRAIN_CONFIG, MetaSystem  # noqa: E402
from oracle import fs2_oracle as O  # noqa: E402


def _rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("dropout", [True, False])
@pytest.mark.parametrize("second_order", [True, False])
def test_training_step_graph_replay_matches_oracle(cuda_device, second_order, dropout):
    """dropout=True (the default, = the reference's learner.train()): every replay of the captured graph draws NEW masks
    (the salt rides in the staged H2D copy) and still matches the oracle evaluated with that step's salt."""
    cfg = O.small_model_config(2, 2)
    algo = copy.deepcopy(DEFAULT_ALGORITHM_CONFIG)
    algo["adapt"]["train"]["steps"] = 2
    algo["adapt"]["test"]["steps"] = 2
    sysm = MetaSystem(None, cfg, DEFAULT_TRAIN_CONFIG, algo, n_speaker=16, device="cuda:0", split=3, second_order=second_order,
                      dropout=dropout, seed=1234)
    P = O.init_params(seed=0, model_config=cfg)
    sysm.load_state_dict({k: v.detach().clone() for k, v in P.items()})
    Pc = {k: v.detach().clone() for k, v in P.items()}          # oracle copy (BatchNorm running stats evolve in both)
    for step in range(3):                                        # step 0 captures the graph, steps 1-2 replay it
        sup, qry = O.synth_task(task=10 + step, shots=3, queries=2, L=12, T=40)
        out = sysm.training_step([([sup], [qry])], step)
        losses, preds, grads = O.maml_task_step(Pc, cfg, sup, qry, 2, 0.001, first_order=not second_order,
                                                drop_seed=(0, sysm.last_salt) if dropout else None)
        got_l = torch.stack([out["losses"][i] for i in range(6)])
        # 1e-4 without dropout.  With dropout a rectifier unit of the SUPPORT passes can sit within 1e-6 of zero: it gates differently in
        # two fp32 implementations, the inner gradient moves one weight row and the query's duration loss jumps by 6.6e-4 (1.02e-4 of
        # the 6-vector) — task 11 with step 1's salt does exactly that on the CPU alone when the oracle's weights are perturbed by 1e-6
        # (tools/kink_probe.py, profiles/r02_kink_probe.txt), and here when LayerNorm became the out-projection's epilogue.
        assert _rel(got_l, torch.stack(losses)) < (3e-4 if dropout else 1e-4), f"step {step}"
        assert _rel(out["output"][0], preds[0]) < 1e-3 and _rel(out["output"][1], preds[1]) < 1e-3
        assert torch.equal(out["output"][9].cpu(), preds[9])
        got = sysm.maml.task_grads()
        tot = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values()))
        err = torch.sqrt(sum(((got[k].double() - grads[k].double()) ** 2).sum() for k in grads)) / tot
        per = sorted((((got[k].double() - grads[k].double()).norm() / grads[k].double().norm()).item(), k) for k in grads
                     if grads[k].double().norm() > 1e-4 * tot)
        med = per[len(per) // 2][0]
        print(f"[systems] so={second_order} dropout={dropout} step {step}: grad err total {err.item():.2e}, per-tensor median "
              f"{med:.2e}, worst {per[-1][0]:.2e} {per[-1][1]}")
        # total: a single ReLU unit whose pre-activation is ~1e-6 from 0 gates differently in two fp32 implementations and
        # moves one weight row by O(1) (seen CPU-vs-CPU as well; more frequent with dropout's 2x / 1.25x rescaling), so
        # the total is held to 1e-2 with dropout while the MEDIAN tensor (robust to one flipped unit) is held to 1e-3
        assert err.item() < (1e-2 if dropout else 2e-3), f"step {step}: outer gradient rel err {err.item():.2e}"
        assert med < 1e-3, f"step {step}: median per-tensor gradient rel err {med:.2e}"
        sysm.be.zero_(sysm.maml.g_outer)
    # BatchNorm running statistics advanced identically (3 steps x (2 support + 1 query) forwards)
    sd = sysm.state_dict()
    assert _rel(sd["postnet.convolutions.1.1.running_var"], Pc["postnet.convolutions.1.1.running_var"]) < 1e-4
    assert int(sd["postnet.convolutions.0.1.num_batches_tracked"]) == 9
