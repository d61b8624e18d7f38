"""SURVEY §8 rows f3/f5 host logic on the CPU: `MetaSystem.test_step` (few-shot adaptation inference, BASELINE configs[4]
structure: rolling first-order adaptation, eval-mode step_0, train-mode recon / free-running synth afterwards) driven through
the CPU restatement of the op set, against `oracle.fs2_oracle.test_time_adaptation` (the restatement of
lightning/systems/base_adaptor.py:139-189).  Integer paths (rounded durations, mel lengths, masks) must match exactly."""
import copy

import pytest
import torch

from meta_tts_b200.systems import DEFAULT_ALGORITHM_CONFIG, DEFAULT_TRAIN_CONFIG, MetaSystem
from oracle import fs2_oracle as O
from oracle.ops_reference import RefOps

from tests_helpers_adapt import check_outputs, rel, talkative_params  # noqa: E402

CFG = O.small_model_config(1, 1)


@pytest.mark.parametrize("dropout", [False, True])
def test_test_step_matches_oracle(dropout):
    algo = copy.deepcopy(DEFAULT_ALGORITHM_CONFIG)
    algo["adapt"]["train"]["steps"] = 2
    algo["adapt"]["test"] = {"steps": 6, "saving_steps": [2, 6]}
    sysm = MetaSystem(None, CFG, DEFAULT_TRAIN_CONFIG, algo, n_speaker=16, device="cpu", use_cuda_graph=False,
                      backend=RefOps(split=3), dropout=dropout, seed=3)
    P = talkative_params(CFG)
    sysm.load_state_dict({k: v.detach().clone() for k, v in P.items()})
    sup, qry = O.synth_task(task=2, shots=3, queries=1, L=7, T=20, ragged=True)
    outs = sysm.test_step([([sup], [qry])], 0)
    assert len(outs) == 1
    Pc = {k: v.detach().clone() for k, v in P.items()}
    ref, theta = O.test_time_adaptation(Pc, CFG, sup, qry, 2, 6, saving_steps=(2, 6), drop_seed=(0, sysm.last_salt) if dropout else None)
    assert ref["step_0"]["synth"]["output"][0].shape[1] > 7, "free-running output should be longer than the phoneme count"
    check_outputs(outs[0], ref)
    fw = sysm.maml.fast_weights(1)
    # zero-initialised parameters (LN / BN biases) are pure accumulated gradients: fp32 summation-order noise (see
    # tests/test_engine_gpu.py) — everything else agrees to ~1e-6
    assert max(rel(fw[k], theta[k]) for k in theta) < 1e-2
    assert sorted(rel(fw[k], theta[k]) for k in theta)[len(theta) // 2] < 1e-5
    # BatchNorm running statistics advanced exactly as in the reference (train-mode forwards only)
    for i in range(5):
        k = f"postnet.convolutions.{i}.1.running_mean"
        assert rel(sysm.maml.consts[k], Pc[k]) < 1e-4
    assert sysm.maml.bn_batches == int(Pc["postnet.convolutions.0.1.num_batches_tracked"])


def test_one_shot_protocol_and_controls():
    """`1-shot: True` (base_adaptor.py:144-151): one evaluation per support utterance; p/e/d controls scale the predictions."""
    algo = copy.deepcopy(DEFAULT_ALGORITHM_CONFIG)
    algo["adapt"]["train"]["steps"] = 1
    algo["adapt"]["test"] = {"steps": 1, "saving_steps": [1], "1-shot": True}
    sysm = MetaSystem(None, CFG, DEFAULT_TRAIN_CONFIG, algo, n_speaker=16, device="cpu", use_cuda_graph=False,
                      backend=RefOps(split=3), dropout=False)
    P = talkative_params(CFG)
    sysm.load_state_dict({k: v.detach().clone() for k, v in P.items()})
    sup, qry = O.synth_task(task=4, shots=2, queries=1, L=6, T=15, ragged=True)
    outs = sysm.test_step([([sup], [qry])], 0)
    assert len(outs) == 2 and all("step_1" in o for o in outs)
    # controls
    from meta_tts_b200.maml import batch_from_tuple
    bt = batch_from_tuple(qry, "cpu", spk_ids=sup[2], average_spk=True, targets=False)
    out = sysm.maml.predict(bt, adapted=False, free_running=True, eval_mode=True, p_control=1.2, e_control=0.8, d_control=1.5)
    with torch.no_grad():
        # (state_dict(): the test steps above advanced the BatchNorm running statistics that eval mode reads)
        ref = O.fs2_forward({k: v.detach().clone() for k, v in sysm.state_dict().items()}, CFG, sup[2], *qry[3:6], p_control=1.2, e_control=0.8,
                            d_control=1.5, average_spk_emb=True, training=False)
    assert torch.equal(out["d_rounded"], ref[5]) and torch.equal(out["mel_len"], ref[9])
    assert rel(out["pitch"], ref[2]) < 1e-4 and rel(out["energy"], ref[3]) < 1e-4 and rel(out["postnet"], ref[1]) < 1e-3


def test_forward_learner_signature_and_modes():
    """base_adaptor.py:41-95 as a public call: meta parameters vs the learner returned by adapt(), teacher forced vs free running,
    eval vs train mode of the un-adapted learner."""
    algo = copy.deepcopy(DEFAULT_ALGORITHM_CONFIG)
    algo["adapt"]["train"]["steps"] = 2
    algo["adapt"]["test"] = {"steps": 2}
    sysm = MetaSystem(None, CFG, DEFAULT_TRAIN_CONFIG, algo, n_speaker=16, device="cpu", use_cuda_graph=False, backend=RefOps(split=3), dropout=False)
    P = talkative_params(CFG)
    sysm.load_state_dict({k: v.detach().clone() for k, v in P.items()})
    sup, qry = O.synth_task(task=2, shots=3, queries=1, L=7, T=20, ragged=True)
    # un-adapted, eval mode, teacher forced and free running
    sysm.eval()
    with torch.no_grad():
        ref_t = O.fs2_forward({k: v.detach().clone() for k, v in P.items()}, CFG, sup[2], *qry[3:], average_spk_emb=True, training=False)
        ref_f = O.fs2_forward({k: v.detach().clone() for k, v in P.items()}, CFG, sup[2], *qry[3:6], d_control=1.3, average_spk_emb=True, training=False)
    out_t = sysm.forward_learner(sysm.learner, sup[2], *qry[3:], average_spk_emb=True)
    out_f = sysm.forward_learner(None, sup[2], *qry[3:6], d_control=1.3, average_spk_emb=True)
    for out, ref in ((out_t, ref_t), (out_f, ref_f)):
        assert len(out) == 10 and all(rel(out[i], ref[i]) < 1e-4 for i in range(5))
        assert torch.equal(out[6], ref[6]) and torch.equal(out[7], ref[7]) and torch.equal(out[9], ref[9])
    assert torch.equal(out_f[5], ref_f[5])
    # adapted learner (2 first-order... here second-order-capable adapt(), same weights): train mode, batch statistics
    sysm.train()
    learner = sysm.adapt([([sup], [qry])], 2, train=False)
    assert learner == 2
    Pc = {k: v.detach().clone() for k, v in P.items()}
    ref, theta = O.test_time_adaptation(Pc, CFG, sup, qry, 2, 2, saving_steps=(2,))
    out = sysm.forward_learner(learner, sup[2], *qry[3:], average_spk_emb=True)
    assert rel(out[1], ref["step_2"]["recon"]["output"][1]) < 1e-3
