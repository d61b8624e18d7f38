"""SURVEY §8 row f5 (checkpoint compatibility): key order of the reference model, Lightning-format save / load round trip
incl. the torch-Adam optimizer state, and the reference's `on_load_checkpoint` / `on_test_start` rules (system.py:115-212)."""
import copy
import os

import numpy as np
import torch

from meta_tts_b200 import checkpoint as CK
from meta_tts_b200.systems import DEFAULT_ALGORITHM_CONFIG, DEFAULT_MODEL_CONFIG, DEFAULT_TRAIN_CONFIG, MetaSystem
from oracle import fs2_oracle as O
from oracle.ops_reference import RefOps

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "fs2_golden.npz"), allow_pickle=False)
CFG = O.small_model_config(1, 1)


def _system(n_speaker=16, pre=None, algo=None):
    algo = copy.deepcopy(algo or DEFAULT_ALGORITHM_CONFIG)
    algo["adapt"]["train"]["steps"] = 1
    algo["adapt"]["test"]["steps"] = 1
    return MetaSystem(pre, CFG, DEFAULT_TRAIN_CONFIG, algo, n_speaker=n_speaker, device="cpu", use_cuda_graph=False,
                      backend=RefOps(split=3), dropout=False)


def test_key_order_matches_real_reference_model():
    """`sd_keys` was recorded from the REAL reference FastSpeech2 (oracle/make_golden.py)."""
    assert CK.reference_state_dict_keys(DEFAULT_MODEL_CONFIG) == [str(k) for k in G["sd_keys"]]
    P = O.init_params(seed=0)
    assert sorted(CK.reference_state_dict_keys(DEFAULT_MODEL_CONFIG)) == sorted(P.keys())
    params = CK.reference_parameter_keys(DEFAULT_MODEL_CONFIG)
    assert len(params) == 235 - 15 and "encoder.position_enc" in params and not any("running" in k for k in params)


def test_save_load_round_trip_with_optimizer_state():
    a = _system()
    P = O.init_params(seed=0, model_config=CFG)
    a.load_state_dict(P)
    for t in range(2):
        sup, qry = O.synth_task(task=t + 2, shots=2, queries=2, L=5, T=12, ragged=True)
        a.training_step([([sup], [qry])], t)
        a.optimizer_step()
    ck = a.save_checkpoint()
    assert ck["global_step"] == 2 and all(k.startswith("model.") for k in ck["state_dict"])
    st = ck["optimizer_states"][0]
    assert set(st["param_groups"][0]) >= {"lr", "betas", "eps", "weight_decay", "params"} and len(st["param_groups"][0]["params"]) == len(CK.reference_parameter_keys(CFG))
    # torch.optim.Adam accepts it for a parameter list in reference order (the frozen tables have no state)
    keys = CK.reference_parameter_keys(CFG)
    sd = {k[6:]: v for k, v in ck["state_dict"].items()}
    plist = [torch.nn.Parameter(sd[k].clone().float(), requires_grad=sd[k].is_floating_point()) for k in keys]
    opt = torch.optim.Adam(plist, lr=1e-3, betas=(0.9, 0.98), eps=1e-9)
    opt.load_state_dict(copy.deepcopy(st))
    i = keys.index("mel_linear.weight")
    assert torch.equal(opt.state[plist[i]]["exp_avg"], st["state"][i]["exp_avg"]) and keys.index("encoder.position_enc") not in st["state"]
    b = _system()
    b.load_state_dict(P)
    b.load_checkpoint(copy.deepcopy(ck))
    assert b.maml.opt_step == 2 and b.global_step == 2
    for name in ("theta", "adam_m", "adam_v"):
        assert torch.equal(getattr(a.maml, name), getattr(b.maml, name)), name
    sup, qry = O.synth_task(task=7, shots=2, queries=2, L=5, T=12, ragged=True)
    la = a.training_step([([sup], [qry])], 0)["losses"]
    lb = b.training_step([([sup], [qry])], 0)["losses"]
    a.optimizer_step()
    b.optimizer_step()
    assert all(float(la[i]) == float(lb[i]) for i in range(6)) and torch.equal(a.maml.theta, b.maml.theta)   # resume == continue


def test_on_load_checkpoint_rules():
    P = O.init_params(seed=0, model_config=CFG, n_speaker=2390)
    # (1) old key name + unknown key + missing key
    s = _system(n_speaker=2390)
    s.load_state_dict(P)
    sd = {("model." + k): v.clone() for k, v in P.items()}
    sd["model.speaker_emb.weight"] = sd.pop("model.speaker_emb.model.weight") + 1.0
    sd["model.some_old_module.weight"] = torch.zeros(3)
    del sd["model.mel_linear.bias"]
    ck = {"global_step": 7, "state_dict": sd, "optimizer_states": [{"state": {}, "param_groups": []}]}
    s.load_checkpoint(ck)
    ch = s.checkpoint_changes
    assert ch["replace"] == [["model.speaker_emb.weight", "model.speaker_emb.model.weight"]]
    assert ch["drop"] == ["model.some_old_module.weight"] and ch["miss"] == ["model.mel_linear.bias"]
    assert "optimizer_states" not in ck and s.test_global_step == 7
    got = s.state_dict()
    assert torch.equal(got["speaker_emb.model.weight"], P["speaker_emb.model.weight"] + 1.0)
    assert torch.equal(got["mel_linear.bias"], P["mel_linear.bias"])                      # kept
    # (2) LibriTTS train-clean-100 checkpoint (326 speakers) into the all-LibriTTS table (2390): rows [:247] and [-79:]
    s = _system(n_speaker=2390, pre={"dataset": "LibriTTS"})
    s.load_state_dict(P)
    small = torch.randn(326, P["speaker_emb.model.weight"].shape[1])
    sd = {("model." + k): v.clone() for k, v in P.items()}
    sd["model.speaker_emb.model.weight"] = small
    s.load_checkpoint({"global_step": 1, "state_dict": sd})
    w = s.state_dict()["speaker_emb.model.weight"]
    assert torch.equal(w[:247], small[:247]) and torch.equal(w[-79:], small[-79:])
    assert torch.equal(w[247:-79], P["speaker_emb.model.weight"][247:-79]) and s.checkpoint_changes["skip"][0][0] == "model.speaker_emb.model.weight"
    # (3) another corpus + avg_train_spk_emb: every row = mean of the 247 training rows
    algo = copy.deepcopy(DEFAULT_ALGORITHM_CONFIG)
    algo["adapt"]["test"]["avg_train_spk_emb"] = True
    s = _system(n_speaker=40, pre={"dataset": "VCTK"}, algo=algo)
    P40 = O.init_params(seed=0, model_config=CFG, n_speaker=40)
    s.load_state_dict(P40)
    s.load_checkpoint({"global_step": 1, "state_dict": sd})
    w = s.state_dict()["speaker_emb.model.weight"]
    assert w.shape[0] == 40 and torch.allclose(w, small[:247].mean(0).expand(40, -1))
    # (4) on_test_start: LibriTTS + avg_train_spk_emb overwrites the 39 test speakers
    s = _system(n_speaker=2390, pre={"dataset": "LibriTTS"}, algo=algo)
    s.load_state_dict(P)
    s.on_test_start()
    w = s.state_dict()["speaker_emb.model.weight"]
    assert torch.allclose(w[-39:], P["speaker_emb.model.weight"][:247].mean(0).expand(39, -1)) and torch.equal(w[:-39], P["speaker_emb.model.weight"][:-39])


def test_exported_optimizer_and_scheduler_state_load_into_torch():
    """the exported `optimizer_states` / `lr_schedulers` entries must load into a REAL torch Adam + LambdaLR
    (what Lightning does on resume) and continue with the same learning rate as this framework's next step."""
    a = _system()
    a.load_state_dict(O.init_params(seed=0, model_config=CFG))
    for t in range(3):
        sup, qry = O.synth_task(task=t + 2, shots=2, queries=2, L=5, T=12, ragged=True)
        a.training_step([([sup], [qry])], t)
        a.optimizer_step()
    ck = a.save_checkpoint()
    keys = CK.reference_parameter_keys(CFG)
    sd = {k[6:]: v for k, v in ck["state_dict"].items()}
    plist = [torch.nn.Parameter(sd[k].clone().float(), requires_grad=sd[k].is_floating_point()) for k in keys]
    opt = torch.optim.Adam(plist, lr=CFG["transformer"]["encoder_hidden"] ** -0.5, betas=(0.9, 0.98), eps=1e-9)
    # the reference's scheduler lambda (lightning/scheduler.py:11-23), restated
    lam = lambda step: min((step + 1) ** -0.5, 4000 ** -1.5 * (step + 1))  # noqa: E731
    sch = torch.optim.lr_scheduler.LambdaLR(opt, lam)
    opt.load_state_dict(copy.deepcopy(ck["optimizer_states"][0]))
    sch.load_state_dict(copy.deepcopy(ck["lr_schedulers"][0]))          # raised KeyError('lr_lambdas') before the fix
    assert sch.last_epoch == 3 and sch.base_lrs == [CFG["transformer"]["encoder_hidden"] ** -0.5]
    want = a.maml.lr_schedule(3)                                        # the rate this framework uses for step index 3
    assert abs(opt.param_groups[0]["lr"] - want) <= 1e-12 and abs(sch.get_last_lr()[0] - want) <= 1e-12
    assert abs(want - CFG["transformer"]["encoder_hidden"] ** -0.5 * lam(3)) <= 1e-15
