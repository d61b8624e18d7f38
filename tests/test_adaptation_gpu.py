"""SURVEY §8 rows f3/f5 on the B200: `MetaSystem.test_step` (few-shot adaptation inference, BASELINE configs[4]) through the
C ABI against `oracle.fs2_oracle.test_time_adaptation` — eval-mode step_0, rolling first-order adaptation, train-mode
teacher-forced recon and free-running synthesis (dropout off and on).  Outputs within 1e-3 relative fp32 (north_star);
rounded durations / mel lengths / masks bit-exact; plus the full configs[4] size (16-shot, 20 steps, 128 phonemes) with
size-independent properties."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu

from meta_tts_b200.systems import DEFAULT_ALGORITHM_CONFIG, DEFAULT_TRAIN_CONFIG, MetaSystem  # noqa: E402
from oracle import fs2_oracle as O  # noqa: E402
from tests_helpers_adapt import check_outputs, rel, talkative_params  # noqa: E402


def _system(cfg, steps, test_steps, saving, dropout, one_shot=False):
    algo = copy.deepcopy(DEFAULT_ALGORITHM_CONFIG)
    algo["adapt"]["train"]["steps"] = steps
    algo["adapt"]["test"] = {"steps": test_steps, "saving_steps": list(saving), "1-shot": one_shot}
    return MetaSystem(None, cfg, DEFAULT_TRAIN_CONFIG, algo, n_speaker=16, device="cuda:0", use_cuda_graph=False, dropout=dropout, seed=3)


@pytest.mark.parametrize("dropout", [False, True])
def test_small_model_test_step(cuda_device, dropout):
    cfg = O.small_model_config(1, 1)
    P = talkative_params(cfg)
    sysm = _system(cfg, 2, 6, (2, 6), dropout)
    sysm.load_state_dict({k: v.detach().clone() for k, v in P.items()})
    sup, qry = O.synth_task(task=2, shots=3, queries=1, L=7, T=20, ragged=True)
    outs = sysm.test_step([([sup], [qry])], 0)
    torch.cuda.synchronize()
    Pc = {k: v.detach().clone() for k, v in P.items()}
    ref, theta = O.test_time_adaptation(Pc, cfg, sup, qry, 2, 6, saving_steps=(2, 6), drop_seed=(0, sysm.last_salt) if dropout else None)
    check_outputs(outs[0], ref, verbose=f"small dropout={dropout}")
    fw = sysm.maml.fast_weights(1)
    assert sorted(rel(fw[k], theta[k]) for k in theta)[len(theta) // 2] < 1e-4
    for i in range(5):
        assert rel(sysm.maml.consts[f"postnet.convolutions.{i}.1.running_var"], Pc[f"postnet.convolutions.{i}.1.running_var"]) < 1e-4


def test_base_model_test_step_ragged(cuda_device):
    """Base model (4 + 6 layers), 4-shot ragged support, 10 first-order steps in 2 rounds, dropout on."""
    cfg = O.BASE_MODEL_CONFIG
    P = talkative_params(cfg)
    sysm = _system(cfg, 5, 10, (5, 10), True)
    sysm.load_state_dict({k: v.detach().clone() for k, v in P.items()})
    sup, qry = O.synth_task(task=6, shots=4, queries=1, L=24, T=90, ragged=True)
    outs = sysm.test_step([([sup], [qry])], 0)
    torch.cuda.synchronize()
    Pc = {k: v.detach().clone() for k, v in P.items()}
    ref, _ = O.test_time_adaptation(Pc, cfg, sup, qry, 5, 10, saving_steps=(5, 10), drop_seed=(0, sysm.last_salt))
    check_outputs(outs[0], ref, verbose="base ragged K=10 dropout")


def test_config5_full_size(cuda_device):
    """BASELINE configs[4]: 20 first-order inner steps on a 16-shot support set (128 phonemes -> 864 frames), then free-running
    synthesis of the query and Griffin-Lim decode.  The oracle's 20 full-size steps take minutes on the host, so the step_0
    forwards are compared against the oracle and the adapted ones are checked through properties: the support-driven query
    loss goes down, mel_len == sum(d_rounded), padded frames are zero, the waveform has hop * (T - 2) samples."""
    from meta_tts_b200 import audio as PA
    from meta_tts_b200 import ops as _ops

    cfg = O.BASE_MODEL_CONFIG
    P = talkative_params(cfg)
    sysm = _system(cfg, 5, 20, (5, 10, 20), True)
    sysm.load_state_dict({k: v.detach().clone() for k, v in P.items()})
    sup, qry = O.synth_task(task=9, shots=16, queries=1, L=128, T=864)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = _ops.launch_count
    ev0.record()
    out = sysm.test_step([([sup], [qry])], 0)[0]
    ev1.record()
    torch.cuda.synchronize()
    print(f"[config5] test_step (20 first-order steps on 16 x 864 frames + 5 recon + 4 synth forwards): {ev0.elapsed_time(ev1):.1f} ms, "
          f"{_ops.launch_count - n0} launches")
    with torch.no_grad():
        Pc = {k: v.detach().clone() for k, v in P.items()}
        r0 = O.fs2_forward(Pc, cfg, sup[2], *qry[3:], average_spk_emb=True, training=False)
        s0 = O.fs2_forward(Pc, cfg, sup[2], *qry[3:6], average_spk_emb=True, training=False)
    ref0 = {"step_0": {"recon": {"losses": tuple(O.fs2_loss(qry, r0)), "output": r0}, "synth": {"output": s0}}}
    check_outputs({"step_0": out["step_0"]}, ref0, verbose="config5 full size")
    losses = [float(out[f"step_{k}"]["recon"]["losses"][0]) for k in (0, 5, 10, 15, 20)]
    print("[config5] query loss after 0/5/10/15/20 steps:", [f"{v:.4f}" for v in losses])
    assert losses[-1] < losses[1]            # (step_0 is eval mode, the others train mode: compare like with like)
    for k in (5, 10, 20):
        o = out[f"step_{k}"]["synth"]["output"]
        mel_len, d_r = o[9].cpu(), o[5].cpu()
        assert torch.equal(mel_len, d_r.to(torch.int64).clamp(min=0).sum(1))
        assert o[1].shape[1] == min(int(mel_len.max()), 1000)          # train mode: decoder truncates at max_seq_len (Models.py:161-166)
    # vocoder-side decode of the step-20 synthesis (tools.py:18-34)
    mel = out["step_20"]["synth"]["output"][1][0]                      # [T, 80] postnet mel
    T = mel.shape[0]
    stft = PA.TacotronSTFT(1024, 256, 1024, 80, 22050, 0, 8000, device="cuda:0")
    ev0.record()
    wav = PA.inv_mel_spec(mel.t(), None, stft, 60)
    ev1.record()
    torch.cuda.synchronize()
    print(f"[config5] Griffin-Lim x60 on {T} frames: {ev0.elapsed_time(ev1):.1f} ms")
    assert wav.shape == (256 * (T - 2),) and bool(torch.isfinite(torch.from_numpy(wav)).all())


def test_sequences_at_and_beyond_max_seq_len(cuda_device):
    """Maximum sizes: a teacher-forced batch of exactly max_seq_len = 1000 frames, and free-running synthesis that predicts MORE
    than max_seq_len frames — eval mode keeps the length with a recomputed sinusoid table (Models.py:148-156), train mode keeps
    the first 1000 frames while mel_len reports the full sum of durations (Models.py:161-166)."""
    from meta_tts_b200.maml import batch_from_tuple
    cfg = O.small_model_config(1, 1)
    P = talkative_params(cfg)
    P["variance_adaptor.duration_predictor.linear_layer.bias"] = P["variance_adaptor.duration_predictor.linear_layer.bias"] + 1.2   # ~ 11 frames / phoneme
    sysm = _system(cfg, 1, 1, (1,), False)
    sysm.load_state_dict({k: v.detach().clone() for k, v in P.items()})
    sup, qry = O.synth_task(task=4, shots=1, queries=1, L=120, T=1000)
    m = sysm.maml
    out = m.predict(batch_from_tuple(qry, "cuda:0", spk_ids=sup[2], average_spk=True), adapted=False, eval_mode=True)
    with torch.no_grad():
        ref = O.fs2_forward({k: v.detach().clone() for k, v in P.items()}, cfg, sup[2], *qry[3:], average_spk_emb=True, training=False)
    assert rel(out["postnet"], ref[1]) < 1e-3 and out["postnet"].shape[1] == 1000
    bt = batch_from_tuple(qry, "cuda:0", spk_ids=sup[2], average_spk=True, targets=False)
    for train in (False, True):
        sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
        out = m.predict(bt, adapted=False, free_running=True, eval_mode=not train)
        with torch.no_grad():
            ref = O.fs2_forward(sd, cfg, sup[2], *qry[3:6], average_spk_emb=True, training=train)
        total = int(ref[9].max())
        assert total > 1000, total
        assert torch.equal(out["d_rounded"].cpu(), ref[5]) and torch.equal(out["mel_len"].cpu(), ref[9])
        assert out["postnet"].shape == ref[1].shape and out["postnet"].shape[1] == (1000 if train else total)
        print(f"[adapt] beyond max_seq_len ({total} frames), train={train}: postnet rel {rel(out['postnet'], ref[1]):.2e}")
        assert rel(out["postnet"], ref[1]) < 1e-3


def test_against_goldens_from_the_real_reference_modules(cuda_device):
    """Free-running synthesis (eval / train, within and beyond max_seq_len) and the test-step protocol straight against
    tests/golden/synth_golden.npz, produced by the REAL reference modules (oracle/make_golden_synth.py)."""
    import os

    import numpy as np
    from meta_tts_b200.maml import batch_from_tuple
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "synth_golden.npz"), allow_pickle=False)
    cfg = O.BASE_MODEL_CONFIG
    t, S, Q, L, T = [int(v) for v in G["task_cfg"]]
    sup, qry = O.synth_task(task=t, shots=S, queries=Q, L=L, T=T, ragged=True)

    def fresh(steps=2, total=4):
        P = O.init_params(seed=0)
        P["variance_adaptor.duration_predictor.linear_layer.bias"] = P["variance_adaptor.duration_predictor.linear_layer.bias"] + float(G["bias"])
        s = _system(cfg, steps, total, (2, 4), False)
        s.load_state_dict(P)
        return s

    def check(tag, post, d_rounded, mel_len, tol):
        assert tuple(post.shape) == tuple(int(v) for v in G[f"{tag}_shape"]), tag
        assert np.array_equal(d_rounded.float().cpu().numpy(), G[f"{tag}_d_rounded"]) and np.array_equal(mel_len.cpu().numpy(), G[f"{tag}_mel_len"]), tag
        r = rel(post[:, ::8], torch.from_numpy(G[f"{tag}_postnet"]))
        print(f"[adapt] real-reference golden {tag}: T = {post.shape[1]}, postnet rel {r:.2e}")
        assert r < tol, (tag, r)

    bt = batch_from_tuple(qry, "cuda:0", spk_ids=sup[2], average_spk=True, targets=False)
    for tag, train, dc in (("free_eval", False, 1.0), ("free_train", True, 1.0), ("free_eval_long", False, 12.0), ("free_train_long", True, 12.0)):
        out = fresh().maml.predict(bt, adapted=False, free_running=True, eval_mode=not train, d_control=dc)
        check(tag, out["postnet"], out["d_rounded"], out["mel_len"], 1e-3)
    outs = fresh().test_step([([sup], [qry])], 0)[0]
    for k in ("step_0", "step_2", "step_4"):
        o = outs[k]["synth"]["output"]
        check(f"tta_{k}_synth", o[1], o[5], o[9], 1e-3 if k == "step_0" else 5e-3)
        got = torch.stack([x.cpu() for x in outs[k]["recon"]["losses"]])
        assert rel(got, torch.from_numpy(G[f"tta_{k}_losses"]).float()) < (1e-3 if k == "step_0" else 5e-3), k
