"""End-to-end parity of the CUDA path (C ABI kernels driven by the engine) against
(a) the autograd oracle on the same seeded inputs and (b) the golden vectors produced by the REAL
reference modules (tests/golden/fs2_golden.npz).  Tolerance: north_star's 1e-3 relative (fp32) for
mel / pitch / energy / duration outputs in bf16x3 mode; gradients compared against the total
gradient norm.  The LengthRegulator index path is checked bit-exact (mel_len / idx)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from meta_tts_b200.maml import MamlEngine, batch_from_tuple  # noqa: E402
from meta_tts_b200.ops import CudaOps  # noqa: E402
from oracle import fs2_oracle as O  # noqa: E402

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "fs2_golden.npz"), allow_pickle=False)
REL_OUT = 1e-3            # BASELINE.json north_star tolerance


def _rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _engine(P, cfg, split=3, K=2):
    be = CudaOps(split=split)
    m = MamlEngine(be, cfg, n_speaker=16, adapt_modules=O.ADAPT_MODULES, inner_lr=0.001, max_inner_steps=K)
    m.load_state_dict({k: v.detach().clone() for k, v in P.items()})
    return m


def _check_task(m, P, cfg, sup, qry, steps, first_order, grad_tol, label, fast_tol=1e-2, salt=None, median_tol=2e-3):
    """salt = None: dropout off (identity) on both sides.  salt = uint32: train-mode dropout ON; the oracle applies the
    masks of the same counter hash (oracle/fs2_oracle.py drop_keep) the kernels evaluate on the device."""
    Pc = {k: v.detach().clone() for k, v in P.items()}
    losses, preds, grads, fast = O.maml_task_step(Pc, cfg, sup, qry, steps, 0.001, first_order, return_fast_weights=True,
                                                  drop_seed=None if salt is None else (0, salt))
    dev = m.theta.device
    bs = batch_from_tuple(sup, dev)
    bq = batch_from_tuple(qry, dev, spk_ids=sup[2], average_spk=True)
    if salt is not None:
        m.be.drop_salt = torch.tensor([salt - (1 << 32) if salt >= (1 << 31) else salt], dtype=torch.int32, device=dev)
    loss6, out = m.task_step(bs, bq, steps, first_order, drop_base=None if salt is None else 0)
    torch.cuda.synchronize()
    r_loss = _rel(loss6, torch.stack(losses))
    r_mel = _rel(out["mel"].reshape(preds[0].shape), preds[0])
    r_post = _rel(out["postnet"].reshape(preds[1].shape), preds[1])
    r_p, r_e, r_d = _rel(out["pitch"], preds[2]), _rel(out["energy"], preds[3]), _rel(out["logd"], preds[4])
    assert torch.equal(out["mel_len"].cpu(), preds[9]), "LengthRegulator mel_len must be bit-exact"
    fw = m.fast_weights(steps)
    r_fast = max(_rel(fw[k], fast[k]) for k in fast)
    print("[engine]   worst fast weights (rel err):", [(f"{e:.1e}", k) for e, k in sorted(((_rel(fw[k], fast[k]), k) for k in fast), reverse=True)[:4]])
    got = m.task_grads()
    tot_ref = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values()))
    tot_err = torch.sqrt(sum(((got[k].double() - grads[k].double()) ** 2).sum() for k in grads))
    r_grad = (tot_err / tot_ref).item()
    worst = max(((got[k].double() - grads[k].double()).norm() / tot_ref).item() for k in grads)
    print(f"[engine] {label}: loss {r_loss:.2e} mel {r_mel:.2e} post {r_post:.2e} pitch {r_p:.2e} energy {r_e:.2e} "
          f"logd {r_d:.2e} fast {r_fast:.2e} grad(total) {r_grad:.2e} grad(worst tensor/total) {worst:.2e}")
    top = sorted(((((got[k].double() - grads[k].double()).norm() / tot_ref).item(), k) for k in grads), reverse=True)[:4]
    print("[engine]   worst tensors (err / total grad norm):", [(f"{e:.1e}", k) for e, k in top])
    assert max(r_loss, r_mel, r_post, r_p, r_e, r_d) < REL_OUT
    # fast weights of zero-initialised parameters (LN/BN biases) are pure gradients: their relative error is
    # the gradient's, which carries fp32 summation-order noise and the occasional ReLU-kink flip (a unit whose
    # pre-activation is ~1e-6 from 0 gates differently in two fp32 implementations; seen 1 in 20k on CPU too).
    assert r_fast < fast_tol
    assert r_grad < grad_tol
    # robust to ReLU-kink flips: the MEDIAN per-tensor relative error (a flipped unit moves a handful of tensors, not most)
    per = sorted(((got[k].double() - grads[k].double()).norm() / grads[k].double().norm()).item() for k in grads
                 if grads[k].double().norm() > 1e-4 * tot_ref)
    print(f"[engine]   per-tensor gradient rel err: median {per[len(per) // 2]:.2e}, p90 {per[int(0.9 * len(per))]:.2e}")
    assert per[len(per) // 2] < median_tol, f"median per-tensor gradient rel err {per[len(per) // 2]:.2e}"


def test_small_model_all_modes(cuda_device):
    cfg = O.small_model_config(1, 1)
    P = O.init_params(seed=0, model_config=cfg)
    m = _engine(P, cfg)
    sup, qry = O.synth_task(task=3, shots=2, queries=2, L=6, T=18, ragged=True)
    for steps, fo in ((1, True), (1, False), (2, False)):
        m.load_state_dict({k: v.detach().clone() for k, v in P.items()})
        _check_task(m, P, cfg, sup, qry, steps, fo, 1e-3, f"small K={steps} fo={fo}")


def test_dropout_on_parity(cuda_device):
    """Train-mode dropout ACTIVE (encoder/decoder 0.2, variance predictors 0.5, postnet 0.5 — the reference's
    learner.train(), base_adaptor.py:103): outputs, fast weights and second-order gradients still match the oracle,
    because both sides drop the same elements."""
    cfg = O.small_model_config(1, 1)
    P = O.init_params(seed=0, model_config=cfg)
    m = _engine(P, cfg)
    sup, qry = O.synth_task(task=3, shots=2, queries=2, L=6, T=18, ragged=True)
    for steps, fo, salt in ((1, True, 7), (2, False, 0xF00DFACE)):
        m.load_state_dict({k: v.detach().clone() for k, v in P.items()})
        _check_task(m, P, cfg, sup, qry, steps, fo, 2e-3, f"small dropout K={steps} fo={fo}", salt=salt, fast_tol=1e-2)
    # masks really are active and salt-dependent: same task, two salts -> different losses
    dev = m.theta.device
    bs, bq = batch_from_tuple(sup, dev), batch_from_tuple(qry, dev, spk_ids=sup[2], average_spk=True)
    vals = []
    for salt in (1, 2, 2):
        m.load_state_dict({k: v.detach().clone() for k, v in P.items()})
        m.be.drop_salt = torch.tensor([salt], dtype=torch.int32, device=dev)
        vals.append(m.task_step(bs, bq, 1, True, drop_base=0)[0].cpu().clone())
    # (not bit-equal: split-K / channel-sum atomics reorder fp32 additions run to run)
    assert not torch.allclose(vals[0], vals[1], rtol=1e-2) and torch.allclose(vals[1], vals[2], rtol=1e-5)


def test_config2_full_size_dropout_parity(cuda_device):
    """BASELINE configs[1] with dropout ON (what bench.py times)."""
    cfg = O.BASE_MODEL_CONFIG
    P = O.init_params(seed=0)
    m = _engine(P, cfg, K=1)
    sup, qry = O.synth_task(task=0, shots=4, queries=4, L=128, T=864)
    _check_task(m, P, cfg, sup, qry, 1, False, 4e-3, "config2 full size second-order, dropout ON", salt=20260925, fast_tol=2e-2)


def test_base_model_ragged_second_order(cuda_device):
    cfg = O.BASE_MODEL_CONFIG
    P = O.init_params(seed=0)
    m = _engine(P, cfg, K=2)
    sup, qry = O.synth_task(task=1, shots=3, queries=2, L=40, T=150, ragged=True)
    _check_task(m, P, cfg, sup, qry, 2, False, 4e-3, "base ragged K=2 second-order", fast_tol=2e-2)   # pitch-predictor ReLU kink (same on CPU)


def test_golden_reference_task_steps(cuda_device):
    """Against the goldens generated from the real reference modules + restated learn2learn."""
    cfg = O.BASE_MODEL_CONFIG
    for tag in ("maml_so_k2", "maml_fo_k2", "maml_so_k1"):
        K, fo, S, Q, L, T, task = [int(v) for v in G[f"{tag}_cfg"]]
        P = O.init_params(seed=0)
        m = _engine(P, cfg, K=2)
        sup, qry = O.synth_task(task=task, shots=S, queries=Q, L=L, T=T, ragged=True)
        dev = m.theta.device
        loss6, out = m.task_step(batch_from_tuple(sup, dev), batch_from_tuple(qry, dev, spk_ids=sup[2], average_spk=True),
                                 K, bool(fo))
        torch.cuda.synchronize()
        ref_losses = torch.from_numpy(G[f"{tag}_losses"])
        assert _rel(loss6, ref_losses) < REL_OUT
        assert _rel(out["mel"].reshape(G[f"{tag}_mel"].shape), torch.from_numpy(G[f"{tag}_mel"])) < REL_OUT
        got = m.task_grads()
        names = [str(k) for k in G[f"{tag}_grad_names"]]
        gn = torch.tensor([got[k].double().norm().item() for k in names])
        ref_gn = torch.from_numpy(G[f"{tag}_grad_norm"])
        tot = ref_gn.norm()
        assert ((gn - ref_gn).abs().max() / tot).item() < 1e-3
        heads = np.stack([np.pad(got[k].flatten()[:8].numpy(), (0, max(0, 8 - got[k].numel()))) for k in names])
        assert np.abs(heads - G[f"{tag}_grad_head"]).max() / np.abs(G[f"{tag}_grad_head"]).max() < 2e-3
        print(f"[engine] golden {tag}: loss rel {_rel(loss6, ref_losses):.2e}")


def test_config2_full_size_parity(cuda_device):
    """BASELINE configs[1]: MAML 1 inner step, 1 task, 4-shot support / query, 128 phonemes -> 864 frames."""
    cfg = O.BASE_MODEL_CONFIG
    P = O.init_params(seed=0)
    m = _engine(P, cfg, K=1)
    sup, qry = O.synth_task(task=0, shots=4, queries=4, L=128, T=864)
    _check_task(m, P, cfg, sup, qry, 1, False, 1e-3, "config2 full size second-order")


@pytest.mark.parametrize("first_order,salt", [(False, None), (False, 77), (True, 78)])
def test_config3_config4_k5_structure(cuda_device, first_order, salt):
    """BASELINE configs[2] / configs[3] structure: S = Q = 5, K = 5 inner steps (meta_emb_vad.yaml:23-29), second-order
    (config 3) and first-order (config 4), base model, ragged batch — at reduced sequence length so the oracle's
    5-step double backward finishes in seconds.  Exercises all five fast-weight arenas and the 5-deep adjoint recursion."""
    cfg = O.BASE_MODEL_CONFIG
    P = O.init_params(seed=0)
    m = _engine(P, cfg, K=5)
    sup, qry = O.synth_task(task=5, shots=5, queries=5, L=16, T=64, ragged=True)
    # Tolerance: this task step evaluates 14.5 M ReLU inputs, ~25 of them within 1e-6 of zero and ~950 within 1e-5
    # (counted in the oracle); bf16x3 products carry ~4e-6 relative error, so a few units gate differently than in the
    # fp32 oracle.  A flip changes the forward by ~1e-6 but switches that unit's whole gradient on or off; in the small
    # variance predictors (80 rows here) one flip moves a conv weight gradient by ~1e-2 of the total norm (the CPU
    # restatement with exact products matches to 4e-5 on the same inputs).  Hence total < 2e-2 + the median criterion.
    # Which units flip varies run to run (split-K / channel-sum atomics reorder fp32 additions): 9 repeats on one B200
    # gave total 1.5e-4 ... 1.3e-2 and median 3e-5 ... 5.6e-3 for the same inputs (profiles/r01_s2_k5_repeats.log), a flip
    # early in the decoder moving every upstream tensor a little — so the median bar here is 1e-2, not 2e-3.
    _check_task(m, P, cfg, sup, qry, 5, first_order, 2e-2, f"config{'4' if first_order else '3'} structure K=5 salt={salt}",
                fast_tol=2e-2, salt=salt, median_tol=1e-2)


def test_bf16_single_pass_mode_runs(cuda_device):
    """split=1 (plain bf16 operands): documented looser tolerance (bf16 rounding ~4e-3 per product)."""
    cfg = O.small_model_config(1, 1)
    P = O.init_params(seed=0, model_config=cfg)
    m = _engine(P, cfg, split=1)
    sup, qry = O.synth_task(task=3, shots=2, queries=2, L=6, T=18, ragged=True)
    Pc = {k: v.detach().clone() for k, v in P.items()}
    losses, preds, grads = O.maml_task_step(Pc, cfg, sup, qry, 1, 0.001, False)
    dev = m.theta.device
    loss6, out = m.task_step(batch_from_tuple(sup, dev), batch_from_tuple(qry, dev, spk_ids=sup[2], average_spk=True), 1, False)
    torch.cuda.synchronize()
    r = _rel(out["mel"].reshape(preds[0].shape), preds[0])
    print(f"[engine] bf16 single-pass mel rel {r:.2e}, loss rel {_rel(loss6, torch.stack(losses)):.2e}")
    assert r < 3e-2 and _rel(loss6, torch.stack(losses)) < 3e-2


def test_default_precision_policy_config2_full_size(cuda_device):
    """What bench.py times: BASELINE configs[1] with dropout ON under the DEFAULT per-class precision policy (weight-gradient and
    tangent-forward products single-pass bf16, everything that reaches the outputs or a data gradient bf16x3).  Outputs keep the
    north_star bar (1e-3; measured 2e-5).  The outer gradient is held to the tolerance of the all-bf16x3 dropout test above (4e-3):
    with dropout on both land at ~2e-3 of the gradient norm (1.9e-3 strict / 2.1e-3 default policy: ReLU / L1 kink flips dominate);
    without dropout the policy measures 3.1e-4 against 2.2e-4 (profiles/r02_precision_budget.md)."""
    from meta_tts_b200.engine import DEFAULT_SPLIT_POLICY
    cfg = O.BASE_MODEL_CONFIG
    P = O.init_params(seed=0)
    m = _engine(P, cfg, K=1)
    m.engine.g.policy = dict(DEFAULT_SPLIT_POLICY)
    sup, qry = O.synth_task(task=0, shots=4, queries=4, L=128, T=864)
    _check_task(m, P, cfg, sup, qry, 1, False, 4e-3, "config2 full size, DEFAULT precision policy, dropout ON", salt=20260925, fast_tol=2e-2,
                median_tol=3e-3)


def test_default_precision_policy_k5(cuda_device):
    """Five inner steps / five Hessian-vector passes under the default policy (BASELINE configs[2] structure)."""
    from meta_tts_b200.engine import DEFAULT_SPLIT_POLICY
    cfg = O.small_model_config(2, 2)
    P = O.init_params(seed=0, model_config=cfg)
    m = _engine(P, cfg, K=5)
    m.engine.g.policy = dict(DEFAULT_SPLIT_POLICY)
    sup, qry = O.synth_task(task=4, shots=5, queries=5, L=16, T=64, ragged=True)
    _check_task(m, P, cfg, sup, qry, 5, False, 2e-2, "K=5 second order, DEFAULT precision policy", fast_tol=2e-2, median_tol=1e-2)


@pytest.mark.parametrize("first_order", [False, True])
def test_config3_config4_full_size_parity(cuda_device, first_order):
    """BASELINE configs[2] (second order) / configs[3] (first order) at FULL size: one task step with 5-shot support + 5 queries,
    128 phonemes -> 864 frames, K = 5 inner steps, base model, dropout off (the oracle's 5-step double backward takes ~1 min of
    CPU).  Outputs to the north_star bar; the outer gradient to 4e-3 of its norm and 3e-3 median per tensor (five inner steps let
    the ReLU / L1 kink flips documented above reach the fast weights)."""
    cfg = O.BASE_MODEL_CONFIG
    P = O.init_params(seed=0)
    m = _engine(P, cfg, K=5)
    sup, qry = O.synth_task(task=0, shots=5, queries=5, L=128, T=864)
    _check_task(m, P, cfg, sup, qry, 5, first_order, 4e-3, f"config{'4' if first_order else '3'} FULL size K=5", fast_tol=2e-2,
                median_tol=3e-3)


@pytest.mark.parametrize("eval_mode", [False, True])
def test_config1_single_utterance_forward(cuda_device, eval_mode):
    """BASELINE configs[0]: FastSpeech2-base single-utterance forward, 64 phonemes -> 512 mel frames (no MAML): every output of
    `FastSpeech2.forward` (fastspeech2.py:40-112) and the 6 losses against the oracle, in train mode (batch statistics in the
    postnet) and under model.eval() (running statistics)."""
    cfg = O.BASE_MODEL_CONFIG
    P = O.init_params(seed=0)
    m = _engine(P, cfg, K=1)
    b12 = O.synth_batch(1, 64, 512, seed=21, speaker=5)
    with torch.no_grad():
        preds = O.fs2_forward({k: v.detach().clone() for k, v in P.items()}, cfg, *b12[2:], training=not eval_mode)
        losses = O.fs2_loss(b12, preds)
    out = m.engine.forward(m.params(0), batch_from_tuple(b12, m.theta.device), m.engine.new_tape(), update_bn=not eval_mode,
                           eval_mode=eval_mode)
    torch.cuda.synchronize()
    errs = {"mel": _rel(out["mel"].reshape(preds[0].shape), preds[0]), "postnet": _rel(out["postnet"].reshape(preds[1].shape), preds[1]),
            "pitch": _rel(out["pitch"], preds[2]), "energy": _rel(out["energy"], preds[3]), "logd": _rel(out["logd"], preds[4]),
            "loss6": _rel(out["loss6"], torch.stack(losses))}
    print(f"[engine] config1 single utterance 64 -> 512, eval={eval_mode}: " + " ".join(f"{k} {v:.2e}" for k, v in errs.items()))
    assert torch.equal(out["mel_len"].cpu(), preds[9])
    assert max(errs.values()) < REL_OUT
