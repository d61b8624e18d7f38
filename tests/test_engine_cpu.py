"""Host-logic tests (no GPU): the 4-pass engine + MAML recursion driven through the CPU restatement of
the op set (oracle/ops_reference.RefOps) must reproduce the autograd oracle (oracle/fs2_oracle.py).

This validates descriptor construction (TMA coordinate / stride semantics are emulated faithfully by
RefOps.gemm), the backward / tangent orchestration and the adjoint recursion of maml.py; the CUDA
kernels themselves are checked against the same RefOps ops in the -m gpu tests.
"""
import numpy as np
import pytest
import torch

from meta_tts_b200.maml import MamlEngine, batch_from_tuple
from oracle import fs2_oracle as O
from oracle.ops_reference import RefOps

CFG = O.small_model_config(1, 1)


def _rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def setup():
    P = O.init_params(seed=0, model_config=CFG)
    # make BN affine / LN non-trivial so every gradient path is exercised
    g = torch.Generator().manual_seed(5)
    for k, v in P.items():
        if k.endswith(("layer_norm.weight", "layer_norm_1.weight", "layer_norm_2.weight", ".1.weight")):
            v.data.add_(0.1 * torch.randn(v.shape, generator=g))
        if k.endswith(("layer_norm.bias", "layer_norm_1.bias", "layer_norm_2.bias", ".1.bias")):
            v.data.add_(0.1 * torch.randn(v.shape, generator=g))
    return P


def _engine(P, split=3):
    be = RefOps(split=split)
    m = MamlEngine(be, CFG, n_speaker=16, adapt_modules=O.ADAPT_MODULES, inner_lr=0.001, max_inner_steps=2)
    m.load_state_dict({k: v.detach().clone() for k, v in P.items()})
    return m


def test_layout_roundtrip(setup):
    m = _engine(setup)
    sd = m.state_dict()
    for k, v in setup.items():
        assert torch.equal(sd[k].to(v.dtype).reshape(v.shape), v.detach()), k
    assert m.layout.n_adapted_params == sum(setup[k].numel() for k in O.adapted_names(setup))


def test_forward_matches_oracle(setup):
    m = _engine(setup)
    b12 = O.synth_batch(2, 7, 22, seed=3, speaker=2, ragged=True)
    Pc = {k: v.detach().clone() for k, v in setup.items()}
    with torch.no_grad():
        ref = O.fs2_forward(Pc, CFG, *b12[2:])
        rl = O.fs2_loss(b12, ref)
    bt = batch_from_tuple(b12, "cpu")
    out = m.engine.forward(m.params(0), bt, m.tapes[0])
    for name, r in zip(["mel", "postnet", "pitch", "energy", "logd"], ref[:5]):
        assert _rel(out[name].reshape(r.shape), r) < 2e-5, name
    assert _rel(out["loss6"], torch.stack(rl)) < 2e-5
    assert torch.equal(out["mel_len"], ref[9])
    # BatchNorm running statistics updated as nn.BatchNorm1d does
    assert _rel(m.consts["postnet.convolutions.0.1.running_mean"], Pc["postnet.convolutions.0.1.running_mean"]) < 1e-5
    assert _rel(m.consts["postnet.convolutions.2.1.running_var"], Pc["postnet.convolutions.2.1.running_var"]) < 1e-5


@pytest.mark.parametrize("steps,first_order", [(1, True), (1, False), (2, False)])
def test_task_step_matches_oracle(setup, steps, first_order):
    m = _engine(setup)
    sup, qry = O.synth_task(task=3, shots=2, queries=2, L=6, T=18, ragged=True)
    Pc = {k: v.detach().clone() for k, v in setup.items()}
    losses, preds, grads, fast = O.maml_task_step(Pc, CFG, sup, qry, steps, 0.001, first_order, return_fast_weights=True)
    bs = batch_from_tuple(sup, "cpu")
    bq = batch_from_tuple(qry, "cpu", spk_ids=sup[2], average_spk=True)
    loss6, out = m.task_step(bs, bq, steps, first_order)
    assert _rel(loss6, torch.stack(losses)) < 2e-5
    assert _rel(out["mel"].reshape(preds[0].shape), preds[0]) < 2e-5
    fw = m.fast_weights(steps)
    worst = max(_rel(fw[k], fast[k]) for k in fast)
    assert worst < 1e-5, f"fast weights rel err {worst}"
    got = m.task_grads()
    tot_ref = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values()))
    errs = {k: ((got[k].double() - grads[k].double()).norm() / tot_ref).item() for k in grads}
    bad = {k: v for k, v in errs.items() if v > 2e-5}
    rels = {k: _rel(got[k], grads[k]) for k in grads if grads[k].norm() > 1e-6 * tot_ref}
    worst_rel = max(rels.values())
    assert not bad and worst_rel < 1e-3, (sorted(bad.items(), key=lambda kv: -kv[1])[:8],
                                          sorted(rels.items(), key=lambda kv: -kv[1])[:8])
