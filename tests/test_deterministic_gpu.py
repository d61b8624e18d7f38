"""Deterministic mode (mtts_set_deterministic / MTTS_DETERMINISTIC=1, the counterpart of the reference's
Trainer(deterministic=True), main.py:35): no split-K, one contributing CTA per reduced element, ordered embedding scatter.
The whole second-order task step is then bit-reproducible run to run, and still matches the oracle."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from meta_tts_b200 import lib as L  # noqa: E402
from meta_tts_b200.maml import MamlEngine, batch_from_tuple  # noqa: E402
from meta_tts_b200.ops import CudaOps  # noqa: E402
from oracle import fs2_oracle as O  # noqa: E402


@pytest.fixture
def deterministic():
    L.call("mtts_set_deterministic", 1)
    yield
    L.call("mtts_set_deterministic", 0)


def _run(m, P, bs, bq, steps, first_order, salt):
    m.load_state_dict({k: v.detach().clone() for k, v in P.items()})
    if salt is not None:
        m.be.drop_salt = torch.tensor([salt], dtype=torch.int32, device=m.theta.device)
    loss6, out = m.task_step(bs, bq, steps, first_order, drop_base=None if salt is None else 0)
    torch.cuda.synchronize()
    return loss6.clone(), out["mel"].clone(), m.g_task.clone(), [f[0].clone() for f in m.fast[:steps]]


@pytest.mark.parametrize("shape,steps,first_order,salt", [
    ((4, 4, 128, 864), 1, False, None),        # BASELINE configs[1], full size
    ((4, 4, 128, 864), 1, False, 20260925),    # ... with dropout ON (what bench.py times)
    ((5, 5, 16, 64), 5, False, 77),            # configs[2] structure: 5 inner steps, 5 Hessian-vector passes
    ((5, 5, 16, 64), 5, True, None),           # configs[3] structure (first order)
])
def test_task_step_is_bit_reproducible(cuda_device, deterministic, shape, steps, first_order, salt):
    S, Q, Lp, T = shape
    cfg = O.BASE_MODEL_CONFIG
    P = O.init_params(seed=0)
    be = CudaOps(split=3)
    m = MamlEngine(be, cfg, n_speaker=16, adapt_modules=O.ADAPT_MODULES, inner_lr=0.001, max_inner_steps=steps)
    sup, qry = O.synth_task(task=0, shots=S, queries=Q, L=Lp, T=T, ragged=(T < 864))
    dev = m.theta.device
    bs, bq = batch_from_tuple(sup, dev), batch_from_tuple(qry, dev, spk_ids=sup[2], average_spk=True)
    runs = [_run(m, P, bs, bq, steps, first_order, salt) for _ in range(3)]
    for r in runs[1:]:
        assert torch.equal(r[0], runs[0][0]), "losses differ between runs"
        assert torch.equal(r[1], runs[0][1]), "mel differs between runs"
        assert torch.equal(r[2], runs[0][2]), "outer gradient differs between runs"
        for a, b in zip(r[3], runs[0][3]):
            assert torch.equal(a, b), "fast weights differ between runs"


def test_deterministic_mode_matches_oracle(cuda_device, deterministic):
    """Same tolerances as the default path (tests/test_engine_gpu.py::test_config2_full_size_parity): outer gradient 1e-3."""
    cfg = O.BASE_MODEL_CONFIG
    P = O.init_params(seed=0)
    be = CudaOps(split=3)
    m = MamlEngine(be, cfg, n_speaker=16, adapt_modules=O.ADAPT_MODULES, inner_lr=0.001, max_inner_steps=1)
    sup, qry = O.synth_task(task=0, shots=4, queries=4, L=128, T=864)
    losses, preds, grads = O.maml_task_step({k: v.detach().clone() for k, v in P.items()}, cfg, sup, qry, 1, 0.001, False)
    dev = m.theta.device
    loss6, out, g, _ = _run(m, P, batch_from_tuple(sup, dev), batch_from_tuple(qry, dev, spk_ids=sup[2], average_spk=True), 1, False, None)
    got = m.task_grads()
    tot = torch.sqrt(sum((v.double() ** 2).sum() for v in grads.values()))
    err = (torch.sqrt(sum(((got[k].double() - grads[k].double()) ** 2).sum() for k in grads)) / tot).item()
    rel = lambda a, b: ((a.double().cpu() - b.double()).norm() / b.double().norm()).item()  # noqa: E731
    print(f"[deterministic] loss {rel(loss6, torch.stack(losses)):.2e} mel {rel(out.reshape(preds[0].shape), preds[0]):.2e} grad {err:.2e}")
    assert rel(loss6, torch.stack(losses)) < 1e-3 and rel(out.reshape(preds[0].shape), preds[0]) < 1e-3 and err < 1e-3
