"""Per-GEMM-class precision budget: one BASELINE configs[1]-sized second-order task step on the GPU with ONE class
of tensor-core products switched from bf16x3 (hi*hi + hi*lo + lo*hi) to single-pass bf16, against the fp32 autograd oracle:
error of the outputs (the north_star bar: 1e-3 relative) and of the outer gradient (total and median per tensor).

classes: p.* = forward / backward passes, t.* = Hessian-vector (tangent) passes; fwd = Linear / Conv forward products, dgrad = data
gradients, wgrad = weight gradients, bmm = attention products (the fused attention kernels in the p passes).

    python tools/precision_budget.py            # prints a markdown table
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from meta_tts_b200.maml import MamlEngine, batch_from_tuple  # noqa: E402
from meta_tts_b200.ops import CudaOps  # noqa: E402
from oracle import fs2_oracle as O  # noqa: E402

dev = torch.device("cuda:0")
torch.cuda.set_device(0)
cfg = O.BASE_MODEL_CONFIG
P = O.init_params(seed=0)
small = len(sys.argv) > 1 and sys.argv[1] == "small"
sup, qry = O.synth_task(task=0, shots=4, queries=4, L=128 if not small else 32, T=864 if not small else 200)
print("# oracle (CPU fp32 autograd, second order K=1) ...", file=sys.stderr)
losses, preds, grads = O.maml_task_step({k: v.detach().clone() for k, v in P.items()}, cfg, sup, qry, 1, 0.001, False)
tot_ref = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values()))
be = CudaOps(split=3, device="cuda:0")
m = MamlEngine(be, cfg, n_speaker=16, adapt_modules=O.ADAPT_MODULES, inner_lr=0.001, max_inner_steps=1)
bs = batch_from_tuple(sup, dev)
bq = batch_from_tuple(qry, dev, spk_ids=sup[2], average_spk=True)
rel = lambda a, b: ((a.double().cpu() - b.double()).norm() / b.double().norm()).item()  # noqa: E731
CLS = ["p.fwd", "p.dgrad", "p.wgrad", "p.bmm", "t.fwd", "t.dgrad", "t.wgrad", "t.bmm"]
variants = [("all bf16x3 (default)", {})] + [(f"{c} = 1", {c: 1}) for c in CLS] + [
    ("p.wgrad + t.wgrad = 1", {"p.wgrad": 1, "t.wgrad": 1}),
    ("all t.* = 1 (single-pass HVP)", {c: 1 for c in CLS if c.startswith("t.")}),
    ("t.* + p.wgrad = 1", {**{c: 1 for c in CLS if c.startswith("t.")}, "p.wgrad": 1}),
    ("backward only: p.dgrad + p.wgrad + t.* = 1", {**{c: 1 for c in CLS if c.startswith("t.")}, "p.wgrad": 1, "p.dgrad": 1}),
    ("everything = 1 (plain bf16)", {c: 1 for c in CLS}),
    ("POLICY: t.fwd + t.bmm + t.wgrad + p.wgrad = 1", {"t.fwd": 1, "t.bmm": 1, "t.wgrad": 1, "p.wgrad": 1}),
    ("t.fwd + t.bmm + t.wgrad = 1", {"t.fwd": 1, "t.bmm": 1, "t.wgrad": 1}),
    ("t.fwd + t.bmm = 1", {"t.fwd": 1, "t.bmm": 1}),
]
if len(sys.argv) > 1 and sys.argv[-1] == "policy":
    variants = [variants[0]] + variants[-3:]
print("| single-pass bf16 classes | max output rel err | loss rel err | outer grad rel err (total) | median per-tensor grad rel err | p90 |")
print("|---|---|---|---|---|---|")
SALT = 20260925
losses_d = preds_d = grads_d = None
variants += [("(dropout ON) all bf16x3", {}, True), ("(dropout ON) POLICY", {"t.fwd": 1, "t.bmm": 1, "t.wgrad": 1, "p.wgrad": 1}, True)]
for var in variants:
    name, pol, drop = var if len(var) == 3 else (var[0], var[1], False)
    if drop and grads_d is None:
        losses_d, preds_d, grads_d = O.maml_task_step({k: v.detach().clone() for k, v in P.items()}, cfg, sup, qry, 1, 0.001, False,
                                                      drop_seed=(0, SALT))
    if drop:
        losses, preds, grads = losses_d, preds_d, grads_d
        tot_ref = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values()))
        be.drop_salt = torch.tensor([SALT], dtype=torch.int32, device=dev)
    m.load_state_dict({k: v.detach().clone() for k, v in P.items()})
    m.engine.g.policy = dict(pol)
    loss6, out = m.task_step(bs, bq, 1, False, drop_base=0 if drop else None)
    torch.cuda.synchronize()
    r_out = max(rel(out["mel"].reshape(preds[0].shape), preds[0]), rel(out["postnet"].reshape(preds[1].shape), preds[1]),
                rel(out["pitch"], preds[2]), rel(out["energy"], preds[3]), rel(out["logd"], preds[4]))
    r_loss = rel(loss6, torch.stack(losses))
    got = m.task_grads()
    r_grad = (torch.sqrt(sum(((got[k].double() - grads[k].double()) ** 2).sum() for k in grads)) / tot_ref).item()
    per = sorted(((got[k].double() - grads[k].double()).norm() / grads[k].double().norm()).item() for k in grads
                 if grads[k].double().norm() > 1e-4 * tot_ref)
    print(f"| {name} | {r_out:.1e} | {r_loss:.1e} | {r_grad:.1e} | {per[len(per) // 2]:.1e} | {per[int(0.9 * len(per))]:.1e} |", flush=True)
