"""CPU-only probe behind the loss tolerance of tests/test_systems_gpu.py: how far does the query's duration loss of one task step move
when the ORACLE's weights are perturbed by eps (relative, Gaussian)?  A smooth function moves by ~eps; a rectifier unit of the support
passes that sits within eps of zero flips, the inner gradient changes one weight row by O(1) and the loss jumps.

    python tools/kink_probe.py > profiles/r02_kink_probe.txt
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from oracle import fs2_oracle as O  # noqa: E402

cfg = O.small_model_config(2, 2)


def run(salt, eps=None, seed=0, task=11):
    P = O.init_params(seed=0, model_config=cfg)
    if eps is not None:
        g = torch.Generator().manual_seed(seed)
        P = {k: (v * (1 + eps * torch.randn(v.shape, generator=g)) if v.is_floating_point() else v) for k, v in P.items()}
    sup, qry = O.synth_task(task=task, shots=3, queries=2, L=12, T=40)
    losses, _, _ = O.maml_task_step(P, cfg, sup, qry, 2, 0.001, first_order=True, drop_seed=(0, salt))
    return torch.stack([v.double() for v in losses])


print("task 11 of tests/test_systems_gpu.py (3 shots, 2 queries, 12 phonemes, 40 frames, K = 2, dropout on); salts of steps 1 and 3")
for salt in (0xA7689733, 0xA7689735):
    base = run(salt)
    for eps in (3e-7, 1e-6, 3e-6):
        d = [((run(salt, eps, s) - base) / base)[5].item() for s in range(5)]
        print(f"salt {salt:#x} eps {eps:g}: relative change of the duration loss over 5 perturbations: " + " ".join(f"{x:+.1e}" for x in d), flush=True)
