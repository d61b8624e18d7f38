"""Generate tests/golden/imaml_golden.npz from REAL reference code (container only):

  1. `hypertorch/hypergrad/CG_torch.py: cg` — the conjugate-gradient routine iMAML's hypergradient runs on
     (lightning/systems/utils.py:174), on a seeded SPD system given as a list of tensors — pins `oracle.fs2_oracle.cg_solve`
     bit for bit (incl. the early-exit quirk that drops the last update);
  2. one whole iMAML task through the REAL `hypertorch/hypergrad/hypergradients.py: CG` (get_outer_gradients, the
     v - J^T v operator, cg, the final vector-Jacobian product w.r.t. the meta parameters) driving the REAL
     `lightning.model.FastSpeech2` / `FastSpeech2Loss` modules: the proximal inner loop and the fixed-point map of
     lightning/systems/imaml.py:51-112 are written over the reference modules with learn2learn's update restated (l2l is not
     installable offline; `lightning/systems/utils.py: CG` is the same algorithm as hypergrad's CG with the l2l wrapper around
     it) — pins `oracle.fs2_oracle.imaml_task_step` (adapted weights, query losses, hypergradient).

    python -m oracle.make_golden_imaml
"""
import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import fs2_oracle as O  # noqa: E402


def main():
    spec = importlib.util.spec_from_file_location("ref_cg", "/root/reference/hypertorch/hypergrad/CG_torch.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    g = torch.Generator().manual_seed(0)
    shapes = [(5, 3), (7,), (2, 2, 2)]
    n = sum(int(np.prod(s)) for s in shapes)
    M = torch.randn(n, n, generator=g)
    A = M @ M.t() / n + 0.5 * torch.eye(n)
    b = [torch.randn(s, generator=g) for s in shapes]

    def Ax(xs):
        v = A @ torch.cat([x.reshape(-1) for x in xs])
        out, o = [], 0
        for s in shapes:
            k = int(np.prod(s))
            out.append(v[o:o + k].reshape(s))
            o += k
        return out

    out = {"A": A.numpy(), "b": torch.cat([t.reshape(-1) for t in b]).numpy(), "shapes": np.array([list(s) + [0] * (3 - len(s)) for s in shapes])}
    for iters, eps in ((1, 1e-10), (5, 1e-10), (30, 1e-10), (30, 1e-2)):
        x = ref.cg(Ax, b, max_iter=iters, epsilon=eps)
        mine = O.cg_solve(Ax, b, iters, eps)
        assert all(torch.equal(a, c) for a, c in zip(x, mine)), (iters, eps)
        out[f"x_{iters}_{eps:g}"] = torch.cat([t.reshape(-1) for t in x]).numpy()
    out.update(imaml_task_golden())
    path = os.path.join(ROOT, "tests", "golden", "imaml_golden.npz")
    np.savez_compressed(path, **out)
    print("[golden] wrote", path, "; oracle cg_solve == reference CG_torch.cg bit for bit")


IMAML_CASE = dict(task=9, shots=4, queries=2, L=10, T=36, steps=2, lr=0.001, reg=1.0, cg_iters=3, batch_size=2, seed=1234)


def reference_imaml_task(case):
    """imaml.py:51-112 over the real modules + the real hypergrad.CG.  Returns (query losses, {name: hypergradient}, {name: w})."""
    from oracle.make_golden import build_reference_model, l2l_clone_module, reference_forward_learner
    sys.path.insert(0, "/root/reference/hypertorch")
    from hypergrad import hypergradients as HG        # the REAL package

    c = case
    model, loss_fn = build_reference_model(seed=0, model_config=O.small_model_config(1, 1))
    sup, qry = O.synth_task(task=c["task"], shots=c["shots"], queries=c["queries"], L=c["L"], T=c["T"], ragged=True)
    mods = torch.nn.ModuleDict({k: getattr(model, k) for k in O.ADAPT_MODULES})
    learner = l2l_clone_module(mods)
    learner.train()
    lm = {k: learner[k] for k in O.ADAPT_MODULES}
    names = [(mk, n) for mk in O.ADAPT_MODULES for n, p in mods[mk].named_parameters() if p.requires_grad]
    hparams = [mods[mk].get_parameter(n) for mk, n in names]            # the meta parameters (leaves of the real model)

    def set_params(tensors):                                              # l2l update_module: learner._parameters[k] = tensor
        for (mk, n), t in zip(names, tensors):
            mod = lm[mk]
            *path, leaf = n.split(".")
            for a in path:
                mod = mod._modules[a]
            mod._parameters[leaf] = t

    def reg_loss(mb, params, hps):                                        # imaml.py:72-74 / 93-97
        set_params(params)
        preds = reference_forward_learner(model, lm, *mb[2:])
        return loss_fn(mb, preds)[0] + 0.5 * c["reg"] * sum(((b - p) ** 2).sum() for b, p in zip(hps, params))

    torch.manual_seed(c["seed"])
    task = O.SupportTask(sup, c["batch_size"])                            # = lightning/systems/utils.py:78-103 `Task`
    w = [h.detach().clone() for h in hparams]
    for _ in range(c["steps"]):                                           # adapt(): first-order proximal steps
        leaves = [t.detach().requires_grad_(True) for t in w]
        g = torch.autograd.grad(reg_loss(task.next_batch(), leaves, [h.detach() for h in hparams]), leaves)
        w = [(l_ - c["lr"] * gi).detach() for l_, gi in zip(leaves, g)]
    task.reset_iterator()                                                 # imaml.py:116
    valid_box = {}

    def fp_map(params, hps):                                              # imaml.py:83-100
        g = torch.autograd.grad(reg_loss(task.next_batch(), params, hps), params, create_graph=True)
        return [p - c["lr"] * gi for p, gi in zip(params, g)]

    def outer_loss(params, hps):                                          # imaml.py:102-110
        set_params(params)
        preds = reference_forward_learner(model, lm, sup[2], *qry[3:], average_spk_emb=True)
        valid_box["losses"] = loss_fn(qry, preds)
        return valid_box["losses"][0]

    grads = HG.CG(w, hparams, K=c["cg_iters"], fp_map=fp_map, outer_loss=outer_loss, tol=1e-10, set_grad=False, stochastic=True)
    key = lambda mk, n: f"{mk}.{n}"  # noqa: E731
    return (tuple(v.detach() for v in valid_box["losses"]), {key(mk, n): g.detach() for (mk, n), g in zip(names, grads)},
            {key(mk, n): t.detach() for (mk, n), t in zip(names, w)})


def imaml_task_golden():
    c = IMAML_CASE
    losses, grads, w = reference_imaml_task(c)
    cfg = O.small_model_config(1, 1)
    P = O.init_params(seed=0, model_config=cfg)
    sup, qry = O.synth_task(task=c["task"], shots=c["shots"], queries=c["queries"], L=c["L"], T=c["T"], ragged=True)
    torch.manual_seed(c["seed"])
    o_losses, _, o_grads, o_w = O.imaml_task_step(P, cfg, sup, qry, c["steps"], c["lr"], c["reg"], c["cg_iters"], c["batch_size"], stochastic=True)
    assert sorted(o_grads) == sorted(grads)
    tot = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values()))
    err = torch.sqrt(sum(((o_grads[k].double() - grads[k].double()) ** 2).sum() for k in grads)) / tot
    werr = max(((o_w[k].double() - w[k].double()).norm() / w[k].double().norm().clamp_min(1e-30)).item() for k in w)
    lerr = max(abs(a.item() - b.item()) / abs(b.item()) for a, b in zip(o_losses, losses))
    print(f"[golden] iMAML task through the real hypergrad.CG + real modules vs oracle.imaml_task_step: hypergradient rel {err.item():.2e} "
          f"(|g| {tot.item():.4e}), adapted weights {werr:.2e}, query losses {lerr:.2e}")
    assert err < 1e-4 and werr < 1e-5 and lerr < 1e-5
    names = sorted(grads)
    return {"imaml_case": np.array([c[k] for k in ("task", "shots", "queries", "L", "T", "steps", "cg_iters", "batch_size", "seed")]),
            "imaml_lr_reg": np.array([c["lr"], c["reg"]]),
            "imaml_losses": np.array([v.item() for v in losses]),
            "imaml_names": np.array(names),
            "imaml_grad_norm": np.array([grads[k].double().norm().item() for k in names]),
            "imaml_grad_head": np.stack([np.pad(grads[k].flatten()[:8].numpy(), (0, max(0, 8 - grads[k].numel()))) for k in names]),
            "imaml_w_norm": np.array([w[k].double().norm().item() for k in names])}


if __name__ == "__main__":
    main()
