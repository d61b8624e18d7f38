"""iMAML on the B200 engine — drop-in for `lightning/systems/imaml.py: IMAMLSystem` (adapt / meta_learn /
training_step) built on the same kernels as the MAML step: the inner loop is a first-order proximal SGD on support mini-batches,
and the hypergradient comes from K conjugate-gradient iterations whose matrix-vector product `lr * (H + reg I) v` is ONE call of
the engine's forward-over-reverse Hessian-vector pass (`FS2Engine.hvp`) — no autograd graph, no per-tensor lists: every CG
vector is a flat fp32 arena and the update formulas are `axpby` / `dot` kernels.

Reference call stack restated here (file:line):
  IMAMLSystem.adapt            imaml.py:51-76     K x { mini_batch = task.next_batch(); loss + 0.5*reg*||theta0 - w||^2; SGD }
  IMAMLSystem.meta_learn       imaml.py:78-150    CG(...) -> clip by global norm (per rank) -> reduce -> Adam -> LambdaLR
  CG                           systems/utils.py:120-189, hypertorch/hypergrad/CG_torch.py:6-41
  Task                         systems/utils.py:78-116
The reference's final `update_tensor_grads(hparams, grads)` pairs ALL trainable model parameters with the gradients of the
adapted ones only (imaml.py:140-141) and so cannot run unless every module is adapted; here the adapted parameters receive
their hypergradient and the others none (see oracle/fs2_oracle.py: imaml_task_step).  The last `fp_map` evaluation of CG
(utils.py:176-181) only contributes d(proximal term)/d(theta0) = lr*reg*I, so it is folded into the closed form
`grad = lr * reg * v` and not executed (its mini-batch is still drawn, to keep the sampler in step with the reference).
"""
from __future__ import annotations

import time
from typing import Dict, Optional

import torch
from torch.utils.data import BatchSampler, RandomSampler

from .collate import split_reprocess
from .engine import ParamSet
from .maml import batch_from_tuple
from . import systems as S


class Task:
    """systems/utils.py:78-116: support mini-batches (BatchSampler over a RandomSampler, drop_last) — host-side sampling
    with torch's generator, exactly as the reference, so a seeded run draws the same mini-batches."""

    def __init__(self, sup_data, qry_data, batch_size=None, shuffle=True):
        self.sup_data, self.qry_data, self.batch_size = sup_data, qry_data, batch_size
        n = len(sup_data[0])
        self.sup_sampler = BatchSampler(RandomSampler(range(n)) if shuffle else range(n), batch_size=batch_size, drop_last=True)
        self.sup_it = iter(self.sup_sampler)

    def reset_iterator(self):
        self.sup_it = iter(self.sup_sampler)

    def next_batch(self):
        try:
            idxs = next(self.sup_it)
        except StopIteration:
            self.reset_iterator()
            idxs = next(self.sup_it)
        return split_reprocess(self.sup_data, idxs)

    def __iter__(self):
        self.reset_iterator()
        return self

    def __next__(self):
        return split_reprocess(self.sup_data, next(self.sup_it))


def _tapes_for(cache: Dict, eng, key):
    t = cache.get(key)
    if t is None:
        t = cache[key] = (eng.new_tape(), eng.new_tape())          # (primal, tangent) per mini-batch shape
    return t


def _fwd_bwd(maml, cache, b12, P, drop_pass, G):
    bt = batch_from_tuple(b12, maml.theta.device)
    tape, tape_t = _tapes_for(cache, maml.engine, (bt.B, bt.L, bt.T))
    maml.engine.forward(P, bt, tape, drop_pass=drop_pass)
    maml.engine.backward(P, maml.grads(G), bt, tape, 1.0, into_encoder=False)
    return bt, tape, tape_t


def imaml_adapt(maml, task: Task, steps: int, reg_param: float, drop_base: Optional[int] = None, tape_cache: Optional[Dict] = None):
    """imaml.py:51-76: `steps` first-order SGD steps on support mini-batches of L(w) + 0.5*reg*||theta0 - w||^2, the weights
    rolling in fast-weight arena 0 (no history is kept: the implicit hypergradient does not need the trajectory)."""
    be, lay = maml.be, maml.layout
    a0, lr = lay.adapt_begin, maml.lr
    cache = tape_cache if tape_cache is not None else {}
    w = maml.fast[0]
    theta0 = maml.theta[a0:]
    dp = lambda k: None if drop_base is None else drop_base + k  # noqa: E731
    fwd_bwd = lambda b12, P, drop_pass, G: _fwd_bwd(maml, cache, b12, P, drop_pass, G)  # noqa: E731
    g_ad = maml.g_inner[a0:]
    for s in range(steps):
        P = maml.params(0) if s == 0 else maml.params(1)
        be.zero_(g_ad)
        fwd_bwd(task.next_batch(), P, dp(s), maml.g_inner)
        if s > 0:                                            # + reg * (w - theta0)   (zero at the first step)
            be.axpby(reg_param, w[0], 1.0, g_ad)
            be.axpby(-reg_param, theta0, 1.0, g_ad)
        be.sgd_split(theta0 if s == 0 else w[0], g_ad, lr, w[0], w[1], w[2])
    maml.bn_batches += steps


def imaml_hypergradient(maml, task: Task, sup12, qry12, steps: int, reg_param: float, cg_iters: int, stochastic: bool,
                        drop_base: Optional[int] = None, cg_eps: float = 1e-10, tape_cache: Optional[Dict] = None):
    """imaml.py:95-121 + systems/utils.py:120-189 after `imaml_adapt`: query loss at the adapted weights and the CG
    hypergradient.  Leaves the hypergradient in `maml.g_task` (adapted region = lr*reg*v, the rest zero).
    Returns (query loss6 tensor, query prediction dict, query Batch)."""
    be, eng, lay = maml.be, maml.engine, maml.layout
    a0, lr = lay.adapt_begin, maml.lr
    dev = maml.theta.device
    cache = tape_cache if tape_cache is not None else {}
    dp = lambda k: None if drop_base is None else drop_base + k  # noqa: E731
    fwd_bwd = lambda b12, P, drop_pass, G: _fwd_bwd(maml, cache, b12, P, drop_pass, G)  # noqa: E731
    g_ad = maml.g_inner[a0:]
    task.reset_iterator()                                    # imaml.py:116
    PW = maml.params(1) if steps > 0 else maml.params(0)
    # ---- outer loss at w and b = dLq/dw (utils.py:158-159) ----
    bq = batch_from_tuple(qry12, dev, spk_ids=sup12[2], average_spk=True)
    tq, _ = _tapes_for(cache, eng, ("q", bq.B, bq.L, bq.T))
    out = eng.forward(PW, bq, tq, drop_pass=dp(steps))
    maml.bn_batches += 1
    be.zero_(maml.g_task)
    eng.backward(PW, maml.grads(maml.g_task), bq, tq, 1.0, into_encoder=False)
    # ---- CG (CG_torch.py:6-41) on flat arenas: x = 0, r = p = b ----
    n_ad = lay.n_adapt
    if not hasattr(maml, "_cg"):
        maml._cg = [be.zeros((n_ad,)) for _ in range(3)] + [be.zeros((2048,)) for _ in range(3)]     # scalar results: [0] value, [1..] partials
    x, r, p, s_rr, s_pap, s_new = maml._cg
    be.zero_(x)
    r.copy_(maml.g_task[a0:])
    p.copy_(r)
    hv_ad = maml.hv[a0:]
    for j in range(cg_iters):
        b12 = task.next_batch() if stochastic else sup12
        be.zero_(g_ad)
        bt, tape, tape_t = fwd_bwd(b12, PW, dp(steps + 1 + j), maml.g_inner)       # primal pass the HVP differentiates
        maml.bn_batches += 1
        be.split_(p, maml.lam_hi, maml.lam_lo)
        Pd = ParamSet(lay, None, None, None, p, maml.lam_hi, maml.lam_lo, only_adapted=True)
        be.zero_(maml.hv)
        eng.hvp(PW, Pd, maml.grads(maml.hv), bt, tape, tape_t)
        be.axpby(reg_param, p, 1.0, hv_ad)                   # (H + reg I) p
        be.axpby(0.0, p, lr, hv_ad)                          # A p = lr * (H + reg I) p      (v - J_fp^T v, utils.py:163-171)
        be.dot(r, r, s_rr)
        be.dot(p, hv_ad, s_pap)
        rTr, pAp = float(s_rr[0]), float(s_pap[0])                 # the reference syncs here too (float(torch.norm(r_vec)))
        alpha = rTr / pAp
        be.axpby(-alpha, hv_ad, 1.0, r)                      # r <- r - alpha A p
        be.dot(r, r, s_new)
        rr_new = float(s_new[0])
        if rr_new ** 0.5 < cg_eps:
            break                                            # x_last is returned WITHOUT this iteration's update
        be.axpby(alpha, p, 1.0, x)                           # x <- x + alpha p
        be.axpby(1.0, r, rr_new / rTr, p)                    # p <- r + beta p
    if stochastic:
        task.next_batch()                                    # the reference's last fp_map draw (utils.py:176-178)
    # ---- hypergradient: lr * reg * v on the adapted region, nothing elsewhere ----
    be.zero_(maml.g_task)
    be.axpby(lr * reg_param, x, 1.0, maml.g_task[a0:])
    return out["loss6"], out, bq


class IMAMLSystem(S.MetaSystem):
    """Drop-in for `lightning.systems.imaml.IMAMLSystem` (hot path): manual optimisation — `training_step` -> `meta_learn`
    computes the hypergradient, clips it by its global norm on this rank, averages over ranks and steps Adam + LambdaLR
    itself (imaml.py:114-147).  `algorithm_config["adapt"]["imaml"]` = {batch_size, reg_param, K, stochastic}."""

    automatic_optimization = False

    def __init__(self, *args, **kwargs):
        kwargs.setdefault("use_cuda_graph", False)           # mini-batch shapes vary and CG has data-dependent control flow
        super().__init__(*args, **kwargs)
        self._imaml_tapes: Dict = {}

    def adapt(self, batch, adaptation_steps=5, learner=None, task=None, train=True):
        """imaml.py:51-76 -> (number of inner steps held in fast-weight arena 0, task).  (The reference returns the adapted
        learner object; here the engine owns the weights.)"""
        S._assert_meta_batch(batch)
        cfg = self.algorithm_config["adapt"]["imaml"]
        sup12, qry12 = batch[0][0][0], batch[0][1][0]
        if task is None:
            task = Task(sup12, qry12, batch_size=cfg["batch_size"])
        if self.dropout:
            S._set_salt(self, self.next_salt())
        imaml_adapt(self.maml, task, adaptation_steps, float(cfg["reg_param"]), 0 if self.dropout else None, self._imaml_tapes)
        return adaptation_steps, task

    def meta_learn(self, batch, batch_idx, train=True):
        S._assert_meta_batch(batch)
        cfg = self.algorithm_config["adapt"]["imaml"]
        sup12, qry12 = batch[0][0][0], batch[0][1][0]
        m = self.maml
        _, task = self.adapt(batch, self.adaptation_steps, train=train)
        drop_base = 0 if self.dropout else None
        loss6, out, bq = imaml_hypergradient(m, task, sup12, qry12, self.adaptation_steps, float(cfg["reg_param"]),
                                             int(cfg["K"]) if train else 0, bool(cfg["stochastic"]), drop_base,
                                             tape_cache=self._imaml_tapes)
        if train:
            t0 = time.perf_counter()
            be = m.be
            opt = self.train_config["optimizer"]
            max_norm = float(opt.get("grad_clip_thresh", 1.0))
            dist = torch.distributed.is_initialized()
            world = torch.distributed.get_world_size(self.process_group) if dist else 1
            if world > 1:
                # clip on this rank (imaml.py:123-129), then mean-reduce (imaml.py:130), then a plain Adam step
                be.sumsq(m.g_task, m.sumsq)
                coef = min(1.0, max_norm / (float(m.sumsq[0]) ** 0.5 + 1e-6))
                be.zero_(m.g_outer)
                be.axpby(coef / world, m.g_task, 1.0, m.g_outer)
                torch.distributed.all_reduce(m.g_outer, group=self.process_group)
                clip = 0.0
            else:
                m.g_outer.copy_(m.g_task)
                clip = max_norm                              # same formula: g * min(1, max_norm / (norm + 1e-6))
            m.outer_update(1.0, clip, tuple(opt["betas"]), float(opt["eps"]), warmup=int(opt.get("warm_up_step", 4000)),
                           anneal_steps=tuple(opt.get("anneal_steps", ())), anneal_rate=float(opt.get("anneal_rate", 0.3)))
            be.zero_(m.g_outer)
            self.host_prof["optimizer"] += time.perf_counter() - t0
        losses = tuple(loss6.clone()[i] for i in range(6))
        return losses, S._pred10(out, bq, False)

    def training_step(self, batch, batch_idx):
        """imaml.py:152-172: manual optimisation — the optimizer already stepped inside meta_learn."""
        train_loss, predictions = self.meta_learn(batch, batch_idx, train=True)
        return {"loss": train_loss[0], "losses": train_loss, "output": predictions, "_batch": batch[0][1][0]}

    def validation_step(self, batch, batch_idx):
        val_loss, predictions = self.meta_learn(batch, batch_idx, train=False)
        return {"losses": val_loss, "output": predictions, "_batch": batch[0][1][0]}

    def optimizer_step(self):
        raise RuntimeError("IMAMLSystem optimises manually inside training_step (automatic_optimization = False, imaml.py:27)")
