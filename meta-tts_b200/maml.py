"""MAML / FOMAML task step and outer update without an autograd graph.

Reference path: `MetaSystem.training_step` (lightning/systems/meta.py:68-80) ->
`BaseAdaptorSystem.meta_learn/adapt` (lightning/systems/base_adaptor.py:98-124) with learn2learn's
`MAML.clone/adapt` (clone_module, maml_update: p <- p - lr*g) and Lightning's backward / DDP
allreduce / clip / Adam (main.py:57-64, lightning/optimizer.py:6-16, lightning/scheduler.py:6-29).

Here the K inner steps write fast weights theta_1..theta_K into preallocated flat arenas (fused
SGD + bf16 operand split, one kernel), every pass keeps its activations in its own Tape (180 GB of
HBM: nothing is recomputed), and the outer gradient is obtained by the adjoint recursion

    lambda_K = dLq/dtheta (theta_K),   gphi = dLq/dphi
    for k = K-1 .. 0:   [Hv_theta ; Hv_phi] = HVP_k(lambda_{k+1})        (engine.hvp: forward-over-reverse)
                        lambda_k = lambda_{k+1} - lr*Hv_theta ;  gphi -= lr*Hv_phi

which on the flat arena is the single update  G <- G - lr*HV.  FOMAML skips the recursion.
The whole task step is a fixed launch sequence on static buffers => captured once in a CUDA graph.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch

from .engine import Batch, FS2Engine, ParamLayout, ParamSet, Tape, const_names


def batch_from_tuple(b12, device, spk_ids=None, average_spk=False, targets: bool = True, max_T: Optional[int] = None) -> Batch:
    """Reference 12-tuple (lightning/collate.py:47-60) -> device Batch (plumbing: H2D copies).
    targets=False keeps only what free-running synthesis reads (`*qry_batch[3:6]`, base_adaptor.py:161).
    max_T = max_seq_len for TRAIN-mode teacher-forced batches: the reference decoder keeps the first max_seq_len frames
    (Models.py:161-166) and the loss crops the mel targets to the truncated mask (loss.py:42-43), which is the same as running
    the whole path on the batch cropped to max_T frames (durations / mel_lens stay as they are: the length regulator and every
    mask clip to T)."""
    (_, _, spk, texts, src_lens, max_src, mels, mel_lens, max_mel, pitches, energies, durs) = b12
    to = lambda t, dt: torch.as_tensor(t).to(device=device, dtype=dt).contiguous()  # noqa: E731
    if targets and max_T is not None and int(max_mel) > max_T:
        mels, max_mel = torch.as_tensor(mels)[:, :max_T], max_T
    if not targets:
        return Batch(spk_ids=to(spk if spk_ids is None else spk_ids, torch.int64), average_spk=average_spk,
                     texts=to(texts, torch.int64), src_lens=to(src_lens, torch.int64), mels=None, mel_lens=None, pitches=None,
                     energies=None, durations=None, B=int(texts.shape[0]), L=int(max_src), T=0)
    return Batch(spk_ids=to(spk if spk_ids is None else spk_ids, torch.int64), average_spk=average_spk,
                 texts=to(texts, torch.int64), src_lens=to(src_lens, torch.int64), mels=to(mels, torch.float32),
                 mel_lens=to(mel_lens, torch.int64), pitches=to(pitches, torch.float32),
                 energies=to(energies, torch.float32), durations=to(durs, torch.int64),
                 B=int(texts.shape[0]), L=int(max_src), T=int(max_mel))


class MamlEngine:
    """Owns the parameter / gradient arenas and runs task steps on a backend."""

    def __init__(self, be, cfg, n_speaker: int, adapt_modules: Sequence[str], inner_lr: float = 0.001,
                 max_inner_steps: int = 5):
        self.be = be
        self.cfg = cfg
        self.lr = float(inner_lr)
        self.K_max = max_inner_steps
        self.layout = ParamLayout(cfg, n_speaker, adapt_modules)
        lay = self.layout
        n, na = lay.total, lay.n_adapt
        bf = torch.bfloat16
        z = be.zeros
        self.theta = z((n,))
        self.theta_hi = z((n,), bf)
        self.theta_lo = z((n,), bf) if be.split == 3 else None
        self.fast = [(z((na,)), z((na,), bf), z((na,), bf) if be.split == 3 else None) for _ in range(self.K_max)]
        self.g_inner = z((n,))           # inner-step gradient (adapted region used)
        # The 6 query losses ride on the TAIL of the gradient arenas (8 extra floats): the task's accumulation `axpby` and the
        # ONE NCCL allreduce of the outer gradient then also produce the mean losses of all tasks of the step on every rank —
        # meta.py:77-79 `self.log_dict(..., sync_dist=True)` without a second collective.
        self.g_task_full = z((n + 8,))
        self.g_task = self.g_task_full[:n]       # this task's outer gradient (lambda | gphi)
        self.g_outer_full = z((n + 8,))
        self.g_outer = self.g_outer_full[:n]     # accumulated over tasks (allreduce buffer)
        self.hv = z((n,))
        self.g_enc = z((lay.adapt_begin,))   # query-pass gradient of the (non-adapted) prefix, accumulated on the 'enc' branch
        self.lam_hi = z((na,), bf)
        self.lam_lo = z((na,), bf) if be.split == 3 else None
        self.consts: Dict[str, torch.Tensor] = {}
        self.bn_batches = 0
        self.engine: Optional[FS2Engine] = None
        self.tapes: List[Tape] = []
        self.tape_q: Optional[Tape] = None
        self.tape_t: Optional[Tape] = None
        # Adam state (outer optimiser)
        self.adam_m = z((n,))
        self.adam_v = z((n,))
        self.sumsq = z((2048,))          # [0] = |g|^2, [1..] per-CTA partials of the fixed-order reduction (MTTS_SCALAR_WS)
        self.hyper = z((4,))
        self.opt_step = 0

    # ---- parameters -------------------------------------------------------------------------------
    def load_state_dict(self, sd) -> None:
        """Reference state_dict (same keys) -> flat arenas.  Host-side plumbing."""
        lay, be = self.layout, self.be
        flat = torch.zeros(lay.total, dtype=torch.float32)
        lay.pack(sd, flat)
        self.theta.copy_(flat)
        be.split_(self.theta, self.theta_hi, self.theta_lo)
        for name in const_names(self.cfg):
            if name.endswith("num_batches_tracked"):
                self.bn_batches = int(sd[name])
                continue
            t = sd[name].detach().to(torch.float32)
            if name.endswith("position_enc"):
                t = t[0]
            self.consts[name] = t.contiguous().to(self.theta.device)
        self.engine = FS2Engine(be, self.cfg, lay, self.consts)
        self.use_tapes(self.new_tapes())

    def new_tapes(self):
        """(support tapes x K_max, query tape, tangent tape): one set per input-shape signature."""
        tq = self.engine.new_tape()
        tq.t["loss6:f"] = self.g_task_full[self.layout.total:self.layout.total + 6]     # the query losses land on the arena tail
        return ([self.engine.new_tape() for _ in range(self.K_max)], tq, self.engine.new_tape())

    def use_tapes(self, tapes) -> None:
        self.tapes, self.tape_q, self.tape_t = tapes

    def state_dict(self) -> Dict[str, torch.Tensor]:
        sd = self.layout.unpack(self.theta.detach().cpu())
        for name, t in self.consts.items():
            sd[name] = (t[None] if name.endswith("position_enc") else t).detach().cpu().clone()
        for i in range(5):
            sd[f"postnet.convolutions.{i}.1.num_batches_tracked"] = torch.tensor(self.bn_batches)
        return sd

    def params(self, k: int = 0) -> ParamSet:
        """Parameter views at inner step k (k = 0: the meta parameters)."""
        if k == 0:
            return ParamSet(self.layout, self.theta, self.theta_hi, self.theta_lo)
        f = self.fast[k - 1]
        return ParamSet(self.layout, self.theta, self.theta_hi, self.theta_lo, f[0], f[1], f[2])

    def grads(self, flat: torch.Tensor) -> ParamSet:
        return ParamSet(self.layout, flat)

    def fast_weights(self, k: int) -> Dict[str, torch.Tensor]:
        """Adapted parameters after k inner steps, in reference layout (for tests / export)."""
        full = self.theta.detach().clone()
        if k > 0:
            full[self.layout.adapt_begin:] = self.fast[k - 1][0]
        sd = self.layout.unpack(full.cpu())
        return {n: sd[n] for n, e in self.layout.entries.items() if e.adapted}

    # ---- one task ---------------------------------------------------------------------------------
    def adapt(self, sup: Batch, steps: int, start: int = 0, drop_base: Optional[int] = None, second_order: bool = True) -> None:
        """Inner loop (base_adaptor.py:98-112): fast weights theta_{start+1..start+steps}.
        drop_base: None = dropout off; else support pass k uses dropout pass index drop_base + k (learner.train(),
        base_adaptor.py:103)."""
        be, eng, lay = self.be, self.engine, self.layout
        a0 = lay.adapt_begin
        for k in range(start, start + steps):
            P = self.params(k)
            self.tapes[k].attn_emit = second_order       # only a Hessian-vector pass re-reads the attention probabilities
            eng.forward(P, sup, self.tapes[k], drop_pass=None if drop_base is None else drop_base + k)
            g_ad = self.g_inner[a0:]
            be.zero_(g_ad)
            eng.backward(P, self.grads(self.g_inner), sup, self.tapes[k], 1.0, into_encoder=False)
            src = self.theta[a0:] if k == 0 else self.fast[k - 1][0]
            dst = self.fast[k]
            be.sgd_split(src, g_ad, self.lr, dst[0], dst[1], dst[2])      # l2l maml_update fused with operand prep
        self.bn_batches += steps

    def adapt_rolling(self, sup: Batch, steps: int, tape: Tape, fresh: bool, drop_base: Optional[int] = None) -> None:
        """FIRST-ORDER inner steps that keep no history (test-time adaptation, base_adaptor.py:98-112 with train=False ->
        first_order=True, called repeatedly with `learner=learner`, base_adaptor.py:172-174): the fast weights live in arena 0
        and are updated in place, one activation tape is reused for every step.  fresh = start from the meta parameters
        (the reference's `self.learner.clone()`), else continue from the current fast weights."""
        be, eng, lay = self.be, self.engine, self.layout
        a0 = lay.adapt_begin
        dst = self.fast[0]
        tape.attn_emit = False
        for s in range(steps):
            first = fresh and s == 0
            P = self.params(0) if first else self.params(1)
            eng.forward(P, sup, tape, drop_pass=None if drop_base is None else drop_base + s)
            g_ad = self.g_inner[a0:]
            be.zero_(g_ad)
            eng.backward(P, self.grads(self.g_inner), sup, tape, 1.0, into_encoder=False)
            be.sgd_split(self.theta[a0:] if first else dst[0], g_ad, self.lr, dst[0], dst[1], dst[2])
        self.bn_batches += steps

    def predict(self, bt: Batch, adapted: bool, free_running: bool = False, eval_mode: bool = False,
                drop_pass: Optional[int] = None, p_control: float = 1.0, e_control: float = 1.0, d_control: float = 1.0):
        """One forward_learner call outside the training step (base_adaptor.py:160-186): meta parameters (adapted=False) or
        the rolling fast weights of `adapt_rolling`; teacher forced (+ loss) or free running; eval or train mode.
        Shapes differ call to call, so the activations go to a throw-away tape."""
        eng = self.engine
        P = self.params(int(adapted))                          # False / 0: meta parameters; True / k: fast weights after k held steps
        tape = eng.new_tape()
        tape.attn_emit = False
        if free_running:
            out = eng.synthesize(P, bt, tape, p_control, e_control, d_control, update_bn=not eval_mode, drop_pass=drop_pass,
                                 eval_mode=eval_mode)
        else:
            out = eng.forward(P, bt, tape, update_bn=not eval_mode, drop_pass=drop_pass, eval_mode=eval_mode)
        if not eval_mode:
            self.bn_batches += 1
        return out

    def task_step(self, sup: Batch, qry: Batch, steps: int, first_order: bool, accumulate_scale: Optional[float] = None,
                  drop_base: Optional[int] = None, reduce=None):
        """meta_learn (base_adaptor.py:114-124) + the task's outer gradient into self.g_task.
        Returns the query loss 6-vector tensor (device) and the query predictions dict."""
        assert steps <= self.K_max
        be, eng, lay = self.be, self.engine, self.layout
        a0 = lay.adapt_begin
        dq = None if drop_base is None else drop_base + steps
        # the encoder is not adapted: the query's encoder pass does not depend on the inner loop -> 'enc' branch
        self.tape_q.attn_emit = False
        xq = eng.encoder_early(self.params(0), qry, self.tape_q, drop_pass=dq)
        self.adapt(sup, steps, drop_base=drop_base, second_order=not first_order)
        PK = self.params(steps)
        out = eng.forward(PK, qry, self.tape_q, drop_pass=dq, enc=xq)
        self.bn_batches += 1
        be.zero_(self.g_task)
        overlap_enc = (not first_order) and steps > 0 and not lay.entries["encoder.src_word_emb.weight"].adapted
        if overlap_enc:
            # the query's encoder backward overlaps the Hessian-vector passes; it accumulates into its own arena
            # (the adjoint recursion below rewrites all of g_task) and is folded in after the join
            be.zero_(self.g_enc)
        eng.backward(PK, self.grads(self.g_task), qry, self.tape_q, 1.0, into_encoder=True,
                     enc_G=self.grads(self.g_enc) if overlap_enc else None)
        # reduce(tensor): the data-parallel sum over ranks (the step's ONE exchange, issued in two pieces).  The adapted region of the
        # outer gradient is final before the last Hessian-vector pass walks back through the encoder, so its 2/3 of the buffer is
        # reduced on a 'comm' branch UNDER that walk (and under the query-encoder backward on the 'enc' branch); only the encoder
        # third is reduced at the end.  Same arithmetic: every element is accumulated once and reduced once.
        early = reduce is not None and accumulate_scale is not None and not first_order and steps > 0 and a0 > 0
        split_done = [False]

        def reduce_adapted():
            be.axpby(-self.lr, self.hv[a0:], 1.0, self.g_task[a0:])
            be.axpby(accumulate_scale, self.g_task_full[a0:], 1.0, self.g_outer_full[a0:])     # adapted gradient + the 6 losses
            with be.branch("comm"):
                reduce(self.g_outer_full[a0:])
            split_done[0] = True

        if not first_order:
            for k in range(steps - 1, -1, -1):
                be.split_(self.g_task[a0:], self.lam_hi, self.lam_lo)
                Pd = ParamSet(lay, None, None, None, self.g_task[a0:], self.lam_hi, self.lam_lo, only_adapted=True)
                be.zero_(self.hv)
                last = early and k == 0
                eng.hvp(self.params(k), Pd, self.grads(self.hv), sup, self.tapes[k], self.tape_t, before_encoder=reduce_adapted if last else None)
                if last:
                    be.axpby(-self.lr, self.hv[:a0], 1.0, self.g_task[:a0])      # the encoder third (the adapted part went in reduce_adapted)
                else:
                    be.axpby(-self.lr, self.hv, 1.0, self.g_task)          # lambda_k | gphi update in one pass
        if overlap_enc:
            be.join("enc")
            be.axpby(1.0, self.g_enc, 1.0, self.g_task[:a0])
        if split_done[0]:
            be.axpby(accumulate_scale, self.g_task[:a0], 1.0, self.g_outer[:a0])
            reduce(self.g_outer[:a0])
            be.join("comm")
        elif accumulate_scale is not None:
            be.axpby(accumulate_scale, self.g_task_full, 1.0, self.g_outer_full)      # gradient + the 6 losses on the tail
            if reduce is not None:
                reduce(self.g_outer_full)
        return out["loss6"], out

    def task_grads(self) -> Dict[str, torch.Tensor]:
        return self.layout.unpack(self.g_task.detach().cpu())

    # ---- outer update ------------------------------------------------------------------------------
    def lr_schedule(self, step: int, warmup: int = 4000, anneal_steps=(300000, 400000, 500000), anneal_rate=0.3) -> float:
        """lightning/optimizer.py:7 (init lr = d_model^-0.5) x lightning/scheduler.py:11-23."""
        cur = step + 1
        lr = min(cur ** -0.5, warmup ** -1.5 * cur)
        for s in anneal_steps:
            if cur > s:
                lr *= anneal_rate
        return self.cfg["transformer"]["encoder_hidden"] ** -0.5 * lr

    def outer_update(self, gscale: float = 1.0, max_norm: float = 1.0, betas=(0.9, 0.98), eps: float = 1e-9,
                     warmup: int = 4000, anneal_steps=(300000, 400000, 500000), anneal_rate: float = 0.3) -> None:
        """clip_grad_norm_(1.0) + Adam + LambdaLR on the flat arena (one norm pass + one update pass),
        also refreshing the bf16 operand copies of theta."""
        be = self.be
        t = self.opt_step + 1
        if not hasattr(self, "_hyper_ring"):
            pin = self.hyper.is_cuda
            self._hyper_ring = [torch.zeros(4).pin_memory() if pin else torch.zeros(4) for _ in range(8)]
            self._hyper_events = [None] * len(self._hyper_ring)
        slot = self.opt_step % len(self._hyper_ring)
        hb = self._hyper_ring[slot]
        if self._hyper_events[slot] is not None:
            self._hyper_events[slot].synchronize()      # the H2D copy that last read this pinned slot has executed
        hb[0] = self.lr_schedule(self.opt_step, warmup, anneal_steps, anneal_rate)
        hb[1], hb[2] = 1 - betas[0] ** t, 1 - betas[1] ** t
        self.hyper.copy_(hb, non_blocking=True)
        if self.hyper.is_cuda:
            ev = torch.cuda.Event()
            ev.record()
            self._hyper_events[slot] = ev
        be.sumsq(self.g_outer, self.sumsq)
        be.adam_clip(self.theta, self.g_outer, self.adam_m, self.adam_v, self.sumsq, gscale, max_norm, self.hyper,
                     betas[0], betas[1], eps, self.theta_hi, self.theta_lo)
        self.opt_step += 1

    def memory_bytes(self) -> int:
        tapes = sum(t.nbytes() for t in self.tapes + [self.tape_q, self.tape_t, self.engine.scr])
        arenas = sum(t.numel() * t.element_size() for t in
                     [self.theta, self.theta_hi, self.g_inner, self.g_task, self.g_outer, self.hv, self.adam_m, self.adam_v])
        return tapes + arenas
