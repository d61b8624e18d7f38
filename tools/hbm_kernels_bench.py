"""Roofline of the HBM-bound kernels (GPU tool): each kernel is launched back to back on ROTATING buffer sets whose total size
exceeds 2x the 126 MB L2, so every launch streams cold data; achieved GB/s = algorithmic bytes (stated per kernel below) / average
launch time (CUDA events around CUDA-graph replays of the launch sequence), against the measured copy bandwidth in MEASURED_PEAKS.json.  Sizes are those of BASELINE configs[1]
(4 utterances x 864 frames, 256 / 512 / 1024 channels; 34.65 M parameters).  Writes a markdown table (profiles/ material)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from meta_tts_b200.ops import CudaOps  # noqa: E402

DEV = "cuda:0"
L2 = 126e6


def peak_gbs():
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (torch copy, read+write bytes)"
    return 6500.0, "fallback (B200_PROFILING.md)"


def run(name, bytes_per_launch, make, launch, note, rows):
    nset = max(2, int(2.2 * L2 / max(bytes_per_launch, 1)) + 1)
    nset = min(nset, 64)
    sets = [make() for _ in range(nset)]
    for s in sets[:2]:
        launch(s)
    torch.cuda.synchronize()
    # the launches are replayed from a CUDA graph: eager ctypes launches cost ~8 us of host time each, more than the small kernels run
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for s in sets:
            launch(s)
    graph.replay()
    torch.cuda.synchronize()
    reps = max(3, 200 // nset)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / (reps * nset)
    rows.append((name, bytes_per_launch / 1e6, us, bytes_per_launch / (us * 1e-6) / 1e9, nset, note))


def main():
    be = CudaOps(split=3, device=DEV)
    f = lambda *s: torch.randn(*s, device=DEV)  # noqa: E731
    z = lambda *s: torch.zeros(*s, device=DEV)  # noqa: E731
    h = lambda *s: torch.zeros(*s, device=DEV, dtype=torch.bfloat16)  # noqa: E731
    rows = []
    B, T, Lp = 4, 864, 128
    R = B * T
    lens = torch.full((B,), T, dtype=torch.int64, device=DEV)
    for C in (256,):
        n = R * C
        run(f"ln_fwd (residual + LayerNorm + split), {R}x{C}", n * 20,
            lambda: (f(R, C), f(R, C), f(C), f(C), z(R, C), z(R, 2), z(R, C), h(R, C), h(R, C)),
            lambda s: be.ln_fwd(s[0], s[1], s[2], s[3], lens, T, R, C, s[4], s[5], s[6], s[7], s[8]),
            "read y, res (8 B/elt); write z, out, hi, lo (12 B/elt)", rows)
        run(f"ln_bwd (+ dgamma, dbeta, dbias), {R}x{C}", n * 16,
            lambda: (f(R, C), f(R, C), torch.cat([z(R, 1), torch.ones(R, 1, device=DEV)], 1).contiguous(), f(C), z(R, C), h(R, C), h(R, C), z(C), z(C), z(C)),
            lambda s: be.ln_bwd(s[0], s[1], s[2], s[3], lens, T, R, C, False, s[4], s[5], s[6], s[7], s[8], s[9]),
            "read dy, z (8 B/elt); write dz, hi, lo (8 B/elt)", rows)
    Tp = 896
    nz = B * 2
    run(f"softmax fwd, {nz} x {T} rows of {T} keys (ld {Tp})", nz * T * Tp * 8,
        lambda: (f(nz, T, Tp), h(nz, T, Tp), h(nz, T, Tp)),
        lambda s: be.softmax(0, s[0], None, None, None, None, None, lens, nz, 2, T, T, Tp, s[1], s[2]),
        "read S fp32 (4 B/elt); write P hi, lo (4 B/elt)", rows)
    run(f"softmax tangent fwd (mode 1), {nz} x {T} rows of {T} keys (ld {Tp})", nz * T * Tp * 12,
        lambda: (f(nz, T, Tp), h(nz, T, Tp), h(nz, T, Tp), h(nz, T, Tp), h(nz, T, Tp)),
        lambda s: be.softmax(1, s[0], None, s[1], s[2], None, None, lens, nz, 2, T, T, Tp, s[3], s[4]),
        "read dS fp32, P hi, lo (8 B/elt); write hi, lo (4 B/elt)", rows)
    run(f"softmax tangent bwd (mode 2), {nz} x {T} rows of {T} keys (ld {Tp})", nz * T * Tp * 20,
        lambda: (f(nz, T, Tp), f(nz, T, Tp), h(nz, T, Tp), h(nz, T, Tp), h(nz, T, Tp), h(nz, T, Tp), h(nz, T, Tp), h(nz, T, Tp)),
        lambda s: be.softmax(2, s[0], s[1], s[2], s[3], s[4], s[5], lens, nz, 2, T, T, Tp, s[6], s[7]),
        "read A, B fp32, P hi, lo, Pd hi, lo (16 B/elt); write hi, lo (4 B/elt)", rows)
    C = 512
    run(f"bn_fwd (batch stats + tanh + split), {R}x{C}", R * C * 12,
        lambda: (f(R, C), f(C), f(C), z(C), torch.ones(C, device=DEV), z(4 * 512), z(2 * C), z(R, C), h(R, C), h(R, C)),
        lambda s: be.bn_fwd(s[0], s[1], s[2], R, C, True, s[3], s[4], s[5], s[6], s[7], s[8], s[9]),
        "algorithmic: read x once (4 B/elt), write out, hi, lo (8 B/elt); the kernel reads x 3x (two-pass statistics + apply)", rows)
    d = 256
    dur = torch.full((B, Lp), 0, dtype=torch.int64, device=DEV)
    dur[:, :] = 6
    dur[:, :96] = 7                                                 # 96*7 + 32*6 = 864
    idx, _ = be.lr_index(dur, T)
    run(f"length_regulate_fwd, {B}x{Lp} -> {T} frames x {d}", B * (Lp + T) * d * 4,
        lambda: (f(B, Lp, d), z(B, T, d)), lambda s: be.lr_fwd(s[0], idx, s[1]), "read x [B,L,C], write out [B,T,C]", rows)
    run(f"length_regulate_bwd (segment sum), {B}x{T} -> {Lp}", B * (Lp + T) * d * 4,
        lambda: (f(B, T, d), z(B, Lp, d)), lambda s: be.lr_bwd(s[0], dur, Lp, s[1]), "read dy [B,T,C], write dx [B,L,C]", rows)
    run(f"colsum (bias gradient), {R}x1024", R * 1024 * 4,
        lambda: (f(R, 1024), z(1024)), lambda s: be.colsum(s[0], None, None, 1, R, 1024, s[1]), "read dy once", rows)
    n = 23010368
    run("sgd_split (inner SGD + operand split), 23.0 M adapted parameters", n * 16,
        lambda: (f(n), f(n), z(n), h(n), h(n)), lambda s: be.sgd_split(s[0], s[1], 0.001, s[2], s[3], s[4]),
        "read theta, g (8 B); write theta', hi, lo (8 B)", rows)
    n = 34650432
    hyper = torch.tensor([1e-3, 0.1, 0.02, 0.0], device=DEV)
    ss = torch.ones(1, device=DEV)
    run("adam_clip (clip + Adam + split), 34.65 M parameters", n * 32,
        lambda: (f(n), f(n), z(n), z(n), h(n), h(n)),
        lambda s: be.adam_clip(s[0], s[1], s[2], s[3], ss, 1.0, 1.0, hyper, 0.9, 0.98, 1e-9, s[4], s[5]),
        "read p, g, m, v (16 B); write p, m, v, hi, lo (16 B)", rows)
    run("axpby (G <- G - lr*HV), 34.65 M", n * 12, lambda: (f(n), f(n)), lambda s: be.axpby(-1e-3, s[0], 1.0, s[1]),
        "read x, y; write y", rows)
    out1 = z(2048)
    run("sumsq (global gradient norm), 34.65 M", n * 4, lambda: (f(n),), lambda s: be.sumsq(s[0], out1), "read x", rows)
    # vocoder side
    F_, NP, io = 864, 1040, 520
    run(f"stft_recombine (angles of the transform -> [mag cos | mag sin] hi/lo), {F_} frames", F_ * (io * 4 + NP * 4 + NP * 4),
        lambda: (f(F_, io).abs(), f(F_, NP), h(F_, NP), h(F_, NP)),
        lambda s: be.stft_recombine(s[0], None, s[1], F_, 513, NP, io, io, s[2], s[3]), "read mag, ri; write X hi, lo", rows)
    peak, src = peak_gbs()
    lines = ["| kernel | algorithmic MB / launch | us / launch (cold L2) | achieved GB/s | % of measured HBM peak | buffer sets | bytes counted |",
             "|---|---|---|---|---|---|---|"]
    for name, mb, us, gbs, nset, note in rows:
        lines.append(f"| `{name}` | {mb:.1f} | {us:.1f} | {gbs:.0f} | {100 * gbs / peak:.0f} % | {nset} | {note} |")
    print(f"HBM peak used: {peak:.0f} GB/s ({src})\n")
    print("\n".join(lines))
    json.dump([{"kernel": r[0], "mb": r[1], "us": r[2], "gbs": r[3], "frac": r[3] / peak} for r in rows],
              open(os.path.join("gpurun_out", "hbm_kernels.json"), "w"))


if __name__ == "__main__":
    os.makedirs("gpurun_out", exist_ok=True)
    main()
