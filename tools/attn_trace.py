"""Pipeline trace of one CTA of the fused attention kernels (SM-clock timestamps written by a -DMTTS_ATTN_TRACE build):

    python tools/attn_trace.py build      # nvcc -DMTTS_ATTN_TRACE -> meta-tts_b200/libmtts_trace.so   (no GPU needed)
    MTTS_LIB_PATH=meta-tts_b200/libmtts_trace.so python tools/attn_trace.py [fwd|dq|dk|dv]

kinds: 0 ring-1 slot free (producer), 1 ring-2 slot free, 2 MMA thread starts score step, 3 its tiles are in smem, 4 score MMAs issued,
5 MMA thread starts waiting for the softmax of the step, 6 softmax done, 7 accumulate MMAs issued, 8 softmax warp starts waiting for
the scores, 9 scores ready, 10 softmax warp done, 11 misc (0 resident copy done, 1 accumulator ready, 2 CTA end, 3 CTA start)."""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if len(sys.argv) > 1 and sys.argv[1] == "build":
    csrc = os.path.join(ROOT, "meta-tts_b200", "csrc")
    srcs = sorted(os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith(".cu"))
    out = os.path.join(ROOT, "meta-tts_b200", "libmtts_trace.so")
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
           "--expt-relaxed-constexpr", "-DMTTS_ATTN_TRACE", "-shared", "-cudart", "static", "-o", out] + srcs
    subprocess.check_call(cmd)
    print("built", out)
    sys.exit(0)

import torch  # noqa: E402

from meta_tts_b200 import lib as L  # noqa: E402
from meta_tts_b200.ops import CudaOps, split_bf16  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "fwd"
B, H, T, DK = 4, 2, 864, 128
d = H * DK
dev = torch.device("cuda:0")
be = CudaOps(split=3)
Tp, Tl = (T + 7) // 8 * 8, (T + 127) // 128 * 128
qh, ql = split_bf16(torch.randn(B * T, 3 * d, device=dev))
dh, dl = split_bf16(torch.randn(B * T, d, device=dev))
bz = lambda *s: torch.zeros(*s, dtype=torch.bfloat16, device=dev)  # noqa: E731
z = lambda *s: torch.zeros(*s, device=dev)  # noqa: E731
o_h, o_l, lse, dvec, g_h, g_l = bz(B * T, d), bz(B * T, d), z(B, H, Tl), z(B, H, Tl), bz(B * T, 3 * d), bz(B * T, 3 * d)
klens = torch.full((B,), T, dtype=torch.int64, device=dev)
args = (qh, ql, klens, B, H, T, DK, o_h, o_l, lse, dh, dl, dvec, g_h, g_l)
for _ in range(3):
    be.attn_fwd(qh, ql, klens, B, H, T, DK, o_h, o_l, lse)
    be.attn_bwd(L.ATTN_PREP, *args)
    torch.cuda.synchronize()
    if mode != "fwd":
        be.attn_bwd({"dq": L.ATTN_DQ, "dk": L.ATTN_DK, "dv": L.ATTN_DV}[mode], *args)
    torch.cuda.synchronize()
buf = (ctypes.c_longlong * (16 * 128))()
assert L.load().mtts_attn_trace_read(buf) == 0
t = [[buf[k * 128 + i] for i in range(128)] for k in range(16)]
t0 = t[11][3]
us = lambda v: (v - t0) / 1965.0 if v else float("nan")  # noqa: E731
print(f"mode {mode}: CTA start 0, resident copy done {us(t[11][0]):.2f} us, accumulator ready {us(t[11][1]):.2f} us, CTA end {us(t[11][2]):.2f} us")
print("step | r1 free | r2 free | sc start | tiles in | sc issued | wait p | p done | acc issued || softmax: wait | scores | done")
n = max(i for i in range(128) if t[2][i]) + 1
for i in range(n):
    print(f"{i:4d} | " + " | ".join(f"{us(t[k][i]):7.2f}" for k in (0, 1, 2, 3, 4, 5, 6, 7)) + " || " + " | ".join(f"{us(t[k][i]):7.2f}" for k in (8, 9, 10)))
