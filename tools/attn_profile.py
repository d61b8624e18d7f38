"""One launch each of the fused attention kernels (fwd, dQ, dKV) at the configs[1] decoder shape, after a warm-up pass —
the target of `ncu --set full -k regex:mtts_attn --launch-skip 3 --launch-count 3 python tools/attn_profile.py`."""
import sys

import torch

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from meta_tts_b200 import lib as L  # noqa: E402
from meta_tts_b200.ops import CudaOps, split_bf16  # noqa: E402

B, H, T, DK = 4, 2, 864, 128
emit = len(sys.argv) > 1 and sys.argv[1] == "emit"
d = H * DK
dev = torch.device("cuda:0")
be = CudaOps(split=3)
Tp, Tl = (T + 7) // 8 * 8, (T + 127) // 128 * 128
qh, ql = split_bf16(torch.randn(B * T, 3 * d, device=dev))
dh, dl = split_bf16(torch.randn(B * T, d, device=dev))
bz = lambda *s: torch.zeros(*s, dtype=torch.bfloat16, device=dev)  # noqa: E731
z = lambda *s: torch.zeros(*s, device=dev)  # noqa: E731
o_h, o_l, lse, dvec, g_h, g_l = bz(B * T, d), bz(B * T, d), z(B, H, Tl), z(B, H, Tl), bz(B * T, 3 * d), bz(B * T, 3 * d)
p_h, p_l, dP, s_h, s_l = bz(B, H, T, Tp), bz(B, H, T, Tp), z(B, H, T, Tp), bz(B, H, T, Tp), bz(B, H, T, Tp)
klens = torch.tensor([864, 800, 700, 650], dtype=torch.int64, device=dev)
for _ in range(2):
    be.attn_fwd(qh, ql, klens, B, H, T, DK, o_h, o_l, lse, *((p_h, p_l) if emit else (None, None)), Tp)
    args = (qh, ql, klens, B, H, T, DK, o_h, o_l, lse, dh, dl, dvec, g_h, g_l)
    be.attn_bwd(L.ATTN_PREP, *args)
    be.attn_bwd(L.ATTN_DQ, *args, *((dP, s_h, s_l) if emit else (None, None, None)), Tp)
    be.attn_bwd(L.ATTN_DK, *args)
    be.attn_bwd(L.ATTN_DV, *args)
    torch.cuda.synchronize()
print("ok")
