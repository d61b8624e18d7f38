"""Fixed overhead of one mtts_gemm launch inside a CUDA graph: time vs contraction length K for a small
output (3456 x 256) — intercept = prologue + epilogue + launch gap; slope = per-k-iteration cost."""
import sys

import torch

sys.path.insert(0, ".")
from meta_tts_b200 import lib as L  # noqa: E402
from meta_tts_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")


def bf(*s):
    return torch.randn(*s, device=dev).to(torch.bfloat16)


def graph_time(fn, n=40):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (5 * n) * 1e3


M = 3456
for split in (3, 1):
    for N, outs in ((256, "f32"), (256, "hilo"), (768, "hilo"), (1024, "hilo")):
        for bn, pair in ((64, False), (128, False), (128, True), (256, True)):
            if bn > N:
                continue
            row = []
            for K in (64, 256, 1024, 4096):
                x, xl, w, wl = bf(M, K), bf(M, K), bf(N, K), bf(N, K)
                of = torch.empty(M, N, device=dev)
                oh, ol = torch.empty(M, N, device=dev, dtype=torch.bfloat16), torch.empty(M, N, device=dev, dtype=torch.bfloat16)
                lo = (lambda t: t) if split == 3 else (lambda t: None)
                kw = dict(c_f32=of) if outs == "f32" else dict(c_hi=oh, c_lo=lo(ol))

                def fn():
                    ops.gemm(ops.Opnd(x, lo(xl), L.MAJOR_K, (K, M), (1, K)), ops.Opnd(w, lo(wl), L.MAJOR_K, (K, N), (1, K)), M, N, K,
                             ldc=N, split=split, block_n=bn, pair=pair, **kw)
                row.append(graph_time(fn))
            print(f"split={split} N={N:4d} out={outs:4s} bn={bn:3d} pair={int(pair)}  K=64:{row[0]:6.1f}  256:{row[1]:6.1f}  1024:{row[2]:6.1f}  4096:{row[3]:6.1f} us")
# a trivial kernel in the same graph setting
t = torch.zeros(1024, device=dev)
be = ops.CudaOps(3)
print("axpby(1024 elts) in-graph us:", graph_time(lambda: be.axpby(1.0, t, 1.0, t)))
