"""Where does a k-iteration's time go?  Times mtts_gemm (N=256) for K = 256 / 1024 / 4096 under the MTTS_GEMM_DBG
diagnostics (stage cap, MMA off, TMA off, no stores), for 108-CTA and 40-CTA grids (is the TMA-only rate a chip-wide
L2 limit or a per-SM one?)."""
import os
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


N = 256
def main():
    for M in (3456, 1280, 256):
        for split, bn, pair in ((3, 64, False), (3, 256, True), (1, 64, False)):
            for label, dbg in ((("base", 0), ("no-mma", 16), ("no-stores", 128)) if pair else (("base", 0), ("no-mma", 16), ("no-tma", 32), ("no-mma,no-tma", 48), ("no-stores", 128))):
                os.environ["MTTS_GEMM_DBG"] = str(dbg)
                row = []
                for K in (256, 1024, 4096):
                    x, xl, w, wl = bf(M, K), bf(M, K), bf(N, K), bf(N, K)
                    of = torch.empty(M, N, device=dev)
                    lo = (lambda t: t) if split == 3 else (lambda t: None)
    
                    def fn():
                        ops.gemm(ops.Opnd(x, lo(xl), L.MAJOR_K, (K, M), (1, K)), ops.Opnd(w, lo(wl), L.MAJOR_K, (K, N), (1, K)), M, N, K,
                                 ldc=N, split=split, block_n=bn, pair=pair, c_f32=of)
                    row.append(graph_time(fn))
                per = (row[2] - row[1]) / 48
                print(f"M={M:5d} split={split} bn={bn:3d} pair={int(pair)} {label:14s} K=256:{row[0]:6.1f}  1024:{row[1]:6.1f}  4096:{row[2]:6.1f} us"
                      f"   per k-iter {per:.3f} us")
    os.environ["MTTS_GEMM_DBG"] = "0"


if __name__ == "__main__":
    main()
