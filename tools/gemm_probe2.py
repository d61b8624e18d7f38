"""TMA-only pipeline depth / tensor-map rank experiment (1-CTA kernel, M=1280, N=256, bn=64)."""
import os
import sys

import torch

sys.path.insert(0, ".")
from meta_tts_b200 import lib as L  # noqa: E402
from meta_tts_b200 import ops  # noqa: E402
from tools.gemm_probe import bf, graph_time  # noqa: E402

dev = torch.device("cuda:0")
M, N = 1280, 256
for split in (3, 1):
    for label, dbg in (("no-mma s=1", 17), ("no-mma s=2", 18), ("no-mma s=3", 19), ("no-mma s=4", 20), ("no-mma s=6", 22), ("base", 0),
                       ("base 2d-map", 256), ("no-mma 2d-map", 16 + 256), ("no-mma 2d s=1", 17 + 256), ("no-mma 2d s=2", 18 + 256)):
        os.environ["MTTS_GEMM_DBG"] = str(dbg)
        row = []
        for K in (1024, 4096):
            x, xl, w, wl = bf(M, K), bf(M, K), bf(N, K), bf(N, K)
            of = torch.empty(M, N, device=dev)
            lo = (lambda t: t) if split == 3 else (lambda t: None)

            def fn():
                ops.gemm(ops.Opnd(x, lo(xl), L.MAJOR_K, (K, M), (1, K)), ops.Opnd(w, lo(wl), L.MAJOR_K, (K, N), (1, K)), M, N, K,
                         ldc=N, split=split, block_n=64, pair=False, c_f32=of)
            row.append(graph_time(fn))
        print(f"split={split} {label:16s} K=1024:{row[0]:6.1f}  4096:{row[1]:6.1f} us   per k-iter {(row[1] - row[0]) / 48:.3f} us")
os.environ["MTTS_GEMM_DBG"] = "0"
