"""Generate tests/golden/imaml_golden.npz from the REAL `hypertorch/hypergrad/CG_torch.py` (container only): the conjugate-
gradient routine iMAML's hypergradient runs on (lightning/systems/utils.py:174), on a seeded SPD system given as a list of
tensors — pins `oracle.fs2_oracle.cg_solve` (incl. the early-exit quirk that drops the last update).

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
    path = os.path.join(ROOT, "tests", "golden", "imaml_golden.npz")
    np.savez_compressed(path, **out)
    print("[golden] wrote", path, "; oracle cg_solve == reference CG_torch.cg bit for bit")


if __name__ == "__main__":
    main()
