"""Build libmtts.so (hand-written sm_100a CUDA behind a C ABI) in-tree with nvcc.

    python meta-tts_b200/build.py            # incremental build
    python meta-tts_b200/build.py --force

The library is plain CUDA-runtime code: no torch headers, no pybind; Python binds it with ctypes
(`meta-tts_b200/lib.py`).  nvcc cross-compiles for sm_100a without a GPU present.
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libmtts.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stamp(src: str) -> str:
    h = hashlib.sha1()
    for f in [src] + sorted(
        os.path.join(CSRC, x) for x in os.listdir(CSRC) if x.endswith((".cuh", ".h"))
    ) + [os.path.join(HERE, "..", "include", "mtts.h")]:
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _compile(src: str, force: bool) -> tuple[str, bool, str]:
    os.makedirs(OBJ, exist_ok=True)
    obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
    stamp_file = obj + ".stamp"
    stamp = _stamp(src)
    if not force and os.path.exists(obj) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return obj, False, ""
    cmd = [_nvcc(), *NVCC_FLAGS, "-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    with open(stamp_file, "w") as fh:
        fh.write(stamp)
    return obj, True, r.stderr


def build_lib(force: bool = False, verbose: bool = False) -> str:
    srcs = _sources()
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(lambda s: _compile(s, force), srcs))
    objs = [r[0] for r in results]
    rebuilt = any(r[1] for r in results)
    if verbose:
        for _, did, log in results:
            if did and log:
                print(log)
    if rebuilt or not os.path.exists(LIB):
        cmd = [_nvcc(), "-shared", "-o", LIB, *objs, "-Xcompiler", "-fPIC", "-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    path = build_lib(force="--force" in sys.argv, verbose=True)
    print("built", path)
