"""The drop-in boundary on a machine without a GPU: libmtts.so loads, exports every entry point `include/mtts.h` declares, the
ctypes binding covers exactly that set, and the product refuses to run without an sm_100 device (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

from meta_tts_b200 import lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "mtts.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)                      # drop comments
    return sorted(set(re.findall(r"\b(mtts_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(L.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    names = declared_symbols()
    assert len(names) >= 40 and "mtts_gemm" in names and "mtts_stft_polar" in names and "mtts_dot" in names
    handle = ctypes.CDLL(L.LIB_PATH)
    for n in names:
        assert hasattr(handle, n), f"{n} is declared in include/mtts.h but not exported by libmtts.so"
    # the Python binding declares argument types for exactly the compute entry points of the header
    bound = set(L.SIGNATURES) | {"mtts_version", "mtts_last_error"}
    assert bound == set(names), (sorted(bound - set(names)), sorted(set(names) - bound))
    lib = L.load()
    assert lib.mtts_version() >= 100 and isinstance(lib.mtts_last_error(), bytes)


def test_every_declaration_cites_the_reference():
    """Each block of the header names the reference call site it replaces (file:line)."""
    src = open(os.path.join(ROOT, "include", "mtts.h")).read()
    cites = re.findall(r"[A-Za-z_/]+\.py:\d+", src)
    assert len(cites) >= 30
    for must in ("modules.py", "SubLayers.py", "Layers.py", "loss.py", "stft.py", "audio_processing.py", "CG_torch.py", "collate.py"):
        assert any(must in c for c in cites), must


def test_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = L.load()
    assert lib.mtts_check_device() != 0 and len(lib.mtts_last_error()) > 0
    from meta_tts_b200.ops import CudaOps
    with pytest.raises(L.MttsError):
        CudaOps(split=3)
    from meta_tts_b200 import audio
    with pytest.raises(L.MttsError):
        audio.STFT(1024, 256, 1024)
    from meta_tts_b200.systems import MetaSystem
    with pytest.raises(L.MttsError):
        MetaSystem(None, None, None, None, device="cuda:0")


def test_ctypes_structs_match_the_header_layout(tmp_path):
    """sizeof / offsetof of every descriptor struct as gcc lays out include/mtts.h == the ctypes mirror in lib.py (an ABI drift would
    otherwise only show up as wrong numbers on the GPU)."""
    import subprocess

    structs = {"mtts_operand": L.Operand, "mtts_gemm_desc": L.GemmDesc, "mtts_ln_epilogue": L.LnEpilogue, "mtts_attn_desc": L.AttnDesc}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "mtts.h"', "int main(void) {"]
    for cname, ct in structs.items():
        lines.append(f'  printf("{cname} size %zu\\n", sizeof({cname}));')
        for fname, _ in ct._fields_:
            lines.append(f'  printf("{cname} {fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    seen = 0
    for line in out.splitlines():
        cname, field, val = line.split()
        ct = structs[cname]
        want = ctypes.sizeof(ct) if field == "size" else getattr(ct, field).offset
        assert int(val) == want, f"{cname}.{field}: header {val}, ctypes {want}"
        seen += 1
    assert seen == sum(len(ct._fields_) + 1 for ct in structs.values())
