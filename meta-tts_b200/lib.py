"""ctypes binding of libmtts.so (include/mtts.h).  Fails loudly: there is no CPU fallback.

The library is built in-tree by `meta-tts_b200/build.py` (nvcc, sm_100a).  Importing this module
does not need a GPU; calling any compute entry point without one raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MTTS_LIB_PATH") or os.path.join(_HERE, "libmtts.so")     # override: instrumented debug builds

# enums (include/mtts.h)
SRC_ZERO, SRC_Z0, SRC_Z1, SRC_TAP, SRC_KB = 0, 1, 2, 3, 4
MAJOR_K, MAJOR_MN = 0, 1
EPI_RELU, EPI_ACCUM, EPI_GATE, EPI_BIAS_ROW, EPI_ADD_C = 1, 2, 4, 8, 16
UN_LOGCLAMP, UN_EXP, UN_SCALE = 0, 1, 2


class Operand(C.Structure):
    _fields_ = [
        ("hi", C.c_void_p),
        ("lo", C.c_void_p),
        ("major", C.c_int32),
        ("src2", C.c_int32),
        ("src3", C.c_int32),
        ("shift_src", C.c_int32),
        ("shift_base", C.c_int32),
        ("shift_step", C.c_int32),
        ("reserved", C.c_int32),
        ("dims", C.c_int64 * 4),
        ("strides", C.c_int64 * 4),
    ]


class GemmDesc(C.Structure):
    _fields_ = [
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("ntaps", C.c_int32), ("nkb", C.c_int32), ("nz0", C.c_int32), ("nz1", C.c_int32),
        ("split", C.c_int32), ("block_n", C.c_int32), ("ksplit", C.c_int32), ("flags", C.c_int32),
        ("alpha", C.c_float),
        ("a", Operand), ("b", Operand),
        ("c_f32", C.c_void_p), ("c_hi", C.c_void_p), ("c_lo", C.c_void_p),
        ("ldc", C.c_int64), ("c_sz0", C.c_int64), ("c_sz1", C.c_int64),
        ("bias", C.c_void_p), ("bias_sz0", C.c_int64),
        ("gate", C.c_void_p),
        ("a2_hi", C.c_void_p), ("a2_lo", C.c_void_p), ("b2_hi", C.c_void_p), ("b2_lo", C.c_void_p),
        ("pair", C.c_int32), ("reserved2", C.c_int32),
    ]


class LnEpilogue(C.Structure):
    """mtts_ln_epilogue (include/mtts.h): dropout -> + residual -> LayerNorm -> pad-row zeroing as the epilogue of mtts_gemm_ln."""
    _fields_ = [
        ("res", C.c_void_p), ("gamma", C.c_void_p), ("beta", C.c_void_p), ("lens", C.c_void_p),
        ("T", C.c_int32), ("eps", C.c_float),
        ("z_out", C.c_void_p), ("stats", C.c_void_p),
        ("drop_thr", C.c_uint32), ("drop_seed", C.c_uint32), ("drop_scale", C.c_float), ("reserved", C.c_int32),
        ("drop_salt", C.c_void_p),
    ]


ATTN_PREP, ATTN_DQ, ATTN_DK, ATTN_DV = 1, 2, 4, 8
ATTN_DKV = ATTN_DK | ATTN_DV


class AttnDesc(C.Structure):
    """mtts_attn_desc (include/mtts.h)."""
    _fields_ = [
        ("B", C.c_int32), ("H", C.c_int32), ("T", C.c_int32), ("dk", C.c_int32),
        ("Tp", C.c_int32), ("Tl", C.c_int32), ("split", C.c_int32), ("scale", C.c_float),
        ("qkv_hi", C.c_void_p), ("qkv_lo", C.c_void_p),
        ("klens", C.c_void_p),
        ("o_hi", C.c_void_p), ("o_lo", C.c_void_p),
        ("lse", C.c_void_p),
        ("p_hi", C.c_void_p), ("p_lo", C.c_void_p),
        ("do_hi", C.c_void_p), ("do_lo", C.c_void_p),
        ("dvec", C.c_void_p),
        ("dqkv_hi", C.c_void_p), ("dqkv_lo", C.c_void_p),
        ("dp", C.c_void_p), ("ds_hi", C.c_void_p), ("ds_lo", C.c_void_p),
    ]


class MttsError(RuntimeError):
    pass


_lib = None


def load() -> C.CDLL:
    """Load libmtts.so; raise if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MttsError(
            f"{LIB_PATH} is missing: build it with `python meta-tts_b200/build.py` "
            "(or __graft_entry__.build()).  There is no CPU / PyTorch fallback."
        )
    lib = C.CDLL(LIB_PATH)
    lib.mtts_last_error.restype = C.c_char_p
    lib.mtts_version.restype = C.c_int
    _declare(lib)
    _lib = lib
    return lib


# name -> argtypes; every function returns int (0 = ok)
_vp, _i, _i64, _f = C.c_void_p, C.c_int, C.c_int64, C.c_float
_u32 = C.c_uint32
_D2 = [_u32, _u32, _f, _u32, _u32, _f, _vp]      # (thr, seed, scale) x {pre, post}, salt*
_D1 = [_u32, _u32, _f, _vp]
SIGNATURES: dict[str, list] = {
    "mtts_check_device": [],
    "mtts_set_pdl": [_i],
    "mtts_set_deterministic": [_i],
    "mtts_zero": [_vp, _i64, _vp],
    "mtts_gemm": [C.POINTER(GemmDesc), _vp],
    "mtts_gemm_ln": [C.POINTER(GemmDesc), C.POINTER(LnEpilogue), _vp],
    "mtts_attn_fwd": [C.POINTER(AttnDesc), _vp],
    "mtts_attn_bwd": [C.POINTER(AttnDesc), _i, _vp],
    "mtts_pack_rows": [_vp, _vp, _i, _i, _i, _vp, _vp],
    "mtts_length_regulate_index": [_vp, _vp, _i, _i, _i, _vp, _vp, _vp],
    "mtts_length_regulate_fwd": [_vp, _vp, _i, _i, _i, _i, _vp, _vp],
    "mtts_length_regulate_bwd": [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp],
    "mtts_ln_fwd": [_vp, _vp, _vp, _vp, _vp, _i, _i64, _i, _f, _vp, _vp, _vp, _vp, _vp] + _D2 + [_vp],
    "mtts_ln_bwd": [_vp, _vp, _vp, _vp, _vp, _i, _i64, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp] + _D2 + [_vp],
    "mtts_ln_tfwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i64, _i, _vp, _vp, _vp, _vp] + _D2 + [_vp],
    "mtts_ln_tbwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i64, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp] + _D2 + [_vp],
    "mtts_rowdot_fwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i64, _i, _vp, _vp],
    "mtts_rowdot_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i64, _i, _vp, _vp, _vp, _vp],
    "mtts_softmax": [_i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp],
    "mtts_embed_fwd": [_vp, _vp, _vp, _vp, _i, _i64, _i, _vp, _vp, _vp, _vp],
    "mtts_embed_bwd": [_vp, _vp, _i64, _i, _i64, _f, _vp, _vp],
    "mtts_bucketize": [_vp, _vp, _i, _i64, _vp, _vp],
    "mtts_add_rowvec": [_vp, _vp, _i64, _vp, _i, _i, _i, _vp, _vp, _vp, _vp],
    "mtts_spk_embed": [_vp, _vp, _i, _i, _i, _i, _vp, _vp],
    "mtts_spk_embed_bwd": [_vp, _vp, _i, _i, _i, _i, _f, _vp, _vp],
    "mtts_colsum": [_vp, _vp, _vp, _i, _i64, _i, _vp, _vp],
    "mtts_bn_fwd": [_vp, _vp, _vp, _i64, _i, _f, _f, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp] + _D1 + [_vp],
    "mtts_bn_bwd": [_vp, _vp, _vp, _vp, _vp, _i64, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp] + _D1 + [_vp],
    "mtts_bn_tfwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i, _i, _vp, _vp, _vp, _vp, _vp] + _D1 + [_vp],
    "mtts_bn_tbwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp] + _D1 + [_vp],
    "mtts_loss_fwd": [_vp] * 11 + [_i, _i, _i, _i, _vp, _vp, _vp, _vp],
    "mtts_loss_bwd": [_vp] * 11 + [_i, _i, _i, _i, _vp, _f, _i, _vp, _vp, _vp, _vp, _vp, _vp],
    "mtts_split": [_vp, _vp, _vp, _i64, _vp],
    "mtts_sgd_split": [_vp, _vp, _f, _vp, _vp, _vp, _i64, _vp],
    "mtts_axpby": [_f, _vp, _f, _vp, _i64, _vp],
    "mtts_sumsq": [_vp, _i64, _vp, _vp],
    "mtts_dot": [_vp, _vp, _i64, _vp, _vp],
    "mtts_duration_round": [_vp, _f, _i64, _vp, _vp],
    "mtts_bn_eval": [_vp, _vp, _vp, _vp, _vp, _i64, _i, _f, _i, _vp, _vp, _vp, _vp],
    "mtts_unary": [_i, _vp, _i64, _f, _f, _vp, _vp, _vp, _vp],
    "mtts_reflect_pad": [_vp, _i, _i64, _i, _i64, _vp, _vp, _vp, _vp],
    "mtts_stft_polar": [_vp, _i64, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp],
    "mtts_stft_recombine": [_vp, _vp, _vp, _i64, _i, _i, _i, _i, _vp, _vp, _vp],
    "mtts_istft_finish": [_vp, _vp, _f, _f, _i, _i64, _i, _vp, _vp],
    "mtts_adam_clip": [_vp, _vp, _vp, _vp, _vp, _f, _f, _vp, _f, _f, _f, _vp, _vp, _i64, _vp],
}


def _declare(lib: C.CDLL) -> None:
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError => header / library mismatch: fail loudly
        fn.argtypes = argtypes
        fn.restype = C.c_int


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().mtts_last_error().decode(errors="replace")
        raise MttsError(f"libmtts {what} failed (rc={rc}): {msg}")


def call(name: str, *args) -> None:
    lib = load()
    check(getattr(lib, name)(*args), name)
