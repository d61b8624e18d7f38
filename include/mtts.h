/*
 * mtts.h — C ABI of libmtts.so, the sm_100a kernel library behind the B200-native
 * Meta-TTS meta-training step.
 *
 * The reference (SungFeng-Huang/Meta-TTS) has NO FFI / plugin interface: its boundary for the
 * hot path is the Python nn.Module API (transformer/Models.py:73,139; lightning/model/
 * fastspeech2.py:40; lightning/systems/base_adaptor.py:41-124).  Each entry point below therefore
 * cites the reference torch call site(s) it replaces; the Python host in `meta-tts_b200/`
 * (ctypes) re-creates the reference classes on top of these calls.
 *
 * Conventions:
 *   - plain pointers + sizes, no torch types; every pointer is a DEVICE pointer unless noted;
 *   - the caller owns all memory (activations, workspaces, flat weight / gradient arenas);
 *   - every op is asynchronous on the `stream` argument (a cudaStream_t passed as void*),
 *     performs no host sync and is CUDA-graph capturable;
 *   - return 0 on success, negative MTTS_E* otherwise; mtts_last_error() gives the message;
 *   - bf16 tensors are raw uint16 storage; "hi/lo" pairs are the 2-term bf16 split of an fp32
 *     value (x ≈ hi + lo) used by the 3-pass (bf16x3) tensor-core mode.
 */
#ifndef MTTS_H_
#define MTTS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MTTS_VERSION 100

#define MTTS_OK          0
#define MTTS_EINVAL     -1   /* bad shape / alignment / argument          */
#define MTTS_ECUDA      -2   /* CUDA runtime / driver error               */
#define MTTS_EARCH      -3   /* device is not sm_100                      */
#define MTTS_EUNSUPPORTED -4

typedef void* mtts_stream;   /* cudaStream_t */

int         mtts_version(void);
const char* mtts_last_error(void);
/* 0 when a usable sm_100 device is current, MTTS_EARCH otherwise. */
int         mtts_check_device(void);

/* ------------------------------------------------------------------------------------------
 * Generic tcgen05 GEMM:   for every z = (z0, z1):
 *     C_z[M,N] (+)= alpha * sum_{tap < ntaps} sum_{kb < nkb}  A(z,tap,kb)[M,K] * B(z,tap,kb)[N,K]^T
 * operands are bf16 (optionally hi/lo split => 3 MMAs per k-step, fp32-grade result), the
 * accumulator is fp32 in TMEM, operand tiles are fetched with TMA (4-D tensor maps, 128B swizzle,
 * out-of-bounds = 0 which implements conv zero padding and all ragged edges).
 *
 * Replaces (reference call sites): nn.Linear  w_qs/w_ks/w_vs/fc  SubLayers.py:39-41,54;
 * torch.bmm attention Modules.py:16,23; nn.Conv1d k=9/k=1 SubLayers.py:88; Conv k=3
 * modules.py:291-296; mel_linear fastspeech2.py:97; PostNet Conv1d k=5 Layers.py:129-137;
 * and the autograd backward (dgrad / wgrad) of each of them.
 * ------------------------------------------------------------------------------------------ */

enum { MTTS_SRC_ZERO = 0, MTTS_SRC_Z0 = 1, MTTS_SRC_Z1 = 2, MTTS_SRC_TAP = 3, MTTS_SRC_KB = 4 };
enum { MTTS_MAJOR_K = 0, MTTS_MAJOR_MN = 1 };

typedef struct {
  const void* hi;          /* bf16 */
  const void* lo;          /* bf16 or NULL (required when split == 3) */
  int32_t  major;          /* MTTS_MAJOR_K : dim0 = contraction index (contiguous), dim1 = M/N rows
                              MTTS_MAJOR_MN: dim0 = M/N index (contiguous),        dim1 = contraction rows */
  int32_t  src2, src3;     /* loop variable feeding TMA coordinate 2 / 3 (MTTS_SRC_*)            */
  int32_t  shift_src;      /* MTTS_SRC_ZERO | MTTS_SRC_TAP | MTTS_SRC_Z0: row (dim1) shift =
                              shift_base + shift_step * var ; rows outside [0, dims[1]) read 0   */
  int32_t  shift_base, shift_step;
  int32_t  reserved;
  int64_t  dims[4];        /* extents, dims[0] contiguous                                        */
  int64_t  strides[4];     /* element strides (strides[0] must be 1; others multiples of 8)      */
} mtts_operand;

enum {
  MTTS_EPI_RELU     = 1,   /* v = max(v, 0)                                                      */
  MTTS_EPI_ACCUM    = 2,   /* c_f32 += v  (red.global.add.f32; required when ksplit > 1)          */
  MTTS_EPI_GATE     = 4,   /* v = gate[m,n] > 0 ? v : 0   (ReLU backward; gate has C geometry)     */
  MTTS_EPI_BIAS_ROW = 8    /* bias indexed by row m instead of column n                          */
};

typedef struct {
  int32_t M, N, K;         /* per (z, tap, kb) problem size; K = contraction length               */
  int32_t ntaps, nkb, nz0, nz1;
  int32_t split;           /* 1: bf16 single pass;  3: bf16x3 (hi*hi + hi*lo + lo*hi)              */
  int32_t block_n;         /* 0 = auto, else 64 / 128 / 256                                        */
  int32_t ksplit;          /* >= 1: CTAs sharing one output tile along the contraction (ACCUM)     */
  int32_t flags;           /* MTTS_EPI_*                                                           */
  float   alpha;
  mtts_operand a, b;
  float*    c_f32;         /* any of the three outputs may be NULL                                 */
  void*     c_hi;          /* bf16                                                                 */
  void*     c_lo;          /* bf16                                                                 */
  int64_t   ldc, c_sz0, c_sz1;   /* element strides: row, z0, z1 (shared by c_f32/c_hi/c_lo/gate)   */
  const float* bias;       /* [N] (or [M] with MTTS_EPI_BIAS_ROW) or NULL; added after alpha        */
  int64_t   bias_sz0;      /* bias stride per z0 (0 = shared)                                       */
  const void* gate;        /* bf16, C geometry, or NULL                                             */
} mtts_gemm_desc;

int mtts_gemm(const mtts_gemm_desc* d, mtts_stream stream);

/* ------------------------------------------------------------------------------------------
 * LengthRegulator  (reference: lightning/model/modules.py:167-194 + utils/tools.py:304-322)
 *   idx[b,t]  = #{ j : cumsum(max(d[b,:],0))[j] <= t }   (searchsorted right)   -- integer, bit-exact
 *   out[b,t,:] = t < mel_len[b] ? x[b, idx[b,t], :] : 0 ;   mel_len[b] = sum_j max(d[b,j],0)
 * Durations are int64 (targets) — the float (predicted) variant takes max(int(d),0) as the
 * reference does (`max(int(expand_size), 0)`, modules.py:186-187).
 * ------------------------------------------------------------------------------------------ */
int mtts_length_regulate_index(const int64_t* dur_i64, const float* dur_f32 /* one of the two */,
                               int B, int L, int T,
                               int32_t* idx /* [B,T], -1 where t >= mel_len */,
                               int64_t* mel_len /* [B] */, mtts_stream stream);
int mtts_length_regulate_fwd(const float* x /* [B,L,C] */, const int32_t* idx /* [B,T] */,
                             int B, int L, int T, int C, float* out /* [B,T,C] */, mtts_stream stream);
/* dx[b,j,:] = sum_{t: idx[b,t]==j} dy[b,t,:]  (segment sum; deterministic, no atomics) */
int mtts_length_regulate_bwd(const float* dy /* [B,T,C] */, const int64_t* dur_i64, const float* dur_f32,
                             int B, int L, int T, int C, float* dx /* [B,L,C] */, mtts_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* MTTS_H_ */
