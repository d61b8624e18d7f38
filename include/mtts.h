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
/* Programmatic Dependent Launch on/off (default on; env MTTS_PDL=0 disables). */
int         mtts_set_pdl(int on);
/* Deterministic mode on/off (default off; env MTTS_DETERMINISTIC=1 enables): no split-K, every cross-CTA atomic reduction
 * (bias / LayerNorm / BatchNorm channel sums, loss and norm scalars, embedding scatter-add) runs in a fixed order, so results are
 * bit-reproducible run to run — the counterpart of the reference's Trainer(deterministic=True), main.py:35.  ~10x slower (a debugging mode). */
int         mtts_set_deterministic(int on);
/* Zero `bytes` bytes at p on the stream (cudaMemsetAsync; a memset node under graph capture).  Replaces optimizer.zero_grad()
 * (Lightning, main.py:57-64) / the zero-initialised gradient buffers of autograd on the flat gradient arenas. */
int         mtts_zero(void* p, int64_t bytes, mtts_stream stream);

/* ------------------------------------------------------------------------------------------
 * Generic tcgen05 GEMM:   for every z = (z0, z1):
 *     C_z[M,N] (+)= alpha * sum_{tap < ntaps} sum_{kb < nkb}  A(z,tap,kb)[M,K] * B(z,tap,kb)[N,K]^T
 * (epilogue order: v = alpha*acc + bias; +C; ReLU; gate; store)
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
  MTTS_EPI_BIAS_ROW = 8,   /* bias indexed by row m instead of column n                          */
  MTTS_EPI_ADD_C    = 16   /* v += c_f32[m,n] (non-atomic read-modify-write; ksplit must be 1)     */
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
  /* optional SECOND product term accumulated into the same tile: C += A2 * B2^T with the geometry of
   * a / b (only the base pointers differ).  Tangent passes: d(xW^T) = xdot W^T + x Wdot^T in one launch. */
  const void* a2_hi; const void* a2_lo; const void* b2_hi; const void* b2_lo;
  int32_t pair;            /* != 0: 2-CTA kernel (tcgen05 cta_group::2, 256 x block_n tile per CTA pair; block_n 128/256) */
  int32_t reserved2;
} mtts_gemm_desc;

int mtts_gemm(const mtts_gemm_desc* d, mtts_stream stream);

/* The same GEMM with `dropout -> + residual -> LayerNorm -> pad-row zeroing` as its epilogue: what MultiHeadAttention.forward does
 * after `fc` (SubLayers.py:54-55: dropout(fc(o)) + residual, layer_norm), PositionwiseFeedForward.forward after `w_2`
 * (SubLayers.py:88-91) and FFTBlock.forward's masked_fill (Layers.py:25,28) — i.e. mtts_gemm followed by mtts_ln_fwd in ONE launch.
 * The output row (N = 256 columns) is spread over the four 64-wide tiles of a 4-CTA cluster, which exchange per-row partial
 * (mean, M2) through distributed shared memory (Chan's parallel variance: one exchange, no cancellation).
 *   v      = alpha * acc + bias;  v *= keep(drop, m * N + n);  v += res[m, n]         -> z_out (fp32, saved for backward)
 *   stats  = (mean_m, rstd_m) of v over n
 *   c_*    = row m valid (m % T < lens[m / T]) ? (v - mean) * rstd * gamma[n] + beta[n] : 0   (fp32 and / or bf16 hi, lo)
 * Requires N == 256, ldc == N, nz0 == nz1 == 1, ksplit == 1, pair == 0, flags == 0 (block_n is forced to 64);
 * every pointer 16-byte aligned. */
typedef struct {
  const float* res;        /* [M, N] or NULL                                                        */
  const float* gamma;      /* [N]                                                                   */
  const float* beta;       /* [N]                                                                   */
  const int64_t* lens;     /* [M / T] valid rows per sequence, or NULL (all rows valid)             */
  int32_t T;
  float eps;
  float* z_out;            /* [M, N] or NULL                                                        */
  float* stats;            /* [M, 2] or NULL                                                        */
  uint32_t drop_thr, drop_seed;   /* dropout on the GEMM result (thr == 0: off), as mtts_ln_fwd's `pre` site */
  float drop_scale;
  int32_t reserved;
  const uint32_t* drop_salt;
} mtts_ln_epilogue;

int mtts_gemm_ln(const mtts_gemm_desc* d, const mtts_ln_epilogue* ln, mtts_stream stream);

/* Ragged -> padded pack = the padding half of collate on the device.  Replaces utils/tools.py:270-301 (pad_1D / pad_2D:
 * np.pad per utterance + np.stack) as used by lightning/collate.py:22-26: dst[b, t, :] = src[row_off[b] + t, :] for
 * t < row_off[b+1] - row_off[b], else 0; rows are copied as bytes (row_bytes a multiple of 4), so every dtype is exact.
 * src holds the B utterances back to back (sum of lengths rows), row_off is int64[B + 1] on the device. */
int mtts_pack_rows(const void* src, const int64_t* row_off, int B, int Lmax, int row_bytes, void* dst, mtts_stream stream);

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


/* ------------------------------------------------------------------------------------------
 * Row-wise kernels (one warp per token row).  `lens`/`T`: row r belongs to batch b = r / T,
 * frame t = r % T and is a PADDED row when t >= lens[b] (lens == NULL: no padding).
 * *_tfwd = tangent forward (JVP of the forward), *_tbwd = tangent backward (JVP of the backward):
 * together they give exact Hessian-vector products for second-order MAML without autograd.
 * ------------------------------------------------------------------------------------------ */

/* Dropout (train mode): every site is (thr, seed, scale): element idx = row*C + col is KEPT iff
 * (mix32(idx + seed_eff*0x9E3779B9) >> 8) >= thr with thr = floor(p*2^24) and
 * seed_eff = seed + (drop_salt ? *drop_salt : 0)*0x632BE5AB (uint32 arithmetic); kept values are multiplied by
 * scale = 1/(1-p); thr = 0 disables the site.  `seed` is a launch-time scalar naming the site and the pass;
 * `drop_salt` is a DEVICE word the host refreshes every step, so a captured CUDA graph draws new masks per replay.  mix32 is the murmur3 finaliser — the same integer hash is
 * evaluated on the host by the oracle (oracle/fs2_oracle.py: drop_mask), so parity holds WITH dropout.
 * LN kernels have two sites: "pre" on the branch input y before the residual add (nn.Dropout in
 * SubLayers.py:54,90) and "post" on the LN output (modules.py:223,235); BN kernels have one site on the
 * (tanh'd) output (F.dropout(..., 0.5) in Layers.py:133-134). */

/* LayerNorm(y + res) then pad-row zeroing.  nn.LayerNorm SubLayers.py:55,91; modules.py:221,233;
 * masked_fill Layers.py:25,28.  z_out = y + res (saved for backward), stats[r] = (mean, rstd). */
int mtts_ln_fwd(const float* y, const float* res, const float* gamma, const float* beta, const int64_t* lens, int T,
                int64_t R, int C, float eps, float* z_out, float* stats, float* out, void* out_hi, void* out_lo,
                uint32_t pre_thr, uint32_t pre_seed, float pre_scale, uint32_t post_thr, uint32_t post_seed, float post_scale, const uint32_t* drop_salt,
                mtts_stream stream);
/* relu_gate != 0: the LN input z is a ReLU output and the gradient is also passed through the ReLU
 * (dz *= z > 0) — the Conv->ReLU->LayerNorm order of modules.py:209-235.  dgamma/dbeta/dbias are
 * ACCUMULATED (atomicAdd); dbias = column sum of dz (bias of the producing Linear/Conv). */
int mtts_ln_bwd(const float* dy, const float* z, const float* stats, const float* gamma, const int64_t* lens, int T, int64_t R,
                int C, int relu_gate, float* dz, void* dz_hi, void* dz_lo, float* dgamma, float* dbeta, float* dbias,
                uint32_t pre_thr, uint32_t pre_seed, float pre_scale, uint32_t post_thr, uint32_t post_seed, float post_scale, const uint32_t* drop_salt,
                mtts_stream stream);
int mtts_ln_tfwd(const float* ydot, const float* resdot, const float* z, const float* stats, const float* gamma,
                 const float* gdot, const float* bdot, const int64_t* lens, int T, int64_t R, int C, float* zdot_out, float* out,
                 void* out_hi, void* out_lo,
                uint32_t pre_thr, uint32_t pre_seed, float pre_scale, uint32_t post_thr, uint32_t post_seed, float post_scale, const uint32_t* drop_salt, mtts_stream stream);
int mtts_ln_tbwd(const float* dy, const float* ddy, const float* z, const float* zdot, const float* stats, const float* gamma,
                 const float* gdot, const int64_t* lens, int T, int64_t R, int C, int relu_gate, float* ddz, void* ddz_hi,
                 void* ddz_lo, float* ddgamma, float* ddbeta, float* ddbias,
                uint32_t pre_thr, uint32_t pre_seed, float pre_scale, uint32_t post_thr, uint32_t post_seed, float post_scale, const uint32_t* drop_salt, mtts_stream stream);

/* VariancePredictor head: Linear(C,1) + masked_fill(mask, 0)  (modules.py:240-250).
 * hdot != NULL selects the tangent form  out = hdot.w + h.wdot + bdot. */
int mtts_rowdot_fwd(const float* h, const float* hdot, const float* w, const float* wdot, const float* b, const float* bdot,
                    const int64_t* lens, int T, int64_t R, int C, float* out, mtts_stream stream);
/* ddout == NULL: dh = dout*w, dw += sum dout*h, db += sum dout.
 * ddout != NULL: tangent-backward  ddh = ddout*w + dout*wdot, ddw += sum(ddout*h + dout*hdot), ddb += sum ddout. */
int mtts_rowdot_bwd(const float* dout, const float* ddout, const float* h, const float* hdot, const float* w, const float* wdot,
                    const int64_t* lens, int T, int64_t R, int C, float* dh, float* dw, float* db, mtts_stream stream);

/* Masked softmax over attention-score rows S[z][q][0..Lk) (leading dim ld), z = b*H + h, keys
 * j >= klens[b] masked (Modules.py:16-22).  P / outputs are bf16 hi(/lo).
 *   mode 0: P   = softmax(A)                                   A = S
 *   mode 1: out = P*(A - sum(P*A))                             backward (A = dP) and tangent fwd (A = Sdot)
 *   mode 2: out = Pd*(A - d) + P*(Bm - dd)                     tangent backward (A = dP, Bm = ddP) */
int mtts_softmax(int mode, const float* A, const float* Bm, const void* p_hi, const void* p_lo, const void* pd_hi,
                 const void* pd_lo, const int64_t* klens, int nz, int H, int Lq, int Lk, int ld, void* o_hi, void* o_lo,
                 mtts_stream stream);

/* ------------------------------------------------------------------------------------------
 * Fused scaled-dot-product attention (csrc/mtts_attn.cu): scores, key-masked softmax and the value /
 * gradient products in one tcgen05 kernel per pass; the [T,T] score matrix stays in tensor memory.
 * Replaces ScaledDotProductAttention.forward, Modules.py:14-25 (bmm(q, k^T)/temperature,
 * masked_fill(mask, -inf), softmax(dim=2), bmm(attn, v)) as called from MultiHeadAttention.forward,
 * SubLayers.py:43-52, and its autograd backward.
 *
 *   q | k | v live in ONE [B*T, 3*H*dk] bf16 (hi/lo) buffer (the fused w_qs|w_ks|w_vs projection,
 *   SubLayers.py:39-41): column block `which*H + h` (which = 0 q, 1 k, 2 v) holds head h.  dk == 128.
 *   o / do are [B*T, H*dk].  keys j >= klens[b] are masked (klens NULL: none).
 *
 * mtts_attn_fwd:  o = softmax(scale * q k^T) v ;  lse[b,h,t] = log2-domain log-sum-exp of row t's scaled
 *   scores (the backward recomputes P from it).  p_hi != NULL additionally EMITS the probabilities
 *   P[B,H,T,Tp] (hi/lo) for passes that re-read them (the Hessian-vector passes).
 * mtts_attn_bwd(parts):  MTTS_ATTN_PREP  dvec[b,h,t] = sum_c do*o   (softmax-backward row term)
 *                        MTTS_ATTN_DQ    dqkv[.., q block] = scale * dS k,  dS = P*(dP - dvec), dP = do v^T
 *                                        (ds_hi != NULL: EMIT dP (fp32) and dS (hi/lo) as [B,H,T,Tp])
 *                        MTTS_ATTN_DK    dqkv[.., k block] = scale * dS^T q
 *                        MTTS_ATTN_DV    dqkv[.., v block] = P^T do
 *   DQ, DK and DV only depend on PREP, write disjoint column blocks and may run concurrently on different streams.
 * lse / dvec: fp32 [B,H,Tl], Tl >= T rounded up to 128 (Tl % 4 == 0); entries beyond T are never read as
 * meaningful values but must be finite.
 * ------------------------------------------------------------------------------------------ */
enum { MTTS_ATTN_PREP = 1, MTTS_ATTN_DQ = 2, MTTS_ATTN_DK = 4, MTTS_ATTN_DV = 8 };

typedef struct {
  int32_t B, H, T, dk;
  int32_t Tp;              /* leading dimension of the emitted [B,H,T,Tp] tensors (multiple of 8, >= T) */
  int32_t Tl;              /* leading dimension of lse / dvec                                          */
  int32_t split;           /* 1 (bf16) or 3 (bf16x3 hi/lo)                                             */
  float   scale;           /* 1 / temperature = d_k^-0.5 (SubLayers.py:27)                             */
  const void* qkv_hi;  const void* qkv_lo;
  const int64_t* klens;    /* [B] or NULL */
  void* o_hi;  void* o_lo; /* fwd: out; bwd PREP: in */
  float* lse;              /* fwd: out; bwd: in  */
  void* p_hi;  void* p_lo; /* fwd: optional emit */
  const void* do_hi;  const void* do_lo;
  float* dvec;             /* bwd PREP: out; DQ / DKV: in */
  void* dqkv_hi;  void* dqkv_lo;
  float* dp;  void* ds_hi;  void* ds_lo;   /* bwd DQ: optional emit */
} mtts_attn_desc;

int mtts_attn_fwd(const mtts_attn_desc* desc, mtts_stream stream);
int mtts_attn_bwd(const mtts_attn_desc* desc, int parts, mtts_stream stream);

/* ------------------------------------------------------------------------------------------
 * Gathers, broadcasts, column sums
 * ------------------------------------------------------------------------------------------ */
/* out[r,:] = table[idx[r],:] (+ base[r,:]) (+ pos[r % T,:]).  nn.Embedding + position_enc Models.py:89-91;
 * pitch / energy embedding add modules.py:80-100,119-126. */
int mtts_embed_fwd(const int64_t* idx, const float* table, const float* base, const float* pos, int T, int64_t R, int C,
                   float* out, void* hi, void* lo, mtts_stream stream);
/* dtable[idx[r],:] += scale*dy[r,:], rows with idx == skip_idx skipped (padding_idx=0, Models.py:56-58). */
int mtts_embed_bwd(const int64_t* idx, const float* dy, int64_t R, int C, int64_t skip_idx, float scale, float* dtable,
                   mtts_stream stream);
/* torch.bucketize(v, bins) (modules.py:83,94): out[r] = #{bins < v[r]}.  Integer path, bit-exact. */
int mtts_bucketize(const float* v, const float* bins, int nb, int64_t R, int64_t* out, mtts_stream stream);
/* out[b,t,:] = x[b,t,:] + vec[b*vec_bstride + :] (+ pos[t,:]).  Speaker-embedding add
 * base_adaptor.py:69-70,80-84 fused with the decoder's position_enc add Models.py:158-160. */
int mtts_add_rowvec(const float* x, const float* vec, int64_t vec_bstride, const float* pos, int B, int T, int C, float* out,
                    void* hi, void* lo, mtts_stream stream);
/* speaker_emb(speaker_args) (+ mean over the support set, base_adaptor.py:64-67). */
int mtts_spk_embed(const int64_t* ids, const float* table, int n, int C, int average, int n_out, float* out, mtts_stream stream);
int mtts_spk_embed_bwd(const int64_t* ids, const float* dspk, int n, int C, int average, int n_out, float scale, float* dtable,
                       mtts_stream stream);
/* out[z,c] += sum_r src[z,r,c]; src is fp32 or bf16 hi(+lo).  Bias gradients. */
int mtts_colsum(const float* f32, const void* hi, const void* lo, int nb, int64_t R, int C, float* out, mtts_stream stream);

/* ------------------------------------------------------------------------------------------
 * BatchNorm1d in train mode (+ optional tanh), PostNet Layers.py:129-137.  Statistics over all
 * R = B*T rows (padded frames included, as the reference).  stats = (mean[C], rstd[C]).
 * ------------------------------------------------------------------------------------------ */
int mtts_bn_fwd(const float* x, const float* gamma, const float* beta, int64_t R, int C, float eps, float momentum, int tanh_flag,
                float* running_mean, float* running_var, float* ws /*[2C]*/, float* stats /*[2C]*/, float* out, void* hi,
                void* lo, uint32_t drop_thr, uint32_t drop_seed, float drop_scale, const uint32_t* drop_salt, mtts_stream stream);
int mtts_bn_bwd(const float* dout, const float* o, const float* x, const float* stats, const float* gamma, int64_t R, int C,
                int tanh_flag, float* ws /*[2C]*/, float* dx, void* hi, void* lo, float* dgamma, float* dbeta, uint32_t drop_thr, uint32_t drop_seed, float drop_scale, const uint32_t* drop_salt, mtts_stream stream);
int mtts_bn_tfwd(const float* xdot, const float* x, const float* stats, const float* gamma, const float* gdot, const float* bdot,
                 const float* o, int64_t R, int C, int tanh_flag, float* ws /*[2C]*/, float* tsums /*[2C]*/, float* odot, void* hi,
                 void* lo, uint32_t drop_thr, uint32_t drop_seed, float drop_scale, const uint32_t* drop_salt, mtts_stream stream);
int mtts_bn_tbwd(const float* dout, const float* ddout, const float* o, const float* odot, const float* x, const float* xdot,
                 const float* stats, const float* tsums, const float* gamma, const float* gdot, int64_t R, int C, int tanh_flag,
                 float* ws /*[4C]*/, float* ddx, void* hi, void* lo, float* ddgamma, float* ddbeta, uint32_t drop_thr, uint32_t drop_seed, float drop_scale, const uint32_t* drop_salt, mtts_stream stream);

/* ------------------------------------------------------------------------------------------
 * FastSpeech2Loss (lightning/model/loss.py:19-92): masked L1 (mel, postnet mel) + masked MSE
 * (pitch, energy, log-duration), means over valid elements.  out6 = (total, mel, postnet_mel,
 * pitch, energy, duration); counts = (n_valid_mel_elements, n_valid_phonemes).
 * loss_bwd(tangent=0): d total*scale / d predictions.  tangent=1: the JVP of those gradients given
 * prediction tangents in (p, e, logd) — L1 has zero curvature so dmel = dpost = 0.
 * ------------------------------------------------------------------------------------------ */
int mtts_loss_fwd(const float* mel, const float* post, const float* mel_tgt, const int64_t* mel_lens, const float* p,
                  const float* p_tgt, const float* e, const float* e_tgt, const float* logd, const int64_t* dur,
                  const int64_t* src_lens, int B, int T, int L, int NM, float* ws /*[8]*/, float* out6, float* counts /*[2]*/,
                  mtts_stream stream);
int mtts_loss_bwd(const float* mel, const float* post, const float* mel_tgt, const int64_t* mel_lens, const float* p,
                  const float* p_tgt, const float* e, const float* e_tgt, const float* logd, const int64_t* dur,
                  const int64_t* src_lens, int B, int T, int L, int NM, const float* counts, float scale, int tangent, float* dmel,
                  float* dpost, float* dp, float* de, float* dlogd, mtts_stream stream);

/* ------------------------------------------------------------------------------------------
 * Flat-arena elementwise kernels (n % 4 == 0).
 * ------------------------------------------------------------------------------------------ */
int mtts_split(const float* src, void* hi, void* lo, int64_t n, mtts_stream stream);
/* l2l maml_update: out = theta - lr*g, fused with the bf16 hi/lo operand preparation. */
int mtts_sgd_split(const float* theta, const float* g, float lr, float* out, void* hi, void* lo, int64_t n, mtts_stream stream);
int mtts_axpby(float a, const float* x, float b, float* y, int64_t n, mtts_stream stream);
/* out[0] = sum x^2 (the global gradient norm of clip_grad_norm_, main.py:61).  `out` holds MTTS_SCALAR_WS floats: out[1..] are
 * per-CTA partials, added in a fixed order (no atomics) — every rank that holds the same reduced gradient gets the same bits. */
#define MTTS_SCALAR_WS 2048
int mtts_sumsq(const float* x, int64_t n, float* out, mtts_stream stream);
/* out[0] = <x, y>  (conjugate-gradient scalars of the iMAML hypergradient, hypertorch/hypergrad/CG_torch.py:21-35); out as above. */
int mtts_dot(const float* x, const float* y, int64_t n, float* out, mtts_stream stream);
/* clip_grad_norm_(max_norm) + Adam (lightning/optimizer.py:6-16, main.py:61).  sumsq = sum g^2 of
 * the UNSCALED buffer, gscale multiplies g first (1/(n_tasks)); hyper = device (lr, 1-b1^t, 1-b2^t). */
int mtts_adam_clip(float* p, const float* g, float* m, float* v, const float* sumsq, float gscale, float max_norm,
                   const float* hyper, float beta1, float beta2, float eps, void* hi, void* lo, int64_t n, mtts_stream stream);

/* ------------------------------------------------------------------------------------------
 * Free-running synthesis (BASELINE configs[4]; lightning/systems/base_adaptor.py:160-186 calls
 * forward_learner without targets) and eval-mode PostNet.
 * ------------------------------------------------------------------------------------------ */
/* duration_rounded = clamp(round(exp(log_d) - 1) * d_control, min=0)   modules.py:133-137 (round half to even). */
int mtts_duration_round(const float* logd, float d_control, int64_t n, float* out, mtts_stream stream);
/* BatchNorm1d under model.eval(): y = (x - running_mean) / sqrt(running_var + eps) * gamma + beta (+ tanh); no update. */
int mtts_bn_eval(const float* x, const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                 int64_t R, int C, float eps, int tanh_flag, float* out, void* hi, void* lo, mtts_stream stream);
/* op 0: log(max(x, a) * b)  (dynamic_range_compression, audio/audio_processing.py:85-91)
 * op 1: exp(x) * a          (dynamic_range_decompression :94-100, a = scale / C)
 * op 2: x * a               (p_control / e_control, modules.py:86,97).  Any of out / hi(+lo) may be NULL. */
enum { MTTS_UN_LOGCLAMP = 0, MTTS_UN_EXP = 1, MTTS_UN_SCALE = 2 };
int mtts_unary(int op, const float* x, int64_t n, float a, float b, float* out, void* hi, void* lo, mtts_stream stream);

/* ------------------------------------------------------------------------------------------
 * STFT / iSTFT / Griffin-Lim (audio/stft.py:15-127, audio/audio_processing.py:7-82, audio/tools.py:18-34).
 * The Fourier- and mel-basis products run through mtts_gemm (the stride-hop conv1d of STFT.transform is a
 * (n_fft/hop)-tap conv over the hop-reshaped signal [rows, hop]; conv_transpose1d of STFT.inverse is the
 * matching dgrad form, so the overlap-add happens in the accumulator).  Frame-major layouts:
 * ri[r, c] = real part of bin c of frame r, ri[r, im_off + c] = imaginary part; row stride ld.
 * ------------------------------------------------------------------------------------------ */
/* F.pad(mode="reflect") by `pad` on both sides (stft.py:60-65) fused with the operand split; columns
 * [N + 2*pad, ld) are zero-filled (ld = rows * hop of the hop-reshaped view; ld may cut the padded tail short). */
int mtts_reflect_pad(const float* x /* [B,N] */, int B, int64_t N, int pad, int64_t ld, float* out /* [B,ld] or NULL */,
                     void* hi, void* lo, mtts_stream stream);
/* magnitude = sqrt(re^2 + im^2), phase = atan2(im, re) (stft.py:78-79), energy[r] = ||magnitude[r,:]||_2 (stft.py:175).
 * mag / phase / mag_hi / mag_lo rows have stride ldm >= nb, columns [nb, ldm) zero-filled (GEMM-operand ready). */
int mtts_stft_polar(const float* ri, int64_t R, int nb, int ld, int im_off, int ldm, float* mag /* [R,ldm] */,
                    float* phase /* [R,ldm] */, float* energy /* [R] */, void* mag_hi, void* mag_lo, mtts_stream stream);
/* X = [mag*cos(ph) | mag*sin(ph)] (stft.py:86-88) as bf16 hi/lo GEMM operand; ph = phase, or the angles of `ri`
 * when phase == NULL (one Griffin-Lim iteration, audio_processing.py:79-81). */
int mtts_stft_recombine(const float* mag /* [R,ldm] */, const float* phase /* [R,ldm] */, const float* ri, int64_t R, int nb,
                        int ld, int im_off, int ldm, void* hi, void* lo, mtts_stream stream);
/* out[b, i] = ola[b, i + trim] / (wsum[i + trim] > tiny ? wsum[i + trim] : 1) * scale   (stft.py:96-122). */
int mtts_istft_finish(const float* ola /* [B,n] */, const float* wsum /* [n] */, float tiny, float scale, int B, int64_t n,
                      int trim, float* out /* [B, n - 2*trim] */, mtts_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* MTTS_H_ */
