// mtts_gemm.cu — the one tensor-core kernel of libmtts: a warp-specialised tcgen05 + TMA GEMM
//
//     C_z[M,N] (+)= alpha * sum_tap sum_kb  A(z,tap,kb)[M,K] * B(z,tap,kb)[N,K]^T
//
// which, by choice of TMA tensor maps / coordinate sources / operand majors, implements every dense
// contraction of the FastSpeech2 hot path: Linear fwd/dgrad/wgrad, Conv1d (k taps as shifted TMA
// loads with hardware zero fill = implicit im2col) fwd/dgrad/wgrad, and the batched attention
// products QK^T, PV and their backward / tangent forms.  See include/mtts.h for the reference
// call sites (SubLayers.py:39-41,54,88; Modules.py:16,23; modules.py:291-296; Layers.py:129-137).
//
// Layout of one CTA (320 threads, 1 CTA = 1 output tile of 128 x BN over its k-range):
//   warp 0      TMA producer  (one elected lane)        global -> smem ring (128B-swizzled tiles)
//   warp 1      MMA issuer    (one elected lane)        tcgen05.mma  smem x smem -> TMEM (fp32)
//   warps 2..9  epilogue      (2 warps per 32-lane TMEM quarter, each half of the columns)  tcgen05.ld -> regs -> global
// Pipelines: full[s]/empty[s] mbarriers for the smem ring, one tmem_full mbarrier MMA -> epilogue.
// mtts_gemm_ln: the 64-wide variant launched as 4-CTA clusters along the row with dropout -> + residual -> LayerNorm -> pad-row
// zeroing as its epilogue (epilogue_ln_tile: row statistics exchanged through distributed shared memory).
//
// bf16x3 mode (SPLIT == 3): operands arrive as hi/lo bf16 pairs; each k-step issues
// hi*hi + hi*lo + lo*hi into the same TMEM accumulator (error ~2^-17, i.e. fp32-grade), which is
// how the 1e-3 (fp32 relative) parity bar is met on bf16 tensor cores.
#include <stdlib.h>
#include "mtts_common.cuh"

namespace {

constexpr int BM = 128;        // UMMA_M (cta_group::1)
constexpr int BK = 64;         // bf16 elements per k-block = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 320;   // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue (two warps per TMEM lane quarter)
constexpr int MAX_SMEM = 227 * 1024;
constexpr int EPI_LN = 1 << 16;    // internal flag (not part of the public MTTS_EPI_* set): the LayerNorm epilogue of mtts_gemm_ln
constexpr int LN_CLUSTER = 4;      // 4 CTAs x 64 columns = the 256-wide row

struct OperandParams {
  int32_t major;
  int32_t src2, src3;
  int32_t shift_src, shift_base, shift_step;
  int32_t rank;              // rank of this operand's tensor maps (2, 3 or 4)
};

struct alignas(64) GemmParams {
  CUtensorMap map_a_hi, map_a_lo, map_b_hi, map_b_lo;
  CUtensorMap map_a2_hi, map_a2_lo, map_b2_hi, map_b2_lo;   // optional second product term (same geometry)
  OperandParams a, b;
  int32_t nterms;
  int32_t M, N, K;
  int32_t ntaps, nkb, nz0, nz1, ksplit;
  int32_t n_tiles;           // tiles along N
  int32_t flags;
  float alpha;
  float* c_f32;
  bf16* c_hi;
  bf16* c_lo;
  int64_t ldc, c_sz0, c_sz1;
  const float* bias;
  int64_t bias_sz0;
  const bf16* gate;
  // LayerNorm epilogue (mtts_gemm_ln; flags & EPI_LN): dropout -> + residual -> LayerNorm -> pad-row zeroing
  const float* ln_res;
  const float* ln_gamma;
  const float* ln_beta;
  const int64_t* ln_lens;
  float* ln_z;
  float* ln_stats;
  int32_t ln_T;
  float ln_eps;
  DropSite ln_drop;
  int32_t dbg;               // MTTS_GEMM_DBG (diagnostics only): bits 0-3 stage cap, 16 skip MMA, 32 skip TMA, 64 epilogue sleeps, 128 no stores
};

// KD = k-blocks (of BK = 64) per pipeline stage.  One stage fill costs ~0.3 us of fixed TMA / barrier time on top of
// its bytes and more than two fills in flight buy nothing (tools/gemm_probe2.py), so narrow tiles — whose MMAs take
// only ~0.25 us per k-block — use KD = 2: half the fills per unit of K.
template <int BN, int SPLIT, int KD = 1>
struct Cfg {
  static constexpr int A_TILE = BM * BK * 2;                  // 16 KB
  static constexpr int B_TILE = BN * BK * 2;
  static constexpr int SUB = (A_TILE + B_TILE) * (SPLIT == 3 ? 2 : 1);   // one k-block of A (hi, lo) and B (hi, lo)
  static constexpr int STAGE = SUB * KD;
  // 64-wide tiles also hold the LayerNorm epilogue's row-statistics exchange area: [4 CTAs][2 column halves][128 rows] x (mean, M2)
  static constexpr int LN_BYTES = BN == 64 ? LN_CLUSTER * 2 * BM * 8 : 0;
  static constexpr int BAR_BYTES = 1024 + LN_BYTES;
  static constexpr int STAGES_RAW = (MAX_SMEM - BAR_BYTES - 1024 /*align slack*/) / STAGE;
  static constexpr int STAGES = STAGES_RAW > 6 ? 6 : STAGES_RAW;
  static constexpr int SMEM = STAGES * STAGE + BAR_BYTES + 1024;
  static_assert(STAGES >= 2, "need at least a double buffer");
};

__device__ __forceinline__ int pick_src(int src, int z0, int z1, int tap, int kb) {
  return src == MTTS_SRC_Z0 ? z0 : src == MTTS_SRC_Z1 ? z1 : src == MTTS_SRC_TAP ? tap : src == MTTS_SRC_KB ? kb : 0;
}

// TMEM accumulator tile (128 lanes x BN columns) -> alpha, bias, +C, ReLU, gate -> global (fp32 / bf16 hi,lo / red.add)
//
// Stores are COALESCED through a per-warp 4 KB staging buffer (`stage`, carved out of the pipeline's shared memory,
// which is idle once the accumulator is complete): tcgen05.ld hands every thread 32 consecutive columns of ITS row,
// so direct stores make each warp instruction touch 32 different rows (16 B of each).  The warp instead writes its
// 32 x 32 chunk to shared memory (XOR-swizzled float4 slots: conflict-free both ways) and reads it back with lane
// l -> (row l / 8 of 4, float4 l % 8): every global instruction then covers 4 rows x 128 contiguous bytes.
template <int BN>
__device__ __forceinline__ void epilogue_tile(const GemmParams& p, uint32_t tmem_base, int q, int lane, int m0, int n0, int z0,
                                              int z1, int c_begin, int c_end, float4* stage) {
  const int row = m0 + q * 32 + lane;
  const bool row_ok = row < p.M;
  const int64_t z_off = int64_t(z0) * p.c_sz0 + int64_t(z1) * p.c_sz1;
  const int64_t c_off = z_off + int64_t(row) * p.ldc;
  const float* bias = p.bias ? p.bias + int64_t(z0) * p.bias_sz0 : nullptr;
  const bool vec_ok = ((p.ldc & 7) == 0) && ((p.c_sz0 & 7) == 0) && ((p.c_sz1 & 7) == 0) &&
                      ((reinterpret_cast<uintptr_t>(p.c_f32) & 15) == 0) && ((reinterpret_cast<uintptr_t>(p.c_hi) & 15) == 0) &&
                      ((reinterpret_cast<uintptr_t>(p.c_lo) & 15) == 0) && ((reinterpret_cast<uintptr_t>(p.gate) & 15) == 0);
  const float bias_row = (bias && (p.flags & MTTS_EPI_BIAS_ROW) && row_ok) ? bias[row] : 0.f;
#pragma unroll 1
  for (int c = c_begin; c < c_end; ++c) {
    __syncwarp();                       // tcgen05.ld is .sync.aligned: reconverge after guards
    uint32_t r[32];
    tmem_ld_32x32(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(c * 32), r);
    tmem_ld_wait();
    const int col0 = n0 + c * 32;
    if (col0 >= p.N || (p.dbg & 128)) continue;          // warp-uniform
    const bool full = (col0 + 32 <= p.N) && vec_ok;      // warp-uniform
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * p.alpha;
    if (row_ok) {
      if (bias) {
        if (p.flags & MTTS_EPI_BIAS_ROW) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += bias_row;
        } else if (col0 + 32 <= p.N && ((reinterpret_cast<uintptr_t>(bias + col0) & 15) == 0)) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + col0 + j));
            v[j] += bb.x; v[j + 1] += bb.y; v[j + 2] += bb.z; v[j + 3] += bb.w;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (col0 + j < p.N) v[j] += __ldg(bias + col0 + j);
        }
      }
      if (p.flags & MTTS_EPI_ADD_C) {
        const float* src = p.c_f32 + c_off + col0;
        if (full) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 cc = *reinterpret_cast<const float4*>(src + j);
            v[j] += cc.x; v[j + 1] += cc.y; v[j + 2] += cc.z; v[j + 3] += cc.w;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (col0 + j < p.N) v[j] += src[j];
        }
      }
      if (p.flags & MTTS_EPI_RELU) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
      }
      if (p.flags & MTTS_EPI_GATE) {
        const bf16* g = p.gate + c_off + col0;
        if (full) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            const uint4 gg = *reinterpret_cast<const uint4*>(g + j);
            const uint32_t w[4] = {gg.x, gg.y, gg.z, gg.w};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              // bf16 > 0  <=>  sign bit clear and magnitude non-zero
              const uint32_t lo16 = w[t] & 0xFFFFu, hi16 = w[t] >> 16;
              if (!((lo16 & 0x8000u) == 0 && (lo16 & 0x7FFFu) != 0)) v[j + 2 * t] = 0.f;
              if (!((hi16 & 0x8000u) == 0 && (hi16 & 0x7FFFu) != 0)) v[j + 2 * t + 1] = 0.f;
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (col0 + j < p.N && !(__bfloat162float(g[j]) > 0.f)) v[j] = 0.f;
        }
      }
    }
    if (full && !(p.dbg & 1024)) {
      // ---- coalesced path: transpose the warp's 32 x 32 chunk through shared memory ----
#pragma unroll
      for (int j = 0; j < 8; ++j) stage[lane * 8 + (j ^ (lane & 7))] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      __syncwarp();
      const int jj = lane & 7;
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int rr = it * 4 + (lane >> 3);
        const float4 t = stage[rr * 8 + (jj ^ (rr & 7))];
        const int grow = m0 + q * 32 + rr;
        if (grow < p.M) {
          const int64_t off = z_off + int64_t(grow) * p.ldc + col0 + jj * 4;
          if (p.c_f32) {
            float* dst = p.c_f32 + off;
            if (p.flags & MTTS_EPI_ACCUM)
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(t.x), "f"(t.y), "f"(t.z), "f"(t.w) : "memory");
            else
              *reinterpret_cast<float4*>(dst) = t;
          }
          if (p.c_hi) {
            bf16 h0, h1, h2, h3, l0, l1, l2, l3;
            split_bf16(t.x, h0, l0); split_bf16(t.y, h1, l1); split_bf16(t.z, h2, l2); split_bf16(t.w, h3, l3);
            uint2 h;
            h.x = uint32_t(__bfloat16_as_ushort(h0)) | (uint32_t(__bfloat16_as_ushort(h1)) << 16);
            h.y = uint32_t(__bfloat16_as_ushort(h2)) | (uint32_t(__bfloat16_as_ushort(h3)) << 16);
            *reinterpret_cast<uint2*>(p.c_hi + off) = h;
            if (p.c_lo) {
              uint2 l;
              l.x = uint32_t(__bfloat16_as_ushort(l0)) | (uint32_t(__bfloat16_as_ushort(l1)) << 16);
              l.y = uint32_t(__bfloat16_as_ushort(l2)) | (uint32_t(__bfloat16_as_ushort(l3)) << 16);
              *reinterpret_cast<uint2*>(p.c_lo + off) = l;
            }
          }
        }
      }
      __syncwarp();                     // the staging buffer is rewritten by the next chunk
      continue;
    }
    if (!row_ok) continue;
    if (p.c_f32) {
      float* dst = p.c_f32 + c_off + col0;
      if (p.flags & MTTS_EPI_ACCUM) {
        if (full) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(v[j]), "f"(v[j + 1]), "f"(v[j + 2]),
                         "f"(v[j + 3])
                         : "memory");
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (col0 + j < p.N) atomicAdd(dst + j, v[j]);
        }
      } else if (full) {
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (col0 + j < p.N) dst[j] = v[j];
      }
    }
    if (p.c_hi) {
      bf16* dh = p.c_hi + c_off + col0;
      bf16* dl = p.c_lo ? p.c_lo + c_off + col0 : nullptr;
      if (full) {
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          uint32_t h[4], l[4];
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            bf16 h0, l0, h1, l1;
            split_bf16(v[j + 2 * t], h0, l0);
            split_bf16(v[j + 2 * t + 1], h1, l1);
            h[t] = uint32_t(__bfloat16_as_ushort(h0)) | (uint32_t(__bfloat16_as_ushort(h1)) << 16);
            l[t] = uint32_t(__bfloat16_as_ushort(l0)) | (uint32_t(__bfloat16_as_ushort(l1)) << 16);
          }
          *reinterpret_cast<uint4*>(dh + j) = make_uint4(h[0], h[1], h[2], h[3]);
          if (dl) *reinterpret_cast<uint4*>(dl + j) = make_uint4(l[0], l[1], l[2], l[3]);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (col0 + j < p.N) {
            bf16 h0, l0;
            split_bf16(v[j], h0, l0);
            dh[j] = h0;
            if (dl) dl[j] = l0;
          }
      }
    }
  }
}

// One warp's 32 x 32 chunk (thread = row, v = its 32 columns) -> global through the warp's staging buffer: every store instruction
// covers 4 rows x 128 contiguous bytes (same transposition as epilogue_tile).
__device__ __forceinline__ void store_chunk_staged(float4* stage, const float (&v)[32], int lane, int row0, int M, int64_t ld, int col0,
                                                   float* f32, bf16* hi, bf16* lo) {
#pragma unroll
  for (int j = 0; j < 8; ++j) stage[lane * 8 + (j ^ (lane & 7))] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  __syncwarp();
  const int jj = lane & 7;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int rr = it * 4 + (lane >> 3);
    const float4 t = stage[rr * 8 + (jj ^ (rr & 7))];
    const int grow = row0 + rr;
    if (grow < M) {
      const int64_t off = int64_t(grow) * ld + col0 + jj * 4;
      if (f32) *reinterpret_cast<float4*>(f32 + off) = t;
      if (hi) {
        bf16 h0, h1, h2, h3, l0, l1, l2, l3;
        split_bf16(t.x, h0, l0); split_bf16(t.y, h1, l1); split_bf16(t.z, h2, l2); split_bf16(t.w, h3, l3);
        uint2 h;
        h.x = uint32_t(__bfloat16_as_ushort(h0)) | (uint32_t(__bfloat16_as_ushort(h1)) << 16);
        h.y = uint32_t(__bfloat16_as_ushort(h2)) | (uint32_t(__bfloat16_as_ushort(h3)) << 16);
        *reinterpret_cast<uint2*>(hi + off) = h;
        if (lo) {
          uint2 l;
          l.x = uint32_t(__bfloat16_as_ushort(l0)) | (uint32_t(__bfloat16_as_ushort(l1)) << 16);
          l.y = uint32_t(__bfloat16_as_ushort(l2)) | (uint32_t(__bfloat16_as_ushort(l3)) << 16);
          *reinterpret_cast<uint2*>(lo + off) = l;
        }
      }
    }
  }
  __syncwarp();                       // the staging buffer is rewritten by the next chunk
}

// LayerNorm epilogue of a 128 x 64 tile whose row continues in the three other CTAs of the cluster (mtts_gemm_ln, include/mtts.h).
// Warp (q, half) owns rows q*32 .. +32 and columns half*32 .. +32 of the tile: 8 partial (mean, M2) per row in the cluster.  Every thread
// writes its partial into the exchange area of ALL four CTAs with st.async, which also counts its bytes on that CTA's `ln_bar`
// (4 CTAs x 256 threads x 8 bytes expected); a CTA reads only its own shared memory, and only after every writer's bytes have
// landed, so no CTA can exit while a peer still needs it.
__device__ __forceinline__ void epilogue_ln_tile(const GemmParams& p, uint32_t tmem_base, int q, int lane, int half, int m0, int n0,
                                                 float4* stage, float2* part, uint64_t* ln_bar, uint64_t* tmem_full_bar) {
  const int rl = q * 32 + lane;
  const int row = m0 + rl;
  const bool row_ok = row < p.M;
  const int col0 = n0 + half * 32;
  // Everything that does not depend on the accumulator is fetched while the main loop runs (these warps are idle until then):
  //   v = (alpha * acc + bias) * keep + res  =  acc * mul + add,   mul = alpha * keep,  add = bias * keep + res
  float mul[32], add[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) { mul[j] = 0.f; add[j] = 0.f; }
  if (row_ok) {
#pragma unroll
    for (int j = 0; j < 32; ++j) mul[j] = p.alpha;
    if (p.ln_drop.thr) {              // element index = m * N + n, as the stand-alone LayerNorm kernel's `pre` site
      const uint32_t e0 = uint32_t(row) * uint32_t(p.N) + uint32_t(col0);
#pragma unroll
      for (int j = 0; j < 32; ++j) mul[j] *= drop_factor(p.ln_drop, e0 + j);
    }
    if (p.bias) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j));
        add[j] = bb.x; add[j + 1] = bb.y; add[j + 2] = bb.z; add[j + 3] = bb.w;
      }
      if (p.ln_drop.thr) {
        const float ia = 1.f / p.alpha;
#pragma unroll
        for (int j = 0; j < 32; ++j) add[j] *= mul[j] * ia;         // bias * keep
      }
    }
    if (p.ln_res) {
      const float4* rs = reinterpret_cast<const float4*>(p.ln_res + int64_t(row) * p.ldc + col0);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 t = rs[j];
        add[4 * j] += t.x; add[4 * j + 1] += t.y; add[4 * j + 2] += t.z; add[4 * j + 3] += t.w;
      }
    }
  }
  if (p.dbg & 64) mbar_wait_sleep(tmem_full_bar, 0); else mbar_wait(tmem_full_bar, 0);
  tc_fence_after();
  uint32_t r[32];
  tmem_ld_32x32(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(half * 32), r);
  tmem_ld_wait();
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = fmaf(__uint_as_float(r[j]), mul[j], add[j]);
  // partial statistics of my 32 columns, centred on their own mean
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 32; ++j) s += v[j];
  const float mi = s * (1.f / 32.f);
  float m2 = 0.f;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const float d = v[j] - mi;
    m2 = fmaf(d, d, m2);
  }
  // st.async: the store itself completes 8 transaction bytes on the destination CTA's mbarrier (which expects all 4 x 256 x 8)
  const uint32_t rank = cluster_ctarank();
  const uint32_t slot = smem_u32(part + (rank * 2 + half) * BM + rl);
  const uint32_t bar = smem_u32(ln_bar);
#pragma unroll
  for (uint32_t t = 0; t < LN_CLUSTER; ++t) st_async_f32x2(mapa_shared(slot, t), mi, m2, mapa_shared(bar, t));
  mbar_wait(ln_bar, 0);
  // Chan et al.: 8 groups of 32 -> mean, M2 of the 256-wide row (fixed order: bit-reproducible)
  float2 pr[2 * LN_CLUSTER];
  float mean = 0.f;
#pragma unroll
  for (int i = 0; i < 2 * LN_CLUSTER; ++i) {
    pr[i] = part[i * BM + rl];
    mean += pr[i].x;
  }
  mean *= 1.f / (2 * LN_CLUSTER);
  float M2 = 0.f;
#pragma unroll
  for (int i = 0; i < 2 * LN_CLUSTER; ++i) {
    const float d = pr[i].x - mean;
    M2 += pr[i].y + 32.f * d * d;
  }
  const float rstd = rsqrtf(M2 * (1.f / (64.f * LN_CLUSTER)) + p.ln_eps);
  if (row_ok && p.ln_stats && rank == 0 && half == 0) {
    p.ln_stats[2 * int64_t(row)] = mean;
    p.ln_stats[2 * int64_t(row) + 1] = rstd;
  }
  if (p.ln_z) store_chunk_staged(stage, v, lane, m0 + q * 32, p.M, p.ldc, col0, p.ln_z, nullptr, nullptr);
  bool valid = row_ok;
  if (valid && p.ln_lens) {
    const int b = row / p.ln_T;
    valid = (row - b * p.ln_T) < p.ln_lens[b];
  }
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    const float4 g = __ldg(reinterpret_cast<const float4*>(p.ln_gamma + col0 + j));
    const float4 b = __ldg(reinterpret_cast<const float4*>(p.ln_beta + col0 + j));
    v[j] = valid ? fmaf((v[j] - mean) * rstd, g.x, b.x) : 0.f;
    v[j + 1] = valid ? fmaf((v[j + 1] - mean) * rstd, g.y, b.y) : 0.f;
    v[j + 2] = valid ? fmaf((v[j + 2] - mean) * rstd, g.z, b.z) : 0.f;
    v[j + 3] = valid ? fmaf((v[j + 3] - mean) * rstd, g.w, b.w) : 0.f;
  }
  store_chunk_staged(stage, v, lane, m0 + q * 32, p.M, p.ldc, col0, p.c_f32, p.c_hi, p.c_lo);
}

// ================================================================================================
template <int BN, int SPLIT, int KD>
__global__ void __launch_bounds__(NUM_THREADS, 1) mtts_gemm_kernel(const __grid_constant__ GemmParams p) {
  using C = Cfg<BN, SPLIT, KD>;
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* bar_base = smem + C::STAGES * C::STAGE;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bar_base);
  uint64_t* empty_bar = full_bar + C::STAGES;
  uint64_t* tmem_full_bar = empty_bar + C::STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  uint64_t* ln_bar = tmem_full_bar + 2;
  const bool ln = (BN == 64) && (p.flags & EPI_LN);       // launched as 4-CTA clusters along the row (blockIdx.x % 4 = n_tile)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int tile = blockIdx.x;
  const int m_tile = tile / p.n_tiles;
  const int n_tile = tile - m_tile * p.n_tiles;
  const int m0 = m_tile * BM;
  const int n0 = n_tile * BN;
  const int z0 = blockIdx.z % p.nz0;
  const int z1 = blockIdx.z / p.nz0;

  // this CTA's slice of the (tap, kb, kc) iteration space
  const int kchunks = ((p.K + BK - 1) / BK + KD - 1) / KD;     // stage fills along K (KD k-blocks each; the tail block is OOB zero-filled)
  const int total_iters = p.nterms * p.ntaps * p.nkb * kchunks;
  const int per_split = (total_iters + p.ksplit - 1) / p.ksplit;
  const int it_begin = blockIdx.y * per_split;
  const int it_end = min(total_iters, it_begin + per_split);
  const int n_iters = max(0, it_end - it_begin);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.map_a_hi);
    tma_prefetch_desc(&p.map_b_hi);
    if (SPLIT == 3) {
      tma_prefetch_desc(&p.map_a_lo);
      tma_prefetch_desc(&p.map_b_lo);
    }
    if (p.nterms > 1) {
      tma_prefetch_desc(&p.map_a2_hi);
      tma_prefetch_desc(&p.map_b2_hi);
      if (SPLIT == 3) {
        tma_prefetch_desc(&p.map_a2_lo);
        tma_prefetch_desc(&p.map_b2_lo);
      }
    }
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    if (ln) {                                    // completes when every epilogue thread of the cluster has stored its 8 bytes here
      mbar_init(ln_bar, 1);
      mbar_arrive_expect_tx(ln_bar, LN_CLUSTER * 256 * 8);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc<BN>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (ln) cluster_sync_all();                    // every CTA's ln_bar is initialised before a peer can arrive on it
  const uint32_t tmem_base = *tmem_slot;
  const int nstages = (p.dbg & 15) ? min(p.dbg & 15, C::STAGES) : C::STAGES;
  pdl_wait();                                    // everything above overlapped the previous kernel's tail

  if (warp == 0) {
    // ===================================== TMA producer ==========================================
    if (lane == 0 && n_iters > 0) {
      int stage = 0;
      uint32_t phase = 0;
      // (term, tap, kb, kc) odometer: no integer divisions on the per-k-iteration critical path
      int kc = it_begin % kchunks, kb, tap, term;
      {
        int rest = it_begin / kchunks;
        kb = rest % p.nkb;
        rest /= p.nkb;
        tap = rest % p.ntaps;
        term = rest / p.ntaps;
      }
      auto advance = [&]() {
        if (++kc == kchunks) {
          kc = 0;
          if (++kb == p.nkb) {
            kb = 0;
            if (++tap == p.ntaps) {
              tap = 0;
              ++term;
            }
          }
        }
      };
      for (int it = it_begin; it < it_end; ++it, advance()) {
        const bool t2 = term > 0;                         // second product term (tangent passes)
        const CUtensorMap* ma_hi = t2 ? &p.map_a2_hi : &p.map_a_hi;
        const CUtensorMap* ma_lo = t2 ? &p.map_a2_lo : &p.map_a_lo;
        const CUtensorMap* mb_hi = t2 ? &p.map_b2_hi : &p.map_b_hi;
        const CUtensorMap* mb_lo = t2 ? &p.map_b2_lo : &p.map_b_lo;

        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (p.dbg & 32) {                                  // diagnostics: no loads, just hand the slot over
          mbar_arrive(&full_bar[stage]);
          if (++stage == nstages) { stage = 0; phase ^= 1; }
          continue;
        }
        mbar_arrive_expect_tx(&full_bar[stage], C::STAGE);

#pragma unroll
        for (int sub = 0; sub < KD; ++sub) {
        const int kcb = (kc * KD + sub) * BK;             // element offset of this k-block along K
        uint8_t* sa = smem + stage * C::STAGE + sub * C::SUB;
        uint8_t* sa_lo = sa + C::A_TILE;
        uint8_t* sb = sa + C::A_TILE * (SPLIT == 3 ? 2 : 1);
        uint8_t* sb_lo = sb + C::B_TILE;

        {  // A operand
          const int shift = p.a.shift_base + p.a.shift_step * pick_src(p.a.shift_src, z0, z1, tap, kb);
          const int c2 = pick_src(p.a.src2, z0, z1, tap, kb);
          const int c3 = pick_src(p.a.src3, z0, z1, tap, kb);
          if (p.a.major == MTTS_MAJOR_K) {
            tma_load_nd(p.a.rank, sa, ma_hi, &full_bar[stage], kcb, m0 + shift, c2, c3);
            if (SPLIT == 3) tma_load_nd(p.a.rank, sa_lo, ma_lo, &full_bar[stage], kcb, m0 + shift, c2, c3);
          } else {
#pragma unroll
            for (int i = 0; i < BM / 64; ++i) {
              tma_load_nd(p.a.rank, sa + i * (BK * 128), ma_hi, &full_bar[stage], m0 + 64 * i, kcb + shift, c2, c3);
              if (SPLIT == 3)
                tma_load_nd(p.a.rank, sa_lo + i * (BK * 128), ma_lo, &full_bar[stage], m0 + 64 * i, kcb + shift, c2, c3);
            }
          }
        }
        {  // B operand
          const int shift = p.b.shift_base + p.b.shift_step * pick_src(p.b.shift_src, z0, z1, tap, kb);
          const int c2 = pick_src(p.b.src2, z0, z1, tap, kb);
          const int c3 = pick_src(p.b.src3, z0, z1, tap, kb);
          if (p.b.major == MTTS_MAJOR_K) {
            tma_load_nd(p.b.rank, sb, mb_hi, &full_bar[stage], kcb, n0 + shift, c2, c3);
            if (SPLIT == 3) tma_load_nd(p.b.rank, sb_lo, mb_lo, &full_bar[stage], kcb, n0 + shift, c2, c3);
          } else {
#pragma unroll
            for (int i = 0; i < BN / 64; ++i) {
              tma_load_nd(p.b.rank, sb + i * (BK * 128), mb_hi, &full_bar[stage], n0 + 64 * i, kcb + shift, c2, c3);
              if (SPLIT == 3)
                tma_load_nd(p.b.rank, sb_lo + i * (BK * 128), mb_lo, &full_bar[stage], n0 + 64 * i, kcb + shift, c2, c3);
            }
          }
        }
        }   // sub
        if (++stage == nstages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer ============================================
    if (lane == 0 && n_iters > 0) {
      const uint32_t idesc = make_idesc_bf16(BN, p.a.major == MTTS_MAJOR_MN, p.b.major == MTTS_MAJOR_MN);
      // per-UMMA_K advance of the descriptor start address and LBO, by operand major
      const uint32_t a_step = (p.a.major == MTTS_MAJOR_K) ? UMMA_K * 2 : UMMA_K * 128;
      const uint32_t b_step = (p.b.major == MTTS_MAJOR_K) ? UMMA_K * 2 : UMMA_K * 128;
      const uint32_t a_lbo = (p.a.major == MTTS_MAJOR_K) ? 16 : BK * 128;
      const uint32_t b_lbo = (p.b.major == MTTS_MAJOR_K) ? 16 : BK * 128;
      int stage = 0;
      uint32_t phase = 0;
      uint32_t accumulate = 0;
      for (int it = 0; it < n_iters; ++it) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (!(p.dbg & 16))
#pragma unroll
        for (int sub = 0; sub < KD; ++sub) {
        const uint32_t sa = smem_u32(smem + stage * C::STAGE + sub * C::SUB);
        const uint32_t sa_lo = sa + C::A_TILE;
        const uint32_t sb = sa + C::A_TILE * (SPLIT == 3 ? 2 : 1);
        const uint32_t sb_lo = sb + C::B_TILE;
#pragma unroll
        for (int kk = 0; kk < BK / UMMA_K; ++kk) {
          const uint64_t da = make_umma_desc(sa + kk * a_step, a_lbo, 1024);
          const uint64_t db = make_umma_desc(sb + kk * b_step, b_lbo, 1024);
          umma_bf16(tmem_base, da, db, idesc, accumulate);
          accumulate = 1;
          if (SPLIT == 3) {
            const uint64_t da_lo = make_umma_desc(sa_lo + kk * a_step, a_lbo, 1024);
            const uint64_t db_lo = make_umma_desc(sb_lo + kk * b_step, b_lbo, 1024);
            umma_bf16(tmem_base, da, db_lo, idesc, 1);
            umma_bf16(tmem_base, da_lo, db, idesc, 1);
          }
        }
        }   // sub
        umma_commit(&empty_bar[stage]);   // frees this smem slot once the MMAs above have read it
        if (++stage == nstages) {
          stage = 0;
          phase ^= 1;
        }
      }
      umma_commit(tmem_full_bar);          // accumulator complete -> epilogue
    }
  } else {
    // ===================================== epilogue ==============================================
    // TMEM lane quarter accessible to a warp is (warp_id % 4); warps 2,3,4,5 -> quarters 2,3,0,1.
    const int q = warp & 3;
    if (n_iters > 0) {
      // two warps share a TMEM lane quarter (warp % 4) and split the BN columns between them
      constexpr int CH = BN / 32;
      const int half = (warp - 2) >> 2;
      const int c_begin = CH >= 2 ? half * (CH / 2) : 0;
      const int c_end = CH >= 2 ? c_begin + CH / 2 : (half == 0 ? CH : 0);
      if (BN == 64 && ln) {
        epilogue_ln_tile(p, tmem_base, q, lane, half, m0, n0, reinterpret_cast<float4*>(smem) + (warp - 2) * 256,
                         reinterpret_cast<float2*>(bar_base + 1024), ln_bar, tmem_full_bar);
      } else {
        if (p.dbg & 64) mbar_wait_sleep(tmem_full_bar, 0); else mbar_wait(tmem_full_bar, 0);
        tc_fence_after();
        epilogue_tile<BN>(p, tmem_base, q, lane, m0, n0, z0, z1, c_begin, c_end,
                          reinterpret_cast<float4*>(smem) + (warp - 2) * 256);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<BN>(tmem_base);
  }
}

// ================================================================================================
// 2-CTA variant: a cluster of two CTAs computes one 256 x BN tile with tcgen05.mma.cta_group::2.
// Each CTA stages its own 128 rows of A and HALF of the B tile (BN/2 rows), so the operand bytes
// fetched from L2 per flop halve compared with two independent 128 x BN tiles — the 1-CTA kernel is
// L2-throughput bound in bf16x3 (profiles/r01_ncu_summary.md).  The leader CTA (cluster rank 0) issues
// the MMAs for both; tcgen05.commit multicasts the smem-slot release and the accumulator-ready signal to
// both CTAs; every CTA runs its own TMA producer (completion bytes land on the leader's full barrier)
// and its own epilogue over its 128 TMEM lanes.  The two CTAs' row tiles are consecutive m-tiles of the
// same z; an odd tail gets a dummy partner (all-OOB A rows, nothing stored).
// ================================================================================================
template <int BN, int SPLIT, int KD = 1>
struct Cfg2 {
  static constexpr int A_TILE = BM * BK * 2;                  // this CTA's 128 rows
  static constexpr int B_TILE = (BN / 2) * BK * 2;            // this CTA's half of the B tile
  static constexpr int SUB = (A_TILE + B_TILE) * (SPLIT == 3 ? 2 : 1);
  static constexpr int STAGE = SUB * KD;                      // KD k-blocks per stage fill (see Cfg)
  static constexpr int BAR_BYTES = 1024;
  static constexpr int STAGES_RAW = (MAX_SMEM - BAR_BYTES - 1024) / STAGE;
  static constexpr int STAGES = STAGES_RAW > 6 ? 6 : STAGES_RAW;
  static constexpr int SMEM = STAGES * STAGE + BAR_BYTES + 1024;
  static_assert(STAGES >= 2, "need at least a double buffer");
};

template <int BN, int SPLIT, int KD>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1) mtts_gemm_pair_kernel(const __grid_constant__ GemmParams p) {
  using C = Cfg2<BN, SPLIT, KD>;
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* bar_base = smem + C::STAGES * C::STAGE;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bar_base);
  uint64_t* empty_bar = full_bar + C::STAGES;
  uint64_t* tmem_full_bar = empty_bar + C::STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();       // 0 = leader

  const int pair = blockIdx.x >> 1;
  const int pair_m = pair / p.n_tiles;
  const int n_tile = pair - pair_m * p.n_tiles;
  const int m0 = (2 * pair_m + int(rank)) * BM;  // may be >= M for the dummy partner of an odd tail
  const int n0 = n_tile * BN;
  const int nb0 = n0 + int(rank) * (BN / 2);     // this CTA's half of the B tile
  const int z0 = blockIdx.z % p.nz0;
  const int z1 = blockIdx.z / p.nz0;

  const int kchunks = ((p.K + BK - 1) / BK + KD - 1) / KD;     // stage fills along K
  const int total_iters = p.nterms * p.ntaps * p.nkb * kchunks;
  const int per_split = (total_iters + p.ksplit - 1) / p.ksplit;
  const int it_begin = blockIdx.y * per_split;
  const int it_end = min(total_iters, it_begin + per_split);
  const int n_iters = max(0, it_end - it_begin);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.map_a_hi);
    tma_prefetch_desc(&p.map_b_hi);
    if (SPLIT == 3) {
      tma_prefetch_desc(&p.map_a_lo);
      tma_prefetch_desc(&p.map_b_lo);
    }
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc_2cta<BN>(tmem_slot);
  }
  tc_fence_before();
  cluster_sync_all();                            // peer barriers initialised before any remote complete_tx / arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                                    // everything above overlapped the previous kernel's tail

  if (warp == 0) {
    // ===================================== TMA producer (both CTAs) ===============================
    if (lane == 0 && n_iters > 0) {
      int stage = 0;
      uint32_t phase = 0;
      // (term, tap, kb, kc) odometer: no integer divisions on the per-k-iteration critical path
      int kc = it_begin % kchunks, kb, tap, term;
      {
        int rest = it_begin / kchunks;
        kb = rest % p.nkb;
        rest /= p.nkb;
        tap = rest % p.ntaps;
        term = rest / p.ntaps;
      }
      auto advance = [&]() {
        if (++kc == kchunks) {
          kc = 0;
          if (++kb == p.nkb) {
            kb = 0;
            if (++tap == p.ntaps) {
              tap = 0;
              ++term;
            }
          }
        }
      };
      for (int it = it_begin; it < it_end; ++it, advance()) {
        const bool t2 = term > 0;                         // second product term (tangent passes)
        const CUtensorMap* ma_hi = t2 ? &p.map_a2_hi : &p.map_a_hi;
        const CUtensorMap* ma_lo = t2 ? &p.map_a2_lo : &p.map_a_lo;
        const CUtensorMap* mb_hi = t2 ? &p.map_b2_hi : &p.map_b_hi;
        const CUtensorMap* mb_lo = t2 ? &p.map_b2_lo : &p.map_b_lo;

        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * C::STAGE);   // bytes of both CTAs

#pragma unroll
        for (int sub = 0; sub < KD; ++sub) {
        const int kcb = (kc * KD + sub) * BK;
        uint8_t* sa = smem + stage * C::STAGE + sub * C::SUB;
        uint8_t* sa_lo = sa + C::A_TILE;
        uint8_t* sb = sa + C::A_TILE * (SPLIT == 3 ? 2 : 1);
        uint8_t* sb_lo = sb + C::B_TILE;
        {
          const int shift = p.a.shift_base + p.a.shift_step * pick_src(p.a.shift_src, z0, z1, tap, kb);
          const int c2 = pick_src(p.a.src2, z0, z1, tap, kb);
          const int c3 = pick_src(p.a.src3, z0, z1, tap, kb);
          if (p.a.major == MTTS_MAJOR_K) {
            tma_load_nd_2cta(p.a.rank, sa, ma_hi, &full_bar[stage], kcb, m0 + shift, c2, c3);
            if (SPLIT == 3) tma_load_nd_2cta(p.a.rank, sa_lo, ma_lo, &full_bar[stage], kcb, m0 + shift, c2, c3);
          } else {
#pragma unroll
            for (int i = 0; i < BM / 64; ++i) {
              tma_load_nd_2cta(p.a.rank, sa + i * (BK * 128), ma_hi, &full_bar[stage], m0 + 64 * i, kcb + shift, c2, c3);
              if (SPLIT == 3)
                tma_load_nd_2cta(p.a.rank, sa_lo + i * (BK * 128), ma_lo, &full_bar[stage], m0 + 64 * i, kcb + shift, c2, c3);
            }
          }
        }
        {
          const int shift = p.b.shift_base + p.b.shift_step * pick_src(p.b.shift_src, z0, z1, tap, kb);
          const int c2 = pick_src(p.b.src2, z0, z1, tap, kb);
          const int c3 = pick_src(p.b.src3, z0, z1, tap, kb);
          if (p.b.major == MTTS_MAJOR_K) {
            tma_load_nd_2cta(p.b.rank, sb, mb_hi, &full_bar[stage], kcb, nb0 + shift, c2, c3);
            if (SPLIT == 3) tma_load_nd_2cta(p.b.rank, sb_lo, mb_lo, &full_bar[stage], kcb, nb0 + shift, c2, c3);
          } else {
#pragma unroll
            for (int i = 0; i < BN / 128; ++i) {
              tma_load_nd_2cta(p.b.rank, sb + i * (BK * 128), mb_hi, &full_bar[stage], nb0 + 64 * i, kcb + shift, c2, c3);
              if (SPLIT == 3)
                tma_load_nd_2cta(p.b.rank, sb_lo + i * (BK * 128), mb_lo, &full_bar[stage], nb0 + 64 * i, kcb + shift, c2, c3);
            }
          }
        }
        }   // sub
        if (++stage == C::STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer (leader CTA only) ===========================
    if (rank == 0 && lane == 0 && n_iters > 0) {
      const uint32_t idesc = make_idesc_bf16(BN, p.a.major == MTTS_MAJOR_MN, p.b.major == MTTS_MAJOR_MN, 256);
      const uint32_t a_step = (p.a.major == MTTS_MAJOR_K) ? UMMA_K * 2 : UMMA_K * 128;
      const uint32_t b_step = (p.b.major == MTTS_MAJOR_K) ? UMMA_K * 2 : UMMA_K * 128;
      const uint32_t a_lbo = (p.a.major == MTTS_MAJOR_K) ? 16 : BK * 128;
      const uint32_t b_lbo = (p.b.major == MTTS_MAJOR_K) ? 16 : BK * 128;
      int stage = 0;
      uint32_t phase = 0;
      uint32_t accumulate = 0;
      for (int it = 0; it < n_iters; ++it) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (!(p.dbg & 16))
#pragma unroll
        for (int sub = 0; sub < KD; ++sub) {
        const uint32_t sa = smem_u32(smem + stage * C::STAGE + sub * C::SUB);
        const uint32_t sa_lo = sa + C::A_TILE;
        const uint32_t sb = sa + C::A_TILE * (SPLIT == 3 ? 2 : 1);
        const uint32_t sb_lo = sb + C::B_TILE;
#pragma unroll
        for (int kk = 0; kk < BK / UMMA_K; ++kk) {
          const uint64_t da = make_umma_desc(sa + kk * a_step, a_lbo, 1024);
          const uint64_t db = make_umma_desc(sb + kk * b_step, b_lbo, 1024);
          umma_bf16_2cta(tmem_base, da, db, idesc, accumulate);
          accumulate = 1;
          if (SPLIT == 3) {
            const uint64_t da_lo = make_umma_desc(sa_lo + kk * a_step, a_lbo, 1024);
            const uint64_t db_lo = make_umma_desc(sb_lo + kk * b_step, b_lbo, 1024);
            umma_bf16_2cta(tmem_base, da, db_lo, idesc, 1);
            umma_bf16_2cta(tmem_base, da_lo, db, idesc, 1);
          }
        }
        }   // sub
        umma_commit_2cta(&empty_bar[stage], 3);   // both CTAs may refill this slot
        if (++stage == C::STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      umma_commit_2cta(tmem_full_bar, 3);          // both CTAs' epilogues
    }
  } else {
    const int q = warp & 3;
    if (n_iters > 0) {
      mbar_wait(tmem_full_bar, 0);
      tc_fence_after();
      // two warps share a TMEM lane quarter (warp % 4) and split the BN columns between them
      constexpr int CH = BN / 32;
      const int half = (warp - 2) >> 2;
      const int c_begin = CH >= 2 ? half * (CH / 2) : 0;
      const int c_end = CH >= 2 ? c_begin + CH / 2 : (half == 0 ? CH : 0);
      epilogue_tile<BN>(p, tmem_base, q, lane, m0, n0, z0, z1, c_begin, c_end,
                        reinterpret_cast<float4*>(smem) + (warp - 2) * 256);
    }
  }

  tc_fence_before();
  cluster_sync_all();                            // no CTA exits while its peer may still signal / read its smem
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta<BN>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(f);
  }
  return fn;
}

int encode_operand_map(CUtensorMap* map, const void* ptr, const mtts_operand& op, int block_mn, const char* name, int rank = 4) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) {
    mtts_set_error("cuTensorMapEncodeTiled entry point unavailable");
    return MTTS_ECUDA;
  }
  MTTS_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "gemm: operand %s base not 16B aligned", name);
  MTTS_REQUIRE(op.strides[0] == 1, "gemm: operand %s strides[0] must be 1", name);
  cuuint64_t gdim[4];
  cuuint64_t gstride[3];
  for (int i = 0; i < 4; ++i) {
    MTTS_REQUIRE(op.dims[i] >= 1, "gemm: operand %s dims[%d]=%lld < 1", name, i, (long long)op.dims[i]);
    gdim[i] = static_cast<cuuint64_t>(op.dims[i]);
  }
  for (int i = 1; i < 4; ++i) {
    int64_t st = op.strides[i];
    if (op.dims[i] == 1 && (st <= 0 || (st % 8) != 0)) st = 8;   // unused dim: any legal stride
    MTTS_REQUIRE(st > 0 && (st % 8) == 0, "gemm: operand %s strides[%d]=%lld must be a positive multiple of 8",
                 name, i, (long long)st);
    gstride[i - 1] = static_cast<cuuint64_t>(st) * 2;
  }
  cuuint32_t box[4] = {64u, op.major == MTTS_MAJOR_K ? static_cast<cuuint32_t>(block_mn) : 64u, 1u, 1u};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(ptr), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    mtts_set_error("cuTensorMapEncodeTiled(%s) failed: CUresult %d (dims %lld %lld %lld %lld strides %lld %lld %lld)",
                   name, (int)r, (long long)op.dims[0], (long long)op.dims[1], (long long)op.dims[2],
                   (long long)op.dims[3], (long long)op.strides[1], (long long)op.strides[2],
                   (long long)op.strides[3]);
    return MTTS_ECUDA;
  }
  return MTTS_OK;
}

template <int BN, int SPLIT, int KD = 1>
int launch_pair(const GemmParams& p, dim3 grid, cudaStream_t stream) {
  using C = Cfg2<BN, SPLIT, KD>;
  static bool configured = false;
  if (!configured) {
    MTTS_CHECK_CUDA(cudaFuncSetAttribute(mtts_gemm_pair_kernel<BN, SPLIT, KD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         C::SMEM));
    configured = true;
  }
  MTTS_CHECK_CUDA(mtts_launch(mtts_gemm_pair_kernel<BN, SPLIT, KD>, dim3(grid), dim3(NUM_THREADS), C::SMEM, stream, p));   // __cluster_dims__(2,1,1)
  MTTS_CHECK_LAUNCH();
  return MTTS_OK;
}

template <int BN, int SPLIT, int KD = 1>
int launch(const GemmParams& p, dim3 grid, cudaStream_t stream) {
  using C = Cfg<BN, SPLIT, KD>;
  static bool configured = false;
  if (!configured) {
    MTTS_CHECK_CUDA(cudaFuncSetAttribute(mtts_gemm_kernel<BN, SPLIT, KD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         C::SMEM));
    configured = true;
  }
  if (p.flags & EPI_LN) {
    // the four 64-wide tiles of a row form one cluster (consecutive blockIdx.x): launch-time cluster dimension + PDL
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = C::SMEM;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = LN_CLUSTER;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = mtts_pdl_enabled() ? 2 : 1;
    MTTS_CHECK_CUDA(cudaLaunchKernelEx(&cfg, mtts_gemm_kernel<BN, SPLIT, KD>, p));
  } else {
    MTTS_CHECK_CUDA(mtts_launch(mtts_gemm_kernel<BN, SPLIT, KD>, dim3(grid), dim3(NUM_THREADS), C::SMEM, stream, p));
  }
  MTTS_CHECK_LAUNCH();
  return MTTS_OK;
}

int gemm_impl(const mtts_gemm_desc* d, const mtts_ln_epilogue* ln, cudaStream_t stream);

}  // namespace

extern "C" int mtts_gemm(const mtts_gemm_desc* d, mtts_stream stream_) {
  return gemm_impl(d, nullptr, static_cast<cudaStream_t>(stream_));
}

extern "C" int mtts_gemm_ln(const mtts_gemm_desc* d, const mtts_ln_epilogue* ln, mtts_stream stream_) {
  MTTS_REQUIRE(d != nullptr && ln != nullptr, "gemm_ln: null descriptor");
  MTTS_REQUIRE(d->N == 64 * LN_CLUSTER && d->ldc == d->N, "gemm_ln: N must be %d with ldc == N (N %d, ldc %lld)", 64 * LN_CLUSTER, d->N,
               static_cast<long long>(d->ldc));
  MTTS_REQUIRE(d->nz0 == 1 && d->nz1 == 1 && d->ksplit <= 1 && d->pair == 0 && d->flags == 0,
               "gemm_ln: needs nz0 == nz1 == 1, ksplit == 1, pair == 0, flags == 0");
  MTTS_REQUIRE(ln->gamma && ln->beta && ln->T > 0, "gemm_ln: gamma / beta / T missing");
  const uintptr_t al = reinterpret_cast<uintptr_t>(d->c_f32) | reinterpret_cast<uintptr_t>(d->c_hi) | reinterpret_cast<uintptr_t>(d->c_lo) |
                       reinterpret_cast<uintptr_t>(d->bias) | reinterpret_cast<uintptr_t>(ln->res) | reinterpret_cast<uintptr_t>(ln->gamma) |
                       reinterpret_cast<uintptr_t>(ln->beta) | reinterpret_cast<uintptr_t>(ln->z_out);
  MTTS_REQUIRE((al & 15) == 0, "gemm_ln: outputs, bias, residual, gamma and beta must be 16-byte aligned");
  return gemm_impl(d, ln, static_cast<cudaStream_t>(stream_));
}

namespace {

int gemm_impl(const mtts_gemm_desc* d, const mtts_ln_epilogue* ln, cudaStream_t stream) {
  MTTS_REQUIRE(d != nullptr, "gemm: null descriptor");
  MTTS_REQUIRE(d->M > 0 && d->N > 0 && d->K > 0, "gemm: bad M/N/K %d %d %d", d->M, d->N, d->K);
  MTTS_REQUIRE(d->ntaps >= 1 && d->nkb >= 1 && d->nz0 >= 1 && d->nz1 >= 1, "gemm: bad loop extents");
  MTTS_REQUIRE(d->split == 1 || d->split == 3, "gemm: split must be 1 or 3");
  MTTS_REQUIRE(d->a.hi && d->b.hi, "gemm: null operand");
  MTTS_REQUIRE(d->split == 1 || (d->a.lo && d->b.lo), "gemm: split=3 needs lo operands");
  MTTS_REQUIRE(d->c_f32 || d->c_hi, "gemm: no output");
  MTTS_REQUIRE(!(d->c_lo && !d->c_hi), "gemm: c_lo without c_hi");
  MTTS_REQUIRE(!(d->flags & MTTS_EPI_GATE) || d->gate, "gemm: GATE flag without gate pointer");
  const int ksplit = (d->ksplit < 1 || mtts_deterministic()) ? 1 : d->ksplit;   // deterministic mode: one CTA per output tile
  MTTS_REQUIRE(ksplit == 1 || ((d->flags & MTTS_EPI_ACCUM) && d->c_f32 && !d->c_hi),
               "gemm: ksplit>1 requires ACCUM into c_f32 only");
  MTTS_REQUIRE(!(d->flags & MTTS_EPI_ACCUM) || d->c_f32, "gemm: ACCUM requires c_f32");
  MTTS_REQUIRE(!(d->flags & MTTS_EPI_ADD_C) || (d->c_f32 && ksplit == 1 && !(d->flags & MTTS_EPI_ACCUM)),
               "gemm: ADD_C requires c_f32, ksplit == 1 and no ACCUM");

  int bn = ln ? 64 : d->block_n;
  if (bn == 0) {
    if (d->N <= 64) bn = 64;
    else if (d->N <= 128) bn = 128;
    else bn = 256;
  }
  MTTS_REQUIRE(bn == 64 || bn == 128 || bn == 256, "gemm: block_n must be 64/128/256");
  const bool pair = d->pair != 0;
  MTTS_REQUIRE(!pair || bn >= 128, "gemm: the 2-CTA kernel needs block_n 128 or 256");

  GemmParams p;
  memset(&p, 0, sizeof(p));
  int rc;
  {
    const char* e = getenv("MTTS_GEMM_DBG");
    p.dbg = e ? atoi(e) : 0;
  }
  // smallest tensor-map rank that covers each operand's geometry (MTTS_GEMM_DBG bit 256 forces rank 4)
  auto op_rank = [&](const mtts_operand& o) {
    if (p.dbg & 256) return 4;
    if (o.dims[3] > 1 || o.src3 != MTTS_SRC_ZERO) return 4;
    if (o.dims[2] > 1 || o.src2 != MTTS_SRC_ZERO) return 3;
    return 2;
  };
  const int ra = op_rank(d->a), rb = op_rank(d->b);
  if ((rc = encode_operand_map(&p.map_a_hi, d->a.hi, d->a, BM, "A.hi", ra)) != MTTS_OK) return rc;
  const int b_rows = pair ? bn / 2 : bn;         // 2-CTA: each CTA stages half of the B tile
  if ((rc = encode_operand_map(&p.map_b_hi, d->b.hi, d->b, b_rows, "B.hi", rb)) != MTTS_OK) return rc;
  if (d->split == 3) {
    if ((rc = encode_operand_map(&p.map_a_lo, d->a.lo, d->a, BM, "A.lo", ra)) != MTTS_OK) return rc;
    if ((rc = encode_operand_map(&p.map_b_lo, d->b.lo, d->b, b_rows, "B.lo", rb)) != MTTS_OK) return rc;
  }
  const bool two = d->a2_hi != nullptr || d->b2_hi != nullptr;
  if (two) {
    MTTS_REQUIRE(d->a2_hi && d->b2_hi && (d->split == 1 || (d->a2_lo && d->b2_lo)), "gemm: incomplete second term");
    if ((rc = encode_operand_map(&p.map_a2_hi, d->a2_hi, d->a, BM, "A2.hi", ra)) != MTTS_OK) return rc;
    if ((rc = encode_operand_map(&p.map_b2_hi, d->b2_hi, d->b, b_rows, "B2.hi", rb)) != MTTS_OK) return rc;
    if (d->split == 3) {
      if ((rc = encode_operand_map(&p.map_a2_lo, d->a2_lo, d->a, BM, "A2.lo", ra)) != MTTS_OK) return rc;
      if ((rc = encode_operand_map(&p.map_b2_lo, d->b2_lo, d->b, b_rows, "B2.lo", rb)) != MTTS_OK) return rc;
    }
  }
  p.nterms = two ? 2 : 1;
  p.a = {d->a.major, d->a.src2, d->a.src3, d->a.shift_src, d->a.shift_base, d->a.shift_step, ra};
  p.b = {d->b.major, d->b.src2, d->b.src3, d->b.shift_src, d->b.shift_base, d->b.shift_step, rb};
  p.M = d->M; p.N = d->N; p.K = d->K;
  p.ntaps = d->ntaps; p.nkb = d->nkb; p.nz0 = d->nz0; p.nz1 = d->nz1;
  // narrow 1-CTA tiles run two k-blocks per stage fill (Cfg::KD)
  const int kd = (((!pair && bn == 64) || (pair && bn == 128)) && mtts_cdiv(d->K, BK) >= 2 && !(p.dbg & 512)) ? 2 : 1;
  const int kchunks = mtts_cdiv(mtts_cdiv(d->K, BK), kd);
  const int total_iters = p.nterms * d->ntaps * d->nkb * kchunks;
  p.ksplit = ksplit > total_iters ? total_iters : ksplit;
  // every split must own >= 1 iteration
  while (p.ksplit > 1 && mtts_cdiv(total_iters, p.ksplit) * (p.ksplit - 1) >= total_iters) --p.ksplit;
  p.n_tiles = mtts_cdiv(d->N, bn);
  p.flags = d->flags;
  p.alpha = d->alpha;
  p.c_f32 = d->c_f32;
  p.c_hi = static_cast<bf16*>(d->c_hi);
  p.c_lo = static_cast<bf16*>(d->c_lo);
  p.ldc = d->ldc; p.c_sz0 = d->c_sz0; p.c_sz1 = d->c_sz1;
  p.bias = d->bias; p.bias_sz0 = d->bias_sz0;
  p.gate = static_cast<const bf16*>(d->gate);
  if (ln) {
    p.flags |= EPI_LN;
    p.ln_res = ln->res; p.ln_gamma = ln->gamma; p.ln_beta = ln->beta; p.ln_lens = ln->lens;
    p.ln_z = ln->z_out; p.ln_stats = ln->stats; p.ln_T = ln->T; p.ln_eps = ln->eps;
    p.ln_drop = DropSite{ln->drop_thr, ln->drop_seed, ln->drop_scale, ln->drop_salt};
  }

  const int m_tiles = mtts_cdiv(d->M, BM);
  dim3 grid(m_tiles * p.n_tiles, p.ksplit, d->nz0 * d->nz1);
  MTTS_REQUIRE(grid.z <= 65535 && grid.y <= 65535, "gemm: grid too large");
  if (pair) {
    grid.x = 2 * mtts_cdiv(m_tiles, 2) * p.n_tiles;
    if (bn == 128 && kd == 2) return d->split == 1 ? launch_pair<128, 1, 2>(p, grid, stream) : launch_pair<128, 3, 2>(p, grid, stream);
    if (d->split == 1) return bn == 128 ? launch_pair<128, 1>(p, grid, stream) : launch_pair<256, 1>(p, grid, stream);
    return bn == 128 ? launch_pair<128, 3>(p, grid, stream) : launch_pair<256, 3>(p, grid, stream);
  }

  if (d->split == 1) {
    if (bn == 64) return kd == 2 ? launch<64, 1, 2>(p, grid, stream) : launch<64, 1>(p, grid, stream);
    if (bn == 128) return launch<128, 1>(p, grid, stream);
    return launch<256, 1>(p, grid, stream);
  } else {
    if (bn == 64) return kd == 2 ? launch<64, 3, 2>(p, grid, stream) : launch<64, 3>(p, grid, stream);
    if (bn == 128) return launch<128, 3>(p, grid, stream);
    return launch<256, 3>(p, grid, stream);
  }
}

}  // namespace
