// mtts_attn.cu — fused scaled-dot-product attention on tcgen05: scores, masked softmax and the value / gradient
// products in ONE kernel per pass, the [T, T] score matrix never leaves the SM.
//
// Replaces, for the forward and backward passes of MultiHeadAttention (SubLayers.py:29-57 ->
// ScaledDotProductAttention.forward, Modules.py:14-25: bmm(q, k^T) / temperature, masked_fill(mask, -inf), softmax(dim=2),
// bmm(attn, v)) the three-launch chain  mtts_gemm (scores) -> mtts_softmax -> mtts_gemm (P V)  and its autograd.
//
// One kernel template, three modes — all are "resident tile x streamed tiles" sweeps of the same shape:
//
//   mode   resident (128 rows)     streamed (NS rows / step)   SS products (smem x smem -> TMEM)     TS products (TMEM x smem -> TMEM)
//   FWD    Q_i                     K_j, V_j    (NS = 64)       S = Q K^T                             O  += P V
//   DQ     Q_i, dO_i               K_j, V_j    (NS = 32)       S = Q K^T, dP = dO V^T                dQ += dS K
//   DKV    K_j, V_j                Q_i, dO_i   (NS = 32)       S^T = K Q^T, dP^T = V dO^T            dV += P^T dO, dK += dS^T Q
//
// The probabilities never touch shared memory: the softmax warps read the fp32 score tile from TMEM (tcgen05.ld, one
// thread per row), and write P (resp. dS) back IN PLACE as packed bf16 hi | lo halves (tcgen05.st), from where the next
// tcgen05.mma reads it as its A operand (A-from-TMEM, "TS" form).  bf16x3: every product is hi*hi + hi*lo + lo*hi.
//
// FWD makes two sweeps over the keys: sweep 0 computes the row statistics (running max / sum, registers only), sweep 1
// recomputes S, writes the NORMALISED P and accumulates O — no accumulator rescaling, and P can be emitted to global
// memory (hi/lo, [B,H,T,Tp]) for the tapes the Hessian-vector passes re-read.  The backward kernels recompute P from
// Q, K and the saved log-sum-exp, FlashAttention-style; D = rowsum(dO * O) comes from a small pre-kernel.
//
// CTA = 192 threads: warp 0 TMA producer, warp 1 MMA issuer, warps 2..5 softmax / epilogue (TMEM lane quarter = warp % 4).
#include <math.h>
#include "mtts_common.cuh"

namespace {

constexpr int BM = 128;            // resident rows = UMMA_M
constexpr int DK = 128;            // head dimension (d_k = d_v), two 64-column k-blocks
constexpr int ATT_THREADS = 192;
constexpr int ATT_MAX_SMEM = 227 * 1024;

enum { ATT_FWD = 0, ATT_DQ = 1, ATT_DKV = 2 };

struct alignas(64) AttnParams {
  CUtensorMap map_qkv_hi, map_qkv_lo, map_do_hi, map_do_lo;
  int32_t B, H, T, Tp, Tl;
  float cs;                  // scale * log2(e): scores are exponentiated in the log2 domain
  float scale;
  const int64_t* klens;
  bf16* o_hi;
  bf16* o_lo;
  float* lse;                // [B,H,Tl]  log2-domain log-sum-exp of the scaled scores
  bf16* p_hi;                // FWD emit
  bf16* p_lo;
  const float* dvec;         // [B,H,Tl]  rowsum(dO * O)
  bf16* dqkv_hi;
  bf16* dqkv_lo;
  float* dp;                 // DQ emit
  bf16* ds_hi;
  bf16* ds_lo;
};

template <int MODE, int SPLIT>
struct ACfg {
  static constexpr int NS = MODE == ATT_FWD ? 64 : 32;             // streamed rows per step
  static constexpr int NRES = MODE == ATT_FWD ? 1 : 2;             // resident operand tiles
  static constexpr int NSS = MODE == ATT_FWD ? 1 : 2;              // SS products per step
  static constexpr int PARTS = SPLIT == 3 ? 2 : 1;                 // hi (, lo)
  static constexpr int RES_KB = BM * 128;                          // one 64-column k-block of a resident tile: [128 rows x 128 B]
  static constexpr int RES_PART = 2 * RES_KB;
  static constexpr int RES_TILE = RES_PART * PARTS;
  static constexpr int RES_BYTES = NRES * RES_TILE;
  static constexpr int X_KB = NS * 128;                            // one k-block / 64-column chunk of a streamed tile
  static constexpr int X_PART = 2 * X_KB;
  static constexpr int X_TILE = X_PART * PARTS;
  static constexpr int STAGE = 2 * X_TILE;                         // two streamed operands per step
  static constexpr int BAR_BYTES = 1024;
  static constexpr int STAGES_RAW = (ATT_MAX_SMEM - 1024 - BAR_BYTES - RES_BYTES) / STAGE;
  static constexpr int STAGES = STAGES_RAW > 4 ? 4 : STAGES_RAW;
  static constexpr int SMEM = RES_BYTES + STAGES * STAGE + BAR_BYTES + 1024;
  static constexpr int TMEM_COLS = MODE == ATT_FWD ? 256 : 512;
  static constexpr int ACC_STRIDE = 64;                            // TMEM columns per score buffer (two buffers: columns 0..127)
  static constexpr int OUT1 = 128, OUT2 = 256;                     // accumulator columns
  static_assert(STAGES >= 2, "need at least a double buffer");
};

// ---- tcgen05 forms not in mtts_common.cuh ------------------------------------------------------------------------------
// D[tmem] (+)= A[tmem] * B[smem]^T : A is read from tensor memory (K-major: lane = row, two bf16 per 32-bit column)
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// thread i of the warp writes 16 / 32 consecutive columns of TMEM lane (base_lane + i)
__device__ __forceinline__ void tmem_st_32x32_x16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      :
      : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// pack 32 fp32 values into 16 + 16 words of bf16 pairs (hi halves, lo halves); element 2j sits in the low 16 bits of word j
template <int SPLIT>
__device__ __forceinline__ void pack_split32(const float* v, uint32_t* hi, uint32_t* lo) {
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    bf16 h0, l0, h1, l1;
    split_bf16(v[2 * j], h0, l0);
    split_bf16(v[2 * j + 1], h1, l1);
    hi[j] = uint32_t(__bfloat16_as_ushort(h0)) | (uint32_t(__bfloat16_as_ushort(h1)) << 16);
    if (SPLIT == 3) lo[j] = uint32_t(__bfloat16_as_ushort(l0)) | (uint32_t(__bfloat16_as_ushort(l1)) << 16);
  }
}
// 16 packed words (32 bf16) -> global, guarded per 8 elements against the row end `ncols` (ncols % 8 == 0, col0 % 8 == 0)
__device__ __forceinline__ void store_bf16x32(bf16* dst, const uint32_t* w, int col0, int ncols) {
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (col0 + 8 * k + 8 <= ncols)
      *reinterpret_cast<uint4*>(dst + 8 * k) = make_uint4(w[4 * k], w[4 * k + 1], w[4 * k + 2], w[4 * k + 3]);
}

// ================================================================================================
template <int MODE, int SPLIT, bool EMIT>
__global__ void __launch_bounds__(ATT_THREADS, 1) mtts_attn_kernel(const __grid_constant__ AttnParams p) {
  using C = ACfg<MODE, SPLIT>;
  constexpr int NS = C::NS;
  constexpr int NSWEEP = MODE == ATT_FWD ? 2 : 1;
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* res = smem;
  uint8_t* stages = smem + C::RES_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(stages + C::STAGES * C::STAGE);
  uint64_t* empty_bar = full_bar + C::STAGES;
  uint64_t* res_full = empty_bar + C::STAGES;
  uint64_t* s_full = res_full + 1;      // [2] score tile(s) of a step are in TMEM
  uint64_t* p_full = s_full + 2;        // [2] the softmax warps are done with the step (P / dS written, or statistics read)
  uint64_t* out_full = p_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(out_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int row0 = blockIdx.x * BM;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int T = p.T;
  const int H = p.H;
  const int n_steps = (T + NS - 1) / NS;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.map_qkv_hi);
    if (SPLIT == 3) tma_prefetch_desc(&p.map_qkv_lo);
    if (MODE != ATT_FWD) {
      tma_prefetch_desc(&p.map_do_hi);
      if (SPLIT == 3) tma_prefetch_desc(&p.map_do_lo);
    }
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(res_full, 1);
    mbar_init(&s_full[0], 1);
    mbar_init(&s_full[1], 1);
    mbar_init(&p_full[0], 4);            // one arrival per softmax warp
    mbar_init(&p_full[1], 4);
    mbar_init(out_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<C::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                                    // everything above overlapped the previous kernel's tail

  // tensor-map coordinate 2 of the q / k / v column blocks of head h ([B*T, 3*H*dk] buffer viewed as [B][3H][T][dk])
  const int zq = h, zk = H + h, zv = 2 * H + h;

  if (warp == 0) {
    // ===================================== TMA producer ==========================================
    if (lane == 0) {
      // resident tile(s): 128 rows starting at row0, loaded in NS-row boxes
      mbar_arrive_expect_tx(res_full, C::RES_BYTES);
#pragma unroll
      for (int r = 0; r < C::NRES; ++r) {
        const bool from_do = (MODE == ATT_DQ && r == 1);
        const int z = from_do ? h : (MODE == ATT_DKV ? (r == 0 ? zk : zv) : zq);
#pragma unroll
        for (int part = 0; part < C::PARTS; ++part) {
          const CUtensorMap* map = from_do ? (part ? &p.map_do_lo : &p.map_do_hi) : (part ? &p.map_qkv_lo : &p.map_qkv_hi);
#pragma unroll
          for (int kb = 0; kb < 2; ++kb)
#pragma unroll
            for (int rb = 0; rb < BM / NS; ++rb)
              tma_load_4d(res + r * C::RES_TILE + part * C::RES_PART + kb * C::RES_KB + rb * C::X_KB, map, res_full, kb * 64,
                          row0 + rb * NS, z, b);
        }
      }
      int slot = 0;
      uint32_t phase = 0;
      for (int sweep = 0; sweep < NSWEEP; ++sweep) {
        const int nops = (MODE == ATT_FWD && sweep == 0) ? 1 : 2;     // the statistics sweep needs K only
        for (int t = 0; t < n_steps; ++t) {
          mbar_wait(&empty_bar[slot], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[slot], nops * C::X_TILE);
          for (int x = 0; x < nops; ++x) {
            const bool from_do = (MODE == ATT_DKV && x == 1);
            const int z = from_do ? h : (MODE == ATT_DKV ? zq : (x == 0 ? zk : zv));
#pragma unroll
            for (int part = 0; part < C::PARTS; ++part) {
              const CUtensorMap* map = from_do ? (part ? &p.map_do_lo : &p.map_do_hi) : (part ? &p.map_qkv_lo : &p.map_qkv_hi);
#pragma unroll
              for (int kb = 0; kb < 2; ++kb)
                tma_load_4d(stages + slot * C::STAGE + x * C::X_TILE + part * C::X_PART + kb * C::X_KB, map, &full_bar[slot],
                            kb * 64, t * NS, z, b);
            }
          }
          if (++slot == C::STAGES) {
            slot = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer ============================================
    if (lane == 0) {
      constexpr uint32_t idesc_ss = make_idesc_bf16(NS, 0, 0, 128);      // [128 x NS] += A[128 x 16] B[NS x 16]^T, both K-major
      constexpr uint32_t idesc_ts = make_idesc_bf16(DK, 0, 1, 128);      // [128 x 128] += A(tmem)[128 x 16] B[16 x 128], B MN-major
      const uint32_t res_a = smem_u32(res);
      const uint32_t stg_a = smem_u32(stages);
      mbar_wait(res_full, 0);
      tc_fence_after();
      int slot_ss = 0, slot_ts = 0;
      uint32_t phase_ss = 0;
      uint32_t g = 0, gp = 0;                     // steps whose SS products were issued / whose softmax was awaited
      uint32_t acc1 = 0, acc2 = 0;                // accumulate flags of the two output accumulators

      auto issue_ss = [&]() {
        mbar_wait(&full_bar[slot_ss], phase_ss);
        tc_fence_after();
        const uint32_t xs = stg_a + slot_ss * C::STAGE;
        const uint32_t acc = tmem_base + (g & 1) * C::ACC_STRIDE;
#pragma unroll
        for (int prod = 0; prod < C::NSS; ++prod) {
          const uint32_t ra = res_a + prod * C::RES_TILE;
          const uint32_t xb = xs + prod * C::X_TILE;
          uint32_t accum = 0;
#pragma unroll
          for (int kb = 0; kb < 2; ++kb)
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const uint64_t da = make_umma_desc(ra + kb * C::RES_KB + kk * 32, 16, 1024);
              const uint64_t db = make_umma_desc(xb + kb * C::X_KB + kk * 32, 16, 1024);
              umma_bf16(acc + prod * NS, da, db, idesc_ss, accum);
              accum = 1;
              if (SPLIT == 3) {
                const uint64_t da_lo = make_umma_desc(ra + C::RES_PART + kb * C::RES_KB + kk * 32, 16, 1024);
                const uint64_t db_lo = make_umma_desc(xb + C::X_PART + kb * C::X_KB + kk * 32, 16, 1024);
                umma_bf16(acc + prod * NS, da, db_lo, idesc_ss, 1);
                umma_bf16(acc + prod * NS, da_lo, db, idesc_ss, 1);
              }
            }
        }
        umma_commit(&s_full[g & 1]);
        ++g;
        if (++slot_ss == C::STAGES) {
          slot_ss = 0;
          phase_ss ^= 1;
        }
      };
      // out[128 x 128] += A(tmem, packed hi | lo at a_col)[128 x NS] * X[NS x 128]   (X streamed tile, MN-major B operand)
      auto issue_ts = [&](uint32_t out_col, uint32_t a_col, uint32_t xb, uint32_t& accum) {
#pragma unroll
        for (int kk = 0; kk < NS / 16; ++kk) {
          const uint64_t db = make_umma_desc(xb + kk * 2048, C::X_KB, 1024);
          umma_bf16_ts(tmem_base + out_col, a_col + kk * 8, db, idesc_ts, accum);
          accum = 1;
          if (SPLIT == 3) {
            const uint64_t db_lo = make_umma_desc(xb + C::X_PART + kk * 2048, C::X_KB, 1024);
            umma_bf16_ts(tmem_base + out_col, a_col + kk * 8, db_lo, idesc_ts, 1);
            umma_bf16_ts(tmem_base + out_col, a_col + NS / 2 + kk * 8, db, idesc_ts, 1);
          }
        }
      };

      for (int sweep = 0; sweep < NSWEEP; ++sweep) {
        const bool main_sweep = (sweep == NSWEEP - 1);
        issue_ss();
        for (int t = 0; t < n_steps; ++t) {
          if (t + 1 < n_steps) issue_ss();        // keeps the tensor pipe busy while the softmax warps work on step t
          mbar_wait(&p_full[gp & 1], (gp >> 1) & 1);
          tc_fence_after();
          if (main_sweep) {
            const uint32_t xs = stg_a + slot_ts * C::STAGE;
            const uint32_t acc = tmem_base + (gp & 1) * C::ACC_STRIDE;
            if constexpr (MODE == ATT_FWD) {
              issue_ts(C::OUT1, acc, xs + C::X_TILE, acc1);                    // O  += P V
            } else if constexpr (MODE == ATT_DQ) {
              issue_ts(C::OUT1, acc + NS, xs, acc1);                           // dQ += dS K
            } else {
              issue_ts(C::OUT1, acc, xs + C::X_TILE, acc1);                    // dV += P^T dO
              issue_ts(C::OUT2, acc + NS, xs, acc2);                           // dK += dS^T Q
            }
          }
          umma_commit(&empty_bar[slot_ts]);       // the slot's tiles have been read by every product of the step
          if (++slot_ts == C::STAGES) slot_ts = 0;
          ++gp;
        }
      }
      umma_commit(out_full);
    }
  } else {
    // ===================================== softmax / epilogue warps ===============================
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    const int r = q * 32 + lane;
    const int row = row0 + r;                     // query row (FWD, DQ) / key row (DKV)
    const uint32_t lane_addr = tmem_base + (uint32_t(q * 32) << 16);
    const int klen = p.klens ? static_cast<int>(min(static_cast<long long>(T), static_cast<long long>(p.klens[b]))) : T;
    const long long zrow = (static_cast<long long>(b) * H + h);
    const float cs = p.cs;
    uint32_t g = 0;
    auto arrive = [&]() {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[g & 1]);
      ++g;
    };

    if constexpr (MODE == ATT_FWD) {
      // ---- sweep 0: row statistics (log2 domain) ----
      float m = -INFINITY, l = 0.f;
      for (int t = 0; t < n_steps; ++t) {
        mbar_wait(&s_full[g & 1], (g >> 1) & 1);
        tc_fence_after();
        const uint32_t acc = lane_addr + (g & 1) * C::ACC_STRIDE;
        uint32_t v0[32], v1[32];
        tmem_ld_32x32(acc, v0);
        tmem_ld_32x32(acc + 32, v1);
        tmem_ld_wait();
        arrive();                                 // the score buffer may be overwritten
        const int c0 = t * NS;
        float mt = m;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          if (c0 + j < klen) mt = fmaxf(mt, __uint_as_float(v0[j]) * cs);
          if (c0 + 32 + j < klen) mt = fmaxf(mt, __uint_as_float(v1[j]) * cs);
        }
        if (mt > -INFINITY) {
          float s = 0.f;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (c0 + j < klen) s += exp2f(fmaf(__uint_as_float(v0[j]), cs, -mt));
            if (c0 + 32 + j < klen) s += exp2f(fmaf(__uint_as_float(v1[j]), cs, -mt));
          }
          l = l * exp2f(m - mt) + s;
          m = mt;
        }
      }
      const float L2 = m + log2f(l);
      if (row < T) p.lse[zrow * p.Tl + row] = L2;
      // ---- sweep 1: normalised P in place, O += P V ----
      for (int t = 0; t < n_steps; ++t) {
        mbar_wait(&s_full[g & 1], (g >> 1) & 1);
        tc_fence_after();
        const uint32_t acc = lane_addr + (g & 1) * C::ACC_STRIDE;
        const int c0 = t * NS;
        uint32_t hi[32], lo[32];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t v[32];
          tmem_ld_32x32(acc + 32 * half, v);
          tmem_ld_wait();
          float pv[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) pv[j] = (c0 + 32 * half + j < klen) ? exp2f(fmaf(__uint_as_float(v[j]), cs, -L2)) : 0.f;
          pack_split32<SPLIT>(pv, hi + 16 * half, lo + 16 * half);
        }
        // both halves have been read: overwrite the fp32 scores with the packed operand  [hi: 32 columns | lo: 32 columns]
        tmem_st_32x32_x16(acc, hi);
        tmem_st_32x32_x16(acc + 16, hi + 16);
        if (SPLIT == 3) {
          tmem_st_32x32_x16(acc + 32, lo);
          tmem_st_32x32_x16(acc + 48, lo + 16);
        }
        tmem_st_wait();
        arrive();
        if (EMIT && row < T) {
          const long long off = (zrow * T + row) * p.Tp + c0;
          store_bf16x32(p.p_hi + off, hi, c0, p.Tp);
          store_bf16x32(p.p_hi + off + 32, hi + 16, c0 + 32, p.Tp);
          if (SPLIT == 3) {
            store_bf16x32(p.p_lo + off, lo, c0, p.Tp);
            store_bf16x32(p.p_lo + off + 32, lo + 16, c0 + 32, p.Tp);
          }
        }
      }
    } else {
      // ---- backward: P recomputed from the saved log-sum-exp, dS = P * (dP - D) ----
      float L2r = 0.f, Dr = 0.f;
      if (MODE == ATT_DQ && row < T) {
        L2r = p.lse[zrow * p.Tl + row];
        Dr = p.dvec[zrow * p.Tl + row];
      }
      const bool key_ok = row < klen;             // DKV: this thread's key is not masked
      for (int t = 0; t < n_steps; ++t) {
        mbar_wait(&s_full[g & 1], (g >> 1) & 1);
        tc_fence_after();
        const uint32_t acc = lane_addr + (g & 1) * C::ACC_STRIDE;
        const int c0 = t * NS;
        uint32_t a1[32], a2[32];
        tmem_ld_32x32(acc, a1);
        tmem_ld_32x32(acc + NS, a2);
        tmem_ld_wait();
        float pv[32], dsv[32];
        if constexpr (MODE == ATT_DQ) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            pv[j] = (c0 + j < klen) ? exp2f(fmaf(__uint_as_float(a1[j]), cs, -L2r)) : 0.f;
            dsv[j] = pv[j] * (__uint_as_float(a2[j]) - Dr);
          }
        } else {
          const float4* L4 = reinterpret_cast<const float4*>(p.lse + zrow * p.Tl + c0);
          const float4* D4 = reinterpret_cast<const float4*>(p.dvec + zrow * p.Tl + c0);
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 lv = __ldg(L4 + j4);
            const float4 dv = __ldg(D4 + j4);
            const float ls[4] = {lv.x, lv.y, lv.z, lv.w};
            const float ds[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int j = 4 * j4 + u;
              pv[j] = (key_ok && c0 + j < T) ? exp2f(fmaf(__uint_as_float(a1[j]), cs, -ls[u])) : 0.f;
              dsv[j] = pv[j] * (__uint_as_float(a2[j]) - ds[u]);
            }
          }
        }
        uint32_t hi[16], lo[16];
        if constexpr (MODE == ATT_DKV) {           // P^T in place of S^T
          pack_split32<SPLIT>(pv, hi, lo);
          tmem_st_32x32_x16(acc, hi);
          if (SPLIT == 3) tmem_st_32x32_x16(acc + 16, lo);
        }
        pack_split32<SPLIT>(dsv, hi, lo);          // dS (dS^T) in place of dP (dP^T)
        tmem_st_32x32_x16(acc + NS, hi);
        if (SPLIT == 3) tmem_st_32x32_x16(acc + NS + 16, lo);
        tmem_st_wait();
        arrive();
        if (EMIT && MODE == ATT_DQ && row < T) {
          const long long off = (zrow * T + row) * p.Tp + c0;
          store_bf16x32(p.ds_hi + off, hi, c0, p.Tp);
          if (SPLIT == 3) store_bf16x32(p.ds_lo + off, lo, c0, p.Tp);
#pragma unroll
          for (int k = 0; k < 8; ++k)
            if (c0 + 4 * k + 4 <= p.Tp)
              *reinterpret_cast<float4*>(p.dp + off + 4 * k) = make_float4(__uint_as_float(a2[4 * k]), __uint_as_float(a2[4 * k + 1]),
                                                                           __uint_as_float(a2[4 * k + 2]), __uint_as_float(a2[4 * k + 3]));
        }
      }
    }

    // ---- epilogue: accumulator rows -> bf16 hi/lo in global memory ----
    mbar_wait(out_full, 0);
    tc_fence_after();
    constexpr int NOUT = MODE == ATT_DKV ? 2 : 1;
#pragma unroll
    for (int o = 0; o < NOUT; ++o) {
      bf16* dst_hi;
      bf16* dst_lo;
      long long ld;
      int colblk;
      float alpha = 1.f;
      if constexpr (MODE == ATT_FWD) {
        dst_hi = p.o_hi; dst_lo = p.o_lo; ld = static_cast<long long>(H) * DK; colblk = h;
      } else {
        dst_hi = p.dqkv_hi; dst_lo = p.dqkv_lo; ld = 3LL * H * DK;
        if (MODE == ATT_DQ) { colblk = zq; alpha = p.scale; }
        else if (o == 0) { colblk = zv; }
        else { colblk = zk; alpha = p.scale; }
      }
      const long long off = (static_cast<long long>(b) * T + row) * ld + static_cast<long long>(colblk) * DK;
#pragma unroll 1
      for (int ch = 0; ch < 4; ++ch) {
        uint32_t v[32];
        tmem_ld_32x32(lane_addr + (o == 0 ? C::OUT1 : C::OUT2) + ch * 32, v);
        tmem_ld_wait();
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]) * alpha;
        uint32_t hi[16], lo[16];
        pack_split32<SPLIT>(f, hi, lo);
        if (row < T) {
          store_bf16x32(dst_hi + off + ch * 32, hi, 0, 32);
          if (SPLIT == 3 && dst_lo) store_bf16x32(dst_lo + off + ch * 32, lo, 0, 32);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<C::TMEM_COLS>(tmem_base);
  }
}

// D[b,h,t] = sum_c dO[b,t,h,c] * O[b,t,h,c]  (one warp per (row, head); the softmax-backward row term of Modules.py:22)
__global__ void __launch_bounds__(256) attn_dvec_kernel(const bf16* __restrict__ do_hi, const bf16* __restrict__ do_lo,
                                                        const bf16* __restrict__ o_hi, const bf16* __restrict__ o_lo, int B, int H, int T,
                                                        int Tl, float* __restrict__ dvec) {
  pdl_enter();
  const int lane = threadIdx.x & 31;
  const long long n = static_cast<long long>(B) * T * H;
  for (long long w = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5); w < n; w += static_cast<long long>(gridDim.x) * 8) {
    const int h = static_cast<int>(w % H);
    const long long bt = w / H;
    const long long i = (bt * H + h) * DK + lane * 4;
    auto ld4 = [&](const bf16* hi, const bf16* lo) {
      const uint2 a = *reinterpret_cast<const uint2*>(hi + i);
      float4 v = make_float4(__uint_as_float(a.x << 16), __uint_as_float(a.x & 0xFFFF0000u), __uint_as_float(a.y << 16),
                             __uint_as_float(a.y & 0xFFFF0000u));
      if (lo) {
        const uint2 c = *reinterpret_cast<const uint2*>(lo + i);
        v.x += __uint_as_float(c.x << 16); v.y += __uint_as_float(c.x & 0xFFFF0000u);
        v.z += __uint_as_float(c.y << 16); v.w += __uint_as_float(c.y & 0xFFFF0000u);
      }
      return v;
    };
    const float4 a = ld4(do_hi, do_lo), c = ld4(o_hi, o_lo);
    const float s = warp_sum(a.x * c.x + a.y * c.y + a.z * c.z + a.w * c.w);
    if (lane == 0) dvec[((bt / T) * H + h) * Tl + (bt % T)] = s;
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled attn_encode_fn() {
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

// [B*T, nblk*128] bf16 buffer viewed as [B][nblk][T][128]; box = 64 columns x box_rows rows
int encode_heads_map(CUtensorMap* map, const void* ptr, int B, int T, int nblk, int box_rows, const char* name) {
  PFN_encodeTiled enc = attn_encode_fn();
  if (!enc) {
    mtts_set_error("cuTensorMapEncodeTiled entry point unavailable");
    return MTTS_ECUDA;
  }
  MTTS_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "attn: %s base not 16B aligned", name);
  const cuuint64_t ld = static_cast<cuuint64_t>(nblk) * DK;
  cuuint64_t gdim[4] = {DK, static_cast<cuuint64_t>(T), static_cast<cuuint64_t>(nblk), static_cast<cuuint64_t>(B)};
  cuuint64_t gstride[3] = {ld * 2, DK * 2, static_cast<cuuint64_t>(T) * ld * 2};
  cuuint32_t box[4] = {64u, static_cast<cuuint32_t>(box_rows), 1u, 1u};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    mtts_set_error("attn: cuTensorMapEncodeTiled(%s) failed: CUresult %d (B %d T %d blocks %d)", name, (int)r, B, T, nblk);
    return MTTS_ECUDA;
  }
  return MTTS_OK;
}

template <int MODE, int SPLIT, bool EMIT>
int launch_attn(const AttnParams& p, cudaStream_t stream) {
  using C = ACfg<MODE, SPLIT>;
  static bool configured = false;
  if (!configured) {
    MTTS_CHECK_CUDA(cudaFuncSetAttribute(mtts_attn_kernel<MODE, SPLIT, EMIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    configured = true;
  }
  dim3 grid(mtts_cdiv(p.T, BM), p.H, p.B);
  MTTS_CHECK_CUDA(mtts_launch(mtts_attn_kernel<MODE, SPLIT, EMIT>, grid, dim3(ATT_THREADS), C::SMEM, stream, p));
  MTTS_CHECK_LAUNCH();
  return MTTS_OK;
}

int fill_common(AttnParams& p, const mtts_attn_desc* d, int box_rows, bool need_do) {
  MTTS_REQUIRE(d != nullptr, "attn: null descriptor");
  MTTS_REQUIRE(d->B > 0 && d->H > 0 && d->T > 0, "attn: bad B/H/T %d %d %d", d->B, d->H, d->T);
  MTTS_REQUIRE(d->dk == DK, "attn: head dimension must be %d (got %d)", DK, d->dk);
  MTTS_REQUIRE(d->split == 1 || d->split == 3, "attn: split must be 1 or 3");
  MTTS_REQUIRE(d->qkv_hi && (d->split == 1 || d->qkv_lo), "attn: missing q/k/v operand");
  MTTS_REQUIRE(d->lse && d->Tl >= ((d->T + 127) / 128) * 128 && (d->Tl & 3) == 0, "attn: lse buffer / Tl (need Tl >= T rounded up to 128)");
  MTTS_REQUIRE(d->H <= 65535 && d->B <= 65535, "attn: grid too large");
  memset(&p, 0, sizeof(p));
  int rc;
  if ((rc = encode_heads_map(&p.map_qkv_hi, d->qkv_hi, d->B, d->T, 3 * d->H, box_rows, "qkv.hi")) != MTTS_OK) return rc;
  if (d->split == 3 && (rc = encode_heads_map(&p.map_qkv_lo, d->qkv_lo, d->B, d->T, 3 * d->H, box_rows, "qkv.lo")) != MTTS_OK) return rc;
  if (need_do) {
    MTTS_REQUIRE(d->do_hi && (d->split == 1 || d->do_lo), "attn: missing dO operand");
    if ((rc = encode_heads_map(&p.map_do_hi, d->do_hi, d->B, d->T, d->H, box_rows, "do.hi")) != MTTS_OK) return rc;
    if (d->split == 3 && (rc = encode_heads_map(&p.map_do_lo, d->do_lo, d->B, d->T, d->H, box_rows, "do.lo")) != MTTS_OK) return rc;
  }
  p.B = d->B; p.H = d->H; p.T = d->T; p.Tp = d->Tp; p.Tl = d->Tl;
  p.scale = d->scale;
  p.cs = d->scale * 1.4426950408889634f;
  p.klens = d->klens;
  p.lse = d->lse;
  return MTTS_OK;
}

}  // namespace

extern "C" int mtts_attn_fwd(const mtts_attn_desc* d, mtts_stream stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  AttnParams p;
  int rc = fill_common(p, d, 64, false);
  if (rc != MTTS_OK) return rc;
  MTTS_REQUIRE(d->o_hi && (d->split == 1 || d->o_lo), "attn_fwd: missing output");
  p.o_hi = static_cast<bf16*>(d->o_hi);
  p.o_lo = static_cast<bf16*>(d->o_lo);
  const bool emit = d->p_hi != nullptr;
  if (emit) {
    MTTS_REQUIRE(d->Tp >= d->T && (d->Tp & 7) == 0 && (d->split == 1 || d->p_lo), "attn_fwd: bad P emit buffers (Tp %d)", d->Tp);
    MTTS_REQUIRE(((reinterpret_cast<uintptr_t>(d->p_hi) | reinterpret_cast<uintptr_t>(d->p_lo)) & 15) == 0, "attn_fwd: P not 16B aligned");
    p.p_hi = static_cast<bf16*>(d->p_hi);
    p.p_lo = static_cast<bf16*>(d->p_lo);
  }
  MTTS_REQUIRE(((reinterpret_cast<uintptr_t>(d->o_hi) | reinterpret_cast<uintptr_t>(d->o_lo)) & 15) == 0, "attn_fwd: O not 16B aligned");
  if (d->split == 3) return emit ? launch_attn<ATT_FWD, 3, true>(p, stream) : launch_attn<ATT_FWD, 3, false>(p, stream);
  return emit ? launch_attn<ATT_FWD, 1, true>(p, stream) : launch_attn<ATT_FWD, 1, false>(p, stream);
}

extern "C" int mtts_attn_bwd(const mtts_attn_desc* d, int parts, mtts_stream stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  AttnParams p;
  int rc = fill_common(p, d, 32, true);
  if (rc != MTTS_OK) return rc;
  MTTS_REQUIRE(parts > 0 && parts < 8, "attn_bwd: parts is a mask of MTTS_ATTN_PREP | MTTS_ATTN_DQ | MTTS_ATTN_DKV");
  MTTS_REQUIRE(d->dvec, "attn_bwd: missing dvec");
  p.dvec = d->dvec;
  if (parts & MTTS_ATTN_PREP) {
    MTTS_REQUIRE(d->o_hi && (d->split == 1 || d->o_lo), "attn_bwd: PREP needs the forward output O");
    const long long n = static_cast<long long>(d->B) * d->T * d->H;
    const int grid = static_cast<int>(n / 8 + 1 < 148 * 8 ? n / 8 + 1 : 148 * 8);
    MTTS_CHECK_CUDA(mtts_launch(attn_dvec_kernel, dim3(grid), dim3(256), 0, stream, static_cast<const bf16*>(d->do_hi),
                                static_cast<const bf16*>(d->split == 3 ? d->do_lo : nullptr), static_cast<const bf16*>(d->o_hi),
                                static_cast<const bf16*>(d->split == 3 ? d->o_lo : nullptr), d->B, d->H, d->T, d->Tl, d->dvec));
    MTTS_CHECK_LAUNCH();
  }
  if (parts & (MTTS_ATTN_DQ | MTTS_ATTN_DKV)) {
    MTTS_REQUIRE(d->dqkv_hi && (d->split == 1 || d->dqkv_lo), "attn_bwd: missing dqkv output");
    MTTS_REQUIRE(((reinterpret_cast<uintptr_t>(d->dqkv_hi) | reinterpret_cast<uintptr_t>(d->dqkv_lo)) & 15) == 0, "attn_bwd: dqkv not 16B aligned");
    p.dqkv_hi = static_cast<bf16*>(d->dqkv_hi);
    p.dqkv_lo = static_cast<bf16*>(d->dqkv_lo);
  }
  if (parts & MTTS_ATTN_DQ) {
    const bool emit = d->ds_hi != nullptr;
    if (emit) {
      MTTS_REQUIRE(d->dp && d->Tp >= d->T && (d->Tp & 7) == 0 && (d->split == 1 || d->ds_lo), "attn_bwd: bad dP / dS emit buffers (Tp %d)", d->Tp);
      MTTS_REQUIRE(((reinterpret_cast<uintptr_t>(d->ds_hi) | reinterpret_cast<uintptr_t>(d->ds_lo) | reinterpret_cast<uintptr_t>(d->dp)) & 15) == 0,
                   "attn_bwd: dP / dS not 16B aligned");
      p.dp = d->dp;
      p.ds_hi = static_cast<bf16*>(d->ds_hi);
      p.ds_lo = static_cast<bf16*>(d->ds_lo);
    }
    if (d->split == 3) rc = emit ? launch_attn<ATT_DQ, 3, true>(p, stream) : launch_attn<ATT_DQ, 3, false>(p, stream);
    else rc = emit ? launch_attn<ATT_DQ, 1, true>(p, stream) : launch_attn<ATT_DQ, 1, false>(p, stream);
    if (rc != MTTS_OK) return rc;
  }
  if (parts & MTTS_ATTN_DKV) {
    rc = d->split == 3 ? launch_attn<ATT_DKV, 3, false>(p, stream) : launch_attn<ATT_DKV, 1, false>(p, stream);
    if (rc != MTTS_OK) return rc;
  }
  return MTTS_OK;
}
