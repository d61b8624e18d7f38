// mtts_attn.cu — fused scaled-dot-product attention on tcgen05: scores, masked softmax and the value / gradient
// products in ONE kernel per pass, the [T, T] score matrix never leaves the SM.
//
// Replaces, for the forward and backward passes of MultiHeadAttention (SubLayers.py:29-57 ->
// ScaledDotProductAttention.forward, Modules.py:14-25: bmm(q, k^T) / temperature, masked_fill(mask, -inf), softmax(dim=2),
// bmm(attn, v)) the three-launch chain  mtts_gemm (scores) -> mtts_softmax -> mtts_gemm (P V)  and its autograd.
//
// One kernel template, four modes — all are "resident 128-row tile x streamed tiles" sweeps of the same shape:
//
//   mode   resident (TMEM, A operand)   streamed (smem rings, NS rows / step)   score products                  accumulated product
//   FWD    Q_i                          K_j, V_j    (NS = 64)                   S = Q K^T                       O  += P V
//   DQ     Q_i, dO_i                    K_j, V_j    (NS = 32)                   S = Q K^T,   dP = dO V^T        dQ += dS K
//   DK     K_j, V_j                     Q_i, dO_i   (NS = 32)                   S^T = K Q^T, dP^T = V dO^T      dK += dS^T Q
//   DV     K_j                          Q_i, dO_i   (NS = 64)                   S^T = K Q^T                     dV += P^T dO
//
// EVERY tcgen05.mma reads its A operand from tensor memory: a shared-memory A operand costs the tensor pipe ~80 cycles per
// instruction whatever N is (measured: the first version of this kernel, A = resident tile in smem, ran 2.4 us per 32-key step for
// 0.6 us of tensor math), while a TMEM A operand leaves only the N x 32 B slice of B to fetch.  The resident tile is copied once
// from global memory into TMEM by the softmax warps (lane = row, two bf16 per 32-bit column = the packing the A operand wants);
// the probabilities never touch shared memory either: the softmax warps read the fp32 score tile from TMEM (tcgen05.ld, one thread
// per row) and write P (resp. dS) back IN PLACE as packed bf16 hi | lo halves (tcgen05.st), from where the next tcgen05.mma
// reads them.  bf16x3: every product is hi*hi + hi*lo + lo*hi.
//
// FWD makes two sweeps over the keys: sweep 0 computes the row statistics (running max / sum, registers only), sweep 1
// recomputes S, writes the NORMALISED P and accumulates O — no accumulator rescaling, and P can be emitted to global
// memory (hi/lo, [B,H,T,Tp]) for the tapes the Hessian-vector passes re-read.  The backward kernels recompute P from
// Q, K and the saved log-sum-exp, FlashAttention-style; D = rowsum(dO * O) comes from a small pre-kernel.  No atomics anywhere:
// the three gradient kernels own disjoint column blocks of dqkv and are bit-reproducible.
//
// CTA = 384 threads: warp 0 TMA producer (both rings), warps 1 and 10 issue the score products of the even / odd steps, warps 2..9
// softmax / epilogue (TMEM lane quarter = warp % 4; the two warps of a quarter split the columns of every score tile), warp 11
// issues the accumulated products.  SEVERAL issuing threads because tcgen05.mma issue blocks while the tensor pipe's short queue is
// full: a single issuer runs in lock-step with the pipe, and every mbarrier wait it makes between two bursts (~0.1 us each, three
// per step) is tensor-pipe idle time (measured with tools/attn_trace.py: 0.3-0.45 us idle per 0.6-0.8 us step).  With the score
// side and the accumulate side on separate threads, ordered only by mbarriers, one side's waits overlap the other side's MMAs.
#include <math.h>
#include "mtts_common.cuh"

// Optional pipeline trace (build with -DMTTS_ATTN_TRACE; tools/attn_trace.py): SM-clock timestamps of CTA (0,0,0).
#ifdef MTTS_ATTN_TRACE
__device__ long long g_attn_trace[16 * 128];
extern "C" int mtts_attn_trace_read(long long* dst) {
  return cudaMemcpyFromSymbol(dst, g_attn_trace, sizeof(g_attn_trace)) == cudaSuccess ? 0 : -2;
}
#define ATT_TRACE(kind, step) do { if (trace_on && (step) < 128) g_attn_trace[(kind) * 128 + (step)] = clock64(); } while (0)
#else
#define ATT_TRACE(kind, step) do { } while (0)
#endif

namespace {

constexpr int BM = 128;            // resident rows = UMMA_M
constexpr int DK = 128;            // head dimension (d_k = d_v), two 64-column k-blocks
constexpr int ATT_THREADS = 384;        // 12 warps
constexpr int ATT_MAX_SMEM = 227 * 1024;

enum { ATT_FWD = 0, ATT_DQ = 1, ATT_DK = 2, ATT_DV = 3 };

struct alignas(64) AttnParams {
  CUtensorMap map_qkv_hi, map_qkv_lo, map_do_hi, map_do_lo;
  int32_t B, H, T, Tp, Tl;
  float cs;                  // scale * log2(e): scores are exponentiated in the log2 domain
  float scale;
  const int64_t* klens;
  const bf16* qkv_hi;        // raw pointers: the resident tile goes global -> registers -> TMEM
  const bf16* qkv_lo;
  const bf16* do_hi;
  const bf16* do_lo;
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
  static constexpr bool WIDE = MODE == ATT_FWD || MODE == ATT_DV;   // one resident tile, one score product per step
  static constexpr int NS = WIDE ? 64 : 32;                        // streamed rows per step
  static constexpr int NRES = WIDE ? 1 : 2;                        // resident operand tiles = score products per step
  static constexpr int PARTS = SPLIT == 3 ? 2 : 1;                 // hi (, lo)
  static constexpr int X_KB = NS * 128;                            // one k-block / 64-column chunk of a streamed tile
  static constexpr int X_PART = 2 * X_KB;
  static constexpr int X_TILE = X_PART * PARTS;
  static constexpr int BAR_BYTES = 3072;                           // mbarriers, TMEM slot, [2][2][128] fp32 statistics exchange
  // two rings of streamed tiles: ring 1 = first operand (K; Q in DK / DV), ring 2 = second operand (V; dO in DK / DV)
  static constexpr int NTILES_RAW = (ATT_MAX_SMEM - 1024 - BAR_BYTES) / X_TILE;
  static constexpr int NTILES = NTILES_RAW > 8 ? 8 : NTILES_RAW;
  static constexpr int S1 = (NTILES + 1) / 2;
  static constexpr int S2 = NTILES - S1;
  static constexpr int SMEM = NTILES * X_TILE + BAR_BYTES + 1024;
  // tensor memory map (512 columns): resident operand(s) | score buffers | accumulator
  static constexpr int RES_COL = 0;                                // tile r: hi at 128 r, lo at 128 r + 64 (64 columns = 128 bf16 each)
  static constexpr int ACC0 = WIDE ? 128 : 256;
  static constexpr int NBUF = WIDE ? 4 : 2;                        // score buffers of 64 columns
  static constexpr int ACC_STRIDE = 64;
  static constexpr int OUT1 = 384;
  static constexpr int TMEM_COLS = 512;
  static_assert(S1 >= 2 && S2 >= 2, "rings must be at least double buffered");
  static_assert(ACC0 + NBUF * ACC_STRIDE <= OUT1, "tensor memory map overlaps");
};

// ---- tcgen05 forms not in mtts_common.cuh ------------------------------------------------------------------------------
// The MMA warp runs CONVERGED (all 32 lanes execute the issue loop, so descriptors stay in uniform registers without the
// per-instruction uniformisation loops the compiler emits inside a single-lane branch); `lead` predicates the instruction itself.
// D[tmem] (+)= A[tmem] * B[smem]^T : A is read from tensor memory (K-major: lane = row, two bf16 per 32-bit column)
__device__ __forceinline__ void umma_ts_pred(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate, uint32_t lead) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.ne.b32 q, %5, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(lead)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pred(uint64_t* bar, uint32_t lead) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "setp.ne.b32 q, %1, 0;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
      :
      : "r"(smem_u32(bar)), "r"(lead)
      : "memory");
}
// thread i of the warp writes 16 consecutive columns of TMEM lane (base_lane + i)
__device__ __forceinline__ void tmem_st_32x32_x16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      :
      : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
// advance the start-address field (bits 0..13, units of 16 B) of a shared-memory descriptor: the sum stays below 2^14 for any
// address inside the 228 KB window, so only the low word changes
__device__ __forceinline__ uint64_t desc_advance(uint64_t d, uint32_t bytes) {
  return (d & 0xFFFFFFFF00000000ull) | static_cast<uint64_t>(static_cast<uint32_t>(d) + (bytes >> 4));
}
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {      // non-blocking probe
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return done != 0;
}
__device__ __forceinline__ uint32_t elect_one_sync() {          // true in exactly one (converged) lane of the warp
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "elect.sync _|P1, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t}"
      : "=r"(pred));
  return pred;
}
__device__ __forceinline__ void tmem_st_32x32_x8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               :
               : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32_x16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// 32-byte store: one full sector per thread (two 16-byte stores from a thread whose rows are strided touch the sector twice)
__device__ __forceinline__ void st_global_256(void* dst, const uint32_t* w) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               :
               : "l"(dst), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float ex2_approx(float x) {          // 2^x, max relative error 2^-22; -inf -> 0
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// pack 2 NW fp32 values into NW + NW words of bf16 pairs (hi halves, lo halves); element 2j sits in the low 16 bits of word j
template <int SPLIT, int NW>
__device__ __forceinline__ void pack_split(const float* v, uint32_t* hi, uint32_t* lo) {
#pragma unroll
  for (int j = 0; j < NW; ++j) {
    const uint32_t h = pack_bf16x2(v[2 * j], v[2 * j + 1]);          // one cvt.rn.bf16x2.f32
    hi[j] = h;
    if (SPLIT == 3) lo[j] = pack_bf16x2(v[2 * j] - __uint_as_float(h << 16), v[2 * j + 1] - __uint_as_float(h & 0xFFFF0000u));
  }
}
// NW packed words (2 NW bf16, NW % 8 == 0) -> global as 32-byte stores, guarded against the row end `ncols` (ncols % 8 == 0,
// col0 % 16 == 0, dst 32-byte aligned when ncols % 16 == 0; a ragged last group falls back to one 16-byte store)
template <int NW>
__device__ __forceinline__ void store_bf16_row(bf16* dst, const uint32_t* w, int col0, int ncols, bool aligned32) {
#pragma unroll
  for (int k = 0; k < NW / 8; ++k) {
    const int c = col0 + 16 * k;
    if (c + 16 <= ncols && aligned32) {
      st_global_256(dst + 16 * k, w + 8 * k);
    } else {
      if (c + 8 <= ncols) *reinterpret_cast<uint4*>(dst + 16 * k) = make_uint4(w[8 * k], w[8 * k + 1], w[8 * k + 2], w[8 * k + 3]);
      if (c + 16 <= ncols) *reinterpret_cast<uint4*>(dst + 16 * k + 8) = make_uint4(w[8 * k + 4], w[8 * k + 5], w[8 * k + 6], w[8 * k + 7]);
    }
  }
}

// ================================================================================================
template <int MODE, int SPLIT, bool EMIT>
__global__ void __launch_bounds__(ATT_THREADS, 1) mtts_attn_kernel(const __grid_constant__ AttnParams p) {
  using C = ACfg<MODE, SPLIT>;
  constexpr int NS = C::NS;
  constexpr bool WIDE = C::WIDE;
  constexpr int NSWEEP = MODE == ATT_FWD ? 2 : 1;
  constexpr int PK = (SPLIT == 3 || EMIT) ? 3 : 1;      // the lo halves of P / dS exist when they are an MMA operand or are emitted
  // WIDE modes:   ring 1 feeds the score product only (free once it has been read), ring 2 feeds the accumulated product.
  // narrow modes: both rings feed the score products, ring 1 also feeds the accumulated product.
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ring1 = smem;
  uint8_t* ring2 = ring1 + C::S1 * C::X_TILE;
  uint64_t* full1 = reinterpret_cast<uint64_t*>(ring2 + C::S2 * C::X_TILE);
  uint64_t* empty1 = full1 + C::S1;
  uint64_t* full2 = empty1 + C::S1;
  uint64_t* empty2 = full2 + C::S2;
  uint64_t* res_ready = empty2 + C::S2;       // the resident operand(s) are in TMEM
  uint64_t* s_full = res_ready + 1;           // [NBUF] score tile(s) of a step are in TMEM
  uint64_t* p_full = s_full + C::NBUF;        // [NBUF] the softmax warps are done with the step (P / dS written, or statistics read)
  uint64_t* buf_free = p_full + C::NBUF;      // [NBUF] the accumulated product has consumed the buffer (statistics sweep: the softmax has read it)
  uint64_t* out_full = buf_free + C::NBUF;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(out_full + 1);
  float* stat_xchg = reinterpret_cast<float*>(tmem_slot + 2);     // [2][128] (m, l) of the other column half (FWD)

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);      // warp-uniform by construction
  const int lane = threadIdx.x & 31;
  const int row0 = blockIdx.x * BM;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int T = p.T;
  const int H = p.H;
  const int n_steps = (T + NS - 1) / NS;
#ifdef MTTS_ATTN_TRACE
  const bool trace_on = (blockIdx.x | blockIdx.y | blockIdx.z) == 0;
#endif

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.map_qkv_hi);
    if (SPLIT == 3) tma_prefetch_desc(&p.map_qkv_lo);
    if (MODE == ATT_DK || MODE == ATT_DV) {
      tma_prefetch_desc(&p.map_do_hi);
      if (SPLIT == 3) tma_prefetch_desc(&p.map_do_lo);
    }
    for (int s = 0; s < C::S1; ++s) {
      mbar_init(&full1[s], 1);
      mbar_init(&empty1[s], 1);
    }
    for (int s = 0; s < C::S2; ++s) {
      mbar_init(&full2[s], 1);
      mbar_init(&empty2[s], 1);
    }
    mbar_init(res_ready, 8);             // one arrival per softmax warp
    for (int s = 0; s < C::NBUF; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&p_full[s], 8);
      mbar_init(&buf_free[s], 1);
    }
    mbar_init(out_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<C::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  pdl_wait();                                    // everything above overlapped the previous kernel's tail
  if (threadIdx.x == 64) ATT_TRACE(11, 3);

  // column blocks of head h: q | k | v of the [B*T, 3*H*dk] buffer (= tensor-map coordinate 2 of its [B][3H][T][dk] view)
  const int zq = h, zk = H + h, zv = 2 * H + h;
  constexpr bool KEY_ROWS = MODE == ATT_DK || MODE == ATT_DV;          // resident rows are keys, streamed rows are queries

  if (warp == 0) {
    // ===================================== TMA producer: both rings, polled by one thread ================================
    // ring 1 = K tiles (Q tiles when the keys are resident), every step of every sweep; ring 2 = V tiles (dO tiles), main sweep only
    if (lane == 0) {
      const int z1 = KEY_ROWS ? zq : zk;
      const int z2 = KEY_ROWS ? h : zv;
      const int n1 = NSWEEP * n_steps, n2 = n_steps;
      int it1 = 0, it2 = 0;
      uint32_t spins = 0;
      while (it1 < n1 || it2 < n2) {
        bool progress = false;
        if (it1 < n1) {
          const int slot = it1 % C::S1;
          if (mbar_test(&empty1[slot], ((it1 / C::S1) & 1) ^ 1)) {
            const int t = it1 >= n_steps ? it1 - n_steps : it1;
            ATT_TRACE(0, it1);
            mbar_arrive_expect_tx(&full1[slot], C::X_TILE);
#pragma unroll
            for (int part = 0; part < C::PARTS; ++part)
#pragma unroll
              for (int kb = 0; kb < 2; ++kb)
                tma_load_4d(ring1 + slot * C::X_TILE + part * C::X_PART + kb * C::X_KB, part ? &p.map_qkv_lo : &p.map_qkv_hi, &full1[slot],
                            kb * 64, t * NS, z1, b);
            ++it1;
            progress = true;
          }
        }
        if (it2 < n2) {
          const int slot = it2 % C::S2;
          if (mbar_test(&empty2[slot], ((it2 / C::S2) & 1) ^ 1)) {
            ATT_TRACE(1, it2);
            mbar_arrive_expect_tx(&full2[slot], C::X_TILE);
#pragma unroll
            for (int part = 0; part < C::PARTS; ++part) {
              const CUtensorMap* map = KEY_ROWS ? (part ? &p.map_do_lo : &p.map_do_hi) : (part ? &p.map_qkv_lo : &p.map_qkv_hi);
#pragma unroll
              for (int kb = 0; kb < 2; ++kb)
                tma_load_4d(ring2 + slot * C::X_TILE + part * C::X_PART + kb * C::X_KB, map, &full2[slot], kb * 64, it2 * NS, z2, b);
            }
            ++it2;
            progress = true;
          }
        }
        if (progress) {
          spins = 0;
        } else if (++spins > (1u << 28)) {
          printf("mtts_attn: producer timeout block(%d,%d,%d)\n", blockIdx.x, blockIdx.y, blockIdx.z);
          __trap();
        }
      }
    }
  } else if (warp == 1 || warp == 10) {
    // ===================================== MMA issuers A0 / A1: score products of the even / odd steps (one elected lane
    // each; elect.sync keeps the descriptor arithmetic on the uniform datapath: ~3.5 instructions per tcgen05.mma instead
    // of ~15 inside a `lane == 0` branch).  Two threads, because each spends ~0.3 us per step in mbarrier round trips. ====
    if (elect_one_sync()) {
      constexpr uint32_t lead = 1u;
      constexpr uint32_t idesc_sc = make_idesc_bf16(NS, 0, 0, 128);      // [128 x NS]  += A(tmem)[128 x 16] B[NS x 16]^T, B K-major
      const uint32_t r1_a = smem_u32(ring1);
      const uint32_t r2_a = smem_u32(ring2);
      mbar_wait(res_ready, 0);
      tc_fence_after();
      // score buffer (g % NBUF) = R1 . X1^T  (, R2 . X2^T in the next NS columns): contraction over d_k = 8 k-steps of 16
      for (uint32_t g = (warp == 1 ? 0u : 1u); g < uint32_t(NSWEEP * n_steps); g += 2) {
        const bool main_sweep = NSWEEP == 1 || g >= uint32_t(n_steps);
        const uint32_t s1 = g % C::S1, s2 = g % C::S2;          // ring 2 is only read here in the narrow modes (one sweep: step = g)
        ATT_TRACE(2, g);
        mbar_wait(&full1[s1], (g / C::S1) & 1);
        if (!WIDE) mbar_wait(&full2[s2], (g / C::S2) & 1);
        mbar_wait(&buf_free[g % C::NBUF], ((g / C::NBUF) & 1) ^ 1);
        tc_fence_after();
        ATT_TRACE(3, g);
        const uint32_t acc = tmem_base + C::ACC0 + (g % C::NBUF) * C::ACC_STRIDE;
#pragma unroll
        for (int prod = 0; prod < C::NRES; ++prod) {
          const uint32_t ra = tmem_base + C::RES_COL + prod * 128;
          const uint64_t db0 = make_umma_desc(prod == 0 ? r1_a + s1 * C::X_TILE : r2_a + s2 * C::X_TILE, 16, 1024);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            const uint32_t boff = (ks >> 2) * C::X_KB + (ks & 3) * 32;
            const uint64_t db = desc_advance(db0, boff);
            umma_ts_pred(acc + prod * NS, ra + ks * 8, db, idesc_sc, ks > 0, lead);
            if (SPLIT == 3) {
              umma_ts_pred(acc + prod * NS, ra + ks * 8, desc_advance(db0, C::X_PART + boff), idesc_sc, 1, lead);
              umma_ts_pred(acc + prod * NS, ra + 64 + ks * 8, db, idesc_sc, 1, lead);
            }
          }
        }
        umma_commit_pred(&s_full[g % C::NBUF], lead);
        ATT_TRACE(4, g);
        // tiles the accumulated product does not read are free as soon as the score products have read them
        if (WIDE || !main_sweep) umma_commit_pred(&empty1[s1], lead);
        if (!WIDE) umma_commit_pred(&empty2[s2], lead);
      }
    }
  } else if (warp == 11) {
    // ===================================== MMA issuer B: accumulated products ==========================================
    if (elect_one_sync()) {
      constexpr uint32_t lead = 1u;
      constexpr uint32_t idesc_ac = make_idesc_bf16(DK, 0, 1, 128);      // [128 x 128] += A(tmem)[128 x 16] B[16 x 128],  B MN-major
      const uint32_t r1_a = smem_u32(ring1);
      const uint32_t r2_a = smem_u32(ring2);
      int s1 = 0, s2 = 0;
      uint32_t ph2 = 0;
      uint32_t accum_out = 0;
      // The two softmax warps of a lane quarter each own half of the tile's columns and write their packed output over their OWN
      // input columns, so k-step kk (16 keys = 8 packed columns) sits at: wide tiles (64 columns, 4 k-steps) hi 32 (kk / 2) +
      // 8 (kk % 2), lo + 16;  narrow tiles (32 columns, 2 k-steps) hi 16 kk, lo + 8.
      // out[128 x 128] += A(tmem, packed hi | lo)[128 x NS] * X[NS x 128]   (X streamed tile, MN-major B operand)
      auto issue_accum = [&](uint32_t a_col, uint32_t xb) {
        const uint64_t db0 = make_umma_desc(xb, C::X_KB, 1024);
#pragma unroll
        for (int kk = 0; kk < NS / 16; ++kk) {
          const uint32_t a_hi = a_col + (WIDE ? 32 * (kk >> 1) + 8 * (kk & 1) : 16 * kk);
          const uint32_t a_lo = a_hi + (WIDE ? 16 : 8);
          const uint64_t db = desc_advance(db0, kk * 2048);
          umma_ts_pred(tmem_base + C::OUT1, a_hi, db, idesc_ac, accum_out, lead);
          accum_out = 1;
          if (SPLIT == 3) {
            umma_ts_pred(tmem_base + C::OUT1, a_hi, desc_advance(db0, C::X_PART + kk * 2048), idesc_ac, 1, lead);
            umma_ts_pred(tmem_base + C::OUT1, a_lo, db, idesc_ac, 1, lead);
          }
        }
      };
      for (uint32_t gp = 0; gp < uint32_t(NSWEEP * n_steps); ++gp) {
        const bool main_sweep = NSWEEP == 1 || gp >= uint32_t(n_steps);
        ATT_TRACE(5, gp);
        mbar_wait(&p_full[gp % C::NBUF], (gp / C::NBUF) & 1);
        tc_fence_after();
        ATT_TRACE(6, gp);
        if (main_sweep) {
          const uint32_t acc = tmem_base + C::ACC0 + (gp % C::NBUF) * C::ACC_STRIDE;
          if constexpr (WIDE) {
            mbar_wait(&full2[s2], ph2);           // V (dO) is first needed here
            tc_fence_after();
            issue_accum(acc, r2_a + s2 * C::X_TILE);                        // O += P V   /   dV += P^T dO
            umma_commit_pred(&empty2[s2], lead);
            if (++s2 == C::S2) { s2 = 0; ph2 ^= 1; }
          } else {
            issue_accum(acc + NS, r1_a + s1 * C::X_TILE);                   // dQ += dS K   /   dK += dS^T Q
            umma_commit_pred(&empty1[s1], lead);
          }
        }
        if (!WIDE && ++s1 == C::S1) s1 = 0;
        umma_commit_pred(&buf_free[gp % C::NBUF], lead);     // arrives once the MMAs above (if any) have read the buffer
        ATT_TRACE(7, gp);
      }
      umma_commit_pred(out_full, lead);
    }
  } else {
    // ===================================== softmax / epilogue warps ===============================
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;             // which half of every tile's columns this warp owns
    const int r = q * 32 + lane;
    const int row = row0 + r;                     // query row (FWD, DQ) / key row (DK, DV)
    const uint32_t lane_addr = tmem_base + (uint32_t(q * 32) << 16);
    const int klen = p.klens ? static_cast<int>(min(static_cast<long long>(T), static_cast<long long>(p.klens[b]))) : T;
    const long long zrow = (static_cast<long long>(b) * H + h);
    const float cs = p.cs;
    const bool tp16 = (p.Tp & 15) == 0;           // emitted rows start 32-byte aligned

    // ---- resident operand(s): global -> registers -> TMEM (lane = row; the raw 32-bit words ARE the packed A-operand layout).
    //      All loads are issued before the first store (one L2 / HBM latency instead of one per 64-byte group). ----
    {
      const long long grow = static_cast<long long>(b) * T + row;
      constexpr int NG = C::NRES * C::PARTS * 2;           // 64-byte groups this thread copies (its 64 of the 128 columns)
      uint4 w[NG][4];
#pragma unroll
      for (int tile = 0; tile < C::NRES; ++tile) {
        const bool from_do = (MODE == ATT_DQ && tile == 1);
        const int blk = from_do ? h : (KEY_ROWS ? (tile == 0 ? zk : zv) : zq);
        const long long ld = from_do ? static_cast<long long>(H) * DK : 3LL * H * DK;
#pragma unroll
        for (int part = 0; part < C::PARTS; ++part) {
          const bf16* src = (from_do ? (part ? p.do_lo : p.do_hi) : (part ? p.qkv_lo : p.qkv_hi)) + grow * ld + static_cast<long long>(blk) * DK;
#pragma unroll
          for (int cc = 0; cc < 2; ++cc)
#pragma unroll
            for (int k = 0; k < 4; ++k)
              w[(tile * C::PARTS + part) * 2 + cc][k] =
                  row < T ? __ldg(reinterpret_cast<const uint4*>(src + (2 * half + cc) * 32 + k * 8)) : make_uint4(0u, 0u, 0u, 0u);
        }
      }
#pragma unroll
      for (int tile = 0; tile < C::NRES; ++tile)
#pragma unroll
        for (int part = 0; part < C::PARTS; ++part)
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {       // 32 bf16 = 64 B = 16 columns per store
            const uint4* v = w[(tile * C::PARTS + part) * 2 + cc];
            const uint32_t ww[16] = {v[0].x, v[0].y, v[0].z, v[0].w, v[1].x, v[1].y, v[1].z, v[1].w,
                                     v[2].x, v[2].y, v[2].z, v[2].w, v[3].x, v[3].y, v[3].z, v[3].w};
            tmem_st_32x32_x16(lane_addr + C::RES_COL + tile * 128 + part * 64 + (2 * half + cc) * 16, ww);
          }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (warp == 2 && lane == 0) ATT_TRACE(11, 0);
      if (lane == 0) mbar_arrive(res_ready);
    }

    uint32_t g = 0;
    auto wait_scores = [&]() -> uint32_t {
      if (warp == 2 && lane == 0) ATT_TRACE(8, g);
      mbar_wait(&s_full[g % C::NBUF], (g / C::NBUF) & 1);
      tc_fence_after();
      if (warp == 2 && lane == 0) ATT_TRACE(9, g);
      return lane_addr + C::ACC0 + (g % C::NBUF) * C::ACC_STRIDE;
    };
    auto arrive = [&]() {
      tc_fence_before();
      __syncwarp();
      if (warp == 2 && lane == 0) ATT_TRACE(10, g);
      if (lane == 0) mbar_arrive(&p_full[g % C::NBUF]);
      ++g;
    };

    if constexpr (MODE == ATT_FWD) {
      // ---- sweep 0: row statistics (log2 domain) over this warp's 32 of every 64 keys ----
      float m = -INFINITY, l = 0.f;
      for (int t = 0; t < n_steps; ++t) {
        const uint32_t acc = wait_scores();
        uint32_t v[32];
        tmem_ld_32x32(acc + 32 * half, v);
        tmem_ld_wait();
        arrive();                                 // the score buffer may be overwritten
        const int c0 = t * NS + 32 * half;
        float mt = m;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (c0 + j < klen) mt = fmaxf(mt, __uint_as_float(v[j]) * cs);
        if (mt > -INFINITY) {
          float s0 = 0.f, s1 = 0.f;
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            if (c0 + j < klen) s0 += ex2_approx(fmaf(__uint_as_float(v[j]), cs, -mt));
            if (c0 + j + 1 < klen) s1 += ex2_approx(fmaf(__uint_as_float(v[j + 1]), cs, -mt));
          }
          l = l * ex2_approx(m - mt) + (s0 + s1);
          m = mt;
        }
      }
      // combine with the warp that owns the other column half (the first column half always holds a valid key: klen >= 1)
      stat_xchg[half * 256 + r] = m;
      stat_xchg[half * 256 + 128 + r] = l;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      {
        const float mo = stat_xchg[(half ^ 1) * 256 + r], lo_ = stat_xchg[(half ^ 1) * 256 + 128 + r];
        const float mm = fmaxf(m, mo);
        l = l * ex2_approx(m - mm) + lo_ * ex2_approx(mo - mm);      // ex2(-inf) = 0 covers a half without valid keys
        m = mm;
      }
      const float L2 = m + log2f(l);
      if (half == 0 && row < T) p.lse[zrow * p.Tl + row] = L2;
      // ---- sweep 1: normalised P in place, O += P V ----
      for (int t = 0; t < n_steps; ++t) {
        const uint32_t acc = wait_scores() + 32 * half;
        const int c0 = t * NS + 32 * half;
        uint32_t v[32];
        tmem_ld_32x32(acc, v);
        tmem_ld_wait();
        float pv[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) pv[j] = (c0 + j < klen) ? ex2_approx(fmaf(__uint_as_float(v[j]), cs, -L2)) : 0.f;
        uint32_t hi[16], lo[16];
        pack_split<PK, 16>(pv, hi, lo);
        tmem_st_32x32_x16(acc, hi);              // this warp's 32 fp32 columns become [hi: 16 columns | lo: 16 columns]
        if (SPLIT == 3) tmem_st_32x32_x16(acc + 16, lo);
        tmem_st_wait();
        arrive();
        if (EMIT && row < T) {
          const long long off = (zrow * T + row) * p.Tp + c0;
          store_bf16_row<16>(p.p_hi + off, hi, c0, p.Tp, tp16);
          if (p.p_lo) store_bf16_row<16>(p.p_lo + off, lo, c0, p.Tp, tp16);
        }
      }
    } else if constexpr (MODE == ATT_DV) {
      // ---- dV: P^T recomputed from the saved log-sum-exp (thread = key, columns = queries) ----
      const bool key_ok = row < klen;
      for (int t = 0; t < n_steps; ++t) {
        const uint32_t acc = wait_scores() + 32 * half;
        const int c0 = t * NS + 32 * half;
        uint32_t v[32];
        tmem_ld_32x32(acc, v);
        tmem_ld_wait();
        const float4* L4 = reinterpret_cast<const float4*>(p.lse + zrow * p.Tl + c0);
        float pv[32];
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const float4 lv = __ldg(L4 + j4);
          const float ls[4] = {lv.x, lv.y, lv.z, lv.w};
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int j = 4 * j4 + u;
            pv[j] = (key_ok && c0 + j < T) ? ex2_approx(fmaf(__uint_as_float(v[j]), cs, -ls[u])) : 0.f;
          }
        }
        uint32_t hi[16], lo[16];
        pack_split<SPLIT, 16>(pv, hi, lo);
        tmem_st_32x32_x16(acc, hi);
        if (SPLIT == 3) tmem_st_32x32_x16(acc + 16, lo);
        tmem_st_wait();
        arrive();
      }
    } else {
      // ---- dQ / dK: P recomputed from the saved log-sum-exp, dS = P * (dP - D); this warp owns 16 of every 32 columns ----
      float L2r = 0.f, Dr = 0.f;
      if (MODE == ATT_DQ && row < T) {
        L2r = p.lse[zrow * p.Tl + row];
        Dr = p.dvec[zrow * p.Tl + row];
      }
      const bool key_ok = row < klen;             // DK: this thread's key is not masked
      for (int t = 0; t < n_steps; ++t) {
        const uint32_t acc = wait_scores() + 16 * half;
        const int c0 = t * NS + 16 * half;
        uint32_t a1[16], a2[16];
        tmem_ld_32x32_x16(acc, a1);
        tmem_ld_32x32_x16(acc + NS, a2);
        tmem_ld_wait();
        float dsv[16];
        if constexpr (MODE == ATT_DQ) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float pj = (c0 + j < klen) ? ex2_approx(fmaf(__uint_as_float(a1[j]), cs, -L2r)) : 0.f;
            dsv[j] = pj * (__uint_as_float(a2[j]) - Dr);
          }
        } else {
          const float4* L4 = reinterpret_cast<const float4*>(p.lse + zrow * p.Tl + c0);
          const float4* D4 = reinterpret_cast<const float4*>(p.dvec + zrow * p.Tl + c0);
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const float4 lv = __ldg(L4 + j4);
            const float4 dv = __ldg(D4 + j4);
            const float ls[4] = {lv.x, lv.y, lv.z, lv.w};
            const float ds[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int j = 4 * j4 + u;
              const float pj = (key_ok && c0 + j < T) ? ex2_approx(fmaf(__uint_as_float(a1[j]), cs, -ls[u])) : 0.f;
              dsv[j] = pj * (__uint_as_float(a2[j]) - ds[u]);
            }
          }
        }
        uint32_t hi[8], lo[8];
        pack_split<PK, 8>(dsv, hi, lo);            // dS (dS^T) over this warp's 16 columns of dP (dP^T): [hi: 8 | lo: 8]
        tmem_st_32x32_x8(acc + NS, hi);
        if (SPLIT == 3) tmem_st_32x32_x8(acc + NS + 8, lo);
        tmem_st_wait();
        arrive();
        if (EMIT && MODE == ATT_DQ && row < T) {
          const long long off = (zrow * T + row) * p.Tp + c0;
          store_bf16_row<8>(p.ds_hi + off, hi, c0, p.Tp, tp16);
          if (p.ds_lo) store_bf16_row<8>(p.ds_lo + off, lo, c0, p.Tp, tp16);
          const bool tp8 = (p.Tp & 7) == 0;
#pragma unroll
          for (int k = 0; k < 2; ++k) {            // dP: 16 fp32 = two 32-byte stores
            if (c0 + 8 * k + 8 <= p.Tp && tp8) {
              st_global_256(p.dp + off + 8 * k, a2 + 8 * k);
            } else {
#pragma unroll
              for (int u = 0; u < 8; ++u)
                if (c0 + 8 * k + u < p.Tp) p.dp[off + 8 * k + u] = __uint_as_float(a2[8 * k + u]);
            }
          }
        }
      }
    }

    // ---- epilogue: accumulator rows -> bf16 hi/lo in global memory (this warp: 64 of the 128 columns) ----
    mbar_wait(out_full, 0);
    tc_fence_after();
    if (warp == 2 && lane == 0) ATT_TRACE(11, 1);
    {
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
        else if (MODE == ATT_DK) { colblk = zk; alpha = p.scale; }
        else { colblk = zv; }
      }
      const long long off = (static_cast<long long>(b) * T + row) * ld + static_cast<long long>(colblk) * DK + 64 * half;
#pragma unroll 1
      for (int ch = 0; ch < 2; ++ch) {
        uint32_t v[32];
        tmem_ld_32x32(lane_addr + C::OUT1 + 64 * half + ch * 32, v);
        tmem_ld_wait();
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]) * alpha;
        uint32_t hi[16], lo[16];
        pack_split<3, 16>(f, hi, lo);            // outputs carry their lo half whenever the caller passes a buffer for it
        if (row < T) {
          store_bf16_row<16>(dst_hi + off + ch * 32, hi, 0, 32, true);
          if (dst_lo) store_bf16_row<16>(dst_lo + off + ch * 32, lo, 0, 32, true);
        }
      }
    }
  }

  if (threadIdx.x == 64) ATT_TRACE(11, 2);
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

int fill_common(AttnParams& p, const mtts_attn_desc* d, int box_rows, bool need_do) {   // box_rows = NS of the mode
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
  p.qkv_hi = static_cast<const bf16*>(d->qkv_hi);
  p.qkv_lo = static_cast<const bf16*>(d->qkv_lo);
  p.do_hi = static_cast<const bf16*>(d->do_hi);
  p.do_lo = static_cast<const bf16*>(d->do_lo);
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
  p.o_lo = static_cast<bf16*>(d->o_lo);          // written whenever given (also in single-pass mode)
  const bool emit = d->p_hi != nullptr;
  if (emit) {
    MTTS_REQUIRE(d->Tp >= d->T && (d->Tp & 7) == 0 && (d->split == 1 || d->p_lo), "attn_fwd: bad P emit buffers (Tp %d)", d->Tp);
    MTTS_REQUIRE(((reinterpret_cast<uintptr_t>(d->p_hi) | reinterpret_cast<uintptr_t>(d->p_lo)) & 31) == 0, "attn_fwd: P not 32B aligned");
    p.p_hi = static_cast<bf16*>(d->p_hi);
    p.p_lo = static_cast<bf16*>(d->p_lo);
  }
  MTTS_REQUIRE(((reinterpret_cast<uintptr_t>(d->o_hi) | reinterpret_cast<uintptr_t>(d->o_lo)) & 31) == 0, "attn_fwd: O not 32B aligned");
  if (d->split == 3) return emit ? launch_attn<ATT_FWD, 3, true>(p, stream) : launch_attn<ATT_FWD, 3, false>(p, stream);
  return emit ? launch_attn<ATT_FWD, 1, true>(p, stream) : launch_attn<ATT_FWD, 1, false>(p, stream);
}

extern "C" int mtts_attn_bwd(const mtts_attn_desc* d, int parts, mtts_stream stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MTTS_REQUIRE(d != nullptr, "attn: null descriptor");
  MTTS_REQUIRE(parts > 0 && parts < 16, "attn_bwd: parts is a mask of MTTS_ATTN_PREP | MTTS_ATTN_DQ | MTTS_ATTN_DK | MTTS_ATTN_DV");
  MTTS_REQUIRE(d->dvec, "attn_bwd: missing dvec");
  MTTS_REQUIRE(d->do_hi && (d->split == 1 || d->do_lo), "attn_bwd: missing dO operand");
  if (parts & MTTS_ATTN_PREP) {
    MTTS_REQUIRE(d->B > 0 && d->H > 0 && d->T > 0 && d->dk == DK && d->Tl >= d->T, "attn_bwd: bad shape");
    MTTS_REQUIRE(d->o_hi && (d->split == 1 || d->o_lo), "attn_bwd: PREP needs the forward output O");
    const long long n = static_cast<long long>(d->B) * d->T * d->H;
    const int grid = static_cast<int>(n / 8 + 1 < 148 * 8 ? n / 8 + 1 : 148 * 8);
    MTTS_CHECK_CUDA(mtts_launch(attn_dvec_kernel, dim3(grid), dim3(256), 0, stream, static_cast<const bf16*>(d->do_hi),
                                static_cast<const bf16*>(d->split == 3 ? d->do_lo : nullptr), static_cast<const bf16*>(d->o_hi),
                                static_cast<const bf16*>(d->split == 3 ? d->o_lo : nullptr), d->B, d->H, d->T, d->Tl, d->dvec));
    MTTS_CHECK_LAUNCH();
  }
  if (parts & (MTTS_ATTN_DQ | MTTS_ATTN_DK | MTTS_ATTN_DV)) {
    MTTS_REQUIRE(d->dqkv_hi && (d->split == 1 || d->dqkv_lo), "attn_bwd: missing dqkv output");
    MTTS_REQUIRE(((reinterpret_cast<uintptr_t>(d->dqkv_hi) | reinterpret_cast<uintptr_t>(d->dqkv_lo)) & 31) == 0, "attn_bwd: dqkv not 32B aligned");
  }
  AttnParams p;
  int rc;
  auto fill = [&](int box_rows) {
    int r = fill_common(p, d, box_rows, true);
    if (r != MTTS_OK) return r;
    p.dvec = d->dvec;
    p.dqkv_hi = static_cast<bf16*>(d->dqkv_hi);
    p.dqkv_lo = static_cast<bf16*>(d->dqkv_lo);
    return MTTS_OK;
  };
  if (parts & MTTS_ATTN_DQ) {
    if ((rc = fill(32)) != MTTS_OK) return rc;
    const bool emit = d->ds_hi != nullptr;
    if (emit) {
      MTTS_REQUIRE(d->dp && d->Tp >= d->T && (d->Tp & 7) == 0 && (d->split == 1 || d->ds_lo), "attn_bwd: bad dP / dS emit buffers (Tp %d)", d->Tp);
      MTTS_REQUIRE(((reinterpret_cast<uintptr_t>(d->ds_hi) | reinterpret_cast<uintptr_t>(d->ds_lo) | reinterpret_cast<uintptr_t>(d->dp)) & 31) == 0,
                   "attn_bwd: dP / dS not 32B aligned");
      p.dp = d->dp;
      p.ds_hi = static_cast<bf16*>(d->ds_hi);
      p.ds_lo = static_cast<bf16*>(d->ds_lo);
    }
    if (d->split == 3) rc = emit ? launch_attn<ATT_DQ, 3, true>(p, stream) : launch_attn<ATT_DQ, 3, false>(p, stream);
    else rc = emit ? launch_attn<ATT_DQ, 1, true>(p, stream) : launch_attn<ATT_DQ, 1, false>(p, stream);
    if (rc != MTTS_OK) return rc;
  }
  if (parts & MTTS_ATTN_DK) {
    if ((rc = fill(32)) != MTTS_OK) return rc;
    rc = d->split == 3 ? launch_attn<ATT_DK, 3, false>(p, stream) : launch_attn<ATT_DK, 1, false>(p, stream);
    if (rc != MTTS_OK) return rc;
  }
  if (parts & MTTS_ATTN_DV) {
    if ((rc = fill(64)) != MTTS_OK) return rc;
    rc = d->split == 3 ? launch_attn<ATT_DV, 3, false>(p, stream) : launch_attn<ATT_DV, 1, false>(p, stream);
    if (rc != MTTS_OK) return rc;
  }
  return MTTS_OK;
}
