// mtts_lr.cu — LengthRegulator: duration-driven row gather (forward) and segment sum (backward).
//
// Reference: lightning/model/modules.py:167-194 (Python double loop, one `.item()` host sync per
// phoneme, `expand` + `cat`) followed by utils/tools.py:304-322 `pad`.  Closed form used here
// (SURVEY.md §8 a11):  idx[b,t] = searchsorted(cumsum(max(int(d[b,:]),0)), t, right=True),
// out[b,t,:] = x[b,idx,:] for t < mel_len[b] = sum_j max(int(d[b,j]),0), else 0.
// The index path is integer arithmetic and bit-exact; the payload is copied unchanged (exact).
//
// HBM-bound: forward moves B*(L_touched + T)*C*4 bytes.  One warp per output row, 128-bit
// vectorised loads/stores, no host sync, shapes static => CUDA-graph capturable.
#include <algorithm>
#include "mtts_common.cuh"

namespace {

__device__ __forceinline__ long long dur_at(const int64_t* di, const float* df, int i) {
  long long v = di ? static_cast<long long>(di[i]) : static_cast<long long>(df[i]);  // int() truncates toward 0
  return v > 0 ? v : 0;
}

// One CTA per batch row: inclusive scan of the clamped durations in shared memory, then every
// thread binary-searches its frames.  L <= 4096.
__global__ void lr_index_kernel(const int64_t* __restrict__ di, const float* __restrict__ df, int L, int T,
                                int32_t* __restrict__ idx, int64_t* __restrict__ mel_len) {
  pdl_enter();
  extern __shared__ long long cum[];   // [L]
  const int b = blockIdx.x;
  // serial-in-chunks inclusive scan (L is small: <= a few hundred phonemes)
  __shared__ long long chunk_tot[32];
  const int nthr = blockDim.x;
  const int per = (L + nthr - 1) / nthr;
  const int beg = threadIdx.x * per;
  const int end = min(L, beg + per);
  long long s = 0;
  for (int i = beg; i < end; ++i) {
    s += dur_at(di, df, b * L + i);
    cum[i] = s;
  }
  // exclusive scan of per-thread totals via warp shuffles + one smem hop
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  long long incl = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    long long n = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += n;
  }
  if (lane == 31) chunk_tot[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    long long t = lane < (nthr >> 5) ? chunk_tot[lane] : 0;
    long long ti = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      long long n = __shfl_up_sync(0xffffffffu, ti, o);
      if (lane >= o) ti += n;
    }
    chunk_tot[lane] = ti - t;   // exclusive
  }
  __syncthreads();
  const long long offset = chunk_tot[warp] + (incl - s);
  for (int i = beg; i < end; ++i) cum[i] += offset;
  __syncthreads();
  const long long total = L > 0 ? cum[L - 1] : 0;
  if (threadIdx.x == 0) mel_len[b] = total;
  for (int t = threadIdx.x; t < T; t += nthr) {
    int r = -1;
    if (t < total) {
      // first j with cum[j] > t   (searchsorted right)
      int lo = 0, hi = L;
      while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (cum[mid] <= t) lo = mid + 1; else hi = mid;
      }
      r = lo;
    }
    idx[b * T + t] = r;
  }
}

// out[b,t,:] = idx>=0 ? x[b,idx,:] : 0.  One warp per (b,t) row, float4 lanes.
__global__ void lr_gather_kernel(const float* __restrict__ x, const int32_t* __restrict__ idx, int L, int T, int C,
                                 long long rows, float* __restrict__ out) {
  pdl_enter();
  const long long row = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const int b = static_cast<int>(row / T);
  const int j = idx[row];
  float* o = out + row * C;
  if ((C & 3) == 0) {
    float4* o4 = reinterpret_cast<float4*>(o);
    if (j >= 0) {
      const float4* s4 = reinterpret_cast<const float4*>(x + (static_cast<long long>(b) * L + j) * C);
      for (int c = lane; c < (C >> 2); c += 32) o4[c] = __ldg(s4 + c);
    } else {
      for (int c = lane; c < (C >> 2); c += 32) o4[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  } else {
    const float* s = x + (static_cast<long long>(b) * L + (j >= 0 ? j : 0)) * C;
    for (int c = lane; c < C; c += 32) o[c] = j >= 0 ? s[c] : 0.f;
  }
}

// dx[b,j,:] = sum over the frames of phoneme j (a contiguous run [start, start+d) clipped to T).  One CTA per (b,j): SEG_FL frame
// lanes x 64 channel threads (4 channels each); frame lane f sums frames start+f, start+f+SEG_FL, ... and the SEG_FL partials are
// added in lane order through shared memory => deterministic, no atomics, and a 75-frame silence costs 19 dependent adds instead
// of 75 (the warp-per-phoneme version took 21 us in the step because of a handful of such runs; the average run is 7 frames).
constexpr int SEG_FL = 4;
constexpr int SEG_THREADS = 64 * SEG_FL;
__global__ void __launch_bounds__(SEG_THREADS) lr_segsum_kernel(const float* __restrict__ dy, const int64_t* __restrict__ di,
                                                               const float* __restrict__ df, int L, int T, int C, long long rows,
                                                               float* __restrict__ dx) {
  pdl_enter();
  extern __shared__ float seg_part[];              // [SEG_FL - 1][C]
  __shared__ long long seg_start;
  const long long row = blockIdx.x;
  const int b = static_cast<int>(row / L);
  const int j = static_cast<int>(row - static_cast<long long>(b) * L);
  const int lane = threadIdx.x & 31;
  if (threadIdx.x < 32) {                          // start = sum_{i<j} d[b,i]  (warp-cooperative)
    long long part = 0;
    for (int i = lane; i < j; i += 32) part += dur_at(di, df, b * L + i);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (lane == 0) seg_start = part;
  }
  __syncthreads();
  const long long start = seg_start;
  long long stop = start + dur_at(di, df, b * L + j);
  if (stop > T) stop = T;
  const int fl = threadIdx.x >> 6, cg = threadIdx.x & 63;
  const float* src = dy + static_cast<long long>(b) * T * C;
  float* o = dx + row * C;
  if ((C & 3) == 0) {
    for (int c = cg * 4; c < C; c += 256) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (long long t = start + fl; t < stop; t += SEG_FL) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(src + t * C + c));
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      if (fl > 0) *reinterpret_cast<float4*>(seg_part + (fl - 1) * C + c) = acc;
      __syncthreads();
      if (fl == 0) {
#pragma unroll
        for (int f = 0; f < SEG_FL - 1; ++f) {
          const float4 v = *reinterpret_cast<const float4*>(seg_part + f * C + c);
          acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        *reinterpret_cast<float4*>(o + c) = acc;
      }
      __syncthreads();                             // seg_part is rewritten by the next channel block
    }
  } else if (fl == 0) {                            // odd channel counts: serial over the run
    for (int c = cg; c < C; c += 64) {
      float a = 0.f;
      for (long long t = start; t < stop; ++t) a += src[t * C + c];
      o[c] = a;
    }
  }
}

// Ragged -> padded pack (collate on the device): dst[b, t, :] = t < len_b ? src[row_off[b] + t, :] : 0, bytes copied
// unchanged (bit-exact for every dtype).  One thread per CHUNK-byte piece of the padded output.
template <typename V>
__global__ void pack_rows_kernel(const V* __restrict__ src, const int64_t* __restrict__ row_off, int Lmax, int vec_per_row,
                                 long long total, V* __restrict__ dst) {
  pdl_enter();
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long row = i / vec_per_row;                 // padded row index b * Lmax + t
    const int v = static_cast<int>(i - row * vec_per_row);
    const int b = static_cast<int>(row / Lmax);
    const int t = static_cast<int>(row - static_cast<long long>(b) * Lmax);
    const long long r0 = row_off[b], r1 = row_off[b + 1];
    V out{};
    if (t < r1 - r0) out = src[(r0 + t) * vec_per_row + v];
    dst[i] = out;
  }
}

}  // namespace

extern "C" int mtts_pack_rows(const void* src, const int64_t* row_off, int B, int Lmax, int row_bytes, void* dst,
                              mtts_stream stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MTTS_REQUIRE(src && row_off && dst && B > 0 && Lmax > 0 && row_bytes > 0, "pack_rows: bad args");
  const uintptr_t al = reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst) | static_cast<uintptr_t>(row_bytes);
  const int threads = 256;
  auto blocks = [&](long long total) { return static_cast<unsigned>(std::min<long long>(mtts_cdiv64(total, threads), 148 * 16)); };
  if ((al & 15) == 0) {
    const int vpr = row_bytes / 16;
    const long long total = static_cast<long long>(B) * Lmax * vpr;
    MTTS_CHECK_CUDA(mtts_launch(pack_rows_kernel<uint4>, dim3(blocks(total)), dim3(threads), 0, stream, static_cast<const uint4*>(src),
                                row_off, Lmax, vpr, total, static_cast<uint4*>(dst)));
  } else if ((al & 7) == 0) {
    const int vpr = row_bytes / 8;
    const long long total = static_cast<long long>(B) * Lmax * vpr;
    MTTS_CHECK_CUDA(mtts_launch(pack_rows_kernel<uint2>, dim3(blocks(total)), dim3(threads), 0, stream, static_cast<const uint2*>(src),
                                row_off, Lmax, vpr, total, static_cast<uint2*>(dst)));
  } else {
    MTTS_REQUIRE((al & 3) == 0, "pack_rows: rows must be a multiple of 4 bytes and 4-byte aligned");
    const int vpr = row_bytes / 4;
    const long long total = static_cast<long long>(B) * Lmax * vpr;
    MTTS_CHECK_CUDA(mtts_launch(pack_rows_kernel<uint32_t>, dim3(blocks(total)), dim3(threads), 0, stream,
                                static_cast<const uint32_t*>(src), row_off, Lmax, vpr, total, static_cast<uint32_t*>(dst)));
  }
  MTTS_CHECK_LAUNCH();
  return MTTS_OK;
}

extern "C" int mtts_length_regulate_index(const int64_t* dur_i64, const float* dur_f32, int B, int L, int T,
                                          int32_t* idx, int64_t* mel_len, mtts_stream stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MTTS_REQUIRE((dur_i64 != nullptr) != (dur_f32 != nullptr), "length_regulate: pass exactly one duration pointer");
  MTTS_REQUIRE(B > 0 && L > 0 && T > 0 && L <= 4096, "length_regulate: bad B/L/T %d %d %d", B, L, T);
  MTTS_CHECK_CUDA(mtts_launch(lr_index_kernel, dim3(B), dim3(256), L * sizeof(long long), stream, dur_i64, dur_f32, L, T, idx, mel_len));
  MTTS_CHECK_LAUNCH();
  return MTTS_OK;
}

extern "C" int mtts_length_regulate_fwd(const float* x, const int32_t* idx, int B, int L, int T, int C, float* out,
                                        mtts_stream stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MTTS_REQUIRE(B > 0 && L > 0 && T > 0 && C > 0, "length_regulate_fwd: bad shape");
  const long long rows = static_cast<long long>(B) * T;
  const int threads = 256;
  const long long blocks = mtts_cdiv64(rows * 32, threads);
  MTTS_CHECK_CUDA(mtts_launch(lr_gather_kernel, dim3(static_cast<unsigned>(blocks)), dim3(threads), 0, stream, x, idx, L, T, C, rows, out));
  MTTS_CHECK_LAUNCH();
  return MTTS_OK;
}

extern "C" int mtts_length_regulate_bwd(const float* dy, const int64_t* dur_i64, const float* dur_f32, int B, int L,
                                        int T, int C, float* dx, mtts_stream stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MTTS_REQUIRE((dur_i64 != nullptr) != (dur_f32 != nullptr), "length_regulate_bwd: pass exactly one duration pointer");
  MTTS_REQUIRE(B > 0 && L > 0 && T > 0 && C > 0, "length_regulate_bwd: bad shape");
  const long long rows = static_cast<long long>(B) * L;
  const size_t smem = sizeof(float) * (SEG_FL - 1) * static_cast<size_t>(C);
  MTTS_REQUIRE(smem <= 48 * 1024, "length_regulate_bwd: C too large (%d)", C);
  MTTS_CHECK_CUDA(mtts_launch(lr_segsum_kernel, dim3(static_cast<unsigned>(rows)), dim3(SEG_THREADS), smem, stream, dy, dur_i64, dur_f32, L, T, C, rows, dx));
  MTTS_CHECK_LAUNCH();
  return MTTS_OK;
}
