// mtts_rowops.cu — row-wise (per token) HBM-bound kernels: LayerNorm (+residual, +pad-row zeroing),
// masked softmax, the Linear(256->1) predictor head — each with forward, backward, tangent-forward
// (JVP) and tangent-backward (JVP of the backward) variants.  The tangent variants are what lets the
// second-order MAML outer gradient be computed without an autograd graph (forward-over-reverse HVP).
//
// Reference call sites: nn.LayerNorm SubLayers.py:55,91 and modules.py:221,233; masked_fill pad
// zeroing Layers.py:25,28; softmax Modules.py:16-22; Linear(256,1)+masked_fill modules.py:240-250.
//
// One warp per row, 128-bit loads/stores (each lane owns 4 consecutive channels per 128-channel
// group), warp-shuffle reductions, per-CTA shared-memory reduction of the per-channel parameter
// gradients followed by one atomicAdd per channel per CTA.  Grids are sized as a multiple of the SM
// count (persistent grid-stride over rows).
#include "mtts_common.cuh"

namespace {

constexpr int ROW_THREADS = 256;   // 8 warps / CTA
constexpr int ROW_WARPS = ROW_THREADS / 32;

// reduces = the kernel also accumulates per-channel sums over its rows (atomics across CTAs): ONE CTA in deterministic mode
inline int row_grid(long long rows, bool reduces = false) {
  if (reduces && mtts_deterministic()) return 1;
  int sms = 148;
  long long need = (rows + ROW_WARPS - 1) / ROW_WARPS;
  long long cap = static_cast<long long>(sms) * 4;
  return static_cast<int>(need < cap ? (need < 1 ? 1 : need) : cap);
}

// one row per warp when that stays under 8 CTAs per SM: a capped grid-stride loop would leave warps with 1 or 2 rows (a 2-row tail)
inline int softmax_grid(long long rows) {
  const long long need = (rows + ROW_WARPS - 1) / ROW_WARPS;
  return need <= 148LL * 8 ? static_cast<int>(need < 1 ? 1 : need) : row_grid(rows);
}

__device__ __forceinline__ bool row_valid(const int64_t* lens, int T, long long r) {
  if (!lens) return true;
  const long long b = r / T;
  const long long t = r - b * T;
  return t < lens[b];
}

template <int NV>
struct RowVec {
  float4 v[NV];
};

template <int NV>
__device__ __forceinline__ void load_row(const float* p, int lane, RowVec<NV>& r) {
#pragma unroll
  for (int i = 0; i < NV; ++i) r.v[i] = *reinterpret_cast<const float4*>(p + i * 128 + lane * 4);
}
template <int NV>
__device__ __forceinline__ void load_row_ldg(const float* p, int lane, RowVec<NV>& r) {
#pragma unroll
  for (int i = 0; i < NV; ++i) r.v[i] = __ldg(reinterpret_cast<const float4*>(p + i * 128 + lane * 4));
}
template <int NV>
__device__ __forceinline__ void zero_row(RowVec<NV>& r) {
#pragma unroll
  for (int i = 0; i < NV; ++i) r.v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}
template <int NV>
__device__ __forceinline__ void store_row(float* p, int lane, const RowVec<NV>& r) {
#pragma unroll
  for (int i = 0; i < NV; ++i) *reinterpret_cast<float4*>(p + i * 128 + lane * 4) = r.v[i];
}
template <int NV>
__device__ __forceinline__ void store_row_split(bf16* hi, bf16* lo, int lane, const RowVec<NV>& r) {
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float a[4] = {r.v[i].x, r.v[i].y, r.v[i].z, r.v[i].w};
    uint16_t h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      bf16 hh, ll;
      split_bf16(a[j], hh, ll);
      h[j] = __bfloat16_as_ushort(hh);
      l[j] = __bfloat16_as_ushort(ll);
    }
    if (hi) *reinterpret_cast<uint2*>(hi + i * 128 + lane * 4) = make_uint2(h[0] | (uint32_t(h[1]) << 16), h[2] | (uint32_t(h[3]) << 16));
    if (lo) *reinterpret_cast<uint2*>(lo + i * 128 + lane * 4) = make_uint2(l[0] | (uint32_t(l[1]) << 16), l[2] | (uint32_t(l[3]) << 16));
  }
}
template <int NV>
__device__ __forceinline__ float row_sum(const RowVec<NV>& r) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) s += (r.v[i].x + r.v[i].y) + (r.v[i].z + r.v[i].w);
  return warp_sum(s);
}
template <int NV>
__device__ __forceinline__ float row_dot(const RowVec<NV>& a, const RowVec<NV>& b) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) s += (a.v[i].x * b.v[i].x + a.v[i].y * b.v[i].y) + (a.v[i].z * b.v[i].z + a.v[i].w * b.v[i].w);
  return warp_sum(s);
}
#define ROW_FOREACH(NVv, i, ...)              \
  _Pragma("unroll") for (int i = 0; i < NVv; ++i) { __VA_ARGS__ }

// elementwise helpers on float4
__device__ __forceinline__ float4 f4_add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 f4_sub(float4 a, float4 b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
__device__ __forceinline__ float4 f4_mul(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
__device__ __forceinline__ float4 f4_scale(float4 a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }
__device__ __forceinline__ float4 f4_fma(float4 a, float4 b, float4 c) {
  return make_float4(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y), fmaf(a.z, b.z, c.z), fmaf(a.w, b.w, c.w));
}
__device__ __forceinline__ float4 f4_relu_gate(float4 v, float4 z) {
  return make_float4(z.x > 0.f ? v.x : 0.f, z.y > 0.f ? v.y : 0.f, z.z > 0.f ? v.z : 0.f, z.w > 0.f ? v.w : 0.f);
}

// multiply a row by its dropout keep*scale factors (element index = r*C + column)
template <int NV>
__device__ __forceinline__ void drop_row(const DropSite& d, long long r, int lane, RowVec<NV>& v) {
  if (d.thr == 0) return;
  const uint32_t base = static_cast<uint32_t>(r * (NV * 128)) + lane * 4;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    v.v[i].x *= drop_factor(d, base + i * 128 + 0);
    v.v[i].y *= drop_factor(d, base + i * 128 + 1);
    v.v[i].z *= drop_factor(d, base + i * 128 + 2);
    v.v[i].w *= drop_factor(d, base + i * 128 + 3);
  }
}

// per-CTA reduction of per-channel accumulators then one atomicAdd per channel
template <int NV, int NACC>
__device__ __forceinline__ void flush_channel_acc(RowVec<NV> (&acc)[NACC], float* const (&dst)[NACC], int lane, int warp) {
  __shared__ float red[ROW_WARPS][NV * 128];
  for (int a = 0; a < NACC; ++a) {
    if (dst[a] == nullptr) continue;      // uniform across the CTA
    __syncthreads();
    store_row<NV>(red[warp], lane, acc[a]);
    __syncthreads();
    for (int c = threadIdx.x; c < NV * 128; c += ROW_THREADS) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < ROW_WARPS; ++w) s += red[w][c];
      atomicAdd(dst[a] + c, s);
    }
  }
}

// ================================================================================================
// LayerNorm
// ================================================================================================
// forward:  z = y (+ res);  xhat = (z - mean) * rstd;  out = valid ? xhat*gamma + beta : 0
template <int NV>
__global__ void __launch_bounds__(ROW_THREADS) ln_fwd_kernel(const float* __restrict__ y, const float* __restrict__ res,
                                                             const float* __restrict__ gamma, const float* __restrict__ beta,
                                                             const int64_t* __restrict__ lens, int T, long long R, float eps,
                                                             float* __restrict__ z_out, float* __restrict__ stats,
                                                             float* __restrict__ out, bf16* __restrict__ out_hi,
                                                             bf16* __restrict__ out_lo, DropSite dpre, DropSite dpost) {
  pdl_enter();
  constexpr int C = NV * 128;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  RowVec<NV> g, b;
  load_row_ldg<NV>(gamma, lane, g);
  load_row_ldg<NV>(beta, lane, b);
  for (long long r = static_cast<long long>(blockIdx.x) * ROW_WARPS + warp; r < R; r += static_cast<long long>(gridDim.x) * ROW_WARPS) {
    RowVec<NV> z;
    load_row<NV>(y + r * C, lane, z);
    drop_row<NV>(dpre, r, lane, z);                 // dropout on the branch (SubLayers.py:54,90) before the residual add
    if (res) {
      RowVec<NV> rr;
      load_row<NV>(res + r * C, lane, rr);
      ROW_FOREACH(NV, i, z.v[i] = f4_add(z.v[i], rr.v[i]);)
    }
    const float mean = row_sum<NV>(z) * (1.f / C);
    RowVec<NV> d;
    ROW_FOREACH(NV, i, d.v[i] = make_float4(z.v[i].x - mean, z.v[i].y - mean, z.v[i].z - mean, z.v[i].w - mean);)
    const float var = row_dot<NV>(d, d) * (1.f / C);
    const float rstd = rsqrtf(var + eps);
    if (z_out) store_row<NV>(z_out + r * C, lane, z);
    if (stats && lane == 0) {
      stats[2 * r] = mean;
      stats[2 * r + 1] = rstd;
    }
    const bool valid = row_valid(lens, T, r);
    RowVec<NV> o;
    if (valid) {
      ROW_FOREACH(NV, i, o.v[i] = f4_fma(f4_scale(d.v[i], rstd), g.v[i], b.v[i]);)
      drop_row<NV>(dpost, r, lane, o);              // dropout on the LN output (modules.py:223,235)
    } else {
      zero_row<NV>(o);
    }
    if (out) store_row<NV>(out + r * C, lane, o);
    if (out_hi) store_row_split<NV>(out_hi + r * C, out_lo ? out_lo + r * C : nullptr, lane, o);
  }
}

// backward: g = valid ? dy*gamma : 0; dz = rstd*(g - mean(g) - xhat*mean(g*xhat)); [dz *= (z>0) if relu_gate]
//           dgamma += sum dy*xhat; dbeta += sum dy; dbias += sum dz (after gate)
template <int NV>
__global__ void __launch_bounds__(ROW_THREADS) ln_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ z,
                                                             const float* __restrict__ stats, const float* __restrict__ gamma,
                                                             const int64_t* __restrict__ lens, int T, long long R,
                                                             int relu_gate, float* __restrict__ dz, bf16* __restrict__ dz_hi,
                                                             bf16* __restrict__ dz_lo, float* __restrict__ dgamma,
                                                             float* __restrict__ dbeta, float* __restrict__ dbias, DropSite dpre,
                                                             DropSite dpost) {
  pdl_enter();
  constexpr int C = NV * 128;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  RowVec<NV> g;
  load_row_ldg<NV>(gamma, lane, g);
  RowVec<NV> acc[3];
  zero_row<NV>(acc[0]);
  zero_row<NV>(acc[1]);
  zero_row<NV>(acc[2]);
  for (long long r = static_cast<long long>(blockIdx.x) * ROW_WARPS + warp; r < R; r += static_cast<long long>(gridDim.x) * ROW_WARPS) {
    const bool valid = row_valid(lens, T, r);
    RowVec<NV> o;
    if (valid) {
      RowVec<NV> zz, d, xh, gg;
      load_row<NV>(z + r * C, lane, zz);
      load_row<NV>(dy + r * C, lane, d);
      drop_row<NV>(dpost, r, lane, d);              // gradient through the output dropout
      const float mean = stats[2 * r], rstd = stats[2 * r + 1];
      ROW_FOREACH(NV, i, xh.v[i] = f4_scale(make_float4(zz.v[i].x - mean, zz.v[i].y - mean, zz.v[i].z - mean, zz.v[i].w - mean), rstd);)
      ROW_FOREACH(NV, i, gg.v[i] = f4_mul(d.v[i], g.v[i]);)
      const float m1 = row_sum<NV>(gg) * (1.f / C);
      const float m2 = row_dot<NV>(gg, xh) * (1.f / C);
      ROW_FOREACH(NV, i, {
        float4 t = make_float4(gg.v[i].x - m1 - xh.v[i].x * m2, gg.v[i].y - m1 - xh.v[i].y * m2,
                               gg.v[i].z - m1 - xh.v[i].z * m2, gg.v[i].w - m1 - xh.v[i].w * m2);
        o.v[i] = f4_scale(t, rstd);
        if (relu_gate) o.v[i] = f4_relu_gate(o.v[i], zz.v[i]);
        acc[0].v[i] = f4_fma(d.v[i], xh.v[i], acc[0].v[i]);
        acc[1].v[i] = f4_add(acc[1].v[i], d.v[i]);
      })
    } else {
      zero_row<NV>(o);
    }
    if (dz) store_row<NV>(dz + r * C, lane, o);       // residual path: un-dropped
    drop_row<NV>(dpre, r, lane, o);                   // branch path (and its bias): through the branch dropout
    ROW_FOREACH(NV, i, acc[2].v[i] = f4_add(acc[2].v[i], o.v[i]);)
    if (dz_hi) store_row_split<NV>(dz_hi + r * C, dz_lo ? dz_lo + r * C : nullptr, lane, o);
  }
  float* const dst[3] = {dgamma, dbeta, dbias};
  flush_channel_acc<NV, 3>(acc, dst, lane, warp);
}

// tangent forward: zdot = ydot (+ resdot); xhd = rstd*(zdot - mean(zdot) - xhat*mean(zdot*xhat));
//                  outdot = valid ? xhd*gamma + xhat*gdot + bdot : 0
template <int NV>
__global__ void __launch_bounds__(ROW_THREADS) ln_tfwd_kernel(const float* __restrict__ ydot, const float* __restrict__ resdot,
                                                              const float* __restrict__ z, const float* __restrict__ stats,
                                                              const float* __restrict__ gamma, const float* __restrict__ gdot,
                                                              const float* __restrict__ bdot, const int64_t* __restrict__ lens,
                                                              int T, long long R, float* __restrict__ zdot_out,
                                                              float* __restrict__ out, bf16* __restrict__ out_hi,
                                                              bf16* __restrict__ out_lo, DropSite dpre, DropSite dpost) {
  pdl_enter();
  constexpr int C = NV * 128;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  RowVec<NV> g, gd, bd;
  load_row_ldg<NV>(gamma, lane, g);
  if (gdot) load_row_ldg<NV>(gdot, lane, gd); else zero_row<NV>(gd);
  if (bdot) load_row_ldg<NV>(bdot, lane, bd); else zero_row<NV>(bd);
  for (long long r = static_cast<long long>(blockIdx.x) * ROW_WARPS + warp; r < R; r += static_cast<long long>(gridDim.x) * ROW_WARPS) {
    RowVec<NV> zd;
    load_row<NV>(ydot + r * C, lane, zd);
    drop_row<NV>(dpre, r, lane, zd);
    if (resdot) {
      RowVec<NV> rr;
      load_row<NV>(resdot + r * C, lane, rr);
      ROW_FOREACH(NV, i, zd.v[i] = f4_add(zd.v[i], rr.v[i]);)
    }
    if (zdot_out) store_row<NV>(zdot_out + r * C, lane, zd);
    const bool valid = row_valid(lens, T, r);
    RowVec<NV> o;
    if (valid) {
      RowVec<NV> zz, xh;
      load_row<NV>(z + r * C, lane, zz);
      const float mean = stats[2 * r], rstd = stats[2 * r + 1];
      ROW_FOREACH(NV, i, xh.v[i] = f4_scale(make_float4(zz.v[i].x - mean, zz.v[i].y - mean, zz.v[i].z - mean, zz.v[i].w - mean), rstd);)
      const float m1 = row_sum<NV>(zd) * (1.f / C);
      const float m2 = row_dot<NV>(zd, xh) * (1.f / C);
      ROW_FOREACH(NV, i, {
        float4 xhd = f4_scale(make_float4(zd.v[i].x - m1 - xh.v[i].x * m2, zd.v[i].y - m1 - xh.v[i].y * m2,
                                          zd.v[i].z - m1 - xh.v[i].z * m2, zd.v[i].w - m1 - xh.v[i].w * m2), rstd);
        o.v[i] = f4_add(f4_fma(xhd, g.v[i], f4_mul(xh.v[i], gd.v[i])), bd.v[i]);
      })
      drop_row<NV>(dpost, r, lane, o);
    } else {
      zero_row<NV>(o);
    }
    if (out) store_row<NV>(out + r * C, lane, o);
    if (out_hi) store_row_split<NV>(out_hi + r * C, out_lo ? out_lo + r * C : nullptr, lane, o);
  }
}

// tangent backward (JVP of ln_bwd):
//   g = dy*gamma, gd = ddy*gamma + dy*gdot;  m1 = mean(g), m2 = mean(g*xhat)
//   xhd as in tfwd; rd = -rstd^2 * mean(xhat*zdot)
//   m1d = mean(gd), m2d = mean(gd*xhat + g*xhd)
//   ddz = rd*(g - m1 - xhat*m2) + rstd*(gd - m1d - xhd*m2 - xhat*m2d)       [then relu gate on z]
//   ddgamma += sum(ddy*xhat + dy*xhd);  ddbeta += sum ddy;  ddbias += sum ddz
template <int NV>
__global__ void __launch_bounds__(ROW_THREADS) ln_tbwd_kernel(const float* __restrict__ dy, const float* __restrict__ ddy,
                                                              const float* __restrict__ z, const float* __restrict__ zdot,
                                                              const float* __restrict__ stats, const float* __restrict__ gamma,
                                                              const float* __restrict__ gdot, const int64_t* __restrict__ lens,
                                                              int T, long long R, int relu_gate, float* __restrict__ ddz,
                                                              bf16* __restrict__ ddz_hi, bf16* __restrict__ ddz_lo,
                                                              float* __restrict__ ddgamma, float* __restrict__ ddbeta,
                                                              float* __restrict__ ddbias, DropSite dpre, DropSite dpost) {
  pdl_enter();
  constexpr int C = NV * 128;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  RowVec<NV> g, gdv;
  load_row_ldg<NV>(gamma, lane, g);
  if (gdot) load_row_ldg<NV>(gdot, lane, gdv); else zero_row<NV>(gdv);
  RowVec<NV> acc[3];
  zero_row<NV>(acc[0]);
  zero_row<NV>(acc[1]);
  zero_row<NV>(acc[2]);
  for (long long r = static_cast<long long>(blockIdx.x) * ROW_WARPS + warp; r < R; r += static_cast<long long>(gridDim.x) * ROW_WARPS) {
    const bool valid = row_valid(lens, T, r);
    RowVec<NV> o;
    if (valid) {
      RowVec<NV> zz, zd, d, dd, xh, xhd, gg, ggd;
      load_row<NV>(z + r * C, lane, zz);
      load_row<NV>(zdot + r * C, lane, zd);
      load_row<NV>(dy + r * C, lane, d);
      load_row<NV>(ddy + r * C, lane, dd);
      drop_row<NV>(dpost, r, lane, d);
      drop_row<NV>(dpost, r, lane, dd);
      const float mean = stats[2 * r], rstd = stats[2 * r + 1];
      ROW_FOREACH(NV, i, xh.v[i] = f4_scale(make_float4(zz.v[i].x - mean, zz.v[i].y - mean, zz.v[i].z - mean, zz.v[i].w - mean), rstd);)
      const float a1 = row_sum<NV>(zd) * (1.f / C);
      const float a2 = row_dot<NV>(zd, xh) * (1.f / C);
      ROW_FOREACH(NV, i, xhd.v[i] = f4_scale(make_float4(zd.v[i].x - a1 - xh.v[i].x * a2, zd.v[i].y - a1 - xh.v[i].y * a2,
                                                         zd.v[i].z - a1 - xh.v[i].z * a2, zd.v[i].w - a1 - xh.v[i].w * a2), rstd);)
      const float rd = -rstd * rstd * a2;
      ROW_FOREACH(NV, i, {
        gg.v[i] = f4_mul(d.v[i], g.v[i]);
        ggd.v[i] = f4_fma(dd.v[i], g.v[i], f4_mul(d.v[i], gdv.v[i]));
      })
      const float m1 = row_sum<NV>(gg) * (1.f / C);
      const float m2 = row_dot<NV>(gg, xh) * (1.f / C);
      const float m1d = row_sum<NV>(ggd) * (1.f / C);
      const float m2d = (row_dot<NV>(ggd, xh) + row_dot<NV>(gg, xhd)) * (1.f / C);
      ROW_FOREACH(NV, i, {
        const float4 G = gg.v[i], GD = ggd.v[i], X = xh.v[i], XD = xhd.v[i];
        float4 t;
        t.x = rd * (G.x - m1 - X.x * m2) + rstd * (GD.x - m1d - XD.x * m2 - X.x * m2d);
        t.y = rd * (G.y - m1 - X.y * m2) + rstd * (GD.y - m1d - XD.y * m2 - X.y * m2d);
        t.z = rd * (G.z - m1 - X.z * m2) + rstd * (GD.z - m1d - XD.z * m2 - X.z * m2d);
        t.w = rd * (G.w - m1 - X.w * m2) + rstd * (GD.w - m1d - XD.w * m2 - X.w * m2d);
        if (relu_gate) t = f4_relu_gate(t, zz.v[i]);
        o.v[i] = t;
        acc[0].v[i] = f4_fma(dd.v[i], X, f4_fma(d.v[i], XD, acc[0].v[i]));
        acc[1].v[i] = f4_add(acc[1].v[i], dd.v[i]);
      })
    } else {
      zero_row<NV>(o);
    }
    if (ddz) store_row<NV>(ddz + r * C, lane, o);
    drop_row<NV>(dpre, r, lane, o);
    ROW_FOREACH(NV, i, acc[2].v[i] = f4_add(acc[2].v[i], o.v[i]);)
    if (ddz_hi) store_row_split<NV>(ddz_hi + r * C, ddz_lo ? ddz_lo + r * C : nullptr, lane, o);
  }
  float* const dst[3] = {ddgamma, ddbeta, ddbias};
  flush_channel_acc<NV, 3>(acc, dst, lane, warp);
}

// ================================================================================================
// Linear(C -> 1) predictor head with row masking  (modules.py:240-250)
//   mode 0 fwd : out[r] = valid ? h[r].w + b : 0
//   mode 1 tfwd: out[r] = valid ? hdot[r].w + h[r].wdot + bdot : 0
// ================================================================================================
template <int NV>
__global__ void __launch_bounds__(ROW_THREADS) rowdot_fwd_kernel(const float* __restrict__ h, const float* __restrict__ hdot,
                                                                 const float* __restrict__ w, const float* __restrict__ wdot,
                                                                 const float* __restrict__ b, const float* __restrict__ bdot,
                                                                 const int64_t* __restrict__ lens, int T, long long R,
                                                                 float* __restrict__ out) {
  pdl_enter();
  constexpr int C = NV * 128;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  RowVec<NV> wv, wd;
  load_row_ldg<NV>(w, lane, wv);
  if (wdot) load_row_ldg<NV>(wdot, lane, wd); else zero_row<NV>(wd);
  for (long long r = static_cast<long long>(blockIdx.x) * ROW_WARPS + warp; r < R; r += static_cast<long long>(gridDim.x) * ROW_WARPS) {
    float v = 0.f;
    if (row_valid(lens, T, r)) {
      RowVec<NV> hh;
      load_row<NV>(h + r * C, lane, hh);
      if (hdot) {     // tangent
        RowVec<NV> hd;
        load_row<NV>(hdot + r * C, lane, hd);
        v = row_dot<NV>(hd, wv) + row_dot<NV>(hh, wd) + (bdot ? bdot[0] : 0.f);
      } else {
        v = row_dot<NV>(hh, wv) + b[0];
      }
    }
    if (lane == 0) out[r] = v;
  }
}
// backward / tangent-backward:
//   dh[r] = valid ? dout[r]*w (+ ddout... see below) : 0
//   mode bwd : dh = dout*w ; dw += sum dout*h ; db += sum dout
//   mode tbwd: ddh = ddout*w + dout*wdot ; ddw += sum(ddout*h + dout*hdot) ; ddb += sum ddout
template <int NV>
__global__ void __launch_bounds__(ROW_THREADS) rowdot_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ ddout,
                                                                 const float* __restrict__ h, const float* __restrict__ hdot,
                                                                 const float* __restrict__ w, const float* __restrict__ wdot,
                                                                 const int64_t* __restrict__ lens, int T, long long R,
                                                                 float* __restrict__ dh, float* __restrict__ dw,
                                                                 float* __restrict__ db) {
  pdl_enter();
  constexpr int C = NV * 128;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool tangent = ddout != nullptr;
  RowVec<NV> wv, wd;
  load_row_ldg<NV>(w, lane, wv);
  if (wdot) load_row_ldg<NV>(wdot, lane, wd); else zero_row<NV>(wd);
  RowVec<NV> acc[1];
  zero_row<NV>(acc[0]);
  float bacc = 0.f;
  for (long long r = static_cast<long long>(blockIdx.x) * ROW_WARPS + warp; r < R; r += static_cast<long long>(gridDim.x) * ROW_WARPS) {
    RowVec<NV> o;
    if (row_valid(lens, T, r)) {
      RowVec<NV> hh;
      load_row<NV>(h + r * C, lane, hh);
      const float s = dout[r];
      if (tangent) {
        const float sd = ddout[r];
        RowVec<NV> hd;
        load_row<NV>(hdot + r * C, lane, hd);
        ROW_FOREACH(NV, i, {
          o.v[i] = f4_add(f4_scale(wv.v[i], sd), f4_scale(wd.v[i], s));
          acc[0].v[i] = f4_add(acc[0].v[i], f4_add(f4_scale(hh.v[i], sd), f4_scale(hd.v[i], s)));
        })
        bacc += sd;
      } else {
        ROW_FOREACH(NV, i, {
          o.v[i] = f4_scale(wv.v[i], s);
          acc[0].v[i] = f4_add(acc[0].v[i], f4_scale(hh.v[i], s));
        })
        bacc += s;
      }
    } else {
      zero_row<NV>(o);
    }
    store_row<NV>(dh + r * C, lane, o);
  }
  float* const dst[1] = {dw};
  flush_channel_acc<NV, 1>(acc, dst, lane, warp);
  if (db) {
    __shared__ float bred[ROW_WARPS];
    if (lane == 0) bred[warp] = bacc;     // bacc identical across lanes
    __syncthreads();
    if (threadIdx.x == 0) {
      float s = 0.f;
      for (int wgt = 0; wgt < ROW_WARPS; ++wgt) s += bred[wgt];
      atomicAdd(db, s);
    }
  }
}

// ================================================================================================
// masked softmax over attention score rows (Modules.py:16-22) and its backward / tangent forms.
// Rows are S[z][q][0..Lk) with leading dimension ld (multiple of 8), z = b*H + h, keys j >= klen[b]
// are masked (-inf  =>  p = 0).  One warp per row, any Lk.
// ================================================================================================
__device__ __forceinline__ float ld_split(const bf16* hi, const bf16* lo, long long i) {
  float v = __bfloat162float(hi[i]);
  if (lo) v += __bfloat162float(lo[i]);
  return v;
}
__device__ __forceinline__ void st_split(bf16* hi, bf16* lo, long long i, float v) {
  bf16 h, l;
  split_bf16(v, h, l);
  hi[i] = h;
  if (lo) lo[i] = l;
}

// mode 0: P = softmax(S)                                  in: S (f32)                 out: hi/lo
// mode 1: dS = P*(dP - sum(P*dP))                         in: P (hi/lo), A = dP (f32) out: hi/lo
//         (also the tangent forward: Pdot = P*(Sdot - sum(P*Sdot)) with A = Sdot)
// mode 2: ddS = Pd*(dP - d) + P*(ddP - dd),  d = sum(P*dP), dd = sum(Pd*dP + P*ddP)
//                                                         in: P, Pd (hi/lo), A = dP, Bm = ddP (f32)
// Single pass over HBM: the row (<= 32*NPL keys) lives in registers, one read + one write per element.
template <int NPL, int MODE>     // MODE is a template parameter so that mode 0/1 do not pay mode 2's register footprint
__global__ void __launch_bounds__(ROW_THREADS) softmax_kernel(int /*mode*/, const float* __restrict__ A, const float* __restrict__ Bm,
                                                              const bf16* __restrict__ p_hi, const bf16* __restrict__ p_lo,
                                                              const bf16* __restrict__ pd_hi, const bf16* __restrict__ pd_lo,
                                                              const int64_t* __restrict__ klens, int H, int Lq, int Lk, int ld,
                                                              long long rows, bf16* __restrict__ o_hi, bf16* __restrict__ o_lo) {
  pdl_enter();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (long long r = static_cast<long long>(blockIdx.x) * ROW_WARPS + warp; r < rows; r += static_cast<long long>(gridDim.x) * ROW_WARPS) {
    const long long zidx = r / Lq;
    const int b = static_cast<int>(zidx / H);
    const int kl = klens ? static_cast<int>(min(static_cast<long long>(Lk), static_cast<long long>(klens[b]))) : Lk;
    const long long base = r * ld;
    float v[NPL];
    if (MODE == 0) {
      float m = -INFINITY;
#pragma unroll
      for (int i = 0; i < NPL; ++i) {
        const int j = lane + 32 * i;
        v[i] = j < kl ? A[base + j] : -INFINITY;
        m = fmaxf(m, v[i]);
      }
      m = warp_max(m);
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < NPL; ++i) {
        v[i] = (lane + 32 * i) < kl ? __expf(v[i] - m) : 0.f;
        s += v[i];
      }
      s = warp_sum(s);
      const float inv = 1.f / s;
#pragma unroll
      for (int i = 0; i < NPL; ++i) {
        const int j = lane + 32 * i;
        if (j < ld) st_split(o_hi, o_lo, base + j, v[i] * inv);
      }
    } else if (MODE == 1) {
      float a[NPL];
      float d = 0.f;
#pragma unroll
      for (int i = 0; i < NPL; ++i) {
        const int j = lane + 32 * i;
        if (j < kl) {
          v[i] = ld_split(p_hi, p_lo, base + j);
          a[i] = A[base + j];
        } else {
          v[i] = 0.f;
          a[i] = 0.f;
        }
        d += v[i] * a[i];
      }
      d = warp_sum(d);
#pragma unroll
      for (int i = 0; i < NPL; ++i) {
        const int j = lane + 32 * i;
        if (j < ld) st_split(o_hi, o_lo, base + j, v[i] * (a[i] - d));
      }
    } else {
      float pd[NPL], u[NPL];           // u = pd*a + p*b  =>  out = u - pd*d - p*dd  (3 row arrays instead of 4)
      float d = 0.f, dd = 0.f;
#pragma unroll
      for (int i = 0; i < NPL; ++i) {
        const int j = lane + 32 * i;
        if (j < kl) {
          v[i] = ld_split(p_hi, p_lo, base + j);
          pd[i] = ld_split(pd_hi, pd_lo, base + j);
          const float a = A[base + j], bb = Bm[base + j];
          d += v[i] * a;
          u[i] = pd[i] * a + v[i] * bb;
        } else {
          v[i] = pd[i] = u[i] = 0.f;
        }
        dd += u[i];
      }
      d = warp_sum(d);
      dd = warp_sum(dd);
#pragma unroll
      for (int i = 0; i < NPL; ++i) {
        const int j = lane + 32 * i;
        if (j < ld) st_split(o_hi, o_lo, base + j, u[i] - pd[i] * d - v[i] * dd);
      }
    }
  }
}

// ---- vectorised variant (ld % 4 == 0, 16-byte aligned bases): lane l owns keys [128*i + 4*l, +4) for i < NQ -------------
// 128-bit loads of the fp32 input and 64-bit loads / stores of the bf16 halves (the scalar kernel issued 2-byte stores:
// 64 B per warp instruction).
__device__ __forceinline__ float4 ld4_split(const bf16* hi, const bf16* lo, long long i) {
  const uint2 h = *reinterpret_cast<const uint2*>(hi + i);
  float4 v = make_float4(__uint_as_float(h.x << 16), __uint_as_float(h.x & 0xFFFF0000u), __uint_as_float(h.y << 16),
                         __uint_as_float(h.y & 0xFFFF0000u));
  if (lo) {
    const uint2 l = *reinterpret_cast<const uint2*>(lo + i);
    v.x += __uint_as_float(l.x << 16); v.y += __uint_as_float(l.x & 0xFFFF0000u);
    v.z += __uint_as_float(l.y << 16); v.w += __uint_as_float(l.y & 0xFFFF0000u);
  }
  return v;
}
__device__ __forceinline__ void st4_split(bf16* hi, bf16* lo, long long i, float4 v) {
  bf16 h0, h1, h2, h3, l0, l1, l2, l3;
  split_bf16(v.x, h0, l0); split_bf16(v.y, h1, l1); split_bf16(v.z, h2, l2); split_bf16(v.w, h3, l3);
  uint2 h, l;
  h.x = uint32_t(__bfloat16_as_ushort(h0)) | (uint32_t(__bfloat16_as_ushort(h1)) << 16);
  h.y = uint32_t(__bfloat16_as_ushort(h2)) | (uint32_t(__bfloat16_as_ushort(h3)) << 16);
  *reinterpret_cast<uint2*>(hi + i) = h;
  if (lo) {
    l.x = uint32_t(__bfloat16_as_ushort(l0)) | (uint32_t(__bfloat16_as_ushort(l1)) << 16);
    l.y = uint32_t(__bfloat16_as_ushort(l2)) | (uint32_t(__bfloat16_as_ushort(l3)) << 16);
    *reinterpret_cast<uint2*>(lo + i) = l;
  }
}
#define F4_EACH(v, EXPR_X, EXPR_Y, EXPR_Z, EXPR_W) { (v).x = EXPR_X; (v).y = EXPR_Y; (v).z = EXPR_Z; (v).w = EXPR_W; }

template <int NQ, int MODE>
__global__ void __launch_bounds__(ROW_THREADS, (MODE == 2 && NQ > 4) ? 2 : 1) softmax_vec_kernel(const float* __restrict__ A, const float* __restrict__ Bm,
                                                                  const bf16* __restrict__ p_hi, const bf16* __restrict__ p_lo,
                                                                  const bf16* __restrict__ pd_hi, const bf16* __restrict__ pd_lo,
                                                                  const int64_t* __restrict__ klens, int H, int Lq, int Lk, int ld,
                                                                  long long rows, bf16* __restrict__ o_hi, bf16* __restrict__ o_lo) {
  pdl_enter();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long r = static_cast<long long>(blockIdx.x) * ROW_WARPS + warp; r < rows; r += static_cast<long long>(gridDim.x) * ROW_WARPS) {
    const long long zidx = r / Lq;
    const int b = static_cast<int>(zidx / H);
    const int kl = klens ? static_cast<int>(min(static_cast<long long>(Lk), static_cast<long long>(klens[b]))) : Lk;
    const long long base = r * ld;
    float4 v[NQ];
    if (MODE == 0) {
      float m = -INFINITY;
#pragma unroll
      for (int i = 0; i < NQ; ++i) {
        const int j = 128 * i + 4 * lane;
        float4 t = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        if (j < ld) t = *reinterpret_cast<const float4*>(A + base + j);
        if (j + 0 >= kl) t.x = -INFINITY;
        if (j + 1 >= kl) t.y = -INFINITY;
        if (j + 2 >= kl) t.z = -INFINITY;
        if (j + 3 >= kl) t.w = -INFINITY;
        v[i] = t;
        m = fmaxf(fmaxf(m, fmaxf(t.x, t.y)), fmaxf(t.z, t.w));
      }
      m = warp_max(m);
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < NQ; ++i) {
        const int j = 128 * i + 4 * lane;
        F4_EACH(v[i], (j + 0 < kl ? __expf(v[i].x - m) : 0.f), (j + 1 < kl ? __expf(v[i].y - m) : 0.f),
                (j + 2 < kl ? __expf(v[i].z - m) : 0.f), (j + 3 < kl ? __expf(v[i].w - m) : 0.f));
        s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
      }
      s = warp_sum(s);
      const float inv = 1.f / s;
#pragma unroll
      for (int i = 0; i < NQ; ++i) {
        const int j = 128 * i + 4 * lane;
        if (j < ld) st4_split(o_hi, o_lo, base + j, make_float4(v[i].x * inv, v[i].y * inv, v[i].z * inv, v[i].w * inv));
      }
    } else if (MODE == 1) {
      float4 a[NQ];
      float d = 0.f;
#pragma unroll
      for (int i = 0; i < NQ; ++i) {
        const int j = 128 * i + 4 * lane;
        v[i] = z4;
        a[i] = z4;
        if (j < kl) {          // groups straddling kl: P is exactly 0 beyond kl (written so by mode 0), so no per-element mask
          v[i] = ld4_split(p_hi, p_lo, base + j);
          a[i] = *reinterpret_cast<const float4*>(A + base + j);
          if (j + 1 >= kl) { v[i].y = 0.f; a[i].y = 0.f; }   // (columns past Lk are never written by the GEMM: may hold NaN)
          if (j + 2 >= kl) { v[i].z = 0.f; a[i].z = 0.f; }
          if (j + 3 >= kl) { v[i].w = 0.f; a[i].w = 0.f; }
        }
        d += (v[i].x * a[i].x + v[i].y * a[i].y) + (v[i].z * a[i].z + v[i].w * a[i].w);
      }
      d = warp_sum(d);
#pragma unroll
      for (int i = 0; i < NQ; ++i) {
        const int j = 128 * i + 4 * lane;
        if (j < ld)
          st4_split(o_hi, o_lo, base + j, make_float4(v[i].x * (a[i].x - d), v[i].y * (a[i].y - d), v[i].z * (a[i].z - d),
                                                       v[i].w * (a[i].w - d)));
      }
    } else {
      float4 pd[NQ], u[NQ];
      float d = 0.f, dd = 0.f;
#pragma unroll
      for (int i = 0; i < NQ; ++i) {
        const int j = 128 * i + 4 * lane;
        v[i] = pd[i] = u[i] = z4;
        if (j < kl) {
          v[i] = ld4_split(p_hi, p_lo, base + j);
          pd[i] = ld4_split(pd_hi, pd_lo, base + j);
          if (j + 1 >= kl) { v[i].y = 0.f; pd[i].y = 0.f; }
          if (j + 2 >= kl) { v[i].z = 0.f; pd[i].z = 0.f; }
          if (j + 3 >= kl) { v[i].w = 0.f; pd[i].w = 0.f; }
          float4 a = *reinterpret_cast<const float4*>(A + base + j);
          float4 bb = *reinterpret_cast<const float4*>(Bm + base + j);
          if (j + 1 >= kl) { a.y = 0.f; bb.y = 0.f; }
          if (j + 2 >= kl) { a.z = 0.f; bb.z = 0.f; }
          if (j + 3 >= kl) { a.w = 0.f; bb.w = 0.f; }
          d += (v[i].x * a.x + v[i].y * a.y) + (v[i].z * a.z + v[i].w * a.w);
          F4_EACH(u[i], pd[i].x * a.x + v[i].x * bb.x, pd[i].y * a.y + v[i].y * bb.y, pd[i].z * a.z + v[i].z * bb.z,
                  pd[i].w * a.w + v[i].w * bb.w);
        }
        dd += (u[i].x + u[i].y) + (u[i].z + u[i].w);
      }
      d = warp_sum(d);
      dd = warp_sum(dd);
#pragma unroll
      for (int i = 0; i < NQ; ++i) {
        const int j = 128 * i + 4 * lane;
        if (j < ld)
          st4_split(o_hi, o_lo, base + j,
                    make_float4(u[i].x - pd[i].x * d - v[i].x * dd, u[i].y - pd[i].y * d - v[i].y * dd,
                                u[i].z - pd[i].z * d - v[i].z * dd, u[i].w - pd[i].w * d - v[i].w * dd));
      }
    }
  }
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
#define DISPATCH_NV(C, CALL)                                      \
  switch (C) {                                                    \
    case 128: { constexpr int NV = 1; CALL; } break;              \
    case 256: { constexpr int NV = 2; CALL; } break;              \
    case 512: { constexpr int NV = 4; CALL; } break;              \
    default:                                                      \
      mtts_set_error("row op: C=%d unsupported (128/256/512)", C); \
      return MTTS_EUNSUPPORTED;                                   \
  }

extern "C" int mtts_ln_fwd(const float* y, const float* res, const float* gamma, const float* beta, const int64_t* lens,
                           int T, int64_t R, int C, float eps, float* z_out, float* stats, float* out, void* out_hi,
                           void* out_lo, uint32_t pre_thr, uint32_t pre_seed, float pre_scale, uint32_t post_thr, uint32_t post_seed, float post_scale, const uint32_t* drop_salt, mtts_stream stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  MTTS_REQUIRE(y && gamma && beta && R > 0, "ln_fwd: bad args");
  DISPATCH_NV(C, MTTS_CHECK_CUDA(mtts_launch(ln_fwd_kernel<NV>, dim3(row_grid(R)), dim3(ROW_THREADS), 0, s, y, res, gamma, beta, lens, T, R, eps, z_out, stats, out,
                                                                        static_cast<bf16*>(out_hi), static_cast<bf16*>(out_lo), DropSite{pre_thr, pre_seed, pre_scale, drop_salt}, DropSite{post_thr, post_seed, post_scale, drop_salt})));
  MTTS_CHECK_LAUNCH();
  return MTTS_OK;
}

extern "C" int mtts_ln_bwd(const float* dy, const float* z, const float* stats, const float* gamma, const int64_t* lens, int T,
                           int64_t R, int C, int relu_gate, float* dz, void* dz_hi, void* dz_lo, float* dgamma, float* dbeta,
                           float* dbias, uint32_t pre_thr, uint32_t pre_seed, float pre_scale, uint32_t post_thr, uint32_t post_seed, float post_scale, const uint32_t* drop_salt, mtts_stream stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  MTTS_REQUIRE(dy && z && stats && gamma && R > 0, "ln_bwd: bad args");
  DISPATCH_NV(C, MTTS_CHECK_CUDA(mtts_launch(ln_bwd_kernel<NV>, dim3(row_grid(R, true)), dim3(ROW_THREADS), 0, s, dy, z, stats, gamma, lens, T, R, relu_gate, dz,
                                                                        static_cast<bf16*>(dz_hi), static_cast<bf16*>(dz_lo),
                                                                        dgamma, dbeta, dbias, DropSite{pre_thr, pre_seed, pre_scale, drop_salt}, DropSite{post_thr, post_seed, post_scale, drop_salt})));
  MTTS_CHECK_LAUNCH();
  return MTTS_OK;
}

extern "C" int mtts_ln_tfwd(const float* ydot, const float* resdot, const float* z, const float* stats, const float* gamma,
                            const float* gdot, const float* bdot, const int64_t* lens, int T, int64_t R, int C,
                            float* zdot_out, float* out, void* out_hi, void* out_lo, uint32_t pre_thr, uint32_t pre_seed, float pre_scale, uint32_t post_thr, uint32_t post_seed, float post_scale, const uint32_t* drop_salt, mtts_stream stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  MTTS_REQUIRE(ydot && z && stats && gamma && R > 0, "ln_tfwd: bad args");
  DISPATCH_NV(C, MTTS_CHECK_CUDA(mtts_launch(ln_tfwd_kernel<NV>, dim3(row_grid(R)), dim3(ROW_THREADS), 0, s, ydot, resdot, z, stats, gamma, gdot, bdot, lens, T, R,
                                                                         zdot_out, out, static_cast<bf16*>(out_hi),
                                                                         static_cast<bf16*>(out_lo), DropSite{pre_thr, pre_seed, pre_scale, drop_salt}, DropSite{post_thr, post_seed, post_scale, drop_salt})));
  MTTS_CHECK_LAUNCH();
  return MTTS_OK;
}

extern "C" int mtts_ln_tbwd(const float* dy, const float* ddy, const float* z, const float* zdot, const float* stats,
                            const float* gamma, const float* gdot, const int64_t* lens, int T, int64_t R, int C, int relu_gate,
                            float* ddz, void* ddz_hi, void* ddz_lo, float* ddgamma, float* ddbeta, float* ddbias, uint32_t pre_thr, uint32_t pre_seed, float pre_scale, uint32_t post_thr, uint32_t post_seed, float post_scale, const uint32_t* drop_salt,
                            mtts_stream stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  MTTS_REQUIRE(dy && ddy && z && zdot && stats && gamma && R > 0, "ln_tbwd: bad args");
  DISPATCH_NV(C, MTTS_CHECK_CUDA(mtts_launch(ln_tbwd_kernel<NV>, dim3(row_grid(R, true)), dim3(ROW_THREADS), 0, s, dy, ddy, z, zdot, stats, gamma, gdot, lens, T, R,
                                                                         relu_gate, ddz, static_cast<bf16*>(ddz_hi),
                                                                         static_cast<bf16*>(ddz_lo), ddgamma, ddbeta, ddbias, DropSite{pre_thr, pre_seed, pre_scale, drop_salt}, DropSite{post_thr, post_seed, post_scale, drop_salt})));
  MTTS_CHECK_LAUNCH();
  return MTTS_OK;
}

extern "C" int mtts_rowdot_fwd(const float* h, const float* hdot, const float* w, const float* wdot, const float* b,
                               const float* bdot, const int64_t* lens, int T, int64_t R, int C, float* out,
                               mtts_stream stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  MTTS_REQUIRE(h && w && out && R > 0 && (hdot || b), "rowdot_fwd: bad args");
  DISPATCH_NV(C, MTTS_CHECK_CUDA(mtts_launch(rowdot_fwd_kernel<NV>, dim3(row_grid(R)), dim3(ROW_THREADS), 0, s, h, hdot, w, wdot, b, bdot, lens, T, R, out)));
  MTTS_CHECK_LAUNCH();
  return MTTS_OK;
}

extern "C" int mtts_rowdot_bwd(const float* dout, const float* ddout, const float* h, const float* hdot, const float* w,
                               const float* wdot, const int64_t* lens, int T, int64_t R, int C, float* dh, float* dw,
                               float* db, mtts_stream stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  MTTS_REQUIRE(dout && h && w && dh && R > 0, "rowdot_bwd: bad args");
  MTTS_REQUIRE(!ddout || hdot, "rowdot_bwd: tangent mode needs hdot");
  DISPATCH_NV(C, MTTS_CHECK_CUDA(mtts_launch(rowdot_bwd_kernel<NV>, dim3(row_grid(R, true)), dim3(ROW_THREADS), 0, s, dout, ddout, h, hdot, w, wdot, lens, T, R, dh, dw, db)));
  MTTS_CHECK_LAUNCH();
  return MTTS_OK;
}

// Forward softmax for rows LONGER than 1024 keys: only reachable under model.eval(), where a sequence beyond max_seq_len keeps
// its length (Models.py:148-156) — three passes over the row (max, sum, write), one warp per row; no backward forms exist
// because train mode truncates to max_seq_len (Models.py:161-166).
__global__ void __launch_bounds__(ROW_THREADS) softmax_long_fwd_kernel(const float* __restrict__ A, const int64_t* __restrict__ klens, int H, int Lq,
                                                                       int Lk, int ld, long long rows, bf16* __restrict__ o_hi,
                                                                       bf16* __restrict__ o_lo) {
  pdl_enter();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (long long r = static_cast<long long>(blockIdx.x) * ROW_WARPS + warp; r < rows; r += static_cast<long long>(gridDim.x) * ROW_WARPS) {
    const int b = static_cast<int>((r / Lq) / H);
    const int kl = klens ? static_cast<int>(min(static_cast<long long>(Lk), static_cast<long long>(klens[b]))) : Lk;
    const long long base = r * ld;
    float m = -INFINITY;
    for (int j = lane; j < kl; j += 32) m = fmaxf(m, A[base + j]);
    m = warp_max(m);
    float s = 0.f;
    for (int j = lane; j < kl; j += 32) s += __expf(A[base + j] - m);
    s = warp_sum(s);
    const float inv = 1.f / s;
    for (int j = lane; j < ld; j += 32) st_split(o_hi, o_lo, base + j, j < kl ? __expf(A[base + j] - m) * inv : 0.f);
  }
}

extern "C" int mtts_softmax(int mode, const float* A, const float* Bm, const void* p_hi, const void* p_lo, const void* pd_hi,
                            const void* pd_lo, const int64_t* klens, int nz, int H, int Lq, int Lk, int ld, void* o_hi,
                            void* o_lo, mtts_stream stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  MTTS_REQUIRE(mode >= 0 && mode <= 2 && A && o_hi && nz > 0 && Lq > 0 && Lk > 0 && ld >= Lk, "softmax: bad args");
  MTTS_REQUIRE(mode == 0 || p_hi, "softmax: mode %d needs P", mode);
  MTTS_REQUIRE(mode != 2 || (pd_hi && Bm), "softmax: mode 2 needs Pdot and ddP");
  const long long rows = static_cast<long long>(nz) * Lq;
  if (ld > 1024) {
    MTTS_REQUIRE(mode == 0, "softmax: rows longer than 1024 keys exist only in eval-mode forwards (train mode truncates to max_seq_len)");
    MTTS_CHECK_CUDA(mtts_launch(softmax_long_fwd_kernel, dim3(row_grid(rows)), dim3(ROW_THREADS), 0, s, A, klens, H, Lq, Lk, ld, rows,
                                static_cast<bf16*>(o_hi), static_cast<bf16*>(o_lo)));
    MTTS_CHECK_LAUNCH();
    return MTTS_OK;
  }
#define SM_LAUNCH_M(NPL, MODE)                                                                                                  \
  MTTS_CHECK_CUDA(mtts_launch(softmax_kernel<NPL, MODE>, dim3(row_grid(rows)), dim3(ROW_THREADS), 0, s, mode, A, Bm, static_cast<const bf16*>(p_hi),                  \
                                                              static_cast<const bf16*>(p_lo), static_cast<const bf16*>(pd_hi), \
                                                              static_cast<const bf16*>(pd_lo), klens, H, Lq, Lk, ld, rows,      \
                                                              static_cast<bf16*>(o_hi), static_cast<bf16*>(o_lo)))
#define SM_LAUNCH(NPL)                   \
  do {                                   \
    if (mode == 0) SM_LAUNCH_M(NPL, 0);  \
    else if (mode == 1) SM_LAUNCH_M(NPL, 1); \
    else SM_LAUNCH_M(NPL, 2);            \
  } while (0)
  // vector path: rows a multiple of 4 keys long with 16-byte aligned bases (the engine pads keys to a multiple of 32)
  const uintptr_t al = reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(Bm) | reinterpret_cast<uintptr_t>(p_hi) |
                       reinterpret_cast<uintptr_t>(p_lo) | reinterpret_cast<uintptr_t>(pd_hi) | reinterpret_cast<uintptr_t>(pd_lo) |
                       reinterpret_cast<uintptr_t>(o_hi) | reinterpret_cast<uintptr_t>(o_lo);
  if ((ld & 3) == 0 && (al & 15) == 0) {
#define SMV_LAUNCH_M(NQ, MODE)                                                                                                   \
  MTTS_CHECK_CUDA(mtts_launch(softmax_vec_kernel<NQ, MODE>, dim3(softmax_grid(rows)), dim3(ROW_THREADS), 0, s, A, Bm, static_cast<const bf16*>(p_hi), \
                              static_cast<const bf16*>(p_lo), static_cast<const bf16*>(pd_hi), static_cast<const bf16*>(pd_lo), klens, H, Lq, \
                              Lk, ld, rows, static_cast<bf16*>(o_hi), static_cast<bf16*>(o_lo)))
#define SMV_LAUNCH(NQ)                        \
  do {                                        \
    if (mode == 0) SMV_LAUNCH_M(NQ, 0);       \
    else if (mode == 1) SMV_LAUNCH_M(NQ, 1);  \
    else SMV_LAUNCH_M(NQ, 2);                 \
  } while (0)
    const int nq = (ld + 127) / 128;
    if (nq <= 1) SMV_LAUNCH(1);
    else if (nq <= 2) SMV_LAUNCH(2);
    else if (nq <= 4) SMV_LAUNCH(4);
    else if (nq <= 6) SMV_LAUNCH(6);
    else if (nq <= 7) SMV_LAUNCH(7);
    else SMV_LAUNCH(8);
#undef SMV_LAUNCH
#undef SMV_LAUNCH_M
    MTTS_CHECK_LAUNCH();
    return MTTS_OK;
  }
  if (ld <= 128) SM_LAUNCH(4);
  else if (ld <= 256) SM_LAUNCH(8);
  else if (ld <= 512) SM_LAUNCH(16);
  else SM_LAUNCH(32);
#undef SM_LAUNCH
#undef SM_LAUNCH_M
  MTTS_CHECK_LAUNCH();
  return MTTS_OK;
}
