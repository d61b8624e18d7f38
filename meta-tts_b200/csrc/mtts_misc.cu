// mtts_misc.cu — the remaining HBM-bound kernels of the FastSpeech2 hot path:
//   embedding gathers (+ positional encoding, + bucketize), speaker-embedding add, column sums (bias
//   gradients), BatchNorm1d with batch statistics (+ tanh) in forward / backward / tangent forms,
//   the FastSpeech2 loss (forward, backward, tangent-backward) and the flat-arena elementwise
//   kernels (bf16 split, fused in-place-style SGD + split, axpby, sum of squares, Adam + clip).
//
// Reference call sites: nn.Embedding + position_enc Models.py:89-91,158-160; bucketize + Embedding
// modules.py:80-100; speaker embedding add base_adaptor.py:64-70,80-84; BatchNorm1d/tanh
// Layers.py:129-137; FastSpeech2Loss loss.py:19-92; l2l maml_update (p - lr*g); Adam + clip
// lightning/optimizer.py:6-16, main.py:61.
#include "mtts_common.cuh"

namespace {

constexpr int EW_THREADS = 256;
inline int ew_grid(long long n, int per_thread = 4) {
  long long need = (n + static_cast<long long>(EW_THREADS) * per_thread - 1) / (static_cast<long long>(EW_THREADS) * per_thread);
  long long cap = 148LL * 8;
  return static_cast<int>(need < 1 ? 1 : (need < cap ? need : cap));
}

__device__ __forceinline__ void store_split4(bf16* hi, bf16* lo, long long i, float4 v) {
  const float a[4] = {v.x, v.y, v.z, v.w};
  uint16_t h[4], l[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    bf16 hh, ll;
    split_bf16(a[j], hh, ll);
    h[j] = __bfloat16_as_ushort(hh);
    l[j] = __bfloat16_as_ushort(ll);
  }
  if (hi) *reinterpret_cast<uint2*>(hi + i) = make_uint2(h[0] | (uint32_t(h[1]) << 16), h[2] | (uint32_t(h[3]) << 16));
  if (lo) *reinterpret_cast<uint2*>(lo + i) = make_uint2(l[0] | (uint32_t(l[1]) << 16), l[2] | (uint32_t(l[3]) << 16));
}

// ================================================================================================
// embedding gather:  out[r,:] = table[idx[r],:] (+ base[r,:]) (+ pos[r % T,:])      C % 4 == 0
// ================================================================================================
__global__ void embed_fwd_kernel(const int64_t* __restrict__ idx, const float* __restrict__ table,
                                 const float* __restrict__ base, const float* __restrict__ pos, int T, long long R, int C,
                                 float* __restrict__ out, bf16* __restrict__ hi, bf16* __restrict__ lo) {
  pdl_enter();
  const int c4 = C >> 2;
  const long long total = R * c4;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / c4;
    const int c = static_cast<int>(i - r * c4) * 4;
    float4 v = __ldg(reinterpret_cast<const float4*>(table + idx[r] * C + c));
    if (base) {
      const float4 b = *reinterpret_cast<const float4*>(base + r * C + c);
      v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
    }
    if (pos) {
      const float4 p = __ldg(reinterpret_cast<const float4*>(pos + (r % T) * C + c));
      v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w;
    }
    if (out) *reinterpret_cast<float4*>(out + r * C + c) = v;
    if (hi) store_split4(hi, lo, r * C + c, v);
  }
}
// dtable[idx[r],:] += scale * dy[r,:]   (rows with idx == skip_idx are skipped: padding_idx)
__global__ void embed_bwd_kernel(const int64_t* __restrict__ idx, const float* __restrict__ dy, long long R, int C,
                                 long long skip_idx, float scale, float* __restrict__ dtable) {
  pdl_enter();
  const long long total = R * C;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / C;
    const int c = static_cast<int>(i - r * C);
    const long long j = idx[r];
    if (j == skip_idx) continue;
    atomicAdd(dtable + j * C + c, scale * dy[i]);
  }
}
// deterministic variant: thread = one column c, which it owns in every table row; rows are added in row order (no atomics)
__global__ void embed_bwd_ordered_kernel(const int64_t* __restrict__ idx, const float* __restrict__ dy, long long R, int C,
                                         long long skip_idx, float scale, float* __restrict__ dtable) {
  pdl_enter();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  for (long long r = 0; r < R; ++r) {
    const long long j = idx[r];
    if (j == skip_idx) continue;
    dtable[j * C + c] += scale * dy[r * C + c];
  }
}
// torch.bucketize(v, bins) (right=False): out = #{ bins < v }  (lower bound)
__global__ void bucketize_kernel(const float* __restrict__ v, const float* __restrict__ bins, int nb, long long R,
                                 int64_t* __restrict__ out) {
  pdl_enter();
  const long long r = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const float x = v[r];
  int lo = 0, hi = nb;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (bins[mid] < x) lo = mid + 1; else hi = mid;
  }
  out[r] = lo;
}

// out[b,t,:] = x[b,t,:] + vec[b*vec_bstride + :] (+ pos[t,:])
__global__ void add_rowvec_kernel(const float* __restrict__ x, const float* __restrict__ vec, long long vec_bstride,
                                  const float* __restrict__ pos, int T, long long R, int C, float* __restrict__ out,
                                  bf16* __restrict__ hi, bf16* __restrict__ lo) {
  pdl_enter();
  const int c4 = C >> 2;
  const long long total = R * c4;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / c4;
    const int c = static_cast<int>(i - r * c4) * 4;
    const long long b = r / T;
    float4 v = *reinterpret_cast<const float4*>(x + r * C + c);
    if (vec) {
      const float4 s = __ldg(reinterpret_cast<const float4*>(vec + b * vec_bstride + c));
      v.x += s.x; v.y += s.y; v.z += s.z; v.w += s.w;
    }
    if (pos) {
      const float4 p = __ldg(reinterpret_cast<const float4*>(pos + (r - b * T) * C + c));
      v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w;
    }
    if (out) *reinterpret_cast<float4*>(out + r * C + c) = v;
    if (hi) store_split4(hi, lo, r * C + c, v);
  }
}

// speaker embedding: out[q,:] = average ? mean_i table[ids[i],:] : table[ids[q],:]
__global__ void spk_embed_kernel(const int64_t* __restrict__ ids, const float* __restrict__ table, int n, int C, int average,
                                 int n_out, float* __restrict__ out) {
  pdl_enter();
  const int q = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float v;
    if (average) {
      float s = 0.f;
      for (int i = 0; i < n; ++i) s += table[ids[i] * C + c];
      v = s / n;
    } else {
      v = table[ids[q] * C + c];
    }
    out[static_cast<long long>(q) * C + c] = v;
  }
}
__global__ void spk_embed_bwd_kernel(const int64_t* __restrict__ ids, const float* __restrict__ dspk, int n, int C,
                                     int average, int n_out, float scale, float* __restrict__ dtable) {
  pdl_enter();
  for (int c = threadIdx.x + blockIdx.x * blockDim.x; c < C; c += blockDim.x * gridDim.x) {
    if (average) {
      float s = 0.f;
      for (int q = 0; q < n_out; ++q) s += dspk[static_cast<long long>(q) * C + c];
      s = s * scale / n;
      for (int i = 0; i < n; ++i) atomicAdd(dtable + ids[i] * C + c, s);
    } else {
      for (int i = 0; i < n; ++i) atomicAdd(dtable + ids[i] * C + c, scale * dspk[static_cast<long long>(i) * C + c]);
    }
  }
}

// ================================================================================================
// column sums:  out[z, c] (+)= sum_{r < R} src[z, r, c]     src fp32 or bf16 hi(+lo)
// grid (col blocks, row chunks, nb); thread = one column, coalesced across threads
// ================================================================================================
__global__ void colsum_kernel(const float* __restrict__ f32, const bf16* __restrict__ hi, const bf16* __restrict__ lo,
                              long long R, int C, int rows_per_cta, float* __restrict__ out) {
  pdl_enter();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const long long zoff = static_cast<long long>(blockIdx.z) * R * C;
  const long long r0 = static_cast<long long>(blockIdx.y) * rows_per_cta;
  const long long r1 = min(R, r0 + rows_per_cta);
  float s = 0.f;
  if (f32) {
    for (long long r = r0; r < r1; ++r) s += f32[zoff + r * C + c];
  } else {
    for (long long r = r0; r < r1; ++r) {
      float v = __bfloat162float(hi[zoff + r * C + c]);
      if (lo) v += __bfloat162float(lo[zoff + r * C + c]);
      s += v;
    }
  }
  atomicAdd(out + static_cast<long long>(blockIdx.z) * C + c, s);
}

// ================================================================================================
// BatchNorm1d, train mode (batch statistics over all R = B*T rows, padded frames included)
// ================================================================================================
// generic per-column multi-sum reducer; MODE selects the summands
//   0: s0 = sum x
//   1: s0 = sum (x - mean)^2                                  (mean = ws_in[c] / R)
//   2: bwd   : g = dout*(1-o^2 | 1); s0 = sum g, s1 = sum g*xh
//   3: tfwd  : s0 = sum xd, s1 = sum xd*xh
//   4: tbwd  : s0 = sum g, s1 = sum g*xh, s2 = sum gd, s3 = sum(gd*xh + g*xhd)
struct BnArgs {
  const float* x; const float* xdot; const float* dout; const float* ddout; const float* o; const float* odot;
  const float* stats;    // [2*C] mean, rstd
  const float* tsums;    // [2*C] mean(xdot), mean(xdot*xhat)
  const float* ws_in;    // mode 1
  long long R; int C; int tanh_flag; int rows_per_cta;
  DropSite drop;       // dropout on the (tanh'd) output; the stored output o is the DROPPED one
};

template <int MODE>
__global__ void bn_reduce_kernel(BnArgs a, float* __restrict__ ws) {
  pdl_enter();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= a.C) return;
  const long long r0 = static_cast<long long>(blockIdx.y) * a.rows_per_cta;
  const long long r1 = min(a.R, r0 + a.rows_per_cta);
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  float mean = 0.f, rstd = 0.f, a1 = 0.f, a2 = 0.f;
  if (MODE == 1) mean = a.ws_in[c] / a.R;
  if (MODE >= 2) { mean = a.stats[c]; rstd = a.stats[a.C + c]; }
  if (MODE == 4) { a1 = a.tsums[c]; a2 = a.tsums[a.C + c]; }
  for (long long r = r0; r < r1; ++r) {
    const long long i = r * a.C + c;
    const float x = a.x[i];
    if (MODE == 0) { s0 += x; continue; }
    if (MODE == 1) { const float d = x - mean; s0 += d * d; continue; }
    const float xh = (x - mean) * rstd;
    if (MODE == 3) { const float xd = a.xdot[i]; s0 += xd; s1 += xd * xh; continue; }
    const float df = drop_factor(a.drop, static_cast<uint32_t>(i));      // 0 or scale
    const float inv = a.drop.thr ? 1.f / a.drop.scale : 1.f;
    float g = a.dout[i] * df;
    float t = 1.f, o = 0.f;
    if (a.tanh_flag) { o = a.o[i] * inv; t = 1.f - o * o; g *= t; }
    s0 += g; s1 += g * xh;
    if (MODE == 4) {
      float gd = a.ddout[i] * df * t;
      if (a.tanh_flag) gd -= 2.f * o * (a.odot[i] * inv) * (a.dout[i] * df);
      const float xhd = rstd * (a.xdot[i] - a1 - xh * a2);
      s2 += gd; s3 += gd * xh + g * xhd;
    }
  }
  atomicAdd(ws + c, s0);
  if (MODE >= 2) atomicAdd(ws + a.C + c, s1);
  if (MODE == 4) { atomicAdd(ws + 2 * a.C + c, s2); atomicAdd(ws + 3 * a.C + c, s3); }
}

// Batch statistics in ONE pass over x: shifted sums  sum(x - s), sum((x - s)^2)  with the per-channel shift s = x[0, c]
// (a value of the same magnitude as the mean, so the textbook cancellation of E[x^2] - E[x]^2 does not arise:
// var = (S2 - S1^2/R)/R loses ~log2(1 + (s - mean)^2/var) bits, a handful for any realistic activation).  Four rows are
// in flight per thread (the row loop is otherwise a chain of dependent-latency loads: 12 us for 7 MB).
__global__ void bn_stats_kernel(const float* __restrict__ x, long long R, int C, int rows_per_cta, float* __restrict__ ws) {
  pdl_enter();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const long long r0 = static_cast<long long>(blockIdx.y) * rows_per_cta;
  const long long r1 = min(R, r0 + rows_per_cta);
  const float sh = x[c];
  float s1 = 0.f, s2 = 0.f;
  long long r = r0;
  for (; r + 4 <= r1; r += 4) {
    const float v0 = x[r * C + c], v1 = x[(r + 1) * C + c], v2 = x[(r + 2) * C + c], v3 = x[(r + 3) * C + c];
    const float d0 = v0 - sh, d1 = v1 - sh, d2 = v2 - sh, d3 = v3 - sh;
    s1 += (d0 + d1) + (d2 + d3);
    s2 += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
  }
  for (; r < r1; ++r) {
    const float d = x[r * C + c] - sh;
    s1 += d;
    s2 += d * d;
  }
  atomicAdd(ws + c, s1);
  atomicAdd(ws + C + c, s2);
}

// forward apply: y = (x-mean)*rstd*gamma + beta; o = tanh? tanh(y) : y ; running stats update by CTA row 0
__global__ void bn_fwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ ws /*shifted sums: sum(x-s), sum((x-s)^2), s = x[0,c]*/,
                                    const float* __restrict__ gamma, const float* __restrict__ beta, long long R, int C,
                                    float eps, float momentum, int tanh_flag, float* __restrict__ running_mean,
                                    float* __restrict__ running_var, float* __restrict__ stats, float* __restrict__ out,
                                    bf16* __restrict__ hi, bf16* __restrict__ lo, int rows_per_cta, DropSite drop) {
  pdl_enter();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float invR = 1.f / static_cast<float>(R);
  const float m1 = ws[c] * invR;
  const float mean = x[c] + m1;
  const float var = fmaxf(ws[C + c] * invR - m1 * m1, 0.f);
  const float rstd = rsqrtf(var + eps);
  if (blockIdx.y == 0) {
    stats[c] = mean;
    stats[C + c] = rstd;
    if (running_mean) {
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
      const float unbiased = R > 1 ? var * (static_cast<float>(R) / static_cast<float>(R - 1)) : var;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * unbiased;
    }
  }
  const float g = gamma[c], b = beta[c];
  const long long r0 = static_cast<long long>(blockIdx.y) * rows_per_cta;
  const long long r1 = min(R, r0 + rows_per_cta);
  auto emit = [&](long long i, float xv) {
    float y = (xv - mean) * rstd * g + b;
    if (tanh_flag) y = tanhf(y);
    y *= drop_factor(drop, static_cast<uint32_t>(i));
    if (out) out[i] = y;
    if (hi) {
      bf16 h, l;
      split_bf16(y, h, l);
      hi[i] = h;
      if (lo) lo[i] = l;
    }
  };
  long long r = r0;
  for (; r + 4 <= r1; r += 4) {                   // four independent loads in flight per thread
    const long long i0 = r * C + c;
    const float v0 = x[i0], v1 = x[i0 + C], v2 = x[i0 + 2LL * C], v3 = x[i0 + 3LL * C];
    emit(i0, v0);
    emit(i0 + C, v1);
    emit(i0 + 2LL * C, v2);
    emit(i0 + 3LL * C, v3);
  }
  for (; r < r1; ++r) emit(r * C + c, x[r * C + c]);
}
// backward apply: dx = rstd*gamma*(g - m1 - xh*m2); dgamma += sum g*xh; dbeta += sum g
__global__ void bn_bwd_apply_kernel(BnArgs a, const float* __restrict__ ws, const float* __restrict__ gamma,
                                    float* __restrict__ dx, bf16* __restrict__ hi, bf16* __restrict__ lo,
                                    float* __restrict__ dgamma, float* __restrict__ dbeta) {
  pdl_enter();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= a.C) return;
  const float mean = a.stats[c], rstd = a.stats[a.C + c];
  const float m1 = ws[c] / a.R, m2 = ws[a.C + c] / a.R;
  if (blockIdx.y == 0) {
    if (dbeta) atomicAdd(dbeta + c, ws[c]);
    if (dgamma) atomicAdd(dgamma + c, ws[a.C + c]);
  }
  const float k = rstd * gamma[c];
  const long long r0 = static_cast<long long>(blockIdx.y) * a.rows_per_cta;
  const long long r1 = min(a.R, r0 + a.rows_per_cta);
  for (long long r = r0; r < r1; ++r) {
    const long long i = r * a.C + c;
    const float xh = (a.x[i] - mean) * rstd;
    float g = a.dout[i] * drop_factor(a.drop, static_cast<uint32_t>(i));
    if (a.tanh_flag) { const float o = a.o[i] * (a.drop.thr ? 1.f / a.drop.scale : 1.f); g *= 1.f - o * o; }
    const float v = k * (g - m1 - xh * m2);
    if (dx) dx[i] = v;
    if (hi) {
      bf16 h, l;
      split_bf16(v, h, l);
      hi[i] = h;
      if (lo) lo[i] = l;
    }
  }
}
// tangent forward apply: xhd = rstd*(xd - a1 - xh*a2); yd = xhd*gamma + xh*gdot + bdot; od = tanh? yd*(1-o^2) : yd
__global__ void bn_tfwd_apply_kernel(BnArgs a, const float* __restrict__ ws, const float* __restrict__ gamma,
                                     const float* __restrict__ gdot, const float* __restrict__ bdot,
                                     float* __restrict__ tsums, float* __restrict__ od, bf16* __restrict__ hi,
                                     bf16* __restrict__ lo) {
  pdl_enter();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= a.C) return;
  const float mean = a.stats[c], rstd = a.stats[a.C + c];
  const float a1 = ws[c] / a.R, a2 = ws[a.C + c] / a.R;
  if (blockIdx.y == 0) { tsums[c] = a1; tsums[a.C + c] = a2; }
  const float g = gamma[c], gd = gdot ? gdot[c] : 0.f, bd = bdot ? bdot[c] : 0.f;
  const long long r0 = static_cast<long long>(blockIdx.y) * a.rows_per_cta;
  const long long r1 = min(a.R, r0 + a.rows_per_cta);
  for (long long r = r0; r < r1; ++r) {
    const long long i = r * a.C + c;
    const float xh = (a.x[i] - mean) * rstd;
    const float xhd = rstd * (a.xdot[i] - a1 - xh * a2);
    float v = xhd * g + xh * gd + bd;
    if (a.tanh_flag) { const float o = a.o[i] * (a.drop.thr ? 1.f / a.drop.scale : 1.f); v *= 1.f - o * o; }
    v *= drop_factor(a.drop, static_cast<uint32_t>(i));
    if (od) od[i] = v;
    if (hi) {
      bf16 h, l;
      split_bf16(v, h, l);
      hi[i] = h;
      if (lo) lo[i] = l;
    }
  }
}
// tangent backward apply
__global__ void bn_tbwd_apply_kernel(BnArgs a, const float* __restrict__ ws, const float* __restrict__ gamma,
                                     const float* __restrict__ gdot, float* __restrict__ ddx, bf16* __restrict__ hi,
                                     bf16* __restrict__ lo, float* __restrict__ ddgamma, float* __restrict__ ddbeta) {
  pdl_enter();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= a.C) return;
  const float mean = a.stats[c], rstd = a.stats[a.C + c];
  const float a1 = a.tsums[c], a2 = a.tsums[a.C + c];
  const float rd = -rstd * rstd * a2;
  const float m1 = ws[c] / a.R, m2 = ws[a.C + c] / a.R, m1d = ws[2 * a.C + c] / a.R, m2d = ws[3 * a.C + c] / a.R;
  if (blockIdx.y == 0) {
    if (ddbeta) atomicAdd(ddbeta + c, ws[2 * a.C + c]);
    if (ddgamma) atomicAdd(ddgamma + c, ws[3 * a.C + c]);
  }
  const float gm = gamma[c], gmd = gdot ? gdot[c] : 0.f;
  const long long r0 = static_cast<long long>(blockIdx.y) * a.rows_per_cta;
  const long long r1 = min(a.R, r0 + a.rows_per_cta);
  for (long long r = r0; r < r1; ++r) {
    const long long i = r * a.C + c;
    const float xh = (a.x[i] - mean) * rstd;
    const float xhd = rstd * (a.xdot[i] - a1 - xh * a2);
    const float df = drop_factor(a.drop, static_cast<uint32_t>(i));
    const float inv = a.drop.thr ? 1.f / a.drop.scale : 1.f;
    float g = a.dout[i] * df, gd = a.ddout[i] * df;
    if (a.tanh_flag) {
      const float o = a.o[i] * inv;
      const float t = 1.f - o * o;
      gd = gd * t - 2.f * o * (a.odot[i] * inv) * g;
      g *= t;
    }
    const float core = g - m1 - xh * m2;
    const float v = gmd * rstd * core + gm * (rd * core + rstd * (gd - m1d - xhd * m2 - xh * m2d));
    if (ddx) ddx[i] = v;
    if (hi) {
      bf16 h, l;
      split_bf16(v, h, l);
      hi[i] = h;
      if (lo) lo[i] = l;
    }
  }
}

inline void bn_launch_dims(long long R, int C, dim3& grid, dim3& block, int& rows_per_cta) {
  block = dim3(128);
  const int col_blocks = mtts_cdiv(C, 128);
  int row_chunks = mtts_deterministic() ? 1 : static_cast<int>(mtts_cdiv64(148LL * 4, col_blocks));   // deterministic: one CTA per column block
  if (row_chunks > R) row_chunks = static_cast<int>(R);
  rows_per_cta = static_cast<int>(mtts_cdiv64(R, row_chunks));
  row_chunks = static_cast<int>(mtts_cdiv64(R, rows_per_cta));
  grid = dim3(col_blocks, row_chunks);
}

// The apply passes have no cross-CTA reduction (the per-channel sums arrive in `ws`): they take a finer row split than the
// reduce passes — ~16 CTAs of 128 threads per SM instead of 4 — so that enough loads are in flight to reach the HBM rate
// (bn_tfwd_apply 18 -> 9 us on 3456 x 512).
inline void bn_apply_dims(long long R, int C, dim3& grid, int& rows_per_cta) {
  const int col_blocks = mtts_cdiv(C, 128);
  long long row_chunks = mtts_cdiv64(148LL * 16, col_blocks);
  if (row_chunks > R) row_chunks = R;
  rows_per_cta = static_cast<int>(mtts_cdiv64(R, row_chunks));
  grid = dim3(col_blocks, static_cast<unsigned>(mtts_cdiv64(R, rows_per_cta)));
}

// ================================================================================================
// FastSpeech2Loss (loss.py:19-92), phoneme-level pitch / energy
// ================================================================================================
struct LossArgs {
  const float* mel; const float* post; const float* mel_tgt; const int64_t* mel_lens;
  const float* p; const float* p_tgt; const float* e; const float* e_tgt; const float* logd; const int64_t* dur;
  const int64_t* src_lens;
  int B, T, L, NM;
};
__device__ __forceinline__ float block_sum(float v) {
  __shared__ float red[32];
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  v = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
  if (warp == 0) v = warp_sum(v);
  return v;   // valid in thread 0
}
// ws[0..4] = sum|mel-t|, sum|post-t|, sum (p-pt)^2, sum (e-et)^2, sum (logd - log(d+1))^2
__global__ void loss_sums_kernel(LossArgs a, float* __restrict__ ws) {
  pdl_enter();
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, s4 = 0.f;
  const long long nmel = static_cast<long long>(a.B) * a.T * a.NM;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < nmel; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long row = i / a.NM;
    const int b = static_cast<int>(row / a.T);
    const int t = static_cast<int>(row - static_cast<long long>(b) * a.T);
    if (t < a.mel_lens[b]) {
      const float tg = a.mel_tgt[i];
      s0 += fabsf(a.mel[i] - tg);
      s1 += fabsf(a.post[i] - tg);
    }
  }
  const long long nsrc = static_cast<long long>(a.B) * a.L;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < nsrc; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int b = static_cast<int>(i / a.L);
    const int l = static_cast<int>(i - static_cast<long long>(b) * a.L);
    if (l < a.src_lens[b]) {
      const float dp = a.p[i] - a.p_tgt[i], de = a.e[i] - a.e_tgt[i];
      const float dd = a.logd[i] - logf(static_cast<float>(a.dur[i]) + 1.f);
      s2 += dp * dp; s3 += de * de; s4 += dd * dd;
    }
  }
  s0 = block_sum(s0); s1 = block_sum(s1); s2 = block_sum(s2); s3 = block_sum(s3); s4 = block_sum(s4);
  if (threadIdx.x == 0) {
    atomicAdd(ws + 0, s0); atomicAdd(ws + 1, s1); atomicAdd(ws + 2, s2); atomicAdd(ws + 3, s3); atomicAdd(ws + 4, s4);
  }
}
// out6 = (total, mel, postnet_mel, pitch, energy, duration); counts[0] = n_mel_elems, counts[1] = n_src
__global__ void loss_finalize_kernel(LossArgs a, const float* __restrict__ ws, float* __restrict__ out6, float* __restrict__ counts) {
  pdl_enter();
  long long nm = 0, ns = 0;
  for (int b = 0; b < a.B; ++b) {
    nm += min(static_cast<long long>(a.T), static_cast<long long>(a.mel_lens[b]));
    ns += min(static_cast<long long>(a.L), static_cast<long long>(a.src_lens[b]));
  }
  const float n_mel = static_cast<float>(nm * a.NM), n_src = static_cast<float>(ns);
  const float mel = ws[0] / n_mel, post = ws[1] / n_mel, pit = ws[2] / n_src, en = ws[3] / n_src, du = ws[4] / n_src;
  out6[0] = mel + post + du + pit + en;
  out6[1] = mel; out6[2] = post; out6[3] = pit; out6[4] = en; out6[5] = du;
  counts[0] = n_mel; counts[1] = n_src;
}
// backward (tangent == 0) or tangent-backward (tangent == 1; inputs are the prediction tangents)
__global__ void loss_bwd_kernel(LossArgs a, const float* __restrict__ counts, float scale, int tangent,
                                float* __restrict__ dmel, float* __restrict__ dpost, float* __restrict__ dp,
                                float* __restrict__ de, float* __restrict__ dlogd) {
  pdl_enter();
  const float n_mel = counts[0], n_src = counts[1];
  const long long nmel = static_cast<long long>(a.B) * a.T * a.NM;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < nmel; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long row = i / a.NM;
    const int b = static_cast<int>(row / a.T);
    const int t = static_cast<int>(row - static_cast<long long>(b) * a.T);
    float gm = 0.f, gp = 0.f;
    if (!tangent && t < a.mel_lens[b]) {
      const float tg = a.mel_tgt[i];
      const float dm = a.mel[i] - tg, dq = a.post[i] - tg;
      gm = (dm > 0.f ? 1.f : (dm < 0.f ? -1.f : 0.f)) * scale / n_mel;
      gp = (dq > 0.f ? 1.f : (dq < 0.f ? -1.f : 0.f)) * scale / n_mel;
    }
    dmel[i] = gm;     // L1 has zero curvature: tangent-backward is 0
    dpost[i] = gp;
  }
  const long long nsrc = static_cast<long long>(a.B) * a.L;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < nsrc; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int b = static_cast<int>(i / a.L);
    const int l = static_cast<int>(i - static_cast<long long>(b) * a.L);
    float gp = 0.f, ge = 0.f, gd = 0.f;
    if (l < a.src_lens[b]) {
      const float k = 2.f * scale / n_src;
      if (tangent) {
        gp = k * a.p[i]; ge = k * a.e[i]; gd = k * a.logd[i];
      } else {
        gp = k * (a.p[i] - a.p_tgt[i]);
        ge = k * (a.e[i] - a.e_tgt[i]);
        gd = k * (a.logd[i] - logf(static_cast<float>(a.dur[i]) + 1.f));
      }
    }
    dp[i] = gp; de[i] = ge; dlogd[i] = gd;
  }
}

// ================================================================================================
// flat-arena elementwise kernels (n % 4 == 0, 16B-aligned)
// ================================================================================================
__global__ void split_kernel(const float* __restrict__ src, bf16* __restrict__ hi, bf16* __restrict__ lo, long long n4) {
  pdl_enter();
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += static_cast<long long>(gridDim.x) * blockDim.x)
    store_split4(hi, lo, i * 4, *reinterpret_cast<const float4*>(src + i * 4));
}
// theta_out = theta_in - lr * g ; (hi, lo) = split(theta_out)      (l2l maml_update + operand prep, fused)
__global__ void sgd_split_kernel(const float* __restrict__ th, const float* __restrict__ g, float lr, float* __restrict__ out,
                                 bf16* __restrict__ hi, bf16* __restrict__ lo, long long n4) {
  pdl_enter();
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 a = *reinterpret_cast<const float4*>(th + i * 4);
    const float4 b = *reinterpret_cast<const float4*>(g + i * 4);
    const float4 v = make_float4(a.x - lr * b.x, a.y - lr * b.y, a.z - lr * b.z, a.w - lr * b.w);
    *reinterpret_cast<float4*>(out + i * 4) = v;
    if (hi) store_split4(hi, lo, i * 4, v);
  }
}
// y = a*x + b*y
__global__ void axpby_kernel(float a, const float* __restrict__ x, float b, float* __restrict__ y, long long n4) {
  pdl_enter();
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 u = *reinterpret_cast<const float4*>(x + i * 4);
    float4 v = *reinterpret_cast<float4*>(y + i * 4);
    v = make_float4(a * u.x + b * v.x, a * u.y + b * v.y, a * u.z + b * v.z, a * u.w + b * v.w);
    *reinterpret_cast<float4*>(y + i * 4) = v;
  }
}
__global__ void sumsq_kernel(const float* __restrict__ x, long long n4, float* __restrict__ out) {
  pdl_enter();
  float s = 0.f;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 u = *reinterpret_cast<const float4*>(x + i * 4);
    s += (u.x * u.x + u.y * u.y) + (u.z * u.z + u.w * u.w);
  }
  s = block_sum(s);
  if (threadIdx.x == 0) out[1 + blockIdx.x] = s;       // per-CTA partial; scalar_finalize_kernel adds them in CTA order
}
__global__ void dot_kernel(const float* __restrict__ x, const float* __restrict__ y, long long n4, float* __restrict__ out) {
  pdl_enter();
  float s = 0.f;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 u = *reinterpret_cast<const float4*>(x + i * 4);
    const float4 w = *reinterpret_cast<const float4*>(y + i * 4);
    s += (u.x * w.x + u.y * w.y) + (u.z * w.z + u.w * w.w);
  }
  s = block_sum(s);
  if (threadIdx.x == 0) out[1 + blockIdx.x] = s;
}
// out[0] = sum of the n per-CTA partials out[1 .. n] in a FIXED order (thread t takes partials t, t + 256, ...; then the fixed
// shuffle tree): the global gradient norm — and with it the clip coefficient and the Adam update — is bit-identical on every
// rank that holds the same reduced gradient, as torch's deterministic sum kernels make it in the reference (main.py:61).
__global__ void scalar_finalize_kernel(float* __restrict__ out, int n) {
  pdl_enter();
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += out[1 + i];
  s = block_sum(s);
  if (threadIdx.x == 0) out[0] = s;
}
// clip-by-global-norm (torch clip_grad_norm_: coef = min(1, max_norm/(norm+1e-6))) + Adam (no weight decay)
// hyper[0] = lr, hyper[1] = bias_correction1 = 1-beta1^t, hyper[2] = bias_correction2 = 1-beta2^t  (device, so graphs replay)
__global__ void adam_clip_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                 float* __restrict__ v, const float* __restrict__ sumsq, float gscale, float max_norm,
                                 const float* __restrict__ hyper, float beta1, float beta2, float eps,
                                 bf16* __restrict__ hi, bf16* __restrict__ lo, long long n4) {
  pdl_enter();
  const float norm = sqrtf(sumsq[0]) * gscale;
  float coef = max_norm > 0.f ? max_norm / (norm + 1e-6f) : 1.f;
  coef = fminf(coef, 1.f) * gscale;
  const float lr = hyper[0], bc1 = hyper[1], bc2 = hyper[2];
  const float step = lr / bc1;
  const float inv_sqrt_bc2 = rsqrtf(bc2);
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float4 P = *reinterpret_cast<float4*>(p + i * 4);
    const float4 G0 = *reinterpret_cast<const float4*>(g + i * 4);
    float4 M = *reinterpret_cast<float4*>(m + i * 4);
    float4 V = *reinterpret_cast<float4*>(v + i * 4);
    float pp[4] = {P.x, P.y, P.z, P.w};
    const float gg[4] = {G0.x * coef, G0.y * coef, G0.z * coef, G0.w * coef};
    float mm[4] = {M.x, M.y, M.z, M.w};
    float vv[4] = {V.x, V.y, V.z, V.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      mm[j] = beta1 * mm[j] + (1.f - beta1) * gg[j];
      vv[j] = beta2 * vv[j] + (1.f - beta2) * gg[j] * gg[j];
      const float denom = sqrtf(vv[j]) * inv_sqrt_bc2 + eps;
      pp[j] -= step * mm[j] / denom;
    }
    P = make_float4(pp[0], pp[1], pp[2], pp[3]);
    *reinterpret_cast<float4*>(p + i * 4) = P;
    *reinterpret_cast<float4*>(m + i * 4) = make_float4(mm[0], mm[1], mm[2], mm[3]);
    *reinterpret_cast<float4*>(v + i * 4) = make_float4(vv[0], vv[1], vv[2], vv[3]);
    if (hi) store_split4(hi, lo, i * 4, P);
  }
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" int mtts_embed_fwd(const int64_t* idx, const float* table, const float* base, const float* pos, int T, int64_t R,
                              int C, float* out, void* hi, void* lo, mtts_stream stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  MTTS_REQUIRE(idx && table && R > 0 && C > 0 && (C % 4) == 0 && (!pos || T > 0), "embed_fwd: bad args");
  MTTS_CHECK_CUDA(mtts_launch(embed_fwd_kernel, dim3(ew_grid(R * (C / 4), 1)), dim3(EW_THREADS), 0, s, idx, table, base, pos, T, R, C, out, static_cast<bf16*>(hi),
                                                                 static_cast<bf16*>(lo)));
  MTTS_CHECK_LAUNCH();
  return MTTS_OK;
}
extern "C" int mtts_embed_bwd(const int64_t* idx, const float* dy, int64_t R, int C, int64_t skip_idx, float scale,
                              float* dtable, mtts_stream stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  MTTS_REQUIRE(idx && dy && dtable && R > 0 && C > 0, "embed_bwd: bad args");
  if (mtts_deterministic())
    MTTS_CHECK_CUDA(mtts_launch(embed_bwd_ordered_kernel, dim3(mtts_cdiv(C, 128)), dim3(128), 0, s, idx, dy, R, C, skip_idx, scale, dtable));
  else
    MTTS_CHECK_CUDA(mtts_launch(embed_bwd_kernel, dim3(ew_grid(R * C, 1)), dim3(EW_THREADS), 0, s, idx, dy, R, C, skip_idx, scale, dtable));
  MTTS_CHECK_LAUNCH();
  return MTTS_OK;
}
extern "C" int mtts_bucketize(const float* v, const float* bins, int nb, int64_t R, int64_t* out, mtts_stream stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  MTTS_REQUIRE(v && bins && out && nb > 0 && R > 0, "bucketize: bad args");
  MTTS_CHECK_CUDA(mtts_launch(bucketize_kernel, dim3(static_cast<unsigned>(mtts_cdiv64(R, 256))), dim3(256), 0, s, v, bins, nb, R, out));
  MTTS_CHECK_LAUNCH();
  return MTTS_OK;
}
extern "C" int mtts_add_rowvec(const float* x, const float* vec, int64_t vec_bstride, const float* pos, int B, int T, int C,
                               float* out, void* hi, void* lo, mtts_stream stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  MTTS_REQUIRE(x && B > 0 && T > 0 && C > 0 && (C % 4) == 0, "add_rowvec: bad args");
  const long long R = static_cast<long long>(B) * T;
  MTTS_CHECK_CUDA(mtts_launch(add_rowvec_kernel, dim3(ew_grid(R * (C / 4), 1)), dim3(EW_THREADS), 0, s, x, vec, vec_bstride, pos, T, R, C, out, static_cast<bf16*>(hi),
                                                                  static_cast<bf16*>(lo)));
  MTTS_CHECK_LAUNCH();
  return MTTS_OK;
}
extern "C" int mtts_spk_embed(const int64_t* ids, const float* table, int n, int C, int average, int n_out, float* out,
                              mtts_stream stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  MTTS_REQUIRE(ids && table && out && n > 0 && n_out > 0 && (average || n_out == n), "spk_embed: bad args");
  MTTS_CHECK_CUDA(mtts_launch(spk_embed_kernel, dim3(n_out), dim3(256), 0, s, ids, table, n, C, average, n_out, out));
  MTTS_CHECK_LAUNCH();
  return MTTS_OK;
}
extern "C" int mtts_spk_embed_bwd(const int64_t* ids, const float* dspk, int n, int C, int average, int n_out, float scale,
                                  float* dtable, mtts_stream stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  MTTS_REQUIRE(ids && dspk && dtable && n > 0 && n_out > 0 && (average || n_out == n), "spk_embed_bwd: bad args");
  MTTS_CHECK_CUDA(mtts_launch(spk_embed_bwd_kernel, dim3(1), dim3(256), 0, s, ids, dspk, n, C, average, n_out, scale, dtable));
  MTTS_CHECK_LAUNCH();
  return MTTS_OK;
}
extern "C" int mtts_colsum(const float* f32, const void* hi, const void* lo, int nb, int64_t R, int C, float* out,
                           mtts_stream stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  MTTS_REQUIRE((f32 || hi) && out && nb > 0 && R > 0 && C > 0, "colsum: bad args");
  const int col_blocks = mtts_cdiv(C, 128);
  int row_chunks = mtts_deterministic() ? 1 : static_cast<int>(mtts_cdiv64(148LL * 4, static_cast<long long>(col_blocks) * nb));
  if (row_chunks < 1) row_chunks = 1;
  if (row_chunks > R) row_chunks = static_cast<int>(R);
  const int rows_per_cta = static_cast<int>(mtts_cdiv64(R, row_chunks));
  row_chunks = static_cast<int>(mtts_cdiv64(R, rows_per_cta));
  MTTS_CHECK_CUDA(mtts_launch(colsum_kernel, dim3(dim3(col_blocks, row_chunks, nb)), dim3(128), 0, s, f32, static_cast<const bf16*>(hi), static_cast<const bf16*>(lo), R,
                                                                  C, rows_per_cta, out));
  MTTS_CHECK_LAUNCH();
  return MTTS_OK;
}

extern "C" int mtts_bn_fwd(const float* x, const float* gamma, const float* beta, int64_t R, int C, float eps, float momentum,
                           int tanh_flag, float* running_mean, float* running_var, float* ws /* [2C] */,
                           float* stats /* [2C] */, float* out, void* hi, void* lo, uint32_t drop_thr, uint32_t drop_seed, float drop_scale, const uint32_t* drop_salt, mtts_stream stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  MTTS_REQUIRE(x && gamma && beta && ws && stats && R > 0 && C > 0, "bn_fwd: bad args");
  dim3 grid, block;
  int rpc;
  bn_launch_dims(R, C, grid, block, rpc);
  BnArgs a{};
  a.x = x; a.R = R; a.C = C; a.tanh_flag = tanh_flag; a.rows_per_cta = rpc; a.ws_in = ws;
  a.drop = DropSite{drop_thr, drop_seed, drop_scale, drop_salt};
  MTTS_CHECK_CUDA(cudaMemsetAsync(ws, 0, sizeof(float) * 2 * C, s));
  MTTS_CHECK_CUDA(mtts_launch(bn_stats_kernel, dim3(grid), dim3(block), 0, s, x, R, C, rpc, ws));
  dim3 agrid;
  int arpc;
  bn_apply_dims(R, C, agrid, arpc);
  MTTS_CHECK_CUDA(mtts_launch(bn_fwd_apply_kernel, dim3(agrid), dim3(block), 0, s, x, ws, gamma, beta, R, C, eps, momentum, tanh_flag, running_mean, running_var, stats,
                                             out, static_cast<bf16*>(hi), static_cast<bf16*>(lo), arpc, a.drop));
  MTTS_CHECK_LAUNCH();
  return MTTS_OK;
}
extern "C" int mtts_bn_bwd(const float* dout, const float* o, const float* x, const float* stats, const float* gamma, int64_t R,
                           int C, int tanh_flag, float* ws /* [2C] */, float* dx, void* hi, void* lo, float* dgamma,
                           float* dbeta, uint32_t drop_thr, uint32_t drop_seed, float drop_scale, const uint32_t* drop_salt, mtts_stream stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  MTTS_REQUIRE(dout && x && stats && gamma && ws && (!tanh_flag || o), "bn_bwd: bad args");
  dim3 grid, block;
  int rpc;
  bn_launch_dims(R, C, grid, block, rpc);
  BnArgs a{};
  a.x = x; a.dout = dout; a.o = o; a.stats = stats; a.R = R; a.C = C; a.tanh_flag = tanh_flag; a.rows_per_cta = rpc;
  a.drop = DropSite{drop_thr, drop_seed, drop_scale, drop_salt};
  MTTS_CHECK_CUDA(cudaMemsetAsync(ws, 0, sizeof(float) * 2 * C, s));
  MTTS_CHECK_CUDA(mtts_launch(bn_reduce_kernel<2>, dim3(grid), dim3(block), 0, s, a, ws));
  dim3 agrid;
  bn_apply_dims(R, C, agrid, a.rows_per_cta);
  MTTS_CHECK_CUDA(mtts_launch(bn_bwd_apply_kernel, dim3(agrid), dim3(block), 0, s, a, ws, gamma, dx, static_cast<bf16*>(hi), static_cast<bf16*>(lo), dgamma, dbeta));
  MTTS_CHECK_LAUNCH();
  return MTTS_OK;
}
extern "C" int mtts_bn_tfwd(const float* xdot, const float* x, const float* stats, const float* gamma, const float* gdot,
                            const float* bdot, const float* o, int64_t R, int C, int tanh_flag, float* ws /* [2C] */,
                            float* tsums /* [2C] */, float* odot, void* hi, void* lo, uint32_t drop_thr, uint32_t drop_seed, float drop_scale, const uint32_t* drop_salt, mtts_stream stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  MTTS_REQUIRE(xdot && x && stats && gamma && ws && tsums && (!tanh_flag || o), "bn_tfwd: bad args");
  dim3 grid, block;
  int rpc;
  bn_launch_dims(R, C, grid, block, rpc);
  BnArgs a{};
  a.x = x; a.xdot = xdot; a.o = o; a.stats = stats; a.R = R; a.C = C; a.tanh_flag = tanh_flag; a.rows_per_cta = rpc;
  a.drop = DropSite{drop_thr, drop_seed, drop_scale, drop_salt};
  MTTS_CHECK_CUDA(cudaMemsetAsync(ws, 0, sizeof(float) * 2 * C, s));
  MTTS_CHECK_CUDA(mtts_launch(bn_reduce_kernel<3>, dim3(grid), dim3(block), 0, s, a, ws));
  dim3 agrid;
  bn_apply_dims(R, C, agrid, a.rows_per_cta);
  MTTS_CHECK_CUDA(mtts_launch(bn_tfwd_apply_kernel, dim3(agrid), dim3(block), 0, s, a, ws, gamma, gdot, bdot, tsums, odot, static_cast<bf16*>(hi), static_cast<bf16*>(lo)));
  MTTS_CHECK_LAUNCH();
  return MTTS_OK;
}
extern "C" int mtts_bn_tbwd(const float* dout, const float* ddout, const float* o, const float* odot, const float* x,
                            const float* xdot, const float* stats, const float* tsums, const float* gamma, const float* gdot,
                            int64_t R, int C, int tanh_flag, float* ws /* [4C] */, float* ddx, void* hi, void* lo,
                            float* ddgamma, float* ddbeta, uint32_t drop_thr, uint32_t drop_seed, float drop_scale, const uint32_t* drop_salt, mtts_stream stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  MTTS_REQUIRE(dout && ddout && x && xdot && stats && tsums && gamma && ws && (!tanh_flag || (o && odot)), "bn_tbwd: bad args");
  dim3 grid, block;
  int rpc;
  bn_launch_dims(R, C, grid, block, rpc);
  BnArgs a{};
  a.x = x; a.xdot = xdot; a.dout = dout; a.ddout = ddout; a.o = o; a.odot = odot; a.stats = stats; a.tsums = tsums;
  a.R = R; a.C = C; a.tanh_flag = tanh_flag; a.rows_per_cta = rpc;
  a.drop = DropSite{drop_thr, drop_seed, drop_scale, drop_salt};
  MTTS_CHECK_CUDA(cudaMemsetAsync(ws, 0, sizeof(float) * 4 * C, s));
  MTTS_CHECK_CUDA(mtts_launch(bn_reduce_kernel<4>, dim3(grid), dim3(block), 0, s, a, ws));
  dim3 agrid;
  bn_apply_dims(R, C, agrid, a.rows_per_cta);
  MTTS_CHECK_CUDA(mtts_launch(bn_tbwd_apply_kernel, dim3(agrid), dim3(block), 0, s, a, ws, gamma, gdot, ddx, static_cast<bf16*>(hi), static_cast<bf16*>(lo), ddgamma, ddbeta));
  MTTS_CHECK_LAUNCH();
  return MTTS_OK;
}

static LossArgs make_loss_args(const float* mel, const float* post, const float* mel_tgt, const int64_t* mel_lens,
                               const float* p, const float* p_tgt, const float* e, const float* e_tgt, const float* logd,
                               const int64_t* dur, const int64_t* src_lens, int B, int T, int L, int NM) {
  LossArgs a;
  a.mel = mel; a.post = post; a.mel_tgt = mel_tgt; a.mel_lens = mel_lens; a.p = p; a.p_tgt = p_tgt; a.e = e; a.e_tgt = e_tgt;
  a.logd = logd; a.dur = dur; a.src_lens = src_lens; a.B = B; a.T = T; a.L = L; a.NM = NM;
  return a;
}
extern "C" int mtts_loss_fwd(const float* mel, const float* post, const float* mel_tgt, const int64_t* mel_lens, const float* p,
                             const float* p_tgt, const float* e, const float* e_tgt, const float* logd, const int64_t* dur,
                             const int64_t* src_lens, int B, int T, int L, int NM, float* ws /* [8] */, float* out6,
                             float* counts /* [2] */, mtts_stream stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  MTTS_REQUIRE(mel && post && mel_tgt && mel_lens && p && p_tgt && e && e_tgt && logd && dur && src_lens && ws && out6 && counts,
               "loss_fwd: null argument");
  LossArgs a = make_loss_args(mel, post, mel_tgt, mel_lens, p, p_tgt, e, e_tgt, logd, dur, src_lens, B, T, L, NM);
  MTTS_CHECK_CUDA(cudaMemsetAsync(ws, 0, sizeof(float) * 8, s));
  MTTS_CHECK_CUDA(mtts_launch(loss_sums_kernel, dim3(mtts_deterministic() ? 1 : ew_grid(static_cast<long long>(B) * T * NM, 4)), dim3(EW_THREADS), 0, s, a, ws));
  MTTS_CHECK_CUDA(mtts_launch(loss_finalize_kernel, dim3(1), dim3(1), 0, s, a, ws, out6, counts));
  MTTS_CHECK_LAUNCH();
  return MTTS_OK;
}
extern "C" int mtts_loss_bwd(const float* mel, const float* post, const float* mel_tgt, const int64_t* mel_lens, const float* p,
                             const float* p_tgt, const float* e, const float* e_tgt, const float* logd, const int64_t* dur,
                             const int64_t* src_lens, int B, int T, int L, int NM, const float* counts, float scale, int tangent,
                             float* dmel, float* dpost, float* dp, float* de, float* dlogd, mtts_stream stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  MTTS_REQUIRE(mel_lens && src_lens && p && e && logd && counts && dmel && dpost && dp && de && dlogd, "loss_bwd: null argument");
  MTTS_REQUIRE(tangent || (mel && post && mel_tgt && p_tgt && e_tgt && dur), "loss_bwd: null argument");
  LossArgs a = make_loss_args(mel, post, mel_tgt, mel_lens, p, p_tgt, e, e_tgt, logd, dur, src_lens, B, T, L, NM);
  MTTS_CHECK_CUDA(mtts_launch(loss_bwd_kernel, dim3(ew_grid(static_cast<long long>(B) * T * NM, 4)), dim3(EW_THREADS), 0, s, a, counts, scale, tangent, dmel, dpost, dp, de,
                                                                                      dlogd));
  MTTS_CHECK_LAUNCH();
  return MTTS_OK;
}

#define REQ_N4(n, p0) MTTS_REQUIRE((n) > 0 && ((n) % 4) == 0 && (reinterpret_cast<uintptr_t>(p0) & 15) == 0, "elementwise: n %% 4 != 0 or misaligned")

extern "C" int mtts_split(const float* src, void* hi, void* lo, int64_t n, mtts_stream stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  REQ_N4(n, src);
  MTTS_CHECK_CUDA(mtts_launch(split_kernel, dim3(ew_grid(n / 4, 2)), dim3(EW_THREADS), 0, s, src, static_cast<bf16*>(hi), static_cast<bf16*>(lo), n / 4));
  MTTS_CHECK_LAUNCH();
  return MTTS_OK;
}
extern "C" int mtts_sgd_split(const float* theta, const float* g, float lr, float* out, void* hi, void* lo, int64_t n,
                              mtts_stream stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  REQ_N4(n, theta);
  MTTS_CHECK_CUDA(mtts_launch(sgd_split_kernel, dim3(ew_grid(n / 4, 2)), dim3(EW_THREADS), 0, s, theta, g, lr, out, static_cast<bf16*>(hi), static_cast<bf16*>(lo), n / 4));
  MTTS_CHECK_LAUNCH();
  return MTTS_OK;
}
extern "C" int mtts_axpby(float a, const float* x, float b, float* y, int64_t n, mtts_stream stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  REQ_N4(n, x);
  MTTS_CHECK_CUDA(mtts_launch(axpby_kernel, dim3(ew_grid(n / 4, 2)), dim3(EW_THREADS), 0, s, a, x, b, y, n / 4));
  MTTS_CHECK_LAUNCH();
  return MTTS_OK;
}
// out: MTTS_SCALAR_WS floats — out[0] the result, out[1..] per-CTA partials (fixed-order two-stage reduction, no atomics)
extern "C" int mtts_sumsq(const float* x, int64_t n, float* out, mtts_stream stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  REQ_N4(n, x);
  int grid = ew_grid(n / 4, 4);
  if (grid > MTTS_SCALAR_WS - 1) grid = MTTS_SCALAR_WS - 1;
  MTTS_CHECK_CUDA(mtts_launch(sumsq_kernel, dim3(grid), dim3(EW_THREADS), 0, s, x, n / 4, out));
  MTTS_CHECK_CUDA(mtts_launch(scalar_finalize_kernel, dim3(1), dim3(256), 0, s, out, grid));
  MTTS_CHECK_LAUNCH();
  return MTTS_OK;
}
extern "C" int mtts_dot(const float* x, const float* y, int64_t n, float* out, mtts_stream stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  REQ_N4(n, x);
  MTTS_REQUIRE(y != nullptr && (reinterpret_cast<uintptr_t>(y) & 15) == 0, "dot: bad second operand");
  int grid = ew_grid(n / 4, 4);
  if (grid > MTTS_SCALAR_WS - 1) grid = MTTS_SCALAR_WS - 1;
  MTTS_CHECK_CUDA(mtts_launch(dot_kernel, dim3(grid), dim3(EW_THREADS), 0, s, x, y, n / 4, out));
  MTTS_CHECK_CUDA(mtts_launch(scalar_finalize_kernel, dim3(1), dim3(256), 0, s, out, grid));
  MTTS_CHECK_LAUNCH();
  return MTTS_OK;
}
extern "C" int mtts_adam_clip(float* p, const float* g, float* m, float* v, const float* sumsq, float gscale, float max_norm,
                              const float* hyper, float beta1, float beta2, float eps, void* hi, void* lo, int64_t n,
                              mtts_stream stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  REQ_N4(n, p);
  MTTS_CHECK_CUDA(mtts_launch(adam_clip_kernel, dim3(ew_grid(n / 4, 2)), dim3(EW_THREADS), 0, s, p, g, m, v, sumsq, gscale, max_norm, hyper, beta1, beta2, eps,
                                                           static_cast<bf16*>(hi), static_cast<bf16*>(lo), n / 4));
  MTTS_CHECK_LAUNCH();
  return MTTS_OK;
}
