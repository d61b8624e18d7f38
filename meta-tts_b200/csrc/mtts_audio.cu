// mtts_audio.cu — free-running synthesis helpers and the STFT / iSTFT / Griffin-Lim elementwise kernels (sm_100a).
//
// The dense parts of the vocoder-side decode (the [1026 x 1024] Fourier-basis products of audio/stft.py:67-72,90-94 and
// the mel-basis products of audio/stft.py:173, audio/tools.py:24) go through mtts_gemm: the strided conv1d of
// STFT.transform is a 4-tap conv over the hop-reshaped signal [rows, 256] and the conv_transpose1d of STFT.inverse is the
// matching 4-tap "dgrad" form, so framing and overlap-add both happen in the TMEM accumulator (nothing is materialised).
// Everything here is HBM-bound row / elementwise work around those GEMMs.
#include "mtts_common.cuh"

namespace {

constexpr int TPB = 256;

// ---- duration rounding (lightning/model/modules.py:133-137) -------------------------------------------------------
// d = clamp(round(exp(log_d) - 1) * d_control, min=0);  torch.round is round-half-to-even = rintf.
__global__ void duration_round_kernel(const float* __restrict__ logd, float d_control, int64_t n, float* __restrict__ out) {
  pdl_enter();
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = rintf(expf(logd[i]) - 1.0f) * d_control;
  out[i] = fmaxf(v, 0.0f);
}

// ---- BatchNorm1d in eval mode (+ tanh): running statistics, no update (Layers.py:129-137 under model.eval()) ---------
__global__ void bn_eval_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                               const float* __restrict__ rmean, const float* __restrict__ rvar, int64_t R, int C, float eps,
                               int tanh_flag, float* __restrict__ out, bf16* __restrict__ hi, bf16* __restrict__ lo) {
  pdl_enter();
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= R * C) return;
  int c = (int)(i % C);
  float a = gamma[c] * (1.0f / sqrtf(rvar[c] + eps));
  float b = beta[c] - rmean[c] * a;
  float v = fmaf(x[i], a, b);
  if (tanh_flag) v = tanhf(v);
  if (out) out[i] = v;
  if (hi) {
    bf16 h, l;
    split_bf16(v, h, l);
    hi[i] = h;
    if (lo) lo[i] = l;
  }
}

// ---- generic scale / elementwise maps used by the mel side ----------------------------------------------------------
// op 0: out = log(max(x, a) * b)      dynamic_range_compression  audio/audio_processing.py:85-91
// op 1: out = exp(x) * a              dynamic_range_decompression audio/audio_processing.py:94-100 (a = scale / C)
// op 2: out = x * a                   p_control / e_control       modules.py:86,97
__global__ void unary_kernel(int op, const float* __restrict__ x, int64_t n, float a, float b, float* __restrict__ out,
                             bf16* __restrict__ hi, bf16* __restrict__ lo) {
  pdl_enter();
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = x[i];
  if (op == 0) v = logf(fmaxf(v, a) * b);
  else if (op == 1) v = expf(v) * a;
  else v = v * a;
  if (out) out[i] = v;
  if (hi) {
    bf16 h, l;
    split_bf16(v, h, l);
    hi[i] = h;
    if (lo) lo[i] = l;
  }
}

// ---- reflect pad + operand split (audio/stft.py:60-65) --------------------------------------------------------------
// out[b, i] = x[b, reflect(i - pad)] for i < N + 2*pad, 0 for N + 2*pad <= i < ld (the hop-reshaped tail).
__global__ void reflect_pad_kernel(const float* __restrict__ x, int64_t N, int pad, int64_t ld, float* __restrict__ out,
                                   bf16* __restrict__ hi, bf16* __restrict__ lo) {
  pdl_enter();
  int b = blockIdx.y;
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= ld) return;
  float v = 0.0f;
  if (i < N + 2 * (int64_t)pad) {
    int64_t j = i - pad;
    if (j < 0) j = -j;
    if (j >= N) j = 2 * (N - 1) - j;
    v = x[b * N + j];
  }
  int64_t o = b * ld + i;
  if (out) out[o] = v;
  bf16 h, l;
  split_bf16(v, h, l);
  hi[o] = h;
  if (lo) lo[o] = l;
}

// ---- magnitude / phase / frame energy of an STFT output row (audio/stft.py:74-81,175) --------------------------------
// ri[r, c] = real, ri[r, im_off + c] = imag (c < nb).  mag / phase rows have stride ldm >= nb (pad columns zeroed) so that
// they can feed mtts_gemm directly (mag_hi / mag_lo: the operand split for the mel-basis product).  One warp per frame row.
__global__ void stft_polar_kernel(const float* __restrict__ ri, int64_t R, int nb, int ld, int im_off, int ldm, float* __restrict__ mag,
                                  float* __restrict__ phase, float* __restrict__ energy, bf16* __restrict__ mag_hi,
                                  bf16* __restrict__ mag_lo) {
  pdl_enter();
  int lane = threadIdx.x & 31;
  int64_t r = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= R) return;
  const float* row = ri + r * ld;
  float ss = 0.0f;
  for (int c = lane; c < ldm; c += 32) {
    float m = 0.0f, ph = 0.0f;
    if (c < nb) {
      float re = row[c], im = row[im_off + c];
      float m2 = re * re + im * im;
      ss += m2;
      m = sqrtf(m2);
      ph = atan2f(im, re);
    }
    if (mag) mag[r * ldm + c] = m;
    if (phase) phase[r * ldm + c] = ph;
    if (mag_hi) {
      bf16 h, l;
      split_bf16(m, h, l);
      mag_hi[r * ldm + c] = h;
      if (mag_lo) mag_lo[r * ldm + c] = l;
    }
  }
  if (energy) {
#pragma unroll
    for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if (lane == 0) energy[r] = sqrtf(ss);          // torch.norm(magnitudes, dim=1)
  }
}

// ---- recombine magnitude and phase into the iSTFT operand (audio/stft.py:86-88) ---------------------------------------
// X[r, c] = mag*cos(ph), X[r, im_off + c] = mag*sin(ph); ph = phase[r, c] or atan2(ri imag, ri real) (Griffin-Lim keeps
// only the angles of the previous transform, audio/audio_processing.py:79-81).  Pad columns are written as zeros.
__global__ void stft_recombine_kernel(const float* __restrict__ mag, const float* __restrict__ phase, const float* __restrict__ ri,
                                      int64_t R, int nb, int ld, int im_off, int ldm, bf16* __restrict__ hi, bf16* __restrict__ lo) {
  pdl_enter();
  int64_t r = blockIdx.x;
  int c = blockIdx.y * blockDim.x + threadIdx.x;
  if (c >= ld) return;
  int k = c >= im_off ? c - im_off : c;
  float v = 0.0f;
  if (k < nb) {
    float ph = phase ? phase[r * ldm + k] : atan2f(ri[r * ld + im_off + k], ri[r * ld + k]);
    float m = mag[r * ldm + k];
    v = c >= im_off ? m * sinf(ph) : m * cosf(ph);
  }
  bf16 h, l;
  split_bf16(v, h, l);
  hi[r * ld + c] = h;
  if (lo) lo[r * ld + c] = l;
}

// ---- window-sum normalisation, hop-ratio scale and centre trim of the overlap-added signal (audio/stft.py:96-122) ------
__global__ void istft_finish_kernel(const float* __restrict__ ola, const float* __restrict__ wsum, float tiny, float scale, int64_t n,
                                    int trim, float* __restrict__ out) {
  pdl_enter();
  int b = blockIdx.y;
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;       // output sample
  int64_t no = n - 2 * (int64_t)trim;
  if (i >= no) return;
  int64_t s = i + trim;
  float v = ola[b * n + s];
  float w = wsum[s];
  if (w > tiny) v /= w;
  out[b * no + i] = v * scale;
}

}  // namespace

extern "C" int mtts_duration_round(const float* logd, float d_control, int64_t n, float* out, mtts_stream stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  MTTS_REQUIRE(logd && out && n > 0, "duration_round: bad args");
  MTTS_CHECK_CUDA(mtts_launch(duration_round_kernel, dim3((unsigned)mtts_cdiv64(n, TPB)), dim3(TPB), 0, s, logd, d_control, n, out));
  MTTS_CHECK_LAUNCH();
  return MTTS_OK;
}

extern "C" int mtts_bn_eval(const float* x, const float* gamma, const float* beta, const float* running_mean,
                            const float* running_var, int64_t R, int C, float eps, int tanh_flag, float* out, void* hi, void* lo,
                            mtts_stream stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  MTTS_REQUIRE(x && gamma && beta && running_mean && running_var && R > 0 && C > 0 && (out || hi), "bn_eval: bad args");
  MTTS_CHECK_CUDA(mtts_launch(bn_eval_kernel, dim3((unsigned)mtts_cdiv64(R * C, TPB)), dim3(TPB), 0, s, x, gamma, beta, running_mean,
                              running_var, R, C, eps, tanh_flag, out, static_cast<bf16*>(hi), static_cast<bf16*>(lo)));
  MTTS_CHECK_LAUNCH();
  return MTTS_OK;
}

extern "C" int mtts_unary(int op, const float* x, int64_t n, float a, float b, float* out, void* hi, void* lo, mtts_stream stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  MTTS_REQUIRE(x && n > 0 && op >= 0 && op <= 2 && (out || hi), "unary: bad args");
  MTTS_CHECK_CUDA(mtts_launch(unary_kernel, dim3((unsigned)mtts_cdiv64(n, TPB)), dim3(TPB), 0, s, op, x, n, a, b, out,
                              static_cast<bf16*>(hi), static_cast<bf16*>(lo)));
  MTTS_CHECK_LAUNCH();
  return MTTS_OK;
}

extern "C" int mtts_reflect_pad(const float* x, int B, int64_t N, int pad, int64_t ld, float* out, void* hi, void* lo,
                                mtts_stream stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  MTTS_REQUIRE(x && hi && B > 0 && N > pad && pad >= 0 && ld > 0, "reflect_pad: bad args (needs N > pad)");
  MTTS_CHECK_CUDA(mtts_launch(reflect_pad_kernel, dim3((unsigned)mtts_cdiv64(ld, TPB), B), dim3(TPB), 0, s, x, N, pad, ld, out,
                              static_cast<bf16*>(hi), static_cast<bf16*>(lo)));
  MTTS_CHECK_LAUNCH();
  return MTTS_OK;
}

extern "C" int mtts_stft_polar(const float* ri, int64_t R, int nb, int ld, int im_off, int ldm, float* mag, float* phase, float* energy,
                               void* mag_hi, void* mag_lo, mtts_stream stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  MTTS_REQUIRE(ri && R > 0 && nb > 0 && im_off >= nb && ld >= im_off + nb && ldm >= nb && (mag || phase || energy || mag_hi),
               "stft_polar: bad args");
  MTTS_CHECK_CUDA(mtts_launch(stft_polar_kernel, dim3((unsigned)mtts_cdiv64(R, TPB / 32)), dim3(TPB), 0, s, ri, R, nb, ld, im_off, ldm, mag,
                              phase, energy, static_cast<bf16*>(mag_hi), static_cast<bf16*>(mag_lo)));
  MTTS_CHECK_LAUNCH();
  return MTTS_OK;
}

extern "C" int mtts_stft_recombine(const float* mag, const float* phase, const float* ri, int64_t R, int nb, int ld, int im_off,
                                   int ldm, void* hi, void* lo, mtts_stream stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  MTTS_REQUIRE(mag && (phase || ri) && hi && R > 0 && R < (1ll << 31) && nb > 0 && im_off >= nb && ld >= im_off + nb && ldm >= nb,
               "stft_recombine: bad args");
  MTTS_CHECK_CUDA(mtts_launch(stft_recombine_kernel, dim3((unsigned)R, (unsigned)mtts_cdiv(ld, TPB)), dim3(TPB), 0, s, mag, phase, ri, R,
                              nb, ld, im_off, ldm, static_cast<bf16*>(hi), static_cast<bf16*>(lo)));
  MTTS_CHECK_LAUNCH();
  return MTTS_OK;
}

extern "C" int mtts_istft_finish(const float* ola, const float* wsum, float tiny, float scale, int B, int64_t n, int trim, float* out,
                                 mtts_stream stream_) {
  cudaStream_t s = static_cast<cudaStream_t>(stream_);
  MTTS_REQUIRE(ola && wsum && out && B > 0 && n > 2 * (int64_t)trim, "istft_finish: bad args");
  MTTS_CHECK_CUDA(mtts_launch(istft_finish_kernel, dim3((unsigned)mtts_cdiv64(n - 2 * trim, TPB), B), dim3(TPB), 0, s, ola, wsum, tiny,
                              scale, n, trim, out));
  MTTS_CHECK_LAUNCH();
  return MTTS_OK;
}
