// SNR mixer of the noise-robustness harness — get_noise_from_sound (recognizer_test.py:426-435) + the addition at :554:
//     RMS_s = sqrt(mean(signal^2));  RMS_n = sqrt(RMS_s^2 / 10^(SNR/10));  out = signal + noise * (RMS_n / RMS_n_current)
// for a batch of clips, quantised back to int16 (round half to even, saturating) so the result feeds K1 directly.
// One CTA per clip: a fixed-order float64 reduction of both sums of squares (deterministic), then the scaled add over the
// same samples (a 5 s clip is 1.3 MB: the second read hits L2).
#include "sia_common.cuh"

using namespace sia;

namespace {

constexpr int kMixThreads = 1024;

__global__ void __launch_bounds__(kMixThreads)
mix_noise_kernel(const int16_t *__restrict__ signal, int64_t signal_stride, const float *__restrict__ noise, int64_t noise_stride,
                 int64_t n, double snr_db, int16_t *__restrict__ out, int64_t out_stride, double *__restrict__ scale_out) {
  __shared__ double s_s[kMixThreads / 32], s_n[kMixThreads / 32];
  __shared__ double s_scale;
  const int16_t *sig = signal + (int64_t)blockIdx.x * signal_stride;
  const float *nz = noise + (int64_t)blockIdx.x * noise_stride;
  int16_t *dst = out + (int64_t)blockIdx.x * out_stride;
  double ss = 0, sn = 0;
  for (int64_t i = threadIdx.x; i < n; i += kMixThreads) {
    const double a = (double)sig[i], b = (double)nz[i];
    ss += a * a; sn += b * b;
  }
#pragma unroll
  for (int d = 16; d; d >>= 1) { ss += __shfl_xor_sync(0xffffffffu, ss, d); sn += __shfl_xor_sync(0xffffffffu, sn, d); }
  if ((threadIdx.x & 31) == 0) { s_s[threadIdx.x >> 5] = ss; s_n[threadIdx.x >> 5] = sn; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, b = 0;
    for (int w = 0; w < kMixThreads / 32; ++w) { a += s_s[w]; b += s_n[w]; }
    const double rms_s = sqrt(a / (double)n);
    const double rms_n = sqrt(rms_s * rms_s / pow(10.0, snr_db / 10.0));
    const double rms_cur = sqrt(b / (double)n);
    s_scale = rms_cur > 0 ? rms_n / rms_cur : 0.0;
    if (scale_out) scale_out[blockIdx.x] = s_scale;
  }
  __syncthreads();
  const double scale = s_scale;
  for (int64_t i = threadIdx.x; i < n; i += kMixThreads) {
    double v = rint((double)sig[i] + (double)nz[i] * scale);
    v = fmin(32767.0, fmax(-32768.0, v));
    dst[i] = (int16_t)v;
  }
}

}  // namespace

extern "C" int sia_mix_noise(int device, const int16_t *d_signal, int64_t signal_stride, const float *d_noise,
                             int64_t noise_stride, int32_t n_clips, int64_t n_samples, double snr_db, int16_t *d_out,
                             int64_t out_stride, double *d_scale_out, void *stream) {
  SIA_REQUIRE(n_clips >= 0 && n_samples > 0, SIA_E_INVALID, "mix_noise: bad sizes");
  if (n_clips == 0) return SIA_OK;
  SIA_REQUIRE(d_signal && d_noise && d_out, SIA_E_INVALID, "NULL argument");
  SIA_REQUIRE(signal_stride >= n_samples && noise_stride >= n_samples && out_stride >= n_samples, SIA_E_INVALID,
              "mix_noise: strides must cover n_samples");
  SIA_CUDA(cudaSetDevice(device));
  mix_noise_kernel<<<(unsigned)n_clips, kMixThreads, 0, (cudaStream_t)stream>>>(d_signal, signal_stride, d_noise, noise_stride,
                                                                              n_samples, snr_db, d_out, out_stride, d_scale_out);
  SIA_CHECK_LAUNCH();
  return SIA_OK;
}
