// Device-wide exclusive scan (reduce / scan-of-block-sums / scan-and-add).  The inputs
// here are per-row and per-peak counts — megabytes next to the gigabytes the spectrogram
// kernels move — so the simple three-launch form is used instead of a chained scan.
#include "sia_common.cuh"

namespace sia {

namespace {
constexpr int kScanThreads = 256;
constexpr int kScanItems = 16;
constexpr int kScanTile = kScanThreads * kScanItems;  // 4096

__device__ __forceinline__ uint64_t warp_incl_scan(uint64_t v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint64_t o = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= d) v += o;
  }
  return v;
}

// block-wide exclusive scan of one value per thread; returns exclusive prefix, *total = block sum
__device__ __forceinline__ uint64_t block_excl_scan(uint64_t v, uint64_t *total) {
  __shared__ uint64_t warp_sums[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  uint64_t incl = warp_incl_scan(v, lane);
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    uint64_t w = lane < nwarps ? warp_sums[lane] : 0;
    uint64_t wi = warp_incl_scan(w, lane);
    warp_sums[lane] = wi - w;                 // exclusive prefix of warp sums
  }
  __syncthreads();
  uint64_t base = warp_sums[warp];
  if (total) {
    // total = prefix of last warp + its sum
    uint64_t t = 0;
    if (threadIdx.x == blockDim.x - 1) t = base + incl;
    __shared__ uint64_t tot;
    if (threadIdx.x == blockDim.x - 1) tot = t;
    __syncthreads();
    *total = tot;
  }
  return base + incl - v;
}

__global__ void __launch_bounds__(kScanThreads)
scan_reduce_kernel(const uint32_t *__restrict__ in, const int64_t *__restrict__ d_n, int64_t n_static,
                   uint64_t *__restrict__ block_sums) {
  const int64_t n = d_n ? min(*d_n, n_static) : n_static;   // a device-side count never exceeds the buffers
  const int64_t base = (int64_t)blockIdx.x * kScanTile;
  uint64_t s = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    int64_t i = base + k * kScanThreads + threadIdx.x;
    if (i < n) s += in[i];
  }
  uint64_t total;
  block_excl_scan(s, &total);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024)
scan_block_sums_kernel(uint64_t *__restrict__ block_sums, int64_t nblocks_max, const int64_t *__restrict__ d_n,
                       int64_t n_static, int64_t *__restrict__ out) {
  const int64_t n = d_n ? min(*d_n, n_static) : n_static;   // a device-side count never exceeds the buffers
  int64_t nblocks = (n + kScanTile - 1) / kScanTile;
  if (nblocks > nblocks_max) nblocks = nblocks_max;
  uint64_t carry = 0;
  for (int64_t base = 0; base < nblocks; base += blockDim.x) {
    int64_t i = base + threadIdx.x;
    uint64_t v = i < nblocks ? block_sums[i] : 0;
    uint64_t total;
    uint64_t ex = block_excl_scan(v, &total);
    if (i < nblocks) block_sums[i] = carry + ex;
    carry += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) out[n] = (int64_t)carry;
}

__global__ void __launch_bounds__(kScanThreads)
scan_tiles_kernel(const uint32_t *__restrict__ in, const int64_t *__restrict__ d_n, int64_t n_static,
                  const uint64_t *__restrict__ block_sums, int64_t *__restrict__ out) {
  __shared__ uint32_t tile[kScanTile];
  const int64_t n = d_n ? min(*d_n, n_static) : n_static;   // a device-side count never exceeds the buffers
  const int64_t base = (int64_t)blockIdx.x * kScanTile;
  if (base >= n) return;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    int j = k * kScanThreads + threadIdx.x;
    int64_t i = base + j;
    tile[j] = i < n ? in[i] : 0u;
  }
  __syncthreads();
  uint32_t v[kScanItems];
  uint64_t s = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    v[k] = tile[threadIdx.x * kScanItems + k];
    s += v[k];
  }
  uint64_t ex = block_excl_scan(s, nullptr) + block_sums[blockIdx.x];
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    int64_t i = base + threadIdx.x * kScanItems + k;
    if (i < n) out[i] = (int64_t)ex;
    ex += v[k];
  }
}
}  // namespace

size_t scan_tmp_bytes(int64_t n) { return (size_t)(ceil_div(n > 0 ? n : 1, kScanTile) + 1) * sizeof(uint64_t); }

static int scan_impl(const uint32_t *d_in, int64_t *d_out, const int64_t *d_n, int64_t n_max, void *d_tmp,
                     cudaStream_t s) {
  if (n_max < 0) { set_error("scan: negative length"); return SIA_E_INVALID; }
  const int64_t nblocks = ceil_div(n_max > 0 ? n_max : 1, kScanTile);
  uint64_t *block_sums = (uint64_t *)d_tmp;
  if (n_max > 0) {
    scan_reduce_kernel<<<(unsigned)nblocks, kScanThreads, 0, s>>>(d_in, d_n, n_max, block_sums);
    SIA_CHECK_LAUNCH();
  }
  scan_block_sums_kernel<<<1, 1024, 0, s>>>(block_sums, n_max > 0 ? nblocks : 0, d_n, n_max, d_out);
  SIA_CHECK_LAUNCH();
  if (n_max > 0) {
    scan_tiles_kernel<<<(unsigned)nblocks, kScanThreads, 0, s>>>(d_in, d_n, n_max, block_sums, d_out);
    SIA_CHECK_LAUNCH();
  }
  return SIA_OK;
}

int exclusive_scan_u32(const uint32_t *d_in, int64_t *d_out, int64_t n, void *d_tmp, cudaStream_t s) {
  return scan_impl(d_in, d_out, nullptr, n, d_tmp, s);
}

int exclusive_scan_u32_dyn(const uint32_t *d_in, int64_t *d_out, const int64_t *d_n, int64_t n_max, void *d_tmp,
                           cudaStream_t s) {
  return scan_impl(d_in, d_out, d_n, n_max, d_tmp, s);
}

}  // namespace sia
