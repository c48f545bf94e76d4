#include "sort.cuh"

namespace sia {

namespace {

constexpr int kSortThreads = 512;
constexpr int kSortItems = 8;
constexpr int kSortTile = kSortThreads * kSortItems;   // 4096 records per block
constexpr int kWarps = kSortThreads / 32;

template <typename Rec> struct RecOps;
template <> struct RecOps<uint64_t> {
  __device__ static __forceinline__ uint32_t digit(const uint64_t &r, int byte) { return (uint32_t)(r >> (8 * byte)) & 0xffu; }
  __device__ static __forceinline__ void fold(const uint64_t &r, uint64_t *o, uint64_t *a) { o[0] |= r; a[0] &= r; }
};
template <> struct RecOps<ulonglong2> {
  __device__ static __forceinline__ uint32_t digit(const ulonglong2 &r, int byte) {
    return byte < 8 ? (uint32_t)(r.x >> (8 * byte)) & 0xffu : (uint32_t)(r.y >> (8 * (byte - 8))) & 0xffu;
  }
  __device__ static __forceinline__ void fold(const ulonglong2 &r, uint64_t *o, uint64_t *a) {
    o[0] |= r.x; a[0] &= r.x; o[1] |= r.y; a[1] &= r.y;
  }
};

// OR and AND of all records -> which bytes vary at all
template <typename Rec>
__global__ void __launch_bounds__(256) bitfold_kernel(const Rec *__restrict__ in, int64_t n, unsigned long long *out4) {
  uint64_t o[2] = {0, 0}, a[2] = {~0ull, ~0ull};
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    RecOps<Rec>::fold(in[i], o, a);
#pragma unroll
  for (int d = 16; d; d >>= 1) {
    o[0] |= __shfl_xor_sync(0xffffffffu, o[0], d); o[1] |= __shfl_xor_sync(0xffffffffu, o[1], d);
    a[0] &= __shfl_xor_sync(0xffffffffu, a[0], d); a[1] &= __shfl_xor_sync(0xffffffffu, a[1], d);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicOr(out4 + 0, o[0]); atomicOr(out4 + 1, o[1]);
    atomicAnd(out4 + 2, a[0]); atomicAnd(out4 + 3, a[1]);
  }
}

template <typename Rec>
__global__ void __launch_bounds__(kSortThreads)
hist_kernel(const Rec *__restrict__ in, int64_t n, int byte, uint32_t *__restrict__ hist, int64_t nblocks) {
  __shared__ uint32_t h[256];
  if (threadIdx.x < 256) h[threadIdx.x] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * kSortTile;
#pragma unroll
  for (int k = 0; k < kSortItems; ++k) {
    const int64_t i = base + k * kSortThreads + threadIdx.x;
    if (i < n) atomicAdd(&h[RecOps<Rec>::digit(in[i], byte)], 1u);
  }
  __syncthreads();
  if (threadIdx.x < 256) hist[(int64_t)threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];
}

template <typename Rec>
__global__ void __launch_bounds__(kSortThreads)
scatter_kernel(const Rec *__restrict__ in, Rec *__restrict__ out, int64_t n, int byte,
               const int64_t *__restrict__ offs, int64_t nblocks) {
  __shared__ uint32_t wcnt[kWarps][256];   // per-warp digit counters, then per-warp bases inside the block
  __shared__ int64_t gbase[256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < kWarps * 256; i += kSortThreads) (&wcnt[0][0])[i] = 0;
  if (threadIdx.x < 256) gbase[threadIdx.x] = offs[(int64_t)threadIdx.x * nblocks + blockIdx.x];
  __syncthreads();
  // each warp owns 8 consecutive batches of 32 consecutive records -> order inside the tile is
  // (warp, batch, lane), which the ranks below preserve (stable)
  const int64_t wbase = (int64_t)blockIdx.x * kSortTile + warp * (32 * kSortItems);
  Rec rec[kSortItems];
  uint32_t dig[kSortItems], rank[kSortItems];
  const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
  for (int k = 0; k < kSortItems; ++k) {
    const int64_t i = wbase + k * 32 + lane;
    const bool valid = i < n;
    const uint32_t active = __ballot_sync(0xffffffffu, valid);
    dig[k] = 0; rank[k] = 0;
    if (valid) {
      rec[k] = in[i];
      dig[k] = RecOps<Rec>::digit(rec[k], byte);
      const uint32_t peers = __match_any_sync(active, dig[k]);
      const uint32_t before = wcnt[warp][dig[k]];
      rank[k] = before + __popc(peers & lt);
      __syncwarp(active);
      if ((peers & lt) == 0) wcnt[warp][dig[k]] = before + __popc(peers);
      __syncwarp(active);
    }
  }
  __syncthreads();
  if (threadIdx.x < 256) {           // exclusive scan over the warps, per digit
    uint32_t run = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) { const uint32_t c = wcnt[w][threadIdx.x]; wcnt[w][threadIdx.x] = run; run += c; }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kSortItems; ++k) {
    const int64_t i = wbase + k * 32 + lane;
    if (i < n) out[gbase[dig[k]] + wcnt[warp][dig[k]] + rank[k]] = rec[k];
  }
}

template <typename Rec>
int sort_impl(Rec *a, Rec *b, int64_t n, int byte_lo, int byte_hi, void *d_tmp, cudaStream_t s, bool *in_b) {
  *in_b = false;
  if (n <= 1) return SIA_OK;
  const int64_t nblocks = ceil_div(n, kSortTile);
  // tmp layout: [4 x u64 fold] [hist u32 256*nblocks] [offs i64 256*nblocks+1] [scan tmp]
  char *p = (char *)d_tmp;
  unsigned long long *fold = (unsigned long long *)p; p += 64;
  uint32_t *hist = (uint32_t *)p; p += (size_t)256 * nblocks * sizeof(uint32_t);
  p = (char *)(((uintptr_t)p + 15) & ~(uintptr_t)15);
  int64_t *offs = (int64_t *)p; p += ((size_t)256 * nblocks + 1) * sizeof(int64_t);
  void *scan_tmp = p;

  unsigned long long init[4] = {0, 0, ~0ull, ~0ull}, got[4];
  SIA_CUDA(cudaMemcpyAsync(fold, init, sizeof init, cudaMemcpyHostToDevice, s));
  bitfold_kernel<Rec><<<(unsigned)std::min<int64_t>(ceil_div(n, 256), kNumSMs * 8), 256, 0, s>>>(a, n, fold);
  SIA_CHECK_LAUNCH();
  SIA_CUDA(cudaMemcpyAsync(got, fold, sizeof got, cudaMemcpyDeviceToHost, s));
  SIA_CUDA(cudaStreamSynchronize(s));
  const uint64_t vary[2] = {got[0] ^ got[2], got[1] ^ got[3]};

  Rec *src = a, *dst = b;
  for (int byte = byte_lo; byte < byte_hi; ++byte) {
    const uint64_t v = byte < 8 ? (vary[0] >> (8 * byte)) & 0xff : (sizeof(Rec) == 16 ? (vary[1] >> (8 * (byte - 8))) & 0xff : 0);
    if (!v) continue;   // every record has the same value in this byte
    hist_kernel<Rec><<<(unsigned)nblocks, kSortThreads, 0, s>>>(src, n, byte, hist, nblocks);
    SIA_CHECK_LAUNCH();
    int rc = exclusive_scan_u32(hist, offs, 256 * nblocks, scan_tmp, s);
    if (rc) return rc;
    scatter_kernel<Rec><<<(unsigned)nblocks, kSortThreads, 0, s>>>(src, dst, n, byte, offs, nblocks);
    SIA_CHECK_LAUNCH();
    std::swap(src, dst);
  }
  *in_b = (src == b);
  return SIA_OK;
}

}  // namespace

size_t radix_sort_tmp_bytes(int64_t n) {
  const int64_t nblocks = ceil_div(n > 0 ? n : 1, kSortTile);
  return 64 + (size_t)256 * nblocks * 4 + 16 + ((size_t)256 * nblocks + 1) * 8 + scan_tmp_bytes(256 * nblocks) + 64;
}

int radix_sort(void *d_a, void *d_b, int64_t n, int rec_bytes, int byte_lo, int byte_hi, void *d_tmp, cudaStream_t s,
               bool *result_in_b) {
  SIA_REQUIRE(rec_bytes == 8 || rec_bytes == 16, SIA_E_INVALID, "radix_sort: records are 8 or 16 bytes");
  SIA_REQUIRE(byte_lo >= 0 && byte_hi <= rec_bytes && byte_lo <= byte_hi, SIA_E_INVALID, "radix_sort: bad byte range");
  bool dummy;
  if (!result_in_b) result_in_b = &dummy;
  if (rec_bytes == 8) return sort_impl<uint64_t>((uint64_t *)d_a, (uint64_t *)d_b, n, byte_lo, byte_hi, d_tmp, s, result_in_b);
  return sort_impl<ulonglong2>((ulonglong2 *)d_a, (ulonglong2 *)d_b, n, byte_lo, byte_hi, d_tmp, s, result_in_b);
}

}  // namespace sia
