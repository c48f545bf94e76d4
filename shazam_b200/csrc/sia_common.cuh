// Shared helpers for the sm_100a kernels of the SIA fingerprint-and-match path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/sia_b200.h"

namespace sia {

void set_error(const std::string &msg);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define SIA_CUDA(expr)                                                            \
  do {                                                                            \
    cudaError_t _e = (expr);                                                      \
    if (_e != cudaSuccess) return ::sia::cuda_fail(_e, #expr, __FILE__, __LINE__); \
  } while (0)

#define SIA_CHECK_LAUNCH() SIA_CUDA(cudaGetLastError())

#define SIA_REQUIRE(cond, code, msg)                 \
  do {                                               \
    if (!(cond)) {                                   \
      ::sia::set_error(msg);                         \
      return (code);                                 \
    }                                                \
  } while (0)

constexpr int kNumSMs = 148;  // B200

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// largest b with starts[b] <= x, for a non-decreasing prefix array starts[0..n] (starts[0] <= x < starts[n])
template <typename T>
__device__ __forceinline__ int find_segment(const T *__restrict__ starts, int n, T x) {
  int lo = 0, hi = n;  // invariant: starts[lo] <= x < starts[hi]
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (starts[mid] <= x) lo = mid; else hi = mid;
  }
  return lo;
}

// ---- device-wide exclusive scan (uint32 in, uint64 out), scan.cu ---------------------------
// d_out has n+1 entries: out[i] = sum(in[0..i)), out[n] = total.  d_tmp: scan_tmp_bytes(n).
size_t scan_tmp_bytes(int64_t n);
int exclusive_scan_u32(const uint32_t *d_in, int64_t *d_out, int64_t n, void *d_tmp, cudaStream_t s);
// same with n read from device memory (*d_n <= n_max); entries beyond *d_n are not written
int exclusive_scan_u32_dyn(const uint32_t *d_in, int64_t *d_out, const int64_t *d_n, int64_t n_max,
                           void *d_tmp, cudaStream_t s);

}  // namespace sia
