// Micro-benchmark: issue cost of FP64 instructions on a B200 SM sub-partition, alone and mixed with integer work.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_rate fp64_rate.cu ; run on a B200.
// Each thread runs 16 independent DFMA chains (+ NINT independent integer adds per DFMA); 4 CTAs x 128 threads per
// SM = 4 warps per sub-partition, like K1.
#include <cstdio>
#include <cuda_runtime.h>

template <int NINT>
__global__ void __launch_bounds__(128) k(double *out, int iters, double a, double b, int m) {
  double x[16];
  int y[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) { x[i] = a + i + threadIdx.x; y[i] = m + i; }
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      x[i] = fma(x[i], a, b);
#pragma unroll
      for (int j = 0; j < NINT; ++j) y[(i + j) & 15] = (y[(i + j) & 15] + m) ^ it;     // one LOP3/IADD3-class op each (2 with the xor)
    }
  }
  double s = 0; int t = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) { s += x[i]; t += y[i]; }
  out[blockIdx.x * 128 + threadIdx.x] = s + t;
}

template <int NINT> void run(int ctas_per_sm) {
  double *d; cudaMalloc(&d, 8 * 148 * 8 * 128);
  const int iters = 8192;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<NINT><<<148 * ctas_per_sm, 128>>>(d, 16, 1.0000001, 1e-9, 3);
  cudaEventRecord(e0);
  k<NINT><<<148 * ctas_per_sm, 128>>>(d, iters, 1.0000001, 1e-9, 3);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  cudaError_t e = cudaGetLastError();
  const double dp_per_smsp = (double)iters * 16 * ctas_per_sm;       // one warp of each CTA per sub-partition
  const double cycles = ms * 1e-3 * 1.965e9;
  printf("DFMA + %d x (IADD, LOP3)  ctas/SM %d: %.3f ms, %.2f cycles per DFMA per sub-partition at 1965 MHz  [%s]\n", NINT,
         ctas_per_sm, ms, cycles / dp_per_smsp, cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  for (int c : {1, 2, 4, 8}) run<0>(c);
  run<1>(4); run<2>(4); run<3>(4); run<4>(4);
  return 0;
}
