// K2 — get_2D_peaks (__init__.py:116-177) on the time-major dB spectrogram.
//
//   local_max = maximum_filter(A, footprint) == A           (:143, scipy mode='reflect')
//   eroded_bg = binary_erosion(A == 0, footprint, border=1) (:147-148)
//   peak      = (local_max XOR eroded_bg) and A > amp_min   (:151,161)
//
// For a max filter 'reflect' equals clipping the window to the array, and eroded_bg can
// only be true where the whole clipped window is 0 (then local_max is true as well), so
//   peak = (A == max(window)) and not all(window == 0) and A > amp_min.
// The all-zero term can only matter when amp_min < 0 and is compiled in only then.
//
// Stage 1 writes a bitmap, 1 bit per spectrogram element.  Three kernels:
//   peaks_square_warp_tma_kernel  the pipeline's kernel (float32 dB, square 21x21 = CONNECTIVITY_MASK 2, the reference
//     default, amp_min >= 0): persistent CTAs walk 64-frame x 96-bin tiles; the 84 x 120 halo tile arrives by one TMA
//     tensor copy (cp.async.bulk.tensor.2d, out-of-range elements arrive as zeros), double-buffered on two mbarriers;
//     each warp owns 16 frames, each lane a float4 column: vertical 21-max in registers (van Herk / Gil-Werman: suffix
//     maxima of one block of rows, prefix maxima of the next, one max to combine), horizontal 21-max across lanes with
//     shuffles; 4 ballots per frame go to a striped bitmap [frame][22 strips][4 words];
//   peaks_square_kernel<T, N, EROSION>  float64 input (the bit-exact parity path) and amp_min < 0: one CTA per
//     64-frame x 128-bin tile + halo in shared memory; the separable filter as a sliding max down the frames with
//     threads along bins, a transposed store, then a sliding max along the bins with threads along frames (register
//     windows, log-doubling); plain bitmap [frame][65 words];
//   peaks_generic_kernel<T>  any half-width <= SIA_MAX_NBHD, square or diamond, brute force from the same tile.
// Stage 2 (peaks_rowcount/extract): bitmap -> per-row popcounts -> scan -> ordered
//   (t asc, f asc) peak lists per track, which is the order generate_hashes' stable
//   sort by time produces (__init__.py:194-195).
#include "sia_common.cuh"
#include "stft.cuh"

#include <cuda.h>
#include <cudaTypedefs.h>
#include <math.h>
#include <math_constants.h>
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>

namespace sia {

namespace {

constexpr int kThreads = 256;
constexpr int RC = 16;  // outputs per thread in the sliding passes

template <typename T> __device__ __forceinline__ T neg_inf();
template <> __device__ __forceinline__ float neg_inf<float>() { return -CUDART_INF_F; }
template <> __device__ __forceinline__ double neg_inf<double>() { return -CUDART_INF; }

// v[0..LEN) -> out[i] = op over v[i .. i+W) for i < OUTN, all indices compile-time
template <typename T, int W, int LEN, int OUTN, bool IS_MAX>
__device__ __forceinline__ void sliding(T (&v)[LEN]) {
  constexpr int P = W >= 16 ? 16 : (W >= 8 ? 8 : (W >= 4 ? 4 : (W >= 2 ? 2 : 1)));
#pragma unroll
  for (int s = 1; s < P; s <<= 1) {
#pragma unroll
    for (int i = 0; i + s < LEN; ++i) v[i] = IS_MAX ? max(v[i], v[i + s]) : min(v[i], v[i + s]);
  }
  if (W > P) {
#pragma unroll
    for (int i = 0; i < OUTN; ++i) {
      T x = v[i + W - P];
      if (W - P > P) {  // W up to 33 = 16 + 17: needs a third term
        T y = v[i + P];
        x = IS_MAX ? max(x, y) : min(x, y);
      }
      v[i] = IS_MAX ? max(v[i], x) : min(v[i], x);
    }
  }
}

// 21-wide window with 3-input max (FMNMX3): 3 -> 9 -> 21, 78 ops per 16 outputs
template <int LEN, int OUTN>
__device__ __forceinline__ void sliding21_max3(float (&v)[LEN]) {
  static_assert(LEN >= OUTN + 20, "window does not fit");
#pragma unroll
  for (int i = 0; i + 2 < LEN; ++i) v[i] = fmaxf(fmaxf(v[i], v[i + 1]), v[i + 2]);          // [i, i+3)
#pragma unroll
  for (int i = 0; i + 8 < LEN; ++i) v[i] = fmaxf(fmaxf(v[i], v[i + 3]), v[i + 6]);          // [i, i+9)
#pragma unroll
  for (int i = 0; i < OUTN; ++i) v[i] = fmaxf(fmaxf(v[i], v[i + 9]), v[i + 12]);             // [i, i+21)
}

// The same 16 windows of 21 out of 36 values, van Herk / Gil-Werman style: suffix maxima of v[0..20], prefix maxima of
// v[21..35], out[i] = max(suffix[i], prefix[i + 20]) — 49 max operations instead of 78 (the two scans advance two
// elements per step with a 3-input max, so the dependent chains are 10 and 7 long).
__device__ __forceinline__ void vanherk21_36(float (&v)[36]) {
  // suffix: v[i] = max(v[i..20]), i = 19 .. 0
#pragma unroll
  for (int i = 18; i >= 0; i -= 2) {
    const float s2 = v[i + 2];
    const float a = fmaxf(v[i + 1], s2);
    v[i] = fmaxf(fmaxf(v[i], v[i + 1]), s2);
    v[i + 1] = a;
  }
  // prefix: v[j] = max(v[21..j]), j = 22 .. 35
#pragma unroll
  for (int j = 22; j + 1 <= 35; j += 2) {
    const float p2 = v[j - 1];
    const float a = fmaxf(v[j], p2);
    v[j + 1] = fmaxf(fmaxf(v[j + 1], v[j]), p2);
    v[j] = a;
  }
#pragma unroll
  for (int i = 1; i < 16; ++i) v[i] = fmaxf(v[i], v[i + 20]);
}

struct TileCoord {
  int trk;
  int64_t row_lo, row_hi;  // global rows of this track [row_lo, row_hi)
  int64_t r0;              // first output row (global) of the tile
  int f0;                  // first output bin of the tile
};

__device__ __forceinline__ TileCoord tile_coord(const int64_t *__restrict__ frame_starts,
                                                const int64_t *__restrict__ ttile_starts, int n_tracks) {
  constexpr int kFTiles = (SIA_NBINS + kPeakTileF - 1) / kPeakTileF;  // 17
  TileCoord tc;
  const int64_t tt = blockIdx.x / kFTiles;
  tc.f0 = (int)(blockIdx.x % kFTiles) * kPeakTileF;
  tc.trk = find_segment(ttile_starts, n_tracks, tt);
  tc.row_lo = frame_starts[tc.trk];
  tc.row_hi = frame_starts[tc.trk + 1];
  tc.r0 = tc.row_lo + (tt - ttile_starts[tc.trk]) * kPeakTileT;
  return tc;
}

// Load the (TT+2N) x (TF+2N) halo tile; out-of-track rows and out-of-range bins get `fill`.
template <typename T, int N, int AS>
__device__ __forceinline__ void load_tile(T *__restrict__ A, const T *__restrict__ spec, const TileCoord &tc, T fill) {
  constexpr int ROWS = kPeakTileT + 2 * N, COLS = kPeakTileF + 2 * N;
  for (int idx = threadIdx.x; idx < ROWS * COLS; idx += kThreads) {
    const int r = idx / COLS, c = idx - r * COLS;
    const int64_t g = tc.r0 - N + r;
    const int f = tc.f0 - N + c;
    T v = fill;
    if (g >= tc.row_lo && g < tc.row_hi && f >= 0 && f < SIA_NBINS) v = spec[g * SIA_F_STRIDE + f];
    A[r * AS + c] = v;
  }
}

// ---- fast path: square footprint, compile-time half-width N -------------------------------------
template <typename T, int N, bool EROSION>
__global__ void __launch_bounds__(kThreads)
peaks_square_kernel(const T *__restrict__ spec, const int64_t *__restrict__ frame_starts,
                    const int64_t *__restrict__ ttile_starts, int n_tracks, double amp_min,
                    uint32_t *__restrict__ bitmap) {
  constexpr int TT = kPeakTileT, TF = kPeakTileF;
  constexpr int ROWS = TT + 2 * N, COLS = TF + 2 * N;
  constexpr int AS = COLS | 1;     // odd row stride: column walks with threads along rows are conflict-free
  constexpr int VS = TT + 1;       // odd stride of the transposed intermediate
  constexpr int W = 2 * N + 1, LEN = RC + 2 * N;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T *A = reinterpret_cast<T *>(smem_raw);            // [ROWS][AS]
  T *VT = A + ROWS * AS;                             // [COLS][VS]   max down the rows, transposed
  T *VTmin = VT + COLS * VS;                         // [COLS][VS]   (EROSION only)
  uint16_t *sbits = reinterpret_cast<uint16_t *>(EROSION ? (VTmin + COLS * VS) : VTmin);  // [TT][TF/16]

  const TileCoord tc = tile_coord(frame_starts, ttile_starts, n_tracks);
  load_tile<T, N, AS>(A, spec, tc, neg_inf<T>());
  __syncthreads();

  // pass V: thread = (bin column c, row chunk q); window slides down the rows
  for (int task = threadIdx.x; task < COLS * (TT / RC); task += kThreads) {
    const int q = task / COLS, c = task - q * COLS;
    T v[LEN];
#pragma unroll
    for (int i = 0; i < LEN; ++i) v[i] = A[(q * RC + i) * AS + c];
    if (EROSION) {
      T u[LEN];
#pragma unroll
      for (int i = 0; i < LEN; ++i) u[i] = v[i] == neg_inf<T>() ? -neg_inf<T>() : v[i];  // outside: +inf for min
      sliding<T, W, LEN, RC, false>(u);
#pragma unroll
      for (int i = 0; i < RC; ++i) VTmin[c * VS + q * RC + i] = u[i];
    }
    sliding<T, W, LEN, RC, true>(v);
#pragma unroll
    for (int i = 0; i < RC; ++i) VT[c * VS + q * RC + i] = v[i];
  }
  __syncthreads();

  // pass H: thread = (frame row r, bin chunk j); window slides along the bins
  for (int task = threadIdx.x; task < TT * (TF / RC); task += kThreads) {
    const int j = task / TT, r = task - j * TT;
    T v[LEN];
#pragma unroll
    for (int i = 0; i < LEN; ++i) v[i] = VT[(j * RC + i) * VS + r];
    sliding<T, W, LEN, RC, true>(v);
    T u[LEN];
    if (EROSION) {
#pragma unroll
      for (int i = 0; i < LEN; ++i) u[i] = VTmin[(j * RC + i) * VS + r];
      sliding<T, W, LEN, RC, false>(u);
    }
    const bool row_ok = tc.r0 + r < tc.row_hi;
    uint32_t bits = 0;
#pragma unroll
    for (int i = 0; i < RC; ++i) {
      const T ctr = A[(r + N) * AS + j * RC + i + N];
      bool pk = row_ok && (tc.f0 + j * RC + i < SIA_NBINS) && ctr == v[i] && (double)ctr > amp_min;
      if (EROSION) pk = pk && !(v[i] == (T)0 && u[i] == (T)0);
      bits |= (uint32_t)pk << i;
    }
    sbits[r * (TF / 16) + j] = (uint16_t)bits;
  }
  __syncthreads();

  // bitmap words: TT rows x TF/32 words per tile
  for (int idx = threadIdx.x; idx < TT * (TF / 32); idx += kThreads) {
    const int r = idx / (TF / 32), w = idx - r * (TF / 32);
    const int64_t g = tc.r0 + r;
    const int gw = tc.f0 / 32 + w;
    if (g < tc.row_hi && gw < SIA_ROW_WORDS) {
      const uint32_t word = (uint32_t)sbits[r * (TF / 16) + 2 * w] | ((uint32_t)sbits[r * (TF / 16) + 2 * w + 1] << 16);
      bitmap[g * kBitmapRowWords + gw] = word;
    }
  }
}

// ---- production path: float32, square 21x21, amp_min >= 0 ------------------------------------------------
// One CTA = 64 frames x 96 bins (3 bitmap words per frame); 4 warps, warp q owns frames 16q..16q+15.
// The 84 x 120 halo tile is staged in shared memory as float4 (conflict-free 128-bit accesses).  Each lane
// owns one float4 column: it slides the 21-frame max down its 4 bins entirely in registers (log-doubling,
// 36 float4 in flight), then takes the 21-bin max across lanes with 10 shuffles of per-lane prefix /
// suffix / block maxima — no second shared-memory pass, no transposes.  Lanes 0..23 produce the 96
// outputs, lanes 24..29 carry the +-12-bin halo columns, lanes 30..31 idle.
constexpr int kW2Rows = kPeakTileT;                                 // 64
constexpr int kW2Bins = 96;
constexpr int kW2Cols4 = 30;                                        // (96 + 2*12) / 4
constexpr int kW2Strips = (SIA_NBINS + kW2Bins - 1) / kW2Bins;      // 22
constexpr int kW2Threads = 128;
constexpr int kW2TileRows = kW2Rows + 20;                           // 84

// The filter proper, on a staged tile A (84 x 30 float4): see the kernel comment above.
__device__ __forceinline__ void warp_tile_peaks(const float4 *__restrict__ A, uint32_t a_base, int64_t r0, int64_t row_hi,
                                                int strip, int f0, float amp_lo, uint32_t *__restrict__ bitmap) {
  const int lane = threadIdx.x & 31, q = threadIdx.x >> 5;
  const int col = lane < kW2Cols4 ? (lane + 3) % kW2Cols4 : 0;      // lanes 0..23 -> columns 3..26 (outputs)
  auto src = [&](int k) { return lane < kW2Cols4 ? (lane + k + kW2Cols4) % kW2Cols4 : lane; };
  const int s_m1 = src(-1), s_m2 = src(-2), s_m3 = src(-3), s_p1 = src(1), s_p2 = src(2), s_p3 = src(3);

  float vx[36], vy[36], vz[36], vw[36];
#pragma unroll
  for (int i = 0; i < 36; ++i) {
    const float4 t = A[(16 * q + i) * kW2Cols4 + col];
    vx[i] = t.x; vy[i] = t.y; vz[i] = t.z; vw[i] = t.w;
  }
  vanherk21_36(vx);
  vanherk21_36(vy);
  vanherk21_36(vz);
  vanherk21_36(vw);

  const unsigned FULL = 0xffffffffu;
  const float amp_up = nextafterf(amp_lo, CUDART_INF_F);
  const int64_t g0 = r0 + 16 * q;                                   // first output frame of this warp
  const int nvalid = (int)min((int64_t)16, row_hi - g0);            // frames of the chunk inside the track
  const int fl = f0 + 4 * lane;
  const bool v0 = lane < 24 && fl < SIA_NBINS, v1 = lane < 24 && fl + 1 < SIA_NBINS;
  const bool v2 = lane < 24 && fl + 2 < SIA_NBINS, v3 = lane < 24 && fl + 3 < SIA_NBINS;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float a0 = vx[i], a1 = vy[i], a2 = vz[i], a3 = vw[i];
    // amp_up (the smallest float above the threshold) rides in the block maximum M, which every window
    // includes, so "centre == window max and centre > amp_min" becomes the single test centre >= h
    const float S2 = fmaxf(a2, a3), S1 = fmaxf(a1, S2), M = fmaxf(fmaxf(a0, S1), amp_up);
    const float P1 = fmaxf(a0, a1), P2 = fmaxf(P1, a2);
    const float Mm1 = __shfl_sync(FULL, M, s_m1), Mp1 = __shfl_sync(FULL, M, s_p1);
    const float Mm2 = __shfl_sync(FULL, M, s_m2), Mp2 = __shfl_sync(FULL, M, s_p2);
    const float S1m2 = __shfl_sync(FULL, S1, s_m2);
    const float S2m3 = __shfl_sync(FULL, S2, s_m3), S3m3 = __shfl_sync(FULL, a3, s_m3);
    const float P2p2 = __shfl_sync(FULL, P2, s_p2);
    const float P0p3 = __shfl_sync(FULL, a0, s_p3), P1p3 = __shfl_sync(FULL, P1, s_p3);
    const float common = fmaxf(fmaxf(Mm1, M), Mp1);
    const float cl = fmaxf(common, Mm2);
    const float h0 = fmaxf(cl, fmaxf(S2m3, P2p2));          // bins -10..+10 around element 0
    const float h1 = fmaxf(cl, fmaxf(S3m3, Mp2));
    const float h2 = fmaxf(cl, fmaxf(Mp2, P0p3));
    const float h3 = fmaxf(fmaxf(common, S1m2), fmaxf(Mp2, P1p3));
    // re-read the centre row from shared memory (volatile: keeps 16 float4 out of the register file)
    float4 c;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];\n"
                 : "=f"(c.x), "=f"(c.y), "=f"(c.z), "=f"(c.w)
                 : "r"(a_base + (uint32_t)((16 * q + i + 10) * kW2Cols4 + col) * 16u));
    // one ballot per element index: word e, bit l  <->  bin f0 + 4*l + e   (striped bitmap layout)
    const uint32_t b0 = __ballot_sync(FULL, c.x >= h0 && v0);
    const uint32_t b1 = __ballot_sync(FULL, c.y >= h1 && v1);
    const uint32_t b2 = __ballot_sync(FULL, c.z >= h2 && v2);
    const uint32_t b3 = __ballot_sync(FULL, c.w >= h3 && v3);
    if (lane == 0 && i < nvalid)
      *reinterpret_cast<uint4 *>(bitmap + (g0 + i) * kBitmapRowWords + 4 * strip) = make_uint4(b0, b1, b2, b3);
  }
}

// Persistent kernel: 2 CTAs per SM walk the tiles with a stride of the grid, two tile buffers
// each; the TMA copy of the next tile is in flight while the current one is filtered, so no warp ever waits for
// the tile it is about to read (the copy engine does the staging, the 128 threads only compute).
constexpr size_t kW2TileBytes = sizeof(float4) * kW2TileRows * kW2Cols4;
constexpr size_t kW2PersistSmem = 2 * kW2TileBytes + 128;

__global__ void __launch_bounds__(kW2Threads, 2)
peaks_square_warp_tma_kernel(const int64_t *__restrict__ frame_starts, const int64_t *__restrict__ ttile_starts, int n_tracks,
                             int64_t n_tiles, float amp_lo, uint32_t *__restrict__ bitmap,
                             const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ __align__(128) unsigned char w2_smem[];
  float4 *bufs[2] = {reinterpret_cast<float4 *>(w2_smem), reinterpret_cast<float4 *>(w2_smem + kW2TileBytes)};
  unsigned long long *bars = reinterpret_cast<unsigned long long *>(w2_smem + 2 * kW2TileBytes);
  const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(bars);
  const uint32_t base0 = (uint32_t)__cvta_generic_to_shared(w2_smem);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(bar0));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(bar0 + 8));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  // thread 0: start the tensor copy of tile `tile` into buffer b
  auto issue = [&](int64_t tile, int b) {
    const int64_t tt = tile / kW2Strips;
    const int strip = (int)(tile - tt * kW2Strips);
    const int trk = find_segment(ttile_starts, n_tracks, tt);
    const int64_t r0 = frame_starts[trk] + (tt - ttile_starts[trk]) * kW2Rows;
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // earlier generic writes (zeroed frames) vs the copy
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar0 + 8 * b), "r"((uint32_t)kW2TileBytes)
                 : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n"
                 ::"r"(base0 + (uint32_t)(b * kW2TileBytes)), "l"(reinterpret_cast<uint64_t>(&tmap)),
                 "r"(strip * kW2Bins - 12), "r"((int)(r0 - 10)), "r"(bar0 + 8 * b) : "memory");
  };
  int64_t tile = blockIdx.x;
  if (threadIdx.x == 0 && tile < n_tiles) issue(tile, 0);
  for (int k = 0; tile < n_tiles; tile += gridDim.x, ++k) {
    const int b = k & 1;
    // the other buffer was released by the barrier that ended the previous iteration
    if (threadIdx.x == 0 && tile + gridDim.x < n_tiles) issue(tile + gridDim.x, b ^ 1);
    const int64_t tt = tile / kW2Strips;
    const int strip = (int)(tile - tt * kW2Strips);
    const int trk = find_segment(ttile_starts, n_tracks, tt);
    const int64_t row_lo = frame_starts[trk], row_hi = frame_starts[trk + 1];
    const int64_t r0 = row_lo + (tt - ttile_starts[trk]) * kW2Rows;
    asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}\n"
                 ::"r"(bar0 + 8 * b), "r"((k >> 1) & 1) : "memory");
    float4 *A = bufs[b];
    const int64_t g_first = r0 - 10;
    if (g_first < row_lo || g_first + kW2TileRows > row_hi) {      // first / last tiles of a track: zero the other tracks' frames
      for (int i = threadIdx.x; i < kW2TileRows * kW2Cols4; i += kW2Threads) {
        const int64_t g = g_first + i / kW2Cols4;
        if (g < row_lo || g >= row_hi) A[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      __syncthreads();
    }
    warp_tile_peaks(A, base0 + (uint32_t)(b * kW2TileBytes), r0, row_hi, strip, strip * kW2Bins, amp_lo, bitmap);
    __syncthreads();                                               // buffer b may be refilled two tiles from now
  }
}

// tensor map of the chunk's spectrogram for the TMA tile loads: [total_frames][2049] float32, row pitch 2080 floats,
// box = one halo tile (84 frames x 120 bins), zero fill outside
int make_spec_tensor_map(const float *spec, int64_t total_frames, CUtensorMap *tm) {
  static PFN_cuTensorMapEncodeTiled encode = nullptr;
  if (!encode) {
    cudaDriverEntryPointQueryResult qres;
    void *fn = nullptr;
    SIA_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    SIA_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, SIA_E_CUDA, "cuTensorMapEncodeTiled is not available");
    encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(fn);
  }
  const cuuint64_t gdim[2] = {(cuuint64_t)SIA_NBINS, (cuuint64_t)total_frames};
  const cuuint64_t gstride[1] = {(cuuint64_t)SIA_F_STRIDE * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)(4 * kW2Cols4), (cuuint32_t)kW2TileRows};
  const cuuint32_t estride[2] = {1, 1};
  const CUresult r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(spec), gdim, gstride, box, estride,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SIA_REQUIRE(r == CUDA_SUCCESS, SIA_E_CUDA, "cuTensorMapEncodeTiled failed for the spectrogram");
  return SIA_OK;
}

int launch_square_warp(const PeaksLaunch &a, cudaStream_t s) {
  // float c > (double) amp_min  <=>  c > amp_lo with amp_lo = amp_min rounded DOWN to float
  float amp_lo = (float)a.amp_min;
  if ((double)amp_lo > a.amp_min) amp_lo = nextafterf(amp_lo, -INFINITY);
  const int64_t blocks = a.total_ttiles * kW2Strips;
  alignas(64) CUtensorMap tm;
  memset(&tm, 0, sizeof tm);
  int rc = make_spec_tensor_map((const float *)a.d_spec, a.total_frames, &tm);
  if (rc) return rc;
  SIA_CUDA(cudaFuncSetAttribute(peaks_square_warp_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kW2PersistSmem));
  const unsigned grid = (unsigned)std::min<int64_t>(blocks, 2 * kNumSMs);
  peaks_square_warp_tma_kernel<<<grid, kW2Threads, kW2PersistSmem, s>>>(a.d_frame_starts, a.d_ttile_starts, a.n_tracks, blocks,
                                                                        amp_lo, a.d_bitmap, tm);
  SIA_CHECK_LAUNCH();
  return SIA_OK;
}

// ---- generic path: any half-width <= SIA_MAX_NBHD, square or diamond, brute force -----------------
template <typename T>
__global__ void __launch_bounds__(kThreads)
peaks_generic_kernel(const T *__restrict__ spec, const int64_t *__restrict__ frame_starts,
                     const int64_t *__restrict__ ttile_starts, int n_tracks, double amp_min, int nb, int diamond,
                     int erosion, uint32_t *__restrict__ bitmap) {
  constexpr int TT = kPeakTileT, TF = kPeakTileF, N = SIA_MAX_NBHD;
  constexpr int ROWS = TT + 2 * N, COLS = TF + 2 * N;
  constexpr int AS = COLS | 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T *A = reinterpret_cast<T *>(smem_raw);

  const TileCoord tc = tile_coord(frame_starts, ttile_starts, n_tracks);
  load_tile<T, N, AS>(A, spec, tc, neg_inf<T>());
  __syncthreads();

  for (int idx = threadIdx.x; idx < TT * TF; idx += kThreads) {   // warp = 32 consecutive bins of one frame
    const int r = idx / TF, c = idx - r * TF;
    const int64_t g = tc.r0 + r;
    const int f = tc.f0 + c;
    const T ctr = A[(r + N) * AS + c + N];
    bool pk = g < tc.row_hi && f < SIA_NBINS && (double)ctr > amp_min;
    if (pk) {
      bool allzero = true;
      for (int dt = -nb; dt <= nb && pk; ++dt) {
        const int wf = diamond ? nb - abs(dt) : nb;
        const T *rowp = A + (r + N + dt) * AS + c + N;
        for (int df = -wf; df <= wf; ++df) {
          const T x = rowp[df];
          if (x > ctr) { pk = false; break; }
          if (x != (T)0 && x != neg_inf<T>()) allzero = false;
        }
      }
      if (erosion && allzero) pk = false;
    }
    const uint32_t word = __ballot_sync(0xffffffffu, pk);
    const int gw = f >> 5;
    if ((threadIdx.x & 31) == 0 && g < tc.row_hi && gw < SIA_ROW_WORDS) bitmap[g * kBitmapRowWords + gw] = word;
  }
}

template <typename T, int N, bool E>
int launch_square(const PeaksLaunch &a, cudaStream_t s) {
  constexpr int ROWS = kPeakTileT + 2 * N, COLS = kPeakTileF + 2 * N, AS = COLS | 1, VS = kPeakTileT + 1;
  const size_t smem = sizeof(T) * (ROWS * AS + COLS * VS * (E ? 2 : 1)) + kPeakTileT * (kPeakTileF / 16) * 2;
  auto kern = peaks_square_kernel<T, N, E>;
  SIA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t blocks = a.total_ttiles * ((SIA_NBINS + kPeakTileF - 1) / kPeakTileF);
  kern<<<(unsigned)blocks, kThreads, smem, s>>>((const T *)a.d_spec, a.d_frame_starts, a.d_ttile_starts, a.n_tracks,
                                              a.amp_min, a.d_bitmap);
  SIA_CHECK_LAUNCH();
  return SIA_OK;
}

template <typename T>
int launch_generic(const PeaksLaunch &a, cudaStream_t s) {
  constexpr int N = SIA_MAX_NBHD;
  constexpr int ROWS = kPeakTileT + 2 * N, COLS = kPeakTileF + 2 * N, AS = COLS | 1;
  const size_t smem = sizeof(T) * ROWS * AS;
  auto kern = peaks_generic_kernel<T>;
  SIA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t blocks = a.total_ttiles * ((SIA_NBINS + kPeakTileF - 1) / kPeakTileF);
  kern<<<(unsigned)blocks, kThreads, smem, s>>>((const T *)a.d_spec, a.d_frame_starts, a.d_ttile_starts, a.n_tracks,
                                              a.amp_min, a.nbhd, a.connectivity == 1, a.amp_min < 0, a.d_bitmap);
  SIA_CHECK_LAUNCH();
  return SIA_OK;
}

// ---- stage 2 ------------------------------------------------------------------------------------------
// Two bitmap layouts, both with a row stride of kBitmapRowWords words:
//   plain   : word w, bit b            <-> bin 32*w + b              (65 words; tile / generic kernels)
//   striped : word 4*s + e, bit l      <-> bin 96*s + 4*l + e        (22 strips x 4 words; warp kernel)
template <bool STRIPED>
__global__ void __launch_bounds__(256)
rowcount_kernel(const uint32_t *__restrict__ bitmap, int64_t total_frames, uint32_t *__restrict__ row_count) {
  const int lane = threadIdx.x & 31;
  const int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (g >= total_frames) return;
  const uint32_t *row = bitmap + g * kBitmapRowWords;
  int c = 0;
  if (STRIPED) {
    if (lane < kW2Strips) {
      const uint4 w = *reinterpret_cast<const uint4 *>(row + 4 * lane);
      c = __popc(w.x) + __popc(w.y) + __popc(w.z) + __popc(w.w);
    }
  } else {
    c = __popc(row[lane]) + __popc(row[lane + 32]);
    if (lane == 0) c += __popc(row[64]);
  }
#pragma unroll
  for (int d = 16; d; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
  if (lane == 0) row_count[g] = (uint32_t)c;
}

template <bool STRIPED>
__global__ void __launch_bounds__(256)
extract_kernel(const uint32_t *__restrict__ bitmap, const int64_t *__restrict__ row_off,
               const int64_t *__restrict__ frame_starts, int n_tracks, int64_t total_frames, int64_t peak_base,
               int32_t *__restrict__ peak_t, int32_t *__restrict__ peak_f, int64_t cap,
               int64_t *__restrict__ track_peak_starts, int32_t *__restrict__ status) {
  const int lane = threadIdx.x & 31;
  const int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (g >= total_frames) return;
  const int trk = find_segment(frame_starts, n_tracks, g);
  const int k = (int)(g - frame_starts[trk]);
  int64_t base = row_off[g] + peak_base;
  if (lane == 0) {
    // prefix offsets never point past the workspace: the peaks beyond `cap` are dropped (status bit 1), and the
    // kernels that follow (pair counting, scans, SHA-1) read these counts from the device
    if (k == 0) track_peak_starts[trk] = min(base, cap);
    if (g == total_frames - 1) track_peak_starts[n_tracks] = min(row_off[total_frames] + peak_base, cap);
  }
  const uint32_t *row = bitmap + g * kBitmapRowWords;
  bool overflow = false;
  auto excl_scan = [&](int cnt, int &total) {
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int o = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += o;
    }
    total = __shfl_sync(0xffffffffu, incl, 31);
    return incl - cnt;
  };
  if (STRIPED) {
    uint4 w = make_uint4(0, 0, 0, 0);
    if (lane < kW2Strips) w = *reinterpret_cast<const uint4 *>(row + 4 * lane);
    const int cnt = __popc(w.x) + __popc(w.y) + __popc(w.z) + __popc(w.w);
    int total;
    int64_t idx = base + excl_scan(cnt, total);
    uint32_t any = w.x | w.y | w.z | w.w;
    while (any) {                                   // ascending bin order: lane-bit l major, element e minor
      const int l = __ffs(any) - 1;
      any &= any - 1;
      const uint32_t nib = ((w.x >> l) & 1u) | (((w.y >> l) & 1u) << 1) | (((w.z >> l) & 1u) << 2) | (((w.w >> l) & 1u) << 3);
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if ((nib >> e) & 1u) {
          if (idx < cap) { peak_t[idx] = k; peak_f[idx] = kW2Bins * lane + 4 * l + e; } else overflow = true;
          ++idx;
        }
    }
  } else {
#pragma unroll
    for (int round = 0; round < 3; ++round) {
      const int w = round * 32 + lane;
      uint32_t word = w < SIA_ROW_WORDS ? row[w] : 0u;
      const int cnt = __popc(word);
      int total;
      int64_t idx = base + excl_scan(cnt, total);
      while (word) {
        const int bit = __ffs(word) - 1;
        word &= word - 1;
        if (idx < cap) { peak_t[idx] = k; peak_f[idx] = w * 32 + bit; } else overflow = true;
        ++idx;
      }
      base += total;
    }
  }
  if (overflow) atomicOr(status, 1);
}

}  // namespace

int peaks_bitmap_launch(const PeaksLaunch &a, cudaStream_t s, bool *striped) {
  *striped = false;
  if (a.total_ttiles == 0) return SIA_OK;
  SIA_REQUIRE(a.nbhd >= 1 && a.nbhd <= SIA_MAX_NBHD, SIA_E_UNSUPPORTED, "peaks: nbhd must be in 1..16");
  SIA_REQUIRE(a.connectivity == 1 || a.connectivity == 2, SIA_E_UNSUPPORTED, "peaks: connectivity must be 1 or 2");
  const bool erosion = a.amp_min < 0;
  if (a.connectivity == 2 && a.nbhd == 10) {
    // (double + erosion would need 255 KB of shared memory: it takes the generic kernel)
    if (a.in_type == SIA_F64 && !erosion) return launch_square<double, 10, false>(a, s);
    if (a.in_type == SIA_F32 && !erosion) {
      *striped = true;
      return launch_square_warp(a, s);
    }
    if (a.in_type == SIA_F32) return launch_square<float, 10, true>(a, s);
  }
  return a.in_type == SIA_F64 ? launch_generic<double>(a, s) : launch_generic<float>(a, s);
}

int peaks_rowcount_launch(const uint32_t *d_bitmap, bool striped, int64_t total_frames, uint32_t *d_row_count,
                          cudaStream_t s) {
  if (total_frames == 0) return SIA_OK;
  const unsigned blocks = (unsigned)ceil_div(total_frames * 32, 256);
  if (striped) rowcount_kernel<true><<<blocks, 256, 0, s>>>(d_bitmap, total_frames, d_row_count);
  else rowcount_kernel<false><<<blocks, 256, 0, s>>>(d_bitmap, total_frames, d_row_count);
  SIA_CHECK_LAUNCH();
  return SIA_OK;
}

int peaks_extract_launch(const uint32_t *d_bitmap, bool striped, const int64_t *d_row_off,
                         const int64_t *d_frame_starts, int n_tracks, int64_t total_frames, int64_t peak_base, int32_t *d_peak_t,
                         int32_t *d_peak_f, int64_t cap_peaks, int64_t *d_track_peak_starts, int32_t *d_status,
                         cudaStream_t s) {
  if (total_frames == 0) return SIA_OK;
  const unsigned blocks = (unsigned)ceil_div(total_frames * 32, 256);
  if (striped)
    extract_kernel<true><<<blocks, 256, 0, s>>>(d_bitmap, d_row_off, d_frame_starts, n_tracks, total_frames, peak_base,
                                                d_peak_t, d_peak_f, cap_peaks, d_track_peak_starts, d_status);
  else
    extract_kernel<false><<<blocks, 256, 0, s>>>(d_bitmap, d_row_off, d_frame_starts, n_tracks, total_frames, peak_base,
                                                 d_peak_t, d_peak_f, cap_peaks, d_track_peak_starts, d_status);
  SIA_CHECK_LAUNCH();
  return SIA_OK;
}

}  // namespace sia
