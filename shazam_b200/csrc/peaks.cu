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
// Stage 1 (this file, peaks_*_kernel): one CTA per 64-frame x 128-bin tile; writes a
//   bitmap, 1 bit per spectrogram element, row-major [frame][65 words].  The square
//   footprint (CONNECTIVITY_MASK = 2, the reference default) is separable: a sliding
//   max down the frames with threads along bins, a transposed store, then a sliding max
//   along the bins with threads along frames — every shared-memory access of both
//   passes is conflict-free, and each thread's window lives in registers (log-doubling:
//   8.4 max ops per output for the 21-wide window).
// Stage 2 (peaks_rowcount/extract): bitmap -> per-row popcounts -> scan -> ordered
//   (t asc, f asc) peak lists per track, which is the order generate_hashes' stable
//   sort by time produces (__init__.py:194-195).
#include "sia_common.cuh"
#include "stft.cuh"

#include <math_constants.h>

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
      bitmap[g * SIA_ROW_WORDS + gw] = word;
    }
  }
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
    if ((threadIdx.x & 31) == 0 && g < tc.row_hi && gw < SIA_ROW_WORDS) bitmap[g * SIA_ROW_WORDS + gw] = word;
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
__global__ void __launch_bounds__(256)
rowcount_kernel(const uint32_t *__restrict__ bitmap, int64_t total_frames, uint32_t *__restrict__ row_count) {
  const int lane = threadIdx.x & 31;
  const int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (g >= total_frames) return;
  const uint32_t *row = bitmap + g * SIA_ROW_WORDS;
  int c = __popc(row[lane]) + __popc(row[lane + 32]);
  if (lane == 0) c += __popc(row[64]);
#pragma unroll
  for (int d = 16; d; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
  if (lane == 0) row_count[g] = (uint32_t)c;
}

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
    if (k == 0) track_peak_starts[trk] = base;
    if (g == total_frames - 1) track_peak_starts[n_tracks] = row_off[total_frames] + peak_base;
  }
  const uint32_t *row = bitmap + g * SIA_ROW_WORDS;
  bool overflow = false;
#pragma unroll
  for (int round = 0; round < 3; ++round) {
    const int w = round * 32 + lane;
    uint32_t word = w < SIA_ROW_WORDS ? row[w] : 0u;
    const int cnt = __popc(word);
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int o = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += o;
    }
    int64_t idx = base + incl - cnt;
    while (word) {
      const int bit = __ffs(word) - 1;
      word &= word - 1;
      if (idx < cap) { peak_t[idx] = k; peak_f[idx] = w * 32 + bit; } else overflow = true;
      ++idx;
    }
    base += __shfl_sync(0xffffffffu, incl, 31);
  }
  if (overflow) atomicOr(status, 1);
}

}  // namespace

int peaks_bitmap_launch(const PeaksLaunch &a, cudaStream_t s) {
  if (a.total_ttiles == 0) return SIA_OK;
  SIA_REQUIRE(a.nbhd >= 1 && a.nbhd <= SIA_MAX_NBHD, SIA_E_UNSUPPORTED, "peaks: nbhd must be in 1..16");
  SIA_REQUIRE(a.connectivity == 1 || a.connectivity == 2, SIA_E_UNSUPPORTED, "peaks: connectivity must be 1 or 2");
  const bool erosion = a.amp_min < 0;
  if (a.connectivity == 2 && a.nbhd == 10) {
    // (double + erosion would need 255 KB of shared memory: it takes the generic kernel)
    if (a.in_type == SIA_F64 && !erosion) return launch_square<double, 10, false>(a, s);
    if (a.in_type == SIA_F32) return erosion ? launch_square<float, 10, true>(a, s) : launch_square<float, 10, false>(a, s);
  }
  return a.in_type == SIA_F64 ? launch_generic<double>(a, s) : launch_generic<float>(a, s);
}

int peaks_rowcount_launch(const uint32_t *d_bitmap, int64_t total_frames, uint32_t *d_row_count, cudaStream_t s) {
  if (total_frames == 0) return SIA_OK;
  rowcount_kernel<<<(unsigned)ceil_div(total_frames * 32, 256), 256, 0, s>>>(d_bitmap, total_frames, d_row_count);
  SIA_CHECK_LAUNCH();
  return SIA_OK;
}

int peaks_extract_launch(const uint32_t *d_bitmap, const int64_t *d_row_off, const int64_t *d_frame_starts,
                         int n_tracks, int64_t total_frames, int64_t peak_base, int32_t *d_peak_t,
                         int32_t *d_peak_f, int64_t cap_peaks, int64_t *d_track_peak_starts, int32_t *d_status,
                         cudaStream_t s) {
  if (total_frames == 0) return SIA_OK;
  extract_kernel<<<(unsigned)ceil_div(total_frames * 32, 256), 256, 0, s>>>(
      d_bitmap, d_row_off, d_frame_starts, n_tracks, total_frames, peak_base, d_peak_t, d_peak_f, cap_peaks,
      d_track_peak_starts, d_status);
  SIA_CHECK_LAUNCH();
  return SIA_OK;
}

}  // namespace sia
