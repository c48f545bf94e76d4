// K4 — the fingerprint index: replaces the MySQL `fingerprints` table
// (mysql_database.py:46-59), its INSERT IGNORE (:62-68,167-181), the
// SELECT ... WHERE hash IN (...) lookup (recognizer.py:60-64,252-259) and the vote inside
// align_matches (recognizer.py:303-310).
//
// Storage: one 16-byte row per fingerprint, sorted by (hash, song_id, offset):
//     .y = digest bytes 0..7 (big-endian, so integer order = byte order)
//     .x = digest bytes 8..9 (16 bits) | song_id (24 bits) | offset (24 bits)
// plus a bucket directory on the top `dir_bits` bits of the digest (SHA-1 output is
// uniform, so buckets are balanced; a heavy key is one long run inside its bucket).
// Equal rows are adjacent after the sort, which is how UNIQUE(song_id, offset, hash)
// (INSERT IGNORE set semantics) is enforced.
//
// Query: the (hash, offset) pairs of a batch of queries are sorted by (query, hash,
// offset) — duplicates drop out, equal hashes of one query become adjacent — each is
// looked up (directory + binary search), the posting runs are expanded into
// (query, song, db_offset - query_offset) keys, sorted, run-length counted into bins,
// and voted: per (query, song) the bin with the largest count (smallest difference on
// ties), per query the top-n songs by that count (ascending song id on ties).
#include "sia_common.cuh"
#include "sort.cuh"

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <string>
#include <vector>

using namespace sia;

namespace {

constexpr int kQidBits = SIA_BINKEY_QID_BITS, kSongBits = SIA_BINKEY_SONG_BITS, kDiffBits = SIA_BINKEY_DIFF_BITS;
constexpr int64_t kMaxQueriesPerPass = 1ll << kQidBits;
constexpr uint64_t kM24 = 0xffffffull;

struct Arena {
  char *base = nullptr;
  size_t cap = 0, used = 0;
  int reserve(size_t bytes) {
    if (bytes <= cap) { used = 0; return SIA_OK; }
    if (base) cudaFree(base);
    base = nullptr; cap = 0; used = 0;
    size_t want = bytes + (bytes >> 3) + (1 << 20);
    SIA_CUDA(cudaMalloc(&base, want));
    cap = want;
    return SIA_OK;
  }
  template <typename T> T *take(size_t n) {
    used = (used + 255) & ~(size_t)255;
    T *p = reinterpret_cast<T *>(base + used);
    used += n * sizeof(T);
    return used <= cap ? p : nullptr;
  }
  void release() { if (base) cudaFree(base); base = nullptr; cap = used = 0; }
};

}  // namespace

namespace {
// Sorted, de-duplicated query entries for queries [q0, q1) (entries [i0, i1) of the caller's arrays).
struct Lookup {
  ulonglong2 *ent = nullptr;
  uint32_t *first = nullptr, *cnt_head = nullptr;
  int64_t *off_all = nullptr, *off_head = nullptr;
  int64_t n = 0, tuples = 0, head_rows = 0, distinct = 0;
};
}  // namespace

// scratch of the handle-less vote entry points, kept per device (cudaMalloc of tens of GB per call is slow)
static Arena g_vote_arena[64];

struct sia_index {
  int device = 0;
  int64_t capacity = 0;
  ulonglong2 *rows = nullptr;   // [capacity]: sorted rows [0, n_rows), then pending rows
  int64_t n_rows = 0, n_pending = 0;
  uint32_t *dir = nullptr;      // [2^dir_bits + 1]
  int dir_bits = 0;
  int32_t *status = nullptr;    // [0] device flags: 1 = song/offset out of range on insert, 2 = query offset out of range,
                                //     8 / 16 = vote table full / bin count overflow (cannot happen: sizing, 32 767-entry rule)
                                // [1] largest song id ever inserted
  int32_t max_song = 0;         // host copy of status[1], refreshed by finalize
  Arena arena;                  // build / lookup scratch
  Arena arena2;                 // per-group vote scratch (sort-based vote)
  Arena arena3;                 // vote tables (hash-table vote)
  // lookup of the last sizing call of sia_index_expand, reused by the call that fills the buffers
  bool cache_valid = false;
  const void *cache_hash = nullptr, *cache_qoff = nullptr, *cache_qid = nullptr;
  int64_t cache_n = 0;
  Lookup cache_lookup;
};

namespace {

__device__ __forceinline__ void load_digest(const uint8_t *__restrict__ h, uint64_t &hi, uint32_t &lo16) {
  const uint16_t *p = reinterpret_cast<const uint16_t *>(h);   // 10*i is 2-byte aligned
  uint32_t w[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) { const uint32_t v = p[k]; w[k] = ((v & 0xff) << 8) | (v >> 8); }   // big-endian pairs
  hi = ((uint64_t)w[0] << 48) | ((uint64_t)w[1] << 32) | ((uint64_t)w[2] << 16) | (uint64_t)w[3];
  lo16 = w[4];
}

__global__ void __launch_bounds__(256)
pack_rows_kernel(const uint8_t *__restrict__ hash, const int32_t *__restrict__ off, const int32_t *__restrict__ song_arr,
                 int32_t song_const, int64_t n, ulonglong2 *__restrict__ out, int32_t *__restrict__ status) {
  int32_t smax = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint64_t hi; uint32_t lo16;
    load_digest(hash + i * SIA_HASH_BYTES, hi, lo16);
    const int32_t song = song_arr ? song_arr[i] : song_const;
    const int32_t o = off[i];
    if (song < 0 || song > (int32_t)kM24 || o < 0 || o > (int32_t)kM24) atomicOr(status, 1);
    else smax = max(smax, song);
    out[i] = make_ulonglong2(((uint64_t)lo16 << 48) | (((uint64_t)song & kM24) << 24) | ((uint64_t)o & kM24), hi);
  }
#pragma unroll
  for (int d = 16; d; d >>= 1) smax = max(smax, __shfl_xor_sync(0xffffffffu, smax, d));
  if ((threadIdx.x & 31) == 0 && smax > 0) atomicMax(status + 1, smax);
}

__global__ void __launch_bounds__(256)
unpack_rows_kernel(const ulonglong2 *__restrict__ rows, int64_t n, uint8_t *__restrict__ hash, int32_t *__restrict__ song,
                   int32_t *__restrict__ off) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const ulonglong2 r = rows[i];
    uint16_t *h = reinterpret_cast<uint16_t *>(hash + i * SIA_HASH_BYTES);
    const uint32_t w[5] = {(uint32_t)(r.y >> 48) & 0xffffu, (uint32_t)(r.y >> 32) & 0xffffu, (uint32_t)(r.y >> 16) & 0xffffu,
                           (uint32_t)r.y & 0xffffu, (uint32_t)(r.x >> 48) & 0xffffu};
#pragma unroll
    for (int k = 0; k < 5; ++k) h[k] = (uint16_t)(((w[k] & 0xff) << 8) | (w[k] >> 8));   // big-endian digest bytes
    song[i] = (int32_t)((r.x >> 24) & 0xffffffu);
    off[i] = (int32_t)(r.x & 0xffffffu);
  }
}

__device__ __forceinline__ bool rec_eq(const ulonglong2 &a, const ulonglong2 &b) { return a.x == b.x && a.y == b.y; }

__global__ void __launch_bounds__(256)
uniq_flag_kernel(const ulonglong2 *__restrict__ r, int64_t n, uint32_t *__restrict__ flag) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    flag[i] = (i == 0 || !rec_eq(r[i], r[i - 1])) ? 1u : 0u;
}

__global__ void __launch_bounds__(256)
keep_flag_kernel(const ulonglong2 *__restrict__ r, int64_t n, const uint32_t *__restrict__ dead_bitmap,
                 uint32_t *__restrict__ flag) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t song = (uint32_t)(r[i].x >> 24) & 0xffffffu;
    flag[i] = (dead_bitmap[song >> 5] >> (song & 31)) & 1u ? 0u : 1u;
  }
}

__global__ void __launch_bounds__(256)
uniq_compact_kernel(const ulonglong2 *__restrict__ r, int64_t n, const uint32_t *__restrict__ flag,
                    const int64_t *__restrict__ pos, ulonglong2 *__restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    if (flag[i]) out[pos[i]] = r[i];
}

__global__ void __launch_bounds__(256)
build_dir_kernel(const ulonglong2 *__restrict__ r, int64_t n, int bits, uint32_t *__restrict__ dir) {
  const int64_t nb = 1ll << bits;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b_prev = i == 0 ? -1 : (int64_t)(r[i - 1].y >> (64 - bits));
    const int64_t b = i == n ? nb : (int64_t)(r[i].y >> (64 - bits));
    for (int64_t k = b_prev + 1; k <= b; ++k) dir[k] = (uint32_t)i;
  }
}

// first row index in [lo, hi) whose (digest) >= (khi, klo16)
__device__ __forceinline__ uint32_t lower_bound_key(const ulonglong2 *__restrict__ r, uint32_t lo, uint32_t hi, uint64_t khi,
                                                    uint32_t klo16) {
  while (lo < hi) {
    const uint32_t mid = lo + ((hi - lo) >> 1);
    const ulonglong2 v = r[mid];
    const bool less = v.y < khi || (v.y == khi && (uint32_t)(v.x >> 48) < klo16);
    if (less) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__device__ __forceinline__ void posting_range(const ulonglong2 *__restrict__ rows, const uint32_t *__restrict__ dir, int bits,
                                              uint64_t khi, uint32_t klo16, uint32_t &first, uint32_t &count) {
  const uint64_t b = khi >> (64 - bits);
  const uint32_t lo = dir[b], hi = dir[b + 1];
  first = lower_bound_key(rows, lo, hi, khi, klo16);
  uint32_t e = first;
  auto match = [&](uint32_t i) { const ulonglong2 v = rows[i]; return v.y == khi && (uint32_t)(v.x >> 48) == klo16; };
  if (e < hi && match(e)) {
    // runs are short: gallop to bracket the end of the equal run, then binary search it
    uint32_t good = e, bad = hi, step = 1;       // rows[good] matches; rows[bad] does not (or bad == hi)
    for (;;) {
      const uint32_t nxt = good + step;
      if (nxt >= hi) break;
      if (!match(nxt)) { bad = nxt; break; }
      good = nxt;
      step <<= 1;
    }
    uint32_t a = good + 1, c = bad;
    while (a < c) {
      const uint32_t mid = a + ((c - a) >> 1);
      if (match(mid)) a = mid + 1; else c = mid;
    }
    e = a;
  }
  count = e - first;
}

// query entry record, sorted by all 16 bytes = (qid, digest, qoff):
//   .y = qid (24) | digest bits 79..40 (40)      .x = digest bits 39..16 (24) | digest lo16 (16) | qoff (24)
__global__ void __launch_bounds__(256)
pack_queries_kernel(const uint8_t *__restrict__ hash, const int32_t *__restrict__ qoff, const int32_t *__restrict__ qid_arr,
                    const int64_t *__restrict__ query_starts, int n_queries, int qid_base, int64_t i0, int64_t n,
                    ulonglong2 *__restrict__ out, int32_t *__restrict__ status) {
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = i0 + k;
    uint64_t hi; uint32_t lo16;
    load_digest(hash + i * SIA_HASH_BYTES, hi, lo16);
    const int32_t q = qid_arr ? qid_arr[i] - qid_base : find_segment(query_starts, n_queries, i) - qid_base;
    const int32_t o = qoff[i];
    if (o < 0 || o > (int32_t)kM24 || q < 0 || q >= (1 << kQidBits)) atomicOr(status, 2);
    out[k] = make_ulonglong2(((hi & kM24) << 40) | ((uint64_t)lo16 << 24) | ((uint64_t)o & kM24),
                             ((uint64_t)q << 40) | (hi >> 24));
  }
}

__device__ __forceinline__ void entry_key(const ulonglong2 &e, uint64_t &khi, uint32_t &klo16, uint32_t &qid, uint32_t &qoff) {
  qid = (uint32_t)(e.y >> 40);
  khi = (e.y << 24) | (e.x >> 40);
  klo16 = (uint32_t)(e.x >> 24) & 0xffffu;
  qoff = (uint32_t)(e.x & kM24);
}

__global__ void __launch_bounds__(256)
lookup_kernel(const ulonglong2 *__restrict__ ent, int64_t n, const ulonglong2 *__restrict__ rows,
              const uint32_t *__restrict__ dir, int bits, int64_t n_rows, uint32_t *__restrict__ first,
              uint32_t *__restrict__ cnt_all, uint32_t *__restrict__ cnt_head) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const ulonglong2 e = ent[i];
    bool dup = false, head = true;
    if (i > 0) {
      const ulonglong2 p = ent[i - 1];
      dup = rec_eq(e, p);
      head = !(p.y == e.y && (p.x >> 24) == (e.x >> 24));   // same (qid, digest) as the previous entry?
    }
    uint32_t f = 0, c = 0;
    if (!dup && n_rows > 0) {
      uint64_t khi; uint32_t klo16, qid, qoff;
      entry_key(e, khi, klo16, qid, qoff);
      posting_range(rows, dir, bits, khi, klo16, f, c);
    }
    first[i] = f;
    cnt_all[i] = c;
    cnt_head[i] = head ? c : 0u;
  }
}

// Expand posting runs into vote keys.  One block handles 256 consecutive entries and spreads
// their postings evenly over its threads (binary search over the block-local offsets).
template <bool HEADS>
__global__ void __launch_bounds__(256)
expand_kernel(const ulonglong2 *__restrict__ ent, int64_t e0, int64_t n, const uint32_t *__restrict__ first,
              const int64_t *__restrict__ off, const ulonglong2 *__restrict__ rows, uint64_t *__restrict__ out,
              int64_t out_base) {
  __shared__ int64_t s_off[257];
  const int64_t b0 = e0 + (int64_t)blockIdx.x * 256;
  const int64_t i = b0 + threadIdx.x;
  const int nloc = (int)min((int64_t)256, e0 + n - b0);
  if (threadIdx.x < nloc) s_off[threadIdx.x] = off[i];
  if (threadIdx.x == 0) s_off[nloc] = off[b0 + nloc];
  __syncthreads();
  const int64_t base = s_off[0], total = s_off[nloc] - base;
  for (int64_t j = threadIdx.x; j < total; j += 256) {
    int lo = 0, hi = nloc;               // largest e with s_off[e] - base <= j
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (s_off[mid] - base <= j) lo = mid; else hi = mid; }
    const int64_t ei = b0 + lo;
    const uint32_t k = (uint32_t)(j - (s_off[lo] - base));
    const ulonglong2 e = ent[ei];
    const ulonglong2 r = rows[first[ei] + k];
    const uint64_t qid = e.y >> 40;
    const uint64_t song = (r.x >> 24) & kM24;
    uint64_t key = (qid << (kSongBits + kDiffBits)) | (song << kDiffBits);
    if (!HEADS) {
      const int32_t diff = (int32_t)(r.x & kM24) - (int32_t)(e.x & kM24);   // db offset - query offset
      key |= (uint64_t)(uint32_t)(diff + SIA_DIFF_BIAS);
    }
    out[base - out_base + j] = key;
  }
}

// run-length / weighted-run reduction of sorted keys: flag run heads, scan, then accumulate
__global__ void __launch_bounds__(256)
run_flag_kernel(const uint64_t *__restrict__ key, int64_t n, int shift, uint32_t *__restrict__ flag) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    flag[i] = (i == 0 || (key[i] >> shift) != (key[i - 1] >> shift)) ? 1u : 0u;
}

// pos[] is the exclusive scan of flag[]: element i belongs to run pos[i+1]-1
__global__ void __launch_bounds__(256)
run_reduce_kernel(const uint64_t *__restrict__ key, const int32_t *__restrict__ weight, int64_t n,
                  const uint32_t *__restrict__ flag, const int64_t *__restrict__ pos, uint64_t *__restrict__ bin_key,
                  int32_t *__restrict__ bin_count) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = pos[i + 1] - 1;
    if (flag[i]) bin_key[b] = key[i];
    atomicAdd(&bin_count[b], weight ? weight[i] : 1);
  }
}

// per (query, song) segment: best = max over its bins of (count << 25 | inverted diff) -> largest count,
// smallest offset difference on ties (Python's max() keeps the first maximum, recognizer.py:308)
__global__ void __launch_bounds__(256)
seg_best_kernel(const uint64_t *__restrict__ bin_key, const int32_t *__restrict__ bin_count, int64_t nbins,
                const uint32_t *__restrict__ flag, const int64_t *__restrict__ pos, uint64_t *__restrict__ seg_key,
                unsigned long long *__restrict__ seg_best) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nbins; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t s = pos[i + 1] - 1;
    const uint64_t k = bin_key[i];
    if (flag[i]) seg_key[s] = k >> kDiffBits;
    const uint64_t inv = ((1ull << kDiffBits) - 1) - (k & ((1ull << kDiffBits) - 1));
    atomicMax(&seg_best[s], ((unsigned long long)(uint32_t)bin_count[i] << kDiffBits) | inv);
  }
}

__device__ __forceinline__ int64_t lower_bound_u64(const uint64_t *__restrict__ a, int64_t n, uint64_t v) {
  int64_t lo = 0, hi = n;
  while (lo < hi) { const int64_t mid = lo + ((hi - lo) >> 1); if (a[mid] < v) lo = mid + 1; else hi = mid; }
  return lo;
}

// one block per query: top-n segments by (count desc, song asc)
__global__ void __launch_bounds__(256)
topn_kernel(const uint64_t *__restrict__ seg_key, const unsigned long long *__restrict__ seg_best, int64_t nseg,
            const uint64_t *__restrict__ row_key, const int32_t *__restrict__ row_count, int64_t nrowbins, int q_lo,
            int qid_base, int topn, int32_t *__restrict__ out_song, int32_t *__restrict__ out_diff, int32_t *__restrict__ out_count,
            int32_t *__restrict__ out_rows, int32_t *__restrict__ out_nres) {
  __shared__ unsigned long long s_best[256];
  __shared__ int64_t s_idx[256];
  __shared__ unsigned long long s_prev;
  const uint64_t q = (uint64_t)blockIdx.x + q_lo;
  const int64_t s0 = lower_bound_u64(seg_key, nseg, q << kSongBits);
  const int64_t s1 = lower_bound_u64(seg_key, nseg, (q + 1) << kSongBits);
  if (threadIdx.x == 0) s_prev = ~0ull;
  __syncthreads();
  int nres = 0;
  for (int r = 0; r < topn; ++r) {
    const unsigned long long prev = s_prev;
    unsigned long long best = 0;
    int64_t bi = -1;
    for (int64_t s = s0 + threadIdx.x; s < s1; s += 256) {
      // rank key: count (high) then inverted song id, so equal counts order by ascending song id
      const unsigned long long c = seg_best[s] >> kDiffBits;
      const unsigned long long k = (c << kSongBits) | (kM24 - (seg_key[s] & kM24));
      if (k < prev && (bi < 0 || k > best)) { best = k; bi = s; }
    }
    s_best[threadIdx.x] = best; s_idx[threadIdx.x] = bi;
    __syncthreads();
    for (int d = 128; d; d >>= 1) {
      if (threadIdx.x < d) {
        const int64_t oi = s_idx[threadIdx.x + d];
        if (oi >= 0 && (s_idx[threadIdx.x] < 0 || s_best[threadIdx.x + d] > s_best[threadIdx.x])) {
          s_best[threadIdx.x] = s_best[threadIdx.x + d]; s_idx[threadIdx.x] = oi;
        }
      }
      __syncthreads();
    }
    const int64_t w = s_idx[0];
    if (w < 0) break;                       // uniform: every thread reads the same shared value
    if (threadIdx.x == 0) {
      const uint64_t sk = seg_key[w];
      const unsigned long long v = seg_best[w];
      const int64_t o = ((int64_t)q + qid_base) * topn + r;
      out_song[o] = (int32_t)(sk & kM24);
      out_count[o] = (int32_t)(v >> kDiffBits);
      out_diff[o] = (int32_t)(((1ull << kDiffBits) - 1) - (v & ((1ull << kDiffBits) - 1))) - SIA_DIFF_BIAS;
      const int64_t rb = lower_bound_u64(row_key, nrowbins, sk << kDiffBits);
      out_rows[o] = (rb < nrowbins && row_key[rb] == (sk << kDiffBits)) ? row_count[rb] : 0;
      s_prev = s_best[0];
    }
    ++nres;
    __syncthreads();
  }
  if (threadIdx.x == 0) out_nres[q + qid_base] = nres;
}

__global__ void split_bins_kernel(const ulonglong2 *__restrict__ rec, int64_t n, uint64_t *__restrict__ key,
                                  int32_t *__restrict__ w) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    key[i] = rec[i].y; w[i] = (int32_t)rec[i].x;
  }
}
__global__ void join_bins_kernel(const uint64_t *__restrict__ key, const int32_t *__restrict__ w, int64_t n,
                                 ulonglong2 *__restrict__ rec) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    rec[i] = make_ulonglong2((uint64_t)(uint32_t)w[i], key[i]);
}

__global__ void select_rows_kernel(const ulonglong2 *__restrict__ ent, int64_t n, const uint32_t *__restrict__ first,
                                   const int64_t *__restrict__ off, const ulonglong2 *__restrict__ rows,
                                   int32_t *__restrict__ o_idx, int32_t *__restrict__ o_song, int32_t *__restrict__ o_off,
                                   int64_t cap) {
  // entries here are packed with qid = position of the hash in the caller's list
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = off[i], c = off[i + 1] - b;
    for (int64_t k = 0; k < c; ++k) {
      if (b + k >= cap) break;
      const ulonglong2 r = rows[first[i] + k];
      o_idx[b + k] = (int32_t)(ent[i].x & kM24);
      o_song[b + k] = (int32_t)((r.x >> 24) & kM24);
      o_off[b + k] = (int32_t)(r.x & kM24);
    }
  }
}

__global__ void gather_offsets_kernel(const int64_t *__restrict__ off_all, const int64_t *__restrict__ off_head,
                                      const int64_t *__restrict__ query_starts, int64_t i0, int nq,
                                      int64_t *__restrict__ out) {
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q <= nq; q += gridDim.x * blockDim.x) {
    const int64_t e = query_starts[q] - i0;
    out[q] = off_all[e];
    out[nq + 1 + q] = off_head[e];
  }
}

// first sorted entry of every query id (entries are sorted by query id first) -> tuple / row offsets per query
__global__ void query_starts_kernel(const ulonglong2 *__restrict__ ent, int64_t n, const int64_t *__restrict__ off_all,
                                    const int64_t *__restrict__ off_head, int nq, int64_t *__restrict__ tuple_starts,
                                    int64_t *__restrict__ row_starts) {
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q <= nq; q += gridDim.x * blockDim.x) {
    int64_t lo = 0, hi = n;
    while (lo < hi) { const int64_t mid = lo + ((hi - lo) >> 1); if ((int64_t)(ent[mid].y >> 40) < q) lo = mid + 1; else hi = mid; }
    tuple_starts[q] = off_all[lo];
    row_starts[q] = off_head[lo];
  }
}


// ---- hash-table vote -------------------------------------------------------------------------
// The sort-free vote behind sia_index_query_batch.  Queries are taken in groups whose tables stay in L2.
// Every query owns two sub-tables inside the group's (zero-initialised) tables:
//   bins : open addressing, linear probing inside the query's slot range; one 64-bit word per slot =
//          key (song 24 | biased diff 25) << 15 | count — the run-length bins of recognizer.py:303-305.
//          A bin holds at most one tuple per query entry, so 15 bits suffice for queries of <= 32767
//          (hash, offset) pairs; larger queries go through the sort-based vote.  key != 0 always
//          (the biased diff is >= 1), so 0 is the empty slot.
//   songs: dense (slot = song id) when the id range is small next to the query's matches, else open
//          addressing on song id; per slot rows (dedup_hashes[song], recognizer.py:259-264) and
//          best = max over the song's bins of (count << 25 | inverted diff) (recognizer.py:308), which is
//          kept current while the bins fill: the thread that brings a bin to count c posts (c, diff).
// Posting runs are expanded straight into the tables — no vote keys are written or sorted.
struct QMeta {
  int64_t bin_base, song_base;    // first slot of the query's sub-tables, relative to the group's tables
  int64_t filt_base;              // first word of the query's duplicate filter (filt_words == 0: no filter)
  uint32_t bin_cap, song_cap, filt_words, pad_;
};
constexpr int kBinCountBits = 15;
constexpr int64_t kMaxHashVoteEntries = (1 << kBinCountBits) - 1;
constexpr int kTopK = 4;          // results extracted per scan of a query's song table
constexpr int kVoteTuplesDefault = 4096; // vote tuples per block of expand_vote_kernel (512 per warp)

__device__ __forceinline__ uint32_t mix32(uint32_t k) {
  k ^= k >> 16; k *= 0x85ebca6bu; k ^= k >> 13; k *= 0xc2b2ae35u; k ^= k >> 16;
  return k;
}
__device__ __forceinline__ uint32_t slot_of(uint32_t h, uint32_t cap) { return (uint32_t)(((uint64_t)h * cap) >> 32); }

// song slot of `song` in query m's song table (dense: the id itself; else find-or-insert by open addressing)
template <bool DENSE>
__device__ __forceinline__ int64_t song_slot(const QMeta &m, uint32_t song, uint32_t *__restrict__ song_key) {
  if (DENSE) return m.song_base + song;
  uint32_t t = slot_of(mix32(song), m.song_cap);
  for (uint32_t probes = 0; probes < m.song_cap; ++probes) {     // bounded: a full table (inconsistent inputs) cannot hang
    const uint32_t old = atomicCAS(&song_key[m.song_base + t], 0u, song + 1u);
    if (old == 0u || old == song + 1u) return m.song_base + t;
    if (++t == m.song_cap) t = 0;
  }
  return -1;
}

// count one vote tuple (song, biased diff) in query m's bin table; returns the bin's count including this tuple
// (0 and flag 8 in *overflow if the table is full, which the sizing rules out unless the duplicate filter
// under-estimated: then the group is redone without the filter)
__device__ __forceinline__ unsigned long long bin_count(const QMeta &m, uint32_t song, uint32_t dbits,
                                                        unsigned long long *__restrict__ bins, uint32_t &fresh,
                                                        int32_t *__restrict__ overflow) {
  const unsigned long long key = ((unsigned long long)song << kDiffBits) | dbits;
  uint32_t s = slot_of(mix32(song * 0x9e3779b1u + dbits), m.bin_cap);
  for (uint32_t probes = 0; probes < m.bin_cap; ++probes) {
    unsigned long long *p = bins + m.bin_base + s;
    const unsigned long long old = atomicCAS(p, 0ull, (key << kBinCountBits) | 1ull);
    if (old == 0ull) { ++fresh; return 1; }
    if ((old >> kBinCountBits) == key) {
      const unsigned long long count = (atomicAdd(p, 1ull) & ((1ull << kBinCountBits) - 1)) + 1;
      if (overflow && count == (1ull << kBinCountBits)) atomicOr(overflow, 16);  // the count field wrapped: redo by sorting
      return count;
    }
    if (++s == m.bin_cap) s = 0;
  }
  if (overflow) atomicOr(overflow, 8);
  return 0;
}

// the song side of a vote tuple whose bin now holds `count`: keep the song's best (count, smallest diff) current
template <bool DENSE>
__device__ __forceinline__ int64_t song_update(const QMeta &m, uint32_t song, uint32_t dbits, unsigned long long count,
                                               uint32_t *__restrict__ song_key, unsigned long long *__restrict__ song_best,
                                               int32_t *__restrict__ overflow) {
  const int64_t ss = song_slot<DENSE>(m, song, song_key);
  if (ss < 0) { if (overflow) atomicOr(overflow, 8); return -1; }
  const unsigned long long val = (count << kDiffBits) | (((1ull << kDiffBits) - 1) - dbits);
  // best only grows: a (possibly stale) read that already covers val makes the atomic unnecessary — most of a
  // song's count-1 bins lose against the first one that was posted
  if (__ldcg(&song_best[ss]) < val) atomicMax(&song_best[ss], val);
  return ss;
}

template <bool DENSE>
__device__ __forceinline__ int64_t vote_insert(const QMeta &m, uint32_t song, uint32_t dbits,
                                               unsigned long long *__restrict__ bins, uint32_t *__restrict__ song_key,
                                               unsigned long long *__restrict__ song_best, uint32_t &fresh,
                                               int32_t *__restrict__ overflow) {
  const unsigned long long count = bin_count(m, song, dbits, bins, fresh, overflow);
  return count ? song_update<DENSE>(m, song, dbits, count, song_key, song_best, overflow) : -1;
}

// duplicate filter: 2 bits per bucket (seen, seen twice), 16 buckets per 32-bit word, ~16 buckets per tuple
__device__ __forceinline__ uint32_t *filter_word(const QMeta &m, uint32_t song, uint32_t dbits, uint32_t *__restrict__ filter,
                                                 uint32_t &seen_bit) {
  const uint32_t h = mix32(song * 0x85ebca6bu ^ (dbits * 0xc2b2ae35u + 0x27d4eb2fu));
  seen_bit = 1u << (2 * (h & 15u));
  return filter + m.filt_base + slot_of(h, m.filt_words);
}

// One block handles tuples_per_block consecutive vote tuples (postings of the entries [e0, e0+n), numbered by the
// exclusive scan off[]), whatever entries they belong to: a heavy key's run is shared by many blocks.  Each warp
// takes an eighth of the block's tuples and walks the entries they belong to: everything that depends on the
// entry (query tables, query offset, head flag, first posting) is warp-uniform and loaded once per entry, the
// lanes then take the entry's postings 32 at a time.
// MODE 0: every tuple is counted in the bin table.  MODE 1 / 2, the two-pass form that keeps the tables in L2:
// pass 1 only marks each tuple's bucket in the query's duplicate filter (seen / seen twice); pass 2 counts a tuple in
// the (then 4x smaller) bin table only if its bucket was seen twice — a tuple alone in its bucket is a bin of count 1.
template <bool DENSE, int MODE>
__global__ void __launch_bounds__(256)
expand_vote_kernel(const ulonglong2 *__restrict__ ent, int64_t e0, int64_t n, const uint32_t *__restrict__ first,
                   const int64_t *__restrict__ off, const uint32_t *__restrict__ cnt_head,
                   const ulonglong2 *__restrict__ rows, const QMeta *__restrict__ meta,
                   unsigned long long *__restrict__ bins, uint32_t *__restrict__ song_key,
                   uint32_t *__restrict__ song_rows, unsigned long long *__restrict__ song_best,
                   uint32_t *__restrict__ filter, unsigned long long *__restrict__ n_bins, int tuples_per_block,
                   int32_t *__restrict__ overflow) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int per_warp = tuples_per_block >> 3;
  const int64_t j_lo = off[e0] + (int64_t)blockIdx.x * tuples_per_block + (int64_t)warp * per_warp;
  const int64_t j_hi = min(off[e0 + n], j_lo + per_warp);
  uint32_t fresh = 0;
  if (j_lo < j_hi) {
    int64_t ei = e0, hi = e0 + n;          // largest entry in [e0, e0+n) with off[ei] <= j_lo (warp-uniform search)
    while (hi - ei > 1) { const int64_t mid = ei + ((hi - ei) >> 1); if (off[mid] <= j_lo) ei = mid; else hi = mid; }
    for (; ei < e0 + n; ++ei) {
      const int64_t o_this = off[ei], o_next = off[ei + 1];
      if (o_this >= j_hi) break;
      if (o_next == o_this) continue;                              // a hash without postings (or a duplicate pair)
      const ulonglong2 e = ent[ei];
      const QMeta m = meta[e.y >> 40];
      const uint32_t qoff = (uint32_t)(e.x & kM24);
      const bool head = cnt_head[ei] != 0;                         // first entry of its (query, hash): rows count once
      const ulonglong2 *__restrict__ run = rows + first[ei];
      const uint32_t k_hi = (uint32_t)(min(j_hi, o_next) - o_this);
      const uint32_t k_lo = (uint32_t)(max(j_lo, o_this) - o_this);
      ulonglong2 r_next = make_ulonglong2(0, 0);
      if (k_lo + lane < k_hi) r_next = run[k_lo + lane];
      for (uint32_t kw = k_lo; kw < k_hi; kw += 32) {
        __syncwarp();                                              // the probe loops below diverge
        const uint32_t k = kw + lane;
        const ulonglong2 r = r_next;
        if (k + 32 < k_hi) r_next = run[k + 32];                   // the next step's posting is in flight during this one
        if (k >= k_hi) continue;
        const uint32_t song = (uint32_t)(r.x >> 24) & 0xffffffu;
        const uint32_t dbits = (uint32_t)(r.x & kM24) - qoff + SIA_DIFF_BIAS;    // db offset - query offset, biased
        if (MODE == 1) {
          uint32_t seen;
          uint32_t *w = filter_word(m, song, dbits, filter, seen);
          const uint32_t old = atomicOr(w, seen);
          if ((old & seen) && !(old & (seen << 1))) atomicOr(w, seen << 1);
          continue;
        }
        unsigned long long count = 1;
        unsigned long long seen_best = 0;                          // a read of the song's best issued before the bin work
        if (MODE == 2) {
          uint32_t seen;
          const uint32_t *w = filter_word(m, song, dbits, filter, seen);
          const uint32_t fw = __ldcg(w);
          if (DENSE) seen_best = __ldcg(&song_best[m.song_base + song]);   // both reads depend on the posting only
          if (fw & (seen << 1)) count = bin_count(m, song, dbits, bins, fresh, overflow);
          else ++fresh;                                             // alone in its bucket: a bin of its own
        } else {
          count = bin_count(m, song, dbits, bins, fresh, overflow);
        }
        int64_t ss = -1;
        if (count) {
          if (MODE == 2 && DENSE) {
            ss = m.song_base + song;
            const unsigned long long val = (count << kDiffBits) | (((1ull << kDiffBits) - 1) - dbits);
            if (seen_best < val) atomicMax(&song_best[ss], val);    // best only grows: a stale read that covers val is enough
          } else {
            ss = song_update<DENSE>(m, song, dbits, count, song_key, song_best, overflow);
          }
        }
        if (ss >= 0 && head) atomicAdd(&song_rows[ss], 1u);
      }
      __syncwarp();
    }
  }
  if (n_bins && MODE != 1) {
#pragma unroll
    for (int d = 16; d; d >>= 1) fresh += __shfl_xor_sync(0xffffffffu, fresh, d);
    if (lane == 0 && fresh) atomicAdd(n_bins, (unsigned long long)fresh);
  }
}

// ---- hash-table vote of exchanged vote keys (sia_vote_tuples) --------------------------------
// keys: query (15) | song (24) | biased diff (25); row keys carry no diff.  Per-query sizes are counted on the
// device, the table layout (QMeta) is built there too; only the largest song id travels to the host.
__global__ void __launch_bounds__(256)
count_keys_kernel(const uint64_t *__restrict__ key, int64_t n, int nq, uint32_t *__restrict__ cnt, int32_t *__restrict__ info) {
  int32_t smax = 0;
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x; i0 < n; i0 += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = i0 + threadIdx.x;
    const bool valid = i < n;
    const uint64_t k = valid ? key[i] : 0;
    const uint32_t q = (uint32_t)(k >> (kSongBits + kDiffBits));
    if (valid && q >= (uint32_t)nq) atomicOr(info + 1, 2);
    smax = max(smax, (int32_t)((k >> kDiffBits) & kM24));
    const uint32_t active = __ballot_sync(0xffffffffu, valid && q < (uint32_t)nq);
    if (valid && q < (uint32_t)nq) {       // keys arrive grouped by query: one atomic per run inside the warp
      const uint32_t peers = __match_any_sync(active, q);
      if ((peers & ((1u << (threadIdx.x & 31)) - 1u)) == 0) atomicAdd(&cnt[q], (uint32_t)__popc(peers));
    }
  }
#pragma unroll
  for (int d = 16; d; d >>= 1) smax = max(smax, __shfl_xor_sync(0xffffffffu, smax, d));
  if ((threadIdx.x & 31) == 0 && smax > 0) atomicMax(info, smax);
}

// single block: exclusive scans of the per-query table sizes -> QMeta
__global__ void __launch_bounds__(1024)
build_meta_kernel(const uint32_t *__restrict__ cnt_t, const uint32_t *__restrict__ cnt_r, int nq, int mult, int64_t dense_span,
                  QMeta *__restrict__ meta, int32_t *__restrict__ info) {
  __shared__ int64_t s_b[1024], s_s[1024];
  const int per = (nq + 1023) / 1024;
  const int a = min(nq, (int)threadIdx.x * per), b = min(nq, a + per);
  int64_t sb = 0, ss = 0;
  for (int q = a; q < b; ++q) {
    if ((uint64_t)mult * cnt_t[q] + 32 > 0xffffffffull) atomicOr(info + 1, 4);
    sb += (int64_t)mult * cnt_t[q] + 32;
    ss += dense_span > 0 ? dense_span : 2 * (int64_t)cnt_r[q] + 32;
  }
  s_b[threadIdx.x] = sb; s_s[threadIdx.x] = ss;
  __syncthreads();
  if (threadIdx.x == 0) {
    int64_t rb = 0, rs = 0;
    for (int i = 0; i < 1024; ++i) { const int64_t x = s_b[i], y = s_s[i]; s_b[i] = rb; s_s[i] = rs; rb += x; rs += y; }
  }
  __syncthreads();
  sb = s_b[threadIdx.x]; ss = s_s[threadIdx.x];
  for (int q = a; q < b; ++q) {
    QMeta m;
    m.bin_base = sb; m.song_base = ss; m.filt_base = 0; m.filt_words = 0; m.pad_ = 0;
    m.bin_cap = (uint32_t)((int64_t)mult * cnt_t[q] + 32);
    m.song_cap = dense_span > 0 ? (uint32_t)dense_span : 2 * cnt_r[q] + 32;
    meta[q] = m;
    sb += m.bin_cap; ss += m.song_cap;
  }
}

template <bool DENSE>
__global__ void __launch_bounds__(256)
vote_keys_kernel(const uint64_t *__restrict__ key, int64_t n, const QMeta *__restrict__ meta,
                 unsigned long long *__restrict__ bins, uint32_t *__restrict__ song_key,
                 unsigned long long *__restrict__ song_best, int32_t *__restrict__ info) {
  uint32_t fresh = 0;
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x; i0 < n; i0 += (int64_t)gridDim.x * blockDim.x) {
    __syncwarp();
    const int64_t i = i0 + threadIdx.x;
    if (i >= n) continue;
    const uint64_t k = key[i];
    const QMeta m = meta[k >> (kSongBits + kDiffBits)];
    vote_insert<DENSE>(m, (uint32_t)(k >> kDiffBits) & 0xffffffu, (uint32_t)k & ((1u << kDiffBits) - 1u), bins, song_key,
                       song_best, fresh, info + 1);
  }
}

template <bool DENSE>
__global__ void __launch_bounds__(256)
row_keys_kernel(const uint64_t *__restrict__ key, int64_t n, const QMeta *__restrict__ meta,
                uint32_t *__restrict__ song_key, uint32_t *__restrict__ song_rows, int32_t *__restrict__ info) {
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x; i0 < n; i0 += (int64_t)gridDim.x * blockDim.x) {
    __syncwarp();
    const int64_t i = i0 + threadIdx.x;
    if (i >= n) continue;
    const uint64_t k = key[i];
    const QMeta m = meta[k >> (kSongBits + kDiffBits)];
    const int64_t ss = song_slot<DENSE>(m, (uint32_t)(k >> kDiffBits) & 0xffffffu, song_key);
    if (ss >= 0) atomicAdd(&song_rows[ss], 1u);
    else atomicOr(info + 1, 8);
  }
}

// one block per query: top-n songs by (count desc, song asc), kTopK results per scan of the query's song table
template <bool DENSE>
__global__ void __launch_bounds__(256)
topn_hash_kernel(const uint32_t *__restrict__ song_key, const uint32_t *__restrict__ song_rows,
                 const unsigned long long *__restrict__ song_best, const QMeta *__restrict__ meta, int q_lo, int qid_base,
                 int topn, int32_t *__restrict__ out_song, int32_t *__restrict__ out_diff, int32_t *__restrict__ out_count,
                 int32_t *__restrict__ out_rows, int32_t *__restrict__ out_nres) {
  __shared__ unsigned long long s_key[8];
  __shared__ uint32_t s_slot[8];
  __shared__ unsigned long long s_win;
  const int q = (int)blockIdx.x + q_lo;
  const QMeta m = meta[q];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned long long prev = ~0ull;
  int nres = 0;
  while (nres < topn) {
    // rank key: count (high) then inverted song id, so equal counts order by ascending song id
    unsigned long long tk[kTopK];
    uint32_t ts[kTopK];
#pragma unroll
    for (int i = 0; i < kTopK; ++i) { tk[i] = 0; ts[i] = 0; }
    for (uint32_t s = threadIdx.x; s < m.song_cap; s += 256) {
      const unsigned long long best = song_best[m.song_base + s];
      if (best == 0ull) continue;          // empty slot / song without a match
      const uint32_t song = DENSE ? s : song_key[m.song_base + s] - 1u;
      unsigned long long k = ((best >> kDiffBits) << kSongBits) | (kM24 - song);
      if (k >= prev || k <= tk[kTopK - 1]) continue;
      uint32_t sl = s;
#pragma unroll
      for (int i = 0; i < kTopK; ++i)
        if (k > tk[i]) { const unsigned long long a = tk[i]; const uint32_t b = ts[i]; tk[i] = k; ts[i] = sl; k = a; sl = b; }
    }
    int got = 0;
    for (int r = 0; r < kTopK && nres < topn; ++r) {
      unsigned long long bk = tk[0];
      uint32_t bs = ts[0];
#pragma unroll
      for (int d = 16; d; d >>= 1) {
        const unsigned long long ok = __shfl_xor_sync(0xffffffffu, bk, d);
        const uint32_t os = __shfl_xor_sync(0xffffffffu, bs, d);
        if (ok > bk) { bk = ok; bs = os; }
      }
      if (lane == 0) { s_key[warp] = bk; s_slot[warp] = bs; }
      __syncthreads();
      if (threadIdx.x == 0) {
        unsigned long long w = 0; uint32_t ws = 0;
        for (int i = 0; i < 8; ++i) if (s_key[i] > w) { w = s_key[i]; ws = s_slot[i]; }
        s_win = w;
        if (w) {
          const unsigned long long v = song_best[m.song_base + ws];
          const int64_t o = ((int64_t)q + qid_base) * topn + nres;
          out_song[o] = (int32_t)(kM24 - (w & kM24));
          out_count[o] = (int32_t)(v >> kDiffBits);
          out_diff[o] = (int32_t)(((1ull << kDiffBits) - 1) - (v & ((1ull << kDiffBits) - 1))) - SIA_DIFF_BIAS;
          out_rows[o] = (int32_t)song_rows[m.song_base + ws];
        }
      }
      __syncthreads();
      const unsigned long long w = s_win;
      if (w == 0) break;
      if (tk[0] == w) {                   // the winner leaves its owner's list
#pragma unroll
        for (int i = 0; i + 1 < kTopK; ++i) { tk[i] = tk[i + 1]; ts[i] = ts[i + 1]; }
        tk[kTopK - 1] = 0;
      }
      prev = w;
      ++nres; ++got;
      __syncthreads();
    }
    if (got < kTopK) break;               // the table is exhausted
  }
  if (threadIdx.x == 0) {
    out_nres[q + qid_base] = nres;
    for (int r = nres; r < topn; ++r) {       // unused slots read as zero whatever ran before (a redone group)
      const int64_t o = ((int64_t)q + qid_base) * topn + r;
      out_song[o] = 0; out_diff[o] = 0; out_count[o] = 0; out_rows[o] = 0;
    }
  }
}

inline unsigned grid_for(int64_t n, int threads = 256) {
  int64_t b = ceil_div(n > 0 ? n : 1, threads);
  return (unsigned)std::min<int64_t>(b, kNumSMs * 32);
}

int set_device(const sia_index *ix) {
  SIA_CUDA(cudaSetDevice(ix->device));
  return SIA_OK;
}

// sorted unique keys + counts from a sorted key array (optionally weighted).  Returns bins on the device.
int reduce_runs(Arena &ar, const uint64_t *d_key, const int32_t *d_weight, int64_t n, uint64_t **bin_key,
                int32_t **bin_count, int64_t *nbins, cudaStream_t s) {
  *nbins = 0; *bin_key = nullptr; *bin_count = nullptr;
  if (n == 0) return SIA_OK;
  uint32_t *flag = ar.take<uint32_t>(n);
  int64_t *pos = ar.take<int64_t>(n + 1);
  void *stmp = ar.take<char>(scan_tmp_bytes(n));
  SIA_REQUIRE(flag && pos && stmp, SIA_E_NOMEM, "index scratch arena too small (reduce_runs)");
  run_flag_kernel<<<grid_for(n), 256, 0, s>>>(d_key, n, 0, flag);
  SIA_CHECK_LAUNCH();
  int rc = exclusive_scan_u32(flag, pos, n, stmp, s);
  if (rc) return rc;
  int64_t nb = 0;
  SIA_CUDA(cudaMemcpyAsync(&nb, pos + n, sizeof nb, cudaMemcpyDeviceToHost, s));
  SIA_CUDA(cudaStreamSynchronize(s));
  uint64_t *bk = ar.take<uint64_t>(nb);
  int32_t *bc = ar.take<int32_t>(nb);
  SIA_REQUIRE(bk && bc, SIA_E_NOMEM, "index scratch arena too small (bins)");
  SIA_CUDA(cudaMemsetAsync(bc, 0, sizeof(int32_t) * nb, s));
  run_reduce_kernel<<<grid_for(n), 256, 0, s>>>(d_key, d_weight, n, flag, pos, bk, bc);
  SIA_CHECK_LAUNCH();
  *bin_key = bk; *bin_count = bc; *nbins = nb;
  return SIA_OK;
}

size_t reduce_runs_bytes(int64_t n) { return (size_t)n * (4 + 8 + 8 + 4) + scan_tmp_bytes(n) + 4096; }

// vote over sorted unique bins (+ sorted unique row bins) for the pass-local queries [q_lo, q_hi)
int vote_sorted(Arena &ar, const uint64_t *bin_key, const int32_t *bin_count, int64_t nbins, const uint64_t *row_key,
                const int32_t *row_count, int64_t nrowbins, int q_lo, int q_hi, int qid_base, int topn, int32_t *o_song,
                int32_t *o_diff, int32_t *o_count, int32_t *o_rows, int32_t *o_nres, cudaStream_t s) {
  if (q_hi <= q_lo) return SIA_OK;
  uint64_t *seg_key = nullptr;
  unsigned long long *seg_best = nullptr;
  int64_t nseg = 0;
  if (nbins > 0) {
    uint32_t *flag = ar.take<uint32_t>(nbins);
    int64_t *pos = ar.take<int64_t>(nbins + 1);
    void *stmp = ar.take<char>(scan_tmp_bytes(nbins));
    SIA_REQUIRE(flag && pos && stmp, SIA_E_NOMEM, "index scratch arena too small (vote)");
    run_flag_kernel<<<grid_for(nbins), 256, 0, s>>>(bin_key, nbins, kDiffBits, flag);
    SIA_CHECK_LAUNCH();
    int rc = exclusive_scan_u32(flag, pos, nbins, stmp, s);
    if (rc) return rc;
    SIA_CUDA(cudaMemcpyAsync(&nseg, pos + nbins, sizeof nseg, cudaMemcpyDeviceToHost, s));
    SIA_CUDA(cudaStreamSynchronize(s));
    seg_key = ar.take<uint64_t>(nseg);
    seg_best = ar.take<unsigned long long>(nseg);
    SIA_REQUIRE(seg_key && seg_best, SIA_E_NOMEM, "index scratch arena too small (segments)");
    SIA_CUDA(cudaMemsetAsync(seg_best, 0, sizeof(unsigned long long) * nseg, s));
    seg_best_kernel<<<grid_for(nbins), 256, 0, s>>>(bin_key, bin_count, nbins, flag, pos, seg_key, seg_best);
    SIA_CHECK_LAUNCH();
  }
  topn_kernel<<<q_hi - q_lo, 256, 0, s>>>(seg_key, seg_best, nseg, row_key, row_count, nrowbins, q_lo, qid_base, topn,
                                          o_song, o_diff, o_count, o_rows, o_nres);
  SIA_CHECK_LAUNCH();
  return SIA_OK;
}

size_t vote_bytes(int64_t nbins) { return (size_t)nbins * (4 + 8 + 8 + 8) + scan_tmp_bytes(nbins) + 4096; }


// Per-query sort of the packed entries in shared memory (sia_index_query_batch: the entries of a pass arrive grouped by
// query, and all entries of a query share the query id in the top bits, so sorting each query's slice by the whole
// 128 bits gives the (query, hash, offset) order of the global sort).  One CTA per query, bitonic network on 16-byte
// keys, slices of up to kSmemSortMax entries (64 KB); a pass with a longer query takes the global LSD radix sort.
constexpr int kSmemSortMax = 4096;
constexpr int kSmemSortThreads = 512;

__device__ __forceinline__ bool rec_less(const ulonglong2 &a, const ulonglong2 &b) {
  return a.y < b.y || (a.y == b.y && a.x < b.x);
}

__global__ void __launch_bounds__(kSmemSortThreads)
sort_queries_kernel(ulonglong2 *__restrict__ ent, const int64_t *__restrict__ query_starts, int64_t i0) {
  extern __shared__ __align__(16) unsigned char sort_smem[];
  ulonglong2 *key = reinterpret_cast<ulonglong2 *>(sort_smem);
  const int64_t s0 = query_starts[blockIdx.x] - i0;
  const int n = (int)(query_starts[blockIdx.x + 1] - i0 - s0);
  if (n < 2) return;
  int P = 2;
  while (P < n) P <<= 1;
  for (int i = threadIdx.x; i < P; i += kSmemSortThreads) key[i] = i < n ? ent[s0 + i] : make_ulonglong2(~0ull, ~0ull);
  __syncthreads();
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < (P >> 1); t += kSmemSortThreads) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));      // the lower index of the t-th pair at distance j
        const int l = i | j;
        const ulonglong2 a = key[i], b = key[l];
        const bool up = (i & k) == 0;
        if (rec_less(b, a) == up) { key[i] = b; key[l] = a; }
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < n; i += kSmemSortThreads) ent[s0 + i] = key[i];
}

int lookup_pass(sia_index *ix, Arena &ar, const uint8_t *d_hash, const int32_t *d_qoff, const int32_t *d_qid,
                const int64_t *d_query_starts, int n_queries, int qid_base, int64_t i0, int64_t n, Lookup &L,
                cudaStream_t s, int64_t max_query_entries = 0) {
  L = Lookup();
  L.n = n;
  if (n == 0) return SIA_OK;
  ulonglong2 *a = ar.take<ulonglong2>(n), *b = ar.take<ulonglong2>(n);
  void *stmp = ar.take<char>(radix_sort_tmp_bytes(n));
  uint32_t *first = ar.take<uint32_t>(n), *c_all = ar.take<uint32_t>(n), *c_head = ar.take<uint32_t>(n);
  int64_t *off_all = ar.take<int64_t>(n + 1), *off_head = ar.take<int64_t>(n + 1);
  SIA_REQUIRE(a && b && stmp && first && c_all && c_head && off_all && off_head, SIA_E_NOMEM,
              "index scratch arena too small (lookup)");
  const bool timing = getenv("SIA_QUERY_TIMING") != nullptr;     // stage times on stderr
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  if (timing) { for (auto &e : ev) cudaEventCreate(&e); cudaEventRecord(ev[0], s); }
  pack_queries_kernel<<<grid_for(n), 256, 0, s>>>(d_hash, d_qoff, d_qid, d_query_starts, n_queries, qid_base, i0, n, a,
                                                 ix->status);
  SIA_CHECK_LAUNCH();
  bool in_b = false;
  int rc = SIA_OK;
  const bool smem_sort = d_query_starts && !d_qid && max_query_entries > 0 && max_query_entries <= kSmemSortMax &&
                         !(getenv("SIA_QUERY_SORT") && std::string(getenv("SIA_QUERY_SORT")) == "radix");
  if (smem_sort) {
    int P = 2;
    while (P < max_query_entries) P <<= 1;
    const size_t smem = (size_t)P * sizeof(ulonglong2);
    SIA_CUDA(cudaFuncSetAttribute(sort_queries_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sort_queries_kernel<<<(unsigned)n_queries, kSmemSortThreads, smem, s>>>(a, d_query_starts, i0);
    SIA_CHECK_LAUNCH();
  } else if ((rc = radix_sort(a, b, n, 16, 0, 16, stmp, s, &in_b))) {
    return rc;
  }
  if (timing) cudaEventRecord(ev[1], s);
  L.ent = in_b ? b : a;
  lookup_kernel<<<grid_for(n), 256, 0, s>>>(L.ent, n, ix->rows, ix->dir, ix->dir_bits, ix->n_rows, first, c_all, c_head);
  SIA_CHECK_LAUNCH();
  if (timing) cudaEventRecord(ev[2], s);
  if ((rc = exclusive_scan_u32(c_all, off_all, n, stmp, s))) return rc;
  if ((rc = exclusive_scan_u32(c_head, off_head, n, stmp, s))) return rc;
  if (timing) {
    cudaEventRecord(ev[3], s);
    cudaEventSynchronize(ev[3]);
    float t1 = 0, t2 = 0, t3 = 0;
    cudaEventElapsedTime(&t1, ev[0], ev[1]); cudaEventElapsedTime(&t2, ev[1], ev[2]); cudaEventElapsedTime(&t3, ev[2], ev[3]);
    fprintf(stderr, "[sia] lookup pass: %lld entries: pack + sort %.2f ms, lookup %.2f ms, scans %.2f ms\n", (long long)n, t1, t2, t3);
    for (auto &e : ev) cudaEventDestroy(e);
  }
  int64_t tot[2];
  SIA_CUDA(cudaMemcpyAsync(&tot[0], off_all + n, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
  SIA_CUDA(cudaMemcpyAsync(&tot[1], off_head + n, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
  SIA_CUDA(cudaStreamSynchronize(s));
  L.first = first; L.cnt_head = c_head; L.off_all = off_all; L.off_head = off_head; L.tuples = tot[0]; L.head_rows = tot[1];
  return SIA_OK;
}

size_t lookup_bytes(int64_t n) { return (size_t)n * (16 * 2 + 4 * 3) + (size_t)(n + 1) * 16 + radix_sort_tmp_bytes(n) + 8192; }

// entries [e0, e0+ne) of a Lookup -> sorted unique vote bins and row bins
int bins_from_entries(sia_index *ix, Arena &ar, const Lookup &L, int64_t e0, int64_t ne, int64_t tuples, int64_t tuple_base,
                      int64_t head_rows, int64_t head_base, uint64_t **bin_key, int32_t **bin_count, int64_t *nbins,
                      uint64_t **row_key, int32_t **row_count, int64_t *nrowbins, cudaStream_t s) {
  int rc;
  *nbins = *nrowbins = 0;
  *bin_key = *row_key = nullptr; *bin_count = *row_count = nullptr;
  for (int pass = 0; pass < 2; ++pass) {
    const int64_t m = pass == 0 ? tuples : head_rows;
    if (m == 0) continue;
    uint64_t *ka = ar.take<uint64_t>(m), *kb = ar.take<uint64_t>(m);
    void *stmp = ar.take<char>(radix_sort_tmp_bytes(m));
    SIA_REQUIRE(ka && kb && stmp, SIA_E_NOMEM, "index scratch arena too small (tuples)");
    const unsigned blocks = (unsigned)ceil_div(ne, 256);
    if (pass == 0)
      expand_kernel<false><<<blocks, 256, 0, s>>>(L.ent, e0, ne, L.first, L.off_all, ix->rows, ka, tuple_base);
    else
      expand_kernel<true><<<blocks, 256, 0, s>>>(L.ent, e0, ne, L.first, L.off_head, ix->rows, ka, head_base);
    SIA_CHECK_LAUNCH();
    bool in_b = false;
    if ((rc = radix_sort(ka, kb, m, 8, 0, 8, stmp, s, &in_b))) return rc;
    uint64_t *sorted = in_b ? kb : ka;
    if (pass == 0) rc = reduce_runs(ar, sorted, nullptr, m, bin_key, bin_count, nbins, s);
    else rc = reduce_runs(ar, sorted, nullptr, m, row_key, row_count, nrowbins, s);
    if (rc) return rc;
  }
  return SIA_OK;
}

size_t bins_bytes(int64_t tuples, int64_t head_rows) {
  return (size_t)tuples * 16 + radix_sort_tmp_bytes(tuples) + reduce_runs_bytes(tuples) + (size_t)head_rows * 16 +
         radix_sort_tmp_bytes(head_rows) + reduce_runs_bytes(head_rows) + 8192;
}

int check_status(sia_index *ix, cudaStream_t s, int mask, const char *msg) {
  int32_t st = 0;
  SIA_CUDA(cudaMemcpyAsync(&st, ix->status, sizeof st, cudaMemcpyDeviceToHost, s));
  SIA_CUDA(cudaStreamSynchronize(s));
  if (st & mask) {
    SIA_CUDA(cudaMemsetAsync(ix->status, 0, sizeof(int32_t), s));
    set_error(msg);
    return SIA_E_INVALID;
  }
  return SIA_OK;
}

}  // namespace

// hash-table vote of concatenated vote keys; SIA_E_UNSUPPORTED = fall back to the sort-based vote
static Arena g_vote_tables[64];
static int vote_tuples_hash(int device, const uint64_t *d_tuple_key, int64_t n_tuples, const uint64_t *d_row_key,
                            int64_t n_rows, int32_t n_queries, int32_t topn, int32_t *d_out_song, int32_t *d_out_diff,
                            int32_t *d_out_count, int32_t *d_out_rows, int32_t *d_out_nres, cudaStream_t s) {
  const int mult = 2;
  if ((int64_t)mult * n_tuples + 32 > 0xffffffffll) return SIA_E_UNSUPPORTED;      // per-query slot ranges are 32 bit
  Arena &ar = g_vote_tables[device];
  // sizes that need no device round trip: bins; the hashed song layout is the fallback if dense is larger
  const int64_t nb = (int64_t)mult * n_tuples + 32ll * n_queries;
  const int64_t ns_hashed = 2 * n_rows + 32ll * n_queries;
  const size_t small = (size_t)n_queries * (8 + sizeof(QMeta)) + 4096;
  int rc = ar.reserve(small + (size_t)nb * 8 + (size_t)ns_hashed * 16 + 4096);
  if (rc) return rc;
  uint32_t *cnt = ar.take<uint32_t>(2 * (size_t)n_queries);
  int32_t *info = ar.take<int32_t>(2);                      // [0] largest song id, [1] flags
  QMeta *meta = ar.take<QMeta>(n_queries);
  unsigned long long *bins = ar.take<unsigned long long>((size_t)nb + 2 * (size_t)ns_hashed);
  SIA_REQUIRE(cnt && info && meta && bins, SIA_E_NOMEM, "vote_tuples: scratch");
  SIA_CUDA(cudaMemsetAsync(cnt, 0, sizeof(uint32_t) * 2 * n_queries, s));
  SIA_CUDA(cudaMemsetAsync(info, 0, 2 * sizeof(int32_t), s));
  SIA_CUDA(cudaMemsetAsync(d_out_nres, 0, sizeof(int32_t) * n_queries, s));
  if (n_tuples) count_keys_kernel<<<grid_for(n_tuples), 256, 0, s>>>(d_tuple_key, n_tuples, n_queries, cnt, info);
  if (n_rows) count_keys_kernel<<<grid_for(n_rows), 256, 0, s>>>(d_row_key, n_rows, n_queries, cnt + n_queries, info);
  SIA_CHECK_LAUNCH();
  int32_t h_info[2] = {0, 0};
  SIA_CUDA(cudaMemcpyAsync(h_info, info, sizeof h_info, cudaMemcpyDeviceToHost, s));
  SIA_CUDA(cudaStreamSynchronize(s));
  SIA_REQUIRE(!(h_info[1] & 2), SIA_E_INVALID, "vote_tuples: query id outside 0..n_queries-1");
  const int64_t span = (int64_t)h_info[0] + 1;
  const bool dense = span * n_queries * 12 <= ns_hashed * 16;    // dense must fit where the hashed layout would
  const int64_t ns = dense ? span * n_queries : ns_hashed;
  unsigned long long *song_best = bins + nb;
  uint32_t *song_rows = reinterpret_cast<uint32_t *>(song_best + ns);
  uint32_t *song_key = song_rows + ns;
  build_meta_kernel<<<1, 1024, 0, s>>>(cnt, cnt + n_queries, n_queries, mult, dense ? span : 0, meta, info);
  SIA_CUDA(cudaMemsetAsync(bins, 0, (size_t)nb * 8 + (size_t)ns * (dense ? 12 : 16), s));
  if (dense) {
    if (n_tuples) vote_keys_kernel<true><<<grid_for(n_tuples), 256, 0, s>>>(d_tuple_key, n_tuples, meta, bins, song_key, song_best, info);
    if (n_rows) row_keys_kernel<true><<<grid_for(n_rows), 256, 0, s>>>(d_row_key, n_rows, meta, song_key, song_rows, info);
    topn_hash_kernel<true><<<n_queries, 256, 0, s>>>(song_key, song_rows, song_best, meta, 0, 0, topn, d_out_song, d_out_diff,
                                                     d_out_count, d_out_rows, d_out_nres);
  } else {
    if (n_tuples) vote_keys_kernel<false><<<grid_for(n_tuples), 256, 0, s>>>(d_tuple_key, n_tuples, meta, bins, song_key, song_best, info);
    if (n_rows) row_keys_kernel<false><<<grid_for(n_rows), 256, 0, s>>>(d_row_key, n_rows, meta, song_key, song_rows, info);
    topn_hash_kernel<false><<<n_queries, 256, 0, s>>>(song_key, song_rows, song_best, meta, 0, 0, topn, d_out_song, d_out_diff,
                                                      d_out_count, d_out_rows, d_out_nres);
  }
  SIA_CHECK_LAUNCH();
  SIA_CUDA(cudaMemcpyAsync(h_info, info, sizeof h_info, cudaMemcpyDeviceToHost, s));
  SIA_CUDA(cudaStreamSynchronize(s));
  if (h_info[1] & (16 | 4 | 8)) return SIA_E_UNSUPPORTED;      // count overflow / oversized query / full table: sort instead
  return SIA_OK;
}

extern "C" {

int sia_index_create(int device, int64_t capacity_rows, sia_index **out) {
  SIA_REQUIRE(out != nullptr, SIA_E_INVALID, "out is NULL");
  *out = nullptr;
  SIA_REQUIRE(capacity_rows > 0 && capacity_rows < 0xfffffff0ll, SIA_E_INVALID,
              "capacity_rows must be in 1 .. 2^32-17 (row ids are 32 bit)");
  int ndev = 0;
  SIA_CUDA(cudaGetDeviceCount(&ndev));
  SIA_REQUIRE(device >= 0 && device < ndev, SIA_E_INVALID, "no such CUDA device");
  SIA_CUDA(cudaSetDevice(device));
  sia_index *ix = new (std::nothrow) sia_index();
  SIA_REQUIRE(ix != nullptr, SIA_E_NOMEM, "out of host memory");
  ix->device = device;
  ix->capacity = capacity_rows;
  cudaError_t e = cudaMalloc(&ix->rows, (size_t)capacity_rows * sizeof(ulonglong2));
  if (e == cudaSuccess) e = cudaMalloc(&ix->status, 2 * sizeof(int32_t));
  if (e == cudaSuccess) e = cudaMemset(ix->status, 0, 2 * sizeof(int32_t));
  if (e == cudaSuccess) e = cudaMalloc(&ix->dir, 2 * sizeof(uint32_t));
  if (e == cudaSuccess) e = cudaMemset(ix->dir, 0, 2 * sizeof(uint32_t));
  if (e != cudaSuccess) {
    int rc = cuda_fail(e, "index allocation", __FILE__, __LINE__);
    sia_index_destroy(ix);
    return rc;
  }
  ix->dir_bits = 0;
  *out = ix;
  return SIA_OK;
}

int sia_index_destroy(sia_index *ix) {
  if (!ix) return SIA_OK;
  cudaSetDevice(ix->device);
  cudaDeviceSynchronize();
  if (ix->rows) cudaFree(ix->rows);
  if (ix->dir) cudaFree(ix->dir);
  if (ix->status) cudaFree(ix->status);
  ix->arena.release();
  ix->arena2.release();
  ix->arena3.release();
  delete ix;
  return SIA_OK;
}

int64_t sia_index_rows(const sia_index *ix) { return ix ? ix->n_rows : 0; }

static int insert_common(sia_index *ix, const int32_t *d_song, int32_t song_const, const uint8_t *d_hash,
                         const int32_t *d_off, int64_t n, cudaStream_t s) {
  SIA_REQUIRE(ix != nullptr, SIA_E_INVALID, "index is NULL");
  SIA_REQUIRE(n >= 0, SIA_E_INVALID, "n < 0");
  if (n == 0) return SIA_OK;
  SIA_REQUIRE(d_hash && d_off, SIA_E_INVALID, "NULL argument");
  int rc = set_device(ix);
  if (rc) return rc;
  if (ix->n_rows + ix->n_pending + n > ix->capacity) {
    set_error("index capacity exceeded (capacity_rows counts stored + pending rows)");
    return SIA_E_CAPACITY;
  }
  pack_rows_kernel<<<grid_for(n), 256, 0, s>>>(d_hash, d_off, d_song, song_const, n,
                                              ix->rows + ix->n_rows + ix->n_pending, ix->status);
  SIA_CHECK_LAUNCH();
  ix->n_pending += n;
  return SIA_OK;
}

int sia_index_insert(sia_index *ix, int32_t song_id, const uint8_t *d_hash, const int32_t *d_off, int64_t n, void *stream) {
  SIA_REQUIRE(song_id >= 0 && song_id <= (int32_t)kM24, SIA_E_INVALID, "song_id outside MEDIUMINT UNSIGNED (0..2^24-1)");
  return insert_common(ix, nullptr, song_id, d_hash, d_off, n, (cudaStream_t)stream);
}

int sia_index_insert_rows(sia_index *ix, const int32_t *d_song, const uint8_t *d_hash, const int32_t *d_off, int64_t n,
                          void *stream) {
  SIA_REQUIRE(d_song != nullptr || n == 0, SIA_E_INVALID, "NULL argument");
  return insert_common(ix, d_song, 0, d_hash, d_off, n, (cudaStream_t)stream);
}

int sia_index_insert_host(sia_index *ix, int32_t song_id, const uint8_t *h_hash, const int32_t *h_off, int64_t n) {
  if (ix) ix->cache_valid = false;
  SIA_REQUIRE(ix != nullptr, SIA_E_INVALID, "index is NULL");
  if (n == 0) return SIA_OK;
  SIA_REQUIRE(h_hash && h_off && n > 0, SIA_E_INVALID, "bad argument");
  int rc = set_device(ix);
  if (rc) return rc;
  if ((rc = ix->arena.reserve((size_t)n * (SIA_HASH_BYTES + 4) + 4096))) return rc;
  uint8_t *dh = ix->arena.take<uint8_t>((size_t)n * SIA_HASH_BYTES);
  int32_t *dof = ix->arena.take<int32_t>(n);
  SIA_CUDA(cudaMemcpy(dh, h_hash, (size_t)n * SIA_HASH_BYTES, cudaMemcpyHostToDevice));
  SIA_CUDA(cudaMemcpy(dof, h_off, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice));
  rc = sia_index_insert(ix, song_id, dh, dof, n, nullptr);
  if (rc) return rc;
  SIA_CUDA(cudaDeviceSynchronize());
  return SIA_OK;
}

int sia_index_finalize(sia_index *ix, int64_t *h_rows) {
  if (ix) ix->cache_valid = false;
  SIA_REQUIRE(ix != nullptr, SIA_E_INVALID, "index is NULL");
  int rc = set_device(ix);
  if (rc) return rc;
  cudaStream_t s = nullptr;
  if ((rc = check_status(ix, s, 1, "insert: song_id or offset outside 0..2^24-1"))) {
    ix->n_pending = 0;   // drop the offending batch
    return rc;
  }
  SIA_CUDA(cudaMemcpy(&ix->max_song, ix->status + 1, sizeof(int32_t), cudaMemcpyDeviceToHost));
  const int64_t n = ix->n_rows + ix->n_pending;
  if (ix->n_pending > 0 && n > 0) {
    ulonglong2 *alt = nullptr;
    SIA_CUDA(cudaMalloc(&alt, (size_t)n * sizeof(ulonglong2)));
    auto fail = [&](int code) { cudaFree(alt); return code; };
    if ((rc = ix->arena.reserve(radix_sort_tmp_bytes(n) + (size_t)n * 12 + scan_tmp_bytes(n) + 8192))) return fail(rc);
    void *stmp = ix->arena.take<char>(radix_sort_tmp_bytes(n));
    bool in_b = false;
    if ((rc = radix_sort(ix->rows, alt, n, 16, 0, 16, stmp, s, &in_b))) return fail(rc);
    ulonglong2 *sorted = in_b ? alt : ix->rows, *other = in_b ? ix->rows : alt;
    // UNIQUE(song_id, offset, hash): drop adjacent duplicates
    uint32_t *flag = ix->arena.take<uint32_t>(n);
    int64_t *pos = ix->arena.take<int64_t>(n + 1);
    void *sc = ix->arena.take<char>(scan_tmp_bytes(n));
    if (!flag || !pos || !sc) { set_error("index scratch arena too small (finalize)"); return fail(SIA_E_NOMEM); }
    uniq_flag_kernel<<<grid_for(n), 256, 0, s>>>(sorted, n, flag);
    if ((rc = exclusive_scan_u32(flag, pos, n, sc, s))) return fail(rc);
    int64_t nu = 0;
    cudaError_t e = cudaMemcpy(&nu, pos + n, sizeof nu, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return fail(cuda_fail(e, "finalize readback", __FILE__, __LINE__));
    if (nu != n) {
      uniq_compact_kernel<<<grid_for(n), 256, 0, s>>>(sorted, n, flag, pos, other);
      std::swap(sorted, other);
    }
    if (sorted != ix->rows) {
      e = cudaMemcpyAsync(ix->rows, sorted, (size_t)nu * sizeof(ulonglong2), cudaMemcpyDeviceToDevice, s);
      if (e != cudaSuccess) return fail(cuda_fail(e, "finalize copy", __FILE__, __LINE__));
    }
    e = cudaDeviceSynchronize();
    cudaFree(alt);
    if (e != cudaSuccess) return cuda_fail(e, "finalize", __FILE__, __LINE__);
    ix->n_rows = nu;
    ix->n_pending = 0;
  }
  // directory: ~8 rows per bucket
  int bits = 10;
  while (bits < 28 && (ix->n_rows >> bits) > 8) ++bits;
  if (bits != ix->dir_bits) {
    if (ix->dir) cudaFree(ix->dir);
    ix->dir = nullptr;
    SIA_CUDA(cudaMalloc(&ix->dir, ((size_t)(1ull << bits) + 1) * sizeof(uint32_t)));
    ix->dir_bits = bits;
  }
  build_dir_kernel<<<grid_for(ix->n_rows + 1), 256, 0, s>>>(ix->rows, ix->n_rows, bits, ix->dir);
  SIA_CHECK_LAUNCH();
  SIA_CUDA(cudaDeviceSynchronize());
  if (h_rows) *h_rows = ix->n_rows;
  return SIA_OK;
}

int sia_index_delete_songs(sia_index *ix, const int32_t *h_song_ids, int32_t n, int64_t *h_rows) {
  if (ix) ix->cache_valid = false;
  SIA_REQUIRE(ix != nullptr, SIA_E_INVALID, "index is NULL");
  SIA_REQUIRE(ix->n_pending == 0, SIA_E_INVALID, "index has pending rows: call sia_index_finalize first");
  if (h_rows) *h_rows = ix->n_rows;
  if (n <= 0 || ix->n_rows == 0) return SIA_OK;
  SIA_REQUIRE(h_song_ids != nullptr, SIA_E_INVALID, "NULL argument");
  int rc = set_device(ix);
  if (rc) return rc;
  cudaStream_t s = nullptr;
  std::vector<uint32_t> bitmap((1u << 24) / 32, 0u);
  for (int i = 0; i < n; ++i) {
    SIA_REQUIRE(h_song_ids[i] >= 0 && h_song_ids[i] <= (int32_t)kM24, SIA_E_INVALID, "song id out of range");
    bitmap[h_song_ids[i] >> 5] |= 1u << (h_song_ids[i] & 31);
  }
  const int64_t nr = ix->n_rows;
  if ((rc = ix->arena.reserve(bitmap.size() * 4 + (size_t)nr * 12 + scan_tmp_bytes(nr) + 8192))) return rc;
  uint32_t *d_bm = ix->arena.take<uint32_t>(bitmap.size());
  uint32_t *flag = ix->arena.take<uint32_t>(nr);
  int64_t *pos = ix->arena.take<int64_t>(nr + 1);
  void *sc = ix->arena.take<char>(scan_tmp_bytes(nr));
  SIA_REQUIRE(d_bm && flag && pos && sc, SIA_E_NOMEM, "index scratch arena too small (delete)");
  SIA_CUDA(cudaMemcpyAsync(d_bm, bitmap.data(), bitmap.size() * 4, cudaMemcpyHostToDevice, s));
  keep_flag_kernel<<<grid_for(nr), 256, 0, s>>>(ix->rows, nr, d_bm, flag);
  SIA_CHECK_LAUNCH();
  if ((rc = exclusive_scan_u32(flag, pos, nr, sc, s))) return rc;
  int64_t keep = 0;
  SIA_CUDA(cudaMemcpy(&keep, pos + nr, sizeof keep, cudaMemcpyDeviceToHost));
  if (keep != nr) {
    ulonglong2 *alt = nullptr;
    SIA_CUDA(cudaMalloc(&alt, (size_t)std::max<int64_t>(keep, 1) * sizeof(ulonglong2)));
    uniq_compact_kernel<<<grid_for(nr), 256, 0, s>>>(ix->rows, nr, flag, pos, alt);
    cudaError_t e = cudaMemcpyAsync(ix->rows, alt, (size_t)keep * sizeof(ulonglong2), cudaMemcpyDeviceToDevice, s);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaFree(alt);
    if (e != cudaSuccess) return cuda_fail(e, "delete_songs", __FILE__, __LINE__);
    ix->n_rows = keep;
    build_dir_kernel<<<grid_for(ix->n_rows + 1), 256, 0, s>>>(ix->rows, ix->n_rows, ix->dir_bits, ix->dir);
    SIA_CHECK_LAUNCH();
    SIA_CUDA(cudaDeviceSynchronize());
  }
  if (h_rows) *h_rows = ix->n_rows;
  return SIA_OK;
}

int sia_index_export(sia_index *ix, int64_t first_row, int64_t n, uint8_t *d_hash, int32_t *d_song, int32_t *d_off,
                     void *stream) {
  SIA_REQUIRE(ix != nullptr, SIA_E_INVALID, "index is NULL");
  SIA_REQUIRE(ix->n_pending == 0, SIA_E_INVALID, "index has pending rows: call sia_index_finalize first");
  SIA_REQUIRE(first_row >= 0 && n >= 0 && first_row + n <= ix->n_rows, SIA_E_INVALID, "export: row range outside the index");
  if (n == 0) return SIA_OK;
  SIA_REQUIRE(d_hash && d_song && d_off, SIA_E_INVALID, "NULL output");
  int rc = set_device(ix);
  if (rc) return rc;
  unpack_rows_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(ix->rows + first_row, n, d_hash, d_song, d_off);
  SIA_CHECK_LAUNCH();
  return SIA_OK;
}

int sia_index_select_host(sia_index *ix, const uint8_t *h_hash, int64_t n, int32_t *h_row_hashidx, int32_t *h_row_song,
                          int32_t *h_row_off, int64_t cap, int64_t *h_nrows) {
  if (ix) ix->cache_valid = false;
  SIA_REQUIRE(ix && h_nrows, SIA_E_INVALID, "NULL argument");
  SIA_REQUIRE(ix->n_pending == 0, SIA_E_INVALID, "index has pending rows: call sia_index_finalize first");
  *h_nrows = 0;
  if (n == 0) return SIA_OK;
  SIA_REQUIRE(h_hash && n > 0 && n <= (int64_t)kM24, SIA_E_INVALID, "select: 1..2^24-1 hashes per call");
  int rc = set_device(ix);
  if (rc) return rc;
  cudaStream_t s = nullptr;
  // position of each hash rides in the qoff field, qid = 0 -> sorted by (digest, position)
  std::vector<int32_t> pos(n);
  for (int64_t i = 0; i < n; ++i) pos[i] = (int32_t)i;
  const size_t in_bytes = (size_t)n * (SIA_HASH_BYTES + 4 + 4) + 4096;
  if ((rc = ix->arena.reserve(in_bytes + lookup_bytes(n) + (size_t)cap * 12 + 4096))) return rc;
  uint8_t *dh = ix->arena.take<uint8_t>((size_t)n * SIA_HASH_BYTES);
  int32_t *dpos = ix->arena.take<int32_t>(n), *dq = ix->arena.take<int32_t>(n);
  SIA_CUDA(cudaMemcpyAsync(dh, h_hash, (size_t)n * SIA_HASH_BYTES, cudaMemcpyHostToDevice, s));
  SIA_CUDA(cudaMemcpyAsync(dpos, pos.data(), (size_t)n * 4, cudaMemcpyHostToDevice, s));
  SIA_CUDA(cudaMemsetAsync(dq, 0, (size_t)n * 4, s));
  Lookup L;
  if ((rc = lookup_pass(ix, ix->arena, dh, dpos, dq, nullptr, 1, 0, 0, n, L, s))) return rc;
  *h_nrows = L.tuples;
  const int64_t m = std::min(L.tuples, cap);
  if (m > 0) {
    SIA_REQUIRE(h_row_hashidx && h_row_song && h_row_off, SIA_E_INVALID, "NULL output");
    int32_t *o1 = ix->arena.take<int32_t>(m), *o2 = ix->arena.take<int32_t>(m), *o3 = ix->arena.take<int32_t>(m);
    SIA_REQUIRE(o1 && o2 && o3, SIA_E_NOMEM, "index scratch arena too small (select)");
    select_rows_kernel<<<grid_for(n), 256, 0, s>>>(L.ent, n, L.first, L.off_all, ix->rows, o1, o2, o3, m);
    SIA_CHECK_LAUNCH();
    SIA_CUDA(cudaMemcpyAsync(h_row_hashidx, o1, (size_t)m * 4, cudaMemcpyDeviceToHost, s));
    SIA_CUDA(cudaMemcpyAsync(h_row_song, o2, (size_t)m * 4, cudaMemcpyDeviceToHost, s));
    SIA_CUDA(cudaMemcpyAsync(h_row_off, o3, (size_t)m * 4, cudaMemcpyDeviceToHost, s));
  }
  SIA_CUDA(cudaStreamSynchronize(s));
  return SIA_OK;
}

int sia_index_query_batch(sia_index *ix, const uint8_t *d_hash, const int32_t *d_qoff, const int64_t *h_query_starts,
                          int32_t n_queries, int32_t topn, int32_t *d_out_song, int32_t *d_out_diff,
                          int32_t *d_out_count, int32_t *d_out_rows, int32_t *d_out_nres, int64_t *h_stats, void *stream) {
  if (ix) ix->cache_valid = false;
  SIA_REQUIRE(ix && h_query_starts, SIA_E_INVALID, "NULL argument");
  SIA_REQUIRE(n_queries >= 0 && topn >= 1, SIA_E_INVALID, "n_queries >= 0 and topn >= 1 required");
  SIA_REQUIRE(ix->n_pending == 0, SIA_E_INVALID, "index has pending rows: call sia_index_finalize first");
  if (h_stats) h_stats[0] = h_stats[1] = h_stats[2] = h_stats[3] = 0;
  if (n_queries == 0) return SIA_OK;
  SIA_REQUIRE(d_out_song && d_out_diff && d_out_count && d_out_rows && d_out_nres, SIA_E_INVALID, "NULL output");
  int rc = set_device(ix);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  for (int q = 0; q < n_queries; ++q)
    SIA_REQUIRE(h_query_starts[q] <= h_query_starts[q + 1], SIA_E_INVALID, "query_starts must be non-decreasing");
  SIA_REQUIRE(h_query_starts[n_queries] == h_query_starts[0] || (d_hash && d_qoff), SIA_E_INVALID, "NULL input");
  SIA_CUDA(cudaMemsetAsync(d_out_nres, 0, sizeof(int32_t) * n_queries, s));

  // SIA_VOTE=sort keeps the sort-based vote (the first implementation; A/B checks); default: hash tables
  // (read per call, so a test can run both paths in one process)
  const bool use_hash = !(getenv("SIA_VOTE") && std::string(getenv("SIA_VOTE")) == "sort");
  const int64_t hash_budget = getenv("SIA_VOTE_GROUP_TUPLES") ? std::max(1ll, atoll(getenv("SIA_VOTE_GROUP_TUPLES")))
                                                               : (512ll << 20);
  // tuples per block of expand_vote_kernel: a multiple of 8 (every warp takes an eighth)
  const int vote_chunk = getenv("SIA_VOTE_CHUNK") ? std::max(256, atoi(getenv("SIA_VOTE_CHUNK")) & ~7) : kVoteTuplesDefault;
  const bool use_filter = !(getenv("SIA_VOTE_FILTER") && atoi(getenv("SIA_VOTE_FILTER")) == 0);
  const int vote_mult = getenv("SIA_VOTE_LOAD") ? std::min(16, std::max(2, atoi(getenv("SIA_VOTE_LOAD")))) : 2;   // bin slots per tuple
  // tuples voted at once: sort path x ~70 B of scratch each; hash path x ~16 B of tables, one launch per group
  int64_t tuple_budget = use_hash ? hash_budget : (96ll << 20);
  if (use_hash) {                            // never ask for more than ~60 % of what the device has left (+ what arena3 holds)
    size_t free_b = 0, total_b = 0;
    SIA_CUDA(cudaMemGetInfo(&free_b, &total_b));
    tuple_budget = std::max<int64_t>(1 << 20, std::min<int64_t>(tuple_budget, (int64_t)((free_b + ix->arena3.cap) * 0.6 / 20)));
  }
  for (int64_t q0 = 0; q0 < n_queries; q0 += kMaxQueriesPerPass) {
    const int nq = (int)std::min<int64_t>(kMaxQueriesPerPass, n_queries - q0);
    const int64_t i0 = h_query_starts[q0], n = h_query_starts[q0 + nq] - i0;
    if (h_stats) h_stats[0] += n;
    if (n == 0) continue;    // out_nres is already 0 for these queries
    if ((rc = ix->arena.reserve((size_t)(nq + 1) * 8 * 3 + lookup_bytes(n) + (1 << 20)))) return rc;
    int64_t *d_qs = ix->arena.take<int64_t>(nq + 1);
    int64_t *d_goff = ix->arena.take<int64_t>(2 * (size_t)(nq + 1));
    SIA_CUDA(cudaMemcpyAsync(d_qs, h_query_starts + q0, sizeof(int64_t) * (nq + 1), cudaMemcpyHostToDevice, s));
    Lookup L;
    const bool timing = getenv("SIA_QUERY_TIMING") != nullptr;     // stage times of this pass on stderr
    const auto h0 = std::chrono::steady_clock::now();
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
    if (timing) { for (auto &e : ev) cudaEventCreate(&e); cudaEventRecord(ev[0], s); }
    int64_t max_entries = 0;
    for (int q = 0; q < nq; ++q) max_entries = std::max(max_entries, h_query_starts[q0 + q + 1] - h_query_starts[q0 + q]);
    if ((rc = lookup_pass(ix, ix->arena, d_hash, d_qoff, nullptr, d_qs, nq, 0, i0, n, L, s, max_entries))) return rc;
    if ((rc = check_status(ix, s, 2, "query: offset outside 0..2^24-1"))) return rc;
    if (timing) cudaEventRecord(ev[1], s);
    if (h_stats) { h_stats[1] += L.head_rows; h_stats[2] += L.tuples; }
    // After the sort query q still owns entries [starts[q]-i0, starts[q+1]-i0) (duplicates stay, with no
    // postings), so the scanned offsets at those positions split the pass into groups that fit the budget.
    std::vector<int64_t> h_goff(2 * (size_t)(nq + 1));
    gather_offsets_kernel<<<grid_for(nq + 1), 256, 0, s>>>(L.off_all, L.off_head, d_qs, i0, nq, d_goff);
    SIA_CHECK_LAUNCH();
    SIA_CUDA(cudaMemcpyAsync(h_goff.data(), d_goff, h_goff.size() * 8, cudaMemcpyDeviceToHost, s));
    SIA_CUDA(cudaStreamSynchronize(s));
    const int64_t *h_off_all = h_goff.data(), *h_off_head = h_goff.data() + nq + 1;
    // Groups of consecutive queries voted together.  Hash-table vote: the group's tables should stay in L2.
    // A query with more entries than a packed bin count can hold is voted alone, by sorting.
    struct Group { int qa, qb; bool sorted; };
    std::vector<Group> groups;
    size_t need = 0;
    auto too_big = [&](int q) { return h_query_starts[q0 + q + 1] - h_query_starts[q0 + q] > kMaxHashVoteEntries; };
    for (int qa = 0; qa < nq;) {
      int qb = qa + 1;
      const bool sorted = !use_hash || too_big(qa);
      while (qb < nq && h_off_all[qb + 1] - h_off_all[qa] <= tuple_budget && !(use_hash && (sorted || too_big(qb)))) ++qb;
      groups.push_back({qa, qb, sorted});
      if (sorted) {
        const int64_t t = h_off_all[qb] - h_off_all[qa], h = h_off_head[qb] - h_off_head[qa];
        need = std::max(need, bins_bytes(t, h) + vote_bytes(t) + (1 << 20));
      }
      qa = qb;
    }
    if (need && (rc = ix->arena2.reserve(need))) return rc;
    // hash-table vote: table layout per group; song tables dense when that is the smaller layout.  With the
    // duplicate filter (default; SIA_VOTE_FILTER=0 turns it off) the bin table holds only tuples whose filter
    // bucket was hit twice: t/2 + 64 slots instead of 2t + 32, next to t + 1 filter words.
    std::vector<QMeta> h_meta(nq), h_meta0(nq);        // filtered layout / plain layout (fallback)
    std::vector<char> dense(groups.size(), 0);
    int64_t max_bytes = 0;
    const int64_t span = (int64_t)ix->max_song + 1;
    auto layout = [&](const Group &g, bool dns, bool filt, std::vector<QMeta> &out, int64_t *nb, int64_t *ns, int64_t *nf) {
      int64_t bb = 0, sb = 0, fb = 0;
      for (int q = g.qa; q < g.qb; ++q) {
        const int64_t t = h_off_all[q + 1] - h_off_all[q], h = h_off_head[q + 1] - h_off_head[q];
        QMeta &m = out[q];
        m.bin_base = bb; m.song_base = sb; m.filt_base = fb; m.pad_ = 0;
        m.bin_cap = filt ? (uint32_t)(t / 2 + 64)
                         : (uint32_t)(std::min<int64_t>(vote_mult, 0xffffff00ll / std::max<int64_t>(t, 1)) * t + 32);
        m.song_cap = dns ? (uint32_t)span : (uint32_t)(2 * h + 32);
        m.filt_words = filt ? (uint32_t)(t + 1) : 0u;
        bb += m.bin_cap; sb += m.song_cap; fb += m.filt_words;
      }
      *nb = bb; *ns = sb; *nf = fb;
    };
    auto table_bytes = [](int64_t nb, int64_t ns, int64_t nf, bool dns) {
      return (size_t)nb * 8 + (size_t)ns * (dns ? 12 : 16) + (size_t)nf * 4;
    };
    bool any_hash = false;
    for (size_t gi = 0; gi < groups.size(); ++gi) {
      const Group &g = groups[gi];
      if (g.sorted) continue;
      any_hash = true;
      int64_t hashed_slots = 0;
      for (int q = g.qa; q < g.qb; ++q) {
        hashed_slots += 2 * (h_off_head[q + 1] - h_off_head[q]) + 32;
        SIA_REQUIRE(h_off_all[q + 1] - h_off_all[q] < (1ll << 30), SIA_E_UNSUPPORTED,
                    "query_batch: more than 2^30 vote tuples in one query");
      }
      dense[gi] = span * (g.qb - g.qa) * 12 <= hashed_slots * 16 * 2;
      int64_t nb, ns, nf;
      layout(g, dense[gi], false, h_meta0, &nb, &ns, &nf);           // the arena must hold the fallback layout too
      max_bytes = std::max<int64_t>(max_bytes, table_bytes(nb, ns, nf, dense[gi]));
      if (use_filter) {
        layout(g, dense[gi], true, h_meta, &nb, &ns, &nf);
        max_bytes = std::max<int64_t>(max_bytes, table_bytes(nb, ns, nf, dense[gi]));
      }
    }
    const int max_bin = any_hash ? 1 : 0;
    char *tables = nullptr;
    unsigned long long *d_nbins = nullptr;      // [groups]: distinct bins per group
    int32_t *d_gflags = nullptr;                // [groups]: 8 = a table filled up (filter under-estimate), 16 = count overflow
    QMeta *d_meta = nullptr, *d_meta0 = nullptr;
    const size_t ng = groups.size();
    if (any_hash) {
      if ((rc = ix->arena3.reserve((size_t)max_bytes + 2 * (size_t)nq * sizeof(QMeta) + ng * 12 + 16384))) return rc;
      d_meta = ix->arena3.take<QMeta>(nq);
      d_meta0 = ix->arena3.take<QMeta>(nq);
      d_nbins = ix->arena3.take<unsigned long long>(ng);
      d_gflags = ix->arena3.take<int32_t>(ng);
      tables = ix->arena3.take<char>((size_t)max_bytes);
      SIA_REQUIRE(d_meta && d_meta0 && d_nbins && d_gflags && tables, SIA_E_NOMEM, "index scratch arena too small (vote tables)");
      SIA_CUDA(cudaMemcpyAsync(d_meta, h_meta.data(), sizeof(QMeta) * nq, cudaMemcpyHostToDevice, s));
      SIA_CUDA(cudaMemcpyAsync(d_meta0, h_meta0.data(), sizeof(QMeta) * nq, cudaMemcpyHostToDevice, s));
      SIA_CUDA(cudaMemsetAsync(d_nbins, 0, sizeof(unsigned long long) * ng, s));
      SIA_CUDA(cudaMemsetAsync(d_gflags, 0, sizeof(int32_t) * ng, s));
    }
    // one group through the tables: filt = two-pass form with the duplicate filter
    auto vote_group = [&](size_t gi, bool filt) -> int {
      const Group &g = groups[gi];
      const int64_t e0 = h_query_starts[q0 + g.qa] - i0, ne = h_query_starts[q0 + g.qb] - i0 - e0;
      const int64_t tuples = h_off_all[g.qb] - h_off_all[g.qa];
      const std::vector<QMeta> &hm = filt ? h_meta : h_meta0;
      const QMeta *dm = filt ? d_meta : d_meta0;
      const QMeta &last = hm[g.qb - 1];
      const int64_t nb = last.bin_base + last.bin_cap, ns = last.song_base + last.song_cap, nf = last.filt_base + last.filt_words;
      // [bins | song_best | song_rows | song_key (open addressing only) | filter]: one zero-fill
      unsigned long long *bins = reinterpret_cast<unsigned long long *>(tables);
      unsigned long long *song_best = bins + nb;
      uint32_t *song_rows = reinterpret_cast<uint32_t *>(song_best + ns);
      uint32_t *song_key = song_rows + ns;
      uint32_t *filter = dense[gi] ? song_key : song_key + ns;
      SIA_CUDA(cudaMemsetAsync(tables, 0, table_bytes(nb, ns, nf, dense[gi]), s));
      const unsigned blocks = (unsigned)ceil_div(tuples, vote_chunk);
      unsigned long long *nbp = d_nbins + gi;
      int32_t *fl = d_gflags + gi;
#define SIA_EXPAND(D, M)                                                                                              \
      expand_vote_kernel<D, M><<<blocks, 256, 0, s>>>(L.ent, e0, ne, L.first, L.off_all, L.cnt_head, ix->rows, dm, bins, \
                                                      song_key, song_rows, song_best, filter, nbp, vote_chunk, fl)
      if (dense[gi]) {
        if (filt) { SIA_EXPAND(true, 1); SIA_EXPAND(true, 2); } else { SIA_EXPAND(true, 0); }
        topn_hash_kernel<true><<<g.qb - g.qa, 256, 0, s>>>(song_key, song_rows, song_best, dm, g.qa, (int)q0, topn,
                                                           d_out_song, d_out_diff, d_out_count, d_out_rows, d_out_nres);
      } else {
        if (filt) { SIA_EXPAND(false, 1); SIA_EXPAND(false, 2); } else { SIA_EXPAND(false, 0); }
        topn_hash_kernel<false><<<g.qb - g.qa, 256, 0, s>>>(song_key, song_rows, song_best, dm, g.qa, (int)q0, topn,
                                                            d_out_song, d_out_diff, d_out_count, d_out_rows, d_out_nres);
      }
#undef SIA_EXPAND
      SIA_CHECK_LAUNCH();
      return SIA_OK;
    };
    for (size_t gi = 0; gi < groups.size(); ++gi) {
      const Group &g = groups[gi];
      const int64_t e0 = h_query_starts[q0 + g.qa] - i0, ne = h_query_starts[q0 + g.qb] - i0 - e0;
      const int64_t tuples = h_off_all[g.qb] - h_off_all[g.qa], heads = h_off_head[g.qb] - h_off_head[g.qa];
      if (g.sorted) {
        ix->arena2.used = 0;
        uint64_t *bk = nullptr, *rk = nullptr;
        int32_t *bc = nullptr, *rcnt = nullptr;
        int64_t nbins = 0, nrowbins = 0;
        if (ne > 0 && (rc = bins_from_entries(ix, ix->arena2, L, e0, ne, tuples, h_off_all[g.qa], heads, h_off_head[g.qa],
                                              &bk, &bc, &nbins, &rk, &rcnt, &nrowbins, s)))
          return rc;
        if (h_stats) h_stats[3] += nbins;
        // keys carry the pass-local query id; outputs are indexed by q0 + q
        if ((rc = vote_sorted(ix->arena2, bk, bc, nbins, rk, rcnt, nrowbins, g.qa, g.qb, (int)q0, topn, d_out_song,
                              d_out_diff, d_out_count, d_out_rows, d_out_nres, s)))
          return rc;
        SIA_CUDA(cudaStreamSynchronize(s));
        continue;
      }
      if (tuples == 0) continue;           // out_nres is already 0 for these queries
      if ((rc = vote_group(gi, use_filter))) return rc;
    }
    if (any_hash) {
      // groups whose filtered bin table filled up (mostly-duplicate tuples: more than a quarter of them shared a
      // bucket) are redone without the filter; their results and bin counts are simply overwritten
      std::vector<int32_t> h_flags(ng);
      std::vector<unsigned long long> h_nb(ng);
      SIA_CUDA(cudaMemcpyAsync(h_flags.data(), d_gflags, sizeof(int32_t) * ng, cudaMemcpyDeviceToHost, s));
      SIA_CUDA(cudaStreamSynchronize(s));
      bool redo = false;
      for (size_t gi = 0; gi < ng; ++gi) {
        if (!(h_flags[gi] & 8)) continue;
        SIA_REQUIRE(use_filter, SIA_E_CUDA, "query: vote table overflow (internal error)");
        redo = true;
        SIA_CUDA(cudaMemsetAsync(d_nbins + gi, 0, sizeof(unsigned long long), s));
        SIA_CUDA(cudaMemsetAsync(d_gflags + gi, 0, sizeof(int32_t), s));
        if ((rc = vote_group(gi, false))) return rc;
      }
      SIA_CUDA(cudaMemcpyAsync(h_nb.data(), d_nbins, sizeof(unsigned long long) * ng, cudaMemcpyDeviceToHost, s));
      if (redo) SIA_CUDA(cudaMemcpyAsync(h_flags.data(), d_gflags, sizeof(int32_t) * ng, cudaMemcpyDeviceToHost, s));
      SIA_CUDA(cudaStreamSynchronize(s));
      for (size_t gi = 0; gi < ng; ++gi) {
        SIA_REQUIRE(!(h_flags[gi] & (8 | 16)), SIA_E_CUDA, "query: vote table overflow (internal error)");
        if (h_stats) h_stats[3] += (int64_t)h_nb[gi];
      }
    }
    if (timing) cudaEventRecord(ev[2], s);
    SIA_CUDA(cudaStreamSynchronize(s));      // the lookup scratch and the tables are reused by the next pass / call
    if (timing) {
      float t_lookup = 0, t_vote = 0;
      cudaEventElapsedTime(&t_lookup, ev[0], ev[1]); cudaEventElapsedTime(&t_vote, ev[1], ev[2]);
      const double host_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - h0).count();
      fprintf(stderr, "[sia] query pass: %d queries, %lld entries, %lld tuples, %zu groups: lookup %.2f ms, vote %.2f ms, host wall %.2f ms\n",
              nq, (long long)n, (long long)L.tuples, groups.size(), t_lookup, t_vote, host_ms);
      for (auto &e : ev) cudaEventDestroy(e);
    }
  }
  return SIA_OK;
}

int sia_index_query_partial(sia_index *ix, const uint8_t *d_hash, const int32_t *d_qoff, const int32_t *d_qid, int64_t n,
                            uint64_t *d_bin_key, int32_t *d_bin_count, int64_t cap_bins, int64_t *h_nbins,
                            uint64_t *d_row_key, int32_t *d_row_count, int64_t cap_rowbins, int64_t *h_nrowbins,
                            void *stream) {
  if (ix) ix->cache_valid = false;
  SIA_REQUIRE(ix && h_nbins && h_nrowbins, SIA_E_INVALID, "NULL argument");
  SIA_REQUIRE(ix->n_pending == 0, SIA_E_INVALID, "index has pending rows: call sia_index_finalize first");
  *h_nbins = *h_nrowbins = 0;
  if (n == 0) return SIA_OK;
  SIA_REQUIRE(d_hash && d_qoff && d_qid && n > 0, SIA_E_INVALID, "bad argument");
  int rc = set_device(ix);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  if ((rc = ix->arena.reserve(lookup_bytes(n) + (1 << 20)))) return rc;
  Lookup L;
  if ((rc = lookup_pass(ix, ix->arena, d_hash, d_qoff, d_qid, nullptr, 0, 0, 0, n, L, s))) return rc;
  if ((rc = check_status(ix, s, 2, "query: offset outside 0..2^24-1 or query id outside 0..32767"))) return rc;
  // grow the arena for the expansion without losing the lookup results: allocate a second arena
  Arena ar2;
  if ((rc = ar2.reserve(bins_bytes(L.tuples, L.head_rows) + (1 << 20)))) return rc;
  uint64_t *bk = nullptr, *rk = nullptr;
  int32_t *bc = nullptr, *rcnt = nullptr;
  int64_t nbins = 0, nrowbins = 0;
  rc = bins_from_entries(ix, ar2, L, 0, n, L.tuples, 0, L.head_rows, 0, &bk, &bc, &nbins, &rk, &rcnt, &nrowbins, s);
  if (!rc) {
    *h_nbins = nbins; *h_nrowbins = nrowbins;
    if (nbins > cap_bins || nrowbins > cap_rowbins) {
      set_error("query_partial: bin output capacity exceeded; *h_nbins / *h_nrowbins hold the required sizes");
      rc = SIA_E_CAPACITY;
    } else {
      cudaError_t e = cudaSuccess;
      if (nbins) {
        e = cudaMemcpyAsync(d_bin_key, bk, nbins * 8, cudaMemcpyDeviceToDevice, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_bin_count, bc, nbins * 4, cudaMemcpyDeviceToDevice, s);
      }
      if (e == cudaSuccess && nrowbins) {
        e = cudaMemcpyAsync(d_row_key, rk, nrowbins * 8, cudaMemcpyDeviceToDevice, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_row_count, rcnt, nrowbins * 4, cudaMemcpyDeviceToDevice, s);
      }
      if (e == cudaSuccess) e = cudaStreamSynchronize(s);
      if (e != cudaSuccess) rc = cuda_fail(e, "query_partial copy", __FILE__, __LINE__);
    }
  }
  cudaStreamSynchronize(s);
  ar2.release();
  return rc;
}

int sia_vote_bins(int device, const uint64_t *d_bin_key, const int32_t *d_bin_count, int64_t nbins,
                  const uint64_t *d_row_key, const int32_t *d_row_count, int64_t nrowbins, int32_t n_queries,
                  int32_t topn, int32_t *d_out_song, int32_t *d_out_diff, int32_t *d_out_count, int32_t *d_out_rows,
                  int32_t *d_out_nres, void *stream) {
  SIA_REQUIRE(n_queries >= 0 && n_queries <= kMaxQueriesPerPass && topn >= 1, SIA_E_INVALID,
              "vote_bins: 0..32768 queries, topn >= 1");
  SIA_REQUIRE(nbins >= 0 && nrowbins >= 0, SIA_E_INVALID, "negative size");
  if (n_queries == 0) return SIA_OK;
  SIA_REQUIRE(d_out_song && d_out_diff && d_out_count && d_out_rows && d_out_nres, SIA_E_INVALID, "NULL output");
  SIA_CUDA(cudaSetDevice(device));
  SIA_REQUIRE(device >= 0 && device < 64, SIA_E_INVALID, "device index");
  cudaStream_t s = (cudaStream_t)stream;
  Arena &ar = g_vote_arena[device];
  int rc = ar.reserve((size_t)(nbins + nrowbins) * (16 * 2 + 12) + radix_sort_tmp_bytes(nbins) +
                      radix_sort_tmp_bytes(nrowbins) + reduce_runs_bytes(nbins) + reduce_runs_bytes(nrowbins) +
                      vote_bytes(nbins) + (1 << 20));
  if (rc) return rc;
  // partial bins from several shards: equal keys must be summed -> sort (key, count) records, then weighted runs
  uint64_t *mk[2] = {nullptr, nullptr};
  int32_t *mc[2] = {nullptr, nullptr};
  int64_t mn[2] = {0, 0};
  for (int pass = 0; pass < 2 && !rc; ++pass) {
    const int64_t m = pass == 0 ? nbins : nrowbins;
    const uint64_t *k = pass == 0 ? d_bin_key : d_row_key;
    const int32_t *c = pass == 0 ? d_bin_count : d_row_count;
    if (m == 0) continue;
    ulonglong2 *ra = ar.take<ulonglong2>(m), *rb = ar.take<ulonglong2>(m);
    uint64_t *sk = ar.take<uint64_t>(m);
    int32_t *sw = ar.take<int32_t>(m);
    void *stmp = ar.take<char>(radix_sort_tmp_bytes(m));
    if (!ra || !rb || !sk || !sw || !stmp) { set_error("vote_bins: scratch"); rc = SIA_E_NOMEM; break; }
    join_bins_kernel<<<grid_for(m), 256, 0, s>>>(k, c, m, ra);
    bool in_b = false;
    if ((rc = radix_sort(ra, rb, m, 16, 8, 16, stmp, s, &in_b))) break;
    split_bins_kernel<<<grid_for(m), 256, 0, s>>>(in_b ? rb : ra, m, sk, sw);
    rc = reduce_runs(ar, sk, sw, m, &mk[pass], &mc[pass], &mn[pass], s);
  }
  if (!rc) {
    cudaMemsetAsync(d_out_nres, 0, sizeof(int32_t) * n_queries, s);
    rc = vote_sorted(ar, mk[0], mc[0], mn[0], mk[1], mc[1], mn[1], 0, n_queries, 0, topn, d_out_song, d_out_diff,
                     d_out_count, d_out_rows, d_out_nres, s);
  }
  cudaStreamSynchronize(s);
  return rc;
}

int sia_index_expand(sia_index *ix, const uint8_t *d_hash, const int32_t *d_qoff, const int32_t *d_qid, int64_t n,
                     int32_t n_queries, uint64_t *d_tuple_key, int64_t cap_tuples, int64_t *h_ntuples,
                     uint64_t *d_row_key, int64_t cap_rows, int64_t *h_nrows, int64_t *d_tuple_starts,
                     int64_t *d_row_starts, void *stream) {
  SIA_REQUIRE(ix && h_ntuples && h_nrows, SIA_E_INVALID, "NULL argument");
  SIA_REQUIRE(ix->n_pending == 0, SIA_E_INVALID, "index has pending rows: call sia_index_finalize first");
  SIA_REQUIRE(n >= 0 && n_queries >= 0 && n_queries <= kMaxQueriesPerPass, SIA_E_INVALID, "expand: bad sizes");
  *h_ntuples = *h_nrows = 0;
  int rc = set_device(ix);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  if (n == 0) {
    if (d_tuple_starts) SIA_CUDA(cudaMemsetAsync(d_tuple_starts, 0, sizeof(int64_t) * (n_queries + 1), s));
    if (d_row_starts) SIA_CUDA(cudaMemsetAsync(d_row_starts, 0, sizeof(int64_t) * (n_queries + 1), s));
    return SIA_OK;
  }
  SIA_REQUIRE(d_hash && d_qoff && d_qid, SIA_E_INVALID, "NULL input");
  Lookup &L = ix->cache_lookup;      // lookup of the sizing call, reused when the same inputs come back
  const bool reuse = ix->cache_valid && ix->cache_hash == d_hash && ix->cache_qoff == d_qoff &&
                     ix->cache_qid == d_qid && ix->cache_n == n && d_tuple_key && d_row_key;
  ix->cache_valid = false;
  if (!reuse) {
    if ((rc = ix->arena.reserve(lookup_bytes(n) + (1 << 20)))) return rc;
    if ((rc = lookup_pass(ix, ix->arena, d_hash, d_qoff, d_qid, nullptr, 0, 0, 0, n, L, s))) return rc;
    if ((rc = check_status(ix, s, 2, "query: offset outside 0..2^24-1 or query id outside 0..32767"))) return rc;
  }
  *h_ntuples = L.tuples;
  *h_nrows = L.head_rows;
  if (!d_tuple_key || !d_row_key) {                                 // sizing call: keep the lookup for the next call
    ix->cache_valid = true; ix->cache_hash = d_hash; ix->cache_qoff = d_qoff;
    ix->cache_qid = d_qid; ix->cache_n = n;
    return SIA_OK;
  }
  if (L.tuples > cap_tuples || L.head_rows > cap_rows) {
    set_error("expand: output capacity exceeded; *h_ntuples / *h_nrows hold the required sizes");
    return SIA_E_CAPACITY;
  }
  SIA_REQUIRE(d_tuple_starts && d_row_starts, SIA_E_INVALID, "NULL output");
  const unsigned blocks = (unsigned)ceil_div(n, 256);
  if (L.tuples) expand_kernel<false><<<blocks, 256, 0, s>>>(L.ent, 0, n, L.first, L.off_all, ix->rows, d_tuple_key, 0);
  if (L.head_rows) expand_kernel<true><<<blocks, 256, 0, s>>>(L.ent, 0, n, L.first, L.off_head, ix->rows, d_row_key, 0);
  query_starts_kernel<<<grid_for(n_queries + 1), 256, 0, s>>>(L.ent, n, L.off_all, L.off_head, n_queries, d_tuple_starts,
                                                            d_row_starts);
  SIA_CHECK_LAUNCH();
  SIA_CUDA(cudaStreamSynchronize(s));      // the lookup scratch is reused by the next call
  return SIA_OK;
}

int sia_vote_tuples(int device, uint64_t *d_tuple_key, int64_t n_tuples, uint64_t *d_row_key, int64_t n_rows,
                    int32_t n_queries, int32_t topn, int32_t *d_out_song, int32_t *d_out_diff, int32_t *d_out_count,
                    int32_t *d_out_rows, int32_t *d_out_nres, void *stream) {
  SIA_REQUIRE(n_queries >= 0 && n_queries <= kMaxQueriesPerPass && topn >= 1, SIA_E_INVALID,
              "vote_tuples: 0..32768 queries, topn >= 1");
  SIA_REQUIRE(n_tuples >= 0 && n_rows >= 0, SIA_E_INVALID, "negative size");
  if (n_queries == 0) return SIA_OK;
  SIA_REQUIRE(d_out_song && d_out_diff && d_out_count && d_out_rows && d_out_nres, SIA_E_INVALID, "NULL output");
  SIA_CUDA(cudaSetDevice(device));
  SIA_REQUIRE(device >= 0 && device < 64, SIA_E_INVALID, "device index");
  cudaStream_t s = (cudaStream_t)stream;
  if (!(getenv("SIA_VOTE") && std::string(getenv("SIA_VOTE")) == "sort")) {
    const int rc_hash = vote_tuples_hash(device, d_tuple_key, n_tuples, d_row_key, n_rows, n_queries, topn, d_out_song,
                                         d_out_diff, d_out_count, d_out_rows, d_out_nres, s);
    if (rc_hash != SIA_E_UNSUPPORTED) return rc_hash;      // else: a bin count beyond 15 bits -> the sort-based vote
    for (int32_t *o : {d_out_song, d_out_diff, d_out_count, d_out_rows})
      SIA_CUDA(cudaMemsetAsync(o, 0, sizeof(int32_t) * (size_t)n_queries * topn, s));
  }
  Arena &ar = g_vote_arena[device];
  int rc = ar.reserve((size_t)(n_tuples + n_rows) * 8 + radix_sort_tmp_bytes(n_tuples) + radix_sort_tmp_bytes(n_rows) +
                      reduce_runs_bytes(n_tuples) + reduce_runs_bytes(n_rows) + vote_bytes(n_tuples) + (1 << 20));
  if (rc) return rc;
  uint64_t *bk[2] = {nullptr, nullptr};
  int32_t *bc[2] = {nullptr, nullptr};
  int64_t nb[2] = {0, 0};
  for (int pass = 0; pass < 2 && !rc; ++pass) {
    const int64_t m = pass == 0 ? n_tuples : n_rows;
    uint64_t *keys = pass == 0 ? d_tuple_key : d_row_key;      // sorted in place (the caller's buffer is scratch)
    if (m == 0) continue;
    uint64_t *alt = ar.take<uint64_t>(m);
    void *stmp = ar.take<char>(radix_sort_tmp_bytes(m));
    if (!alt || !stmp) { set_error("vote_tuples: scratch"); rc = SIA_E_NOMEM; break; }
    bool in_b = false;
    if ((rc = radix_sort(keys, alt, m, 8, 0, 8, stmp, s, &in_b))) break;
    rc = reduce_runs(ar, in_b ? alt : keys, nullptr, m, &bk[pass], &bc[pass], &nb[pass], s);
  }
  if (!rc) {
    cudaMemsetAsync(d_out_nres, 0, sizeof(int32_t) * n_queries, s);
    rc = vote_sorted(ar, bk[0], bc[0], nb[0], bk[1], bc[1], nb[1], 0, n_queries, 0, topn, d_out_song, d_out_diff,
                     d_out_count, d_out_rows, d_out_nres, s);
  }
  cudaStreamSynchronize(s);
  return rc;
}

}  // extern "C"
