// K4, the partitioned vote — align_matches' histogram (recognizer.py:303-310) counted in SHARED memory.
//
// The table vote of index_query.cu counts every (query, song, diff) tuple with atomics on tables that live in HBM:
// a random 32-byte sector per tuple and pass.  Here the tuples of a query are first split by a hash of the SONG id
// into partitions of ~3 600 tuples, so that each partition's bins fit an exact open-addressing table in shared memory:
//
//   scatter  one CTA per block of 8 192 consecutive tuples of ONE query (posting runs walked entry by entry, or a slice
//            of received vote keys): tuple -> (song 24 | diff 25 | head 1), partition = hash(song) * np >> 32, counted in
//            a shared histogram, sorted by partition through a shared staging buffer, and written as one contiguous
//            run per partition into the (query, partition) region (one global atomicAdd per block and partition
//            reserves the room);
//   count    one CTA per region: every tuple inserted into a shared-memory table (one 64-bit word per slot =
//            key 49 | count 15; CAS to claim, 32-bit add to bump), bins that reach count 2 remembered in a list; the
//            region's top-n songs by (count desc, song asc), per song the largest bin, smallest diff on ties, come
//            from that list (or from a scan of the table when fewer than n songs have a repeated bin);
//   merge    one warp per query: top-n over its regions' candidates.  A song's bins all live in ONE partition, so
//            the per-partition winners are exact and so is their merge.
// All traffic is streaming: 8 B read + 8 B written per tuple (scatter), 8 B read (count).  Queries that do not fit —
// a partition above its capacity (one song with thousands of tuples), more than 2 048 partitions — are flagged and
// voted by the table vote instead; nothing is approximated.
#include "index.cuh"

using namespace sia;

namespace {

constexpr int kDiffBits = SIA_KEY_DIFF_BITS, kSongBits = SIA_KEY_SONG_BITS, kQidBits = SIA_KEY_QID_BITS;
constexpr uint64_t kDiffMask = (1ull << kDiffBits) - 1;

constexpr int kBlk = 8192;            // tuples per scatter block
constexpr int kScThreads = 512;
constexpr int kTpt = kBlk / kScThreads;
constexpr int kSlots = 8192;          // slots of the count table (64 KB)
constexpr int kCap = 6144;            // tuples a region can hold (table load <= 0.75)
constexpr int kAvg = 3584;            // tuples per partition the layout aims at
constexpr int kMaxParts = 2048;
constexpr int kDup = 2048;            // repeated bins a region remembers before it scans the table instead
constexpr int kCntThreads = 512;

struct PvQuery {
  int64_t reg_off;                    // first tuple slot of the query's regions
  uint32_t ridx0, np, cap, nt;        // first region index, partitions, tuple slots per region, tuples
};

struct PvTune { uint32_t cap, avg; };

PvTune pv_tune() {
  PvTune v{kCap, kAvg};
  if (const char *e = getenv("SIA_PVOTE_CAP")) {            // tests: small regions force the fallback to the table vote
    const long c = atol(e);
    if (c >= 2 && c <= kCap) { v.cap = (uint32_t)c & ~1u; v.avg = std::max(1u, (uint32_t)(v.cap * 7ull / 12)); }
  }
  return v;
}

__host__ __device__ inline uint32_t pv_parts(uint64_t t, uint32_t cap, uint32_t avg) {
  if (t <= cap) return 1;
  const uint64_t n = (t + avg - 1) / avg;
  return n > (uint64_t)kMaxParts ? (uint32_t)kMaxParts : (uint32_t)n;
}
__host__ __device__ inline uint32_t pv_region_cap(uint64_t t, uint32_t cap) {
  return t <= cap ? (uint32_t)((t + 1) & ~1ull) : cap;        // a single region holds exactly its query
}

__device__ __forceinline__ uint32_t mix32(uint32_t k) {
  k ^= k >> 16; k *= 0x85ebca6bu; k ^= k >> 13; k *= 0xc2b2ae35u; k ^= k >> 16;
  return k;
}
__device__ __forceinline__ uint32_t pv_part(uint32_t song, uint32_t np) { return (uint32_t)(((uint64_t)mix32(song) * np) >> 32); }

// ---- layout ---------------------------------------------------------------------------------------------------
// segments: (query, source) -> a range of tuples / keys.  Entries variant: one source, the query's slice of off_all.
__global__ void pv_segs_entries_kernel(const int64_t *__restrict__ goff, int qa, int nq, int64_t *__restrict__ seg_lo,
                                       uint32_t *__restrict__ seg_cnt) {
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += gridDim.x * blockDim.x) {
    seg_lo[q] = goff[qa + q];
    seg_cnt[q] = (uint32_t)(goff[qa + q + 1] - goff[qa + q]);
  }
}

// keys variant: slot g sorted by query id -> boundaries by binary search; anything inconsistent sets *unsorted
__global__ void pv_segs_keys_kernel(const uint64_t *__restrict__ keys, int64_t cap, const int64_t *__restrict__ counts, int G,
                                    int nq, int64_t *__restrict__ seg_lo, uint32_t *__restrict__ seg_cnt,
                                    uint32_t *__restrict__ unsorted) {
  const uint32_t qmask = (1u << kQidBits) - 1u;
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < nq * G; s += gridDim.x * blockDim.x) {
    const int q = s / G, g = s - q * G;
    int64_t n;
    const uint64_t *__restrict__ kk = slot_keys(keys, cap, counts, g, n);
    int64_t b[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const uint32_t want = (uint32_t)(q + h);
      int64_t lo = 0, hi = n;
      if (want == 0) hi = 0;
      else if (want >= (uint32_t)nq) lo = n;
      while (lo < hi) {
        const int64_t mid = lo + ((hi - lo) >> 1);
        if (((uint32_t)(kk[mid] >> (kSongBits + kDiffBits)) & qmask) < want) lo = mid + 1; else hi = mid;
      }
      b[h] = lo;
    }
    if (b[1] < b[0] || b[1] - b[0] > 0xffffffffll) { atomicOr(unsorted, 1u); b[1] = b[0]; }
    seg_lo[s] = b[0];
    seg_cnt[s] = (uint32_t)(b[1] - b[0]);
  }
}

// single block: per query partitions / region offsets, per segment first scatter block; tot = {regions, blocks}
__global__ void __launch_bounds__(1024)
pv_layout_kernel(const uint32_t *__restrict__ seg_cnt, int nq, int G, PvTune tune, PvQuery *__restrict__ pq,
                 uint32_t *__restrict__ q_ridx0, uint32_t *__restrict__ seg_blk0, uint32_t *__restrict__ tot) {
  __shared__ int64_t s_reg[1024];
  __shared__ uint32_t s_idx[1024], s_blk[1024];
  const int per = (nq + 1023) / 1024;
  const int a = min(nq, (int)threadIdx.x * per), b = min(nq, a + per);
  int64_t reg = 0;
  uint32_t idx = 0, blk = 0;
  for (int q = a; q < b; ++q) {
    uint64_t t = 0;
    for (int g = 0; g < G; ++g) { const uint32_t c = seg_cnt[q * G + g]; t += c; blk += (c + kBlk - 1) / kBlk; }
    const uint32_t np = t ? pv_parts(t, tune.cap, tune.avg) : 0;
    idx += np;
    reg += (int64_t)np * pv_region_cap(t, tune.cap);
  }
  s_reg[threadIdx.x] = reg; s_idx[threadIdx.x] = idx; s_blk[threadIdx.x] = blk;
  __syncthreads();
  if (threadIdx.x == 0) {
    int64_t r = 0; uint32_t i = 0, k = 0;
    for (int j = 0; j < 1024; ++j) {
      const int64_t x = s_reg[j]; const uint32_t y = s_idx[j], z = s_blk[j];
      s_reg[j] = r; s_idx[j] = i; s_blk[j] = k;
      r += x; i += y; k += z;
    }
    tot[0] = i; tot[1] = k;
    q_ridx0[nq] = i; seg_blk0[nq * G] = k;
  }
  __syncthreads();
  reg = s_reg[threadIdx.x]; idx = s_idx[threadIdx.x]; blk = s_blk[threadIdx.x];
  for (int q = a; q < b; ++q) {
    uint64_t t = 0;
    for (int g = 0; g < G; ++g) {
      const uint32_t c = seg_cnt[q * G + g];
      seg_blk0[q * G + g] = blk;
      t += c; blk += (c + kBlk - 1) / kBlk;
    }
    PvQuery m;
    m.np = t ? pv_parts(t, tune.cap, tune.avg) : 0;
    m.cap = pv_region_cap(t, tune.cap);
    m.nt = (uint32_t)min(t, (uint64_t)0xffffffffu);
    m.reg_off = reg; m.ridx0 = idx;
    pq[q] = m;
    q_ridx0[q] = idx;
    idx += m.np; reg += (int64_t)m.np * m.cap;
  }
}

// ---- scatter --------------------------------------------------------------------------------------------------
struct ScatterArgs {
  const uint32_t *seg_blk0; const int64_t *seg_lo; const uint32_t *seg_cnt; int n_seg, G;
  const PvQuery *pq; const uint32_t *tot;
  uint64_t *regions; uint32_t *fill; uint32_t *qover; int q_lo;
  // source 0: posting runs of the query's entries
  const ulonglong2 *ent; const int64_t *first; const int64_t *off; const uint32_t *cnt_head; const uint64_t *post;
  const int64_t *q_ent; int64_t i0;
  // source 1: vote keys in slots
  const uint64_t *keys; int64_t key_cap; const int64_t *counts; uint32_t *unsorted;
};

constexpr size_t kScatterSmem = (size_t)kBlk * 8 + (size_t)kBlk * 2 + (size_t)(2 * kMaxParts + 1) * 4;

template <int SRC>
__global__ void __launch_bounds__(kScThreads, 2) pv_scatter_kernel(const ScatterArgs a) {
  extern __shared__ __align__(16) unsigned char pv_smem[];
  uint64_t *stage = reinterpret_cast<uint64_t *>(pv_smem);              // [kBlk] tuples in arrival order
  uint16_t *idx = reinterpret_cast<uint16_t *>(stage + kBlk);            // [kBlk] arrival index of the t-th tuple in partition order
  uint32_t *hist = reinterpret_cast<uint32_t *>(idx + kBlk);             // [np + 1] counts, then exclusive offsets
  int32_t *gdst = reinterpret_cast<int32_t *>(hist + kMaxParts + 1);     // [np] region slot of partition order position 0
  __shared__ uint32_t s_warp[kScThreads / 32];
  const uint32_t b = blockIdx.x;
  if (b >= a.tot[1]) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int s = find_segment(a.seg_blk0, a.n_seg, b);
  const int ql = s / a.G, g = s - ql * a.G;
  const PvQuery m = a.pq[ql];
  const int64_t lo = a.seg_lo[s];
  const int64_t j0 = lo + (int64_t)(b - a.seg_blk0[s]) * kBlk;
  const int n = (int)min((int64_t)kBlk, lo + (int64_t)a.seg_cnt[s] - j0);
  const uint32_t np = m.np;
  for (uint32_t p = tid; p <= np; p += kScThreads) hist[p] = 0;
  __syncthreads();

  uint32_t pr[kTpt];                       // partition << 13 | rank inside the partition (this block)
  const int iw = warp * (kBlk / (kScThreads / 32));
  if (SRC == 0) {
    int64_t e = 0, onext = 0, adj = 0;
    uint32_t qh = 0;
    bool fresh = true;
    if (iw < n) {
      // entry of the warp's first tuple: largest e in the query's entries with off[e] <= j0 + iw
      const int64_t jw = j0 + iw;
      e = a.q_ent[a.q_lo + ql] - a.i0;
      int64_t hi = a.q_ent[a.q_lo + ql + 1] - a.i0;
      while (hi - e > 1) { const int64_t mid = e + ((hi - e) >> 1); if (a.off[mid] <= jw) e = mid; else hi = mid; }
      onext = a.off[e + 1];
    }
#pragma unroll
    for (int k0 = 0; k0 < kTpt; k0 += 4) {
      uint64_t r[4];
      uint32_t q4[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = iw + (k0 + u) * 32 + lane;
        r[u] = 0; q4[u] = 0;
        if (i < n) {
          const int64_t j = j0 + i;
          while (onext <= j) { ++e; onext = a.off[e + 1]; fresh = true; }
          if (fresh) {
            adj = a.first[e] - a.off[e];
            qh = ((uint32_t)(a.ent[e].x & kM24) << 1) | (a.cnt_head[e] != 0 ? 1u : 0u);
            fresh = false;
          }
          r[u] = __ldcs(a.post + adj + j);
          q4[u] = qh;
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = iw + (k0 + u) * 32 + lane;
        pr[k0 + u] = 0;
        if (i < n) {
          const uint32_t song = (uint32_t)(r[u] >> 24) & 0xffffffu;
          const uint32_t dbits = ((uint32_t)(r[u] & kM24) - (q4[u] >> 1) + SIA_DIFF_BIAS) & (uint32_t)kDiffMask;
          const uint32_t p = pv_part(song, np);
          const uint32_t rk = atomicAdd(&hist[p], 1u);
          stage[i] = ((uint64_t)song << (kDiffBits + 1)) | ((uint64_t)dbits << 1) | (q4[u] & 1u);
          pr[k0 + u] = (p << 13) | rk;
        }
      }
    }
  } else {
    int64_t nk;
    const uint64_t *__restrict__ kk = slot_keys(a.keys, a.key_cap, a.counts, g, nk);
    const uint32_t qmask = (1u << kQidBits) - 1u;
    bool bad = false;
#pragma unroll
    for (int k0 = 0; k0 < kTpt; k0 += 4) {
      uint64_t r[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = iw + (k0 + u) * 32 + lane;
        r[u] = i < n ? __ldcs(kk + j0 + i) : 0ull;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = iw + (k0 + u) * 32 + lane;
        pr[k0 + u] = 0;
        if (i < n) {
          const uint64_t key = r[u];
          bad = bad || ((uint32_t)(key >> (kSongBits + kDiffBits)) & qmask) != (uint32_t)ql;
          const uint32_t song = (uint32_t)(key >> kDiffBits) & 0xffffffu;
          const uint32_t p = pv_part(song, np);
          const uint32_t rk = atomicAdd(&hist[p], 1u);
          stage[i] = ((uint64_t)song << (kDiffBits + 1)) | ((key & kDiffMask) << 1) | (key >> 63);
          pr[k0 + u] = (p << 13) | rk;
        }
      }
    }
    if (bad) atomicOr(a.unsorted, 1u);
  }
  __syncthreads();

  // exclusive scan of the histogram (4 partitions per thread), room reserved in the regions
  {
    const uint32_t p0 = (uint32_t)tid * 4;
    uint32_t v[4], sum = 0;
#pragma unroll
    for (int u = 0; u < 4; ++u) { v[u] = p0 + u < np ? hist[p0 + u] : 0u; sum += v[u]; }
    uint32_t inc = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      const uint32_t w = lane < kScThreads / 32 ? s_warp[lane] : 0u;
      uint32_t winc = w;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, winc, d); if (lane >= d) winc += t; }
      if (lane < kScThreads / 32) s_warp[lane] = winc - w;
    }
    __syncthreads();
    uint32_t base = s_warp[warp] + inc - sum;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (p0 + u < np) {
        const uint32_t c = v[u];
        hist[p0 + u] = base;
        int32_t d = INT32_MIN;
        if (c) {
          const uint32_t old = atomicAdd(&a.fill[m.ridx0 + p0 + u], c);
          if (old + c <= m.cap) d = (int32_t)((p0 + u) * m.cap + old) - (int32_t)base;
          else a.qover[a.q_lo + ql] = 1u;
        }
        gdst[p0 + u] = d;
        base += c;
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kTpt; ++k) {
    const int i = iw + k * 32 + lane;
    if (i < n) idx[hist[pr[k] >> 13] + (pr[k] & 8191u)] = (uint16_t)i;
  }
  __syncthreads();
  uint64_t *__restrict__ reg = a.regions + m.reg_off;
  for (int t = tid; t < n; t += kScThreads) {
    const uint64_t tup = stage[idx[t]];
    const int32_t d = gdst[pv_part((uint32_t)(tup >> (kDiffBits + 1)), np)];
    if (d != INT32_MIN) reg[(int64_t)d + t] = tup;
  }
}

// ---- count ----------------------------------------------------------------------------------------------------
// slot word: (song 24 | diff 25) << 15 | count 15
__device__ __forceinline__ uint64_t pv_rank(uint64_t slot) {          // count desc, song asc, diff asc as ONE maximum
  if (slot == 0ull) return 0ull;
  const uint64_t key = slot >> 15;
  return ((slot & 0x7fffull) << 49) | ((kM24 - (key >> kDiffBits)) << kDiffBits) | (kDiffMask - (key & kDiffMask));
}

__device__ __forceinline__ void pv_insert(uint64_t *tab, uint32_t mask, uint64_t key, uint16_t *dup, uint32_t *ndup,
                                          uint32_t &fresh) {
  uint32_t h = mix32((uint32_t)key * 0x9e3779b1u + (uint32_t)(key >> 32) * 0x85ebca6bu) & mask;
  const uint64_t claim = (key << 15) | 1ull;
  for (;;) {
    uint64_t cur = *reinterpret_cast<volatile uint64_t *>(tab + h);
    if (cur == 0ull) {
      cur = atomicCAS(reinterpret_cast<unsigned long long *>(tab + h), 0ull, (unsigned long long)claim);
      if (cur == 0ull) { ++fresh; return; }
    }
    if ((cur >> 15) == key) {
      const uint32_t old = atomicAdd(reinterpret_cast<uint32_t *>(tab + h), 1u);      // low word: key bits | count
      if ((old & 0x7fffu) == 1u) { const uint32_t d = atomicAdd(ndup, 1u); if (d < (uint32_t)kDup) dup[d] = (uint16_t)h; }
      return;
    }
    h = (h + 1) & mask;
  }
}

template <bool ROWS>
__global__ void __launch_bounds__(kCntThreads, 3)
pv_count_kernel(const uint64_t *__restrict__ regions, const uint32_t *__restrict__ fill, const PvQuery *__restrict__ pq,
                const uint32_t *__restrict__ q_ridx0, int nq, const uint32_t *__restrict__ tot, int q_lo,
                const uint32_t *__restrict__ qover, int topn, uint64_t *__restrict__ cand, uint32_t *__restrict__ cand_rows,
                unsigned long long *__restrict__ n_bins) {
  extern __shared__ __align__(16) unsigned char pv_smem[];
  uint64_t *tab = reinterpret_cast<uint64_t *>(pv_smem);
  __shared__ uint16_t s_dup[kDup];
  __shared__ uint32_t s_ndup;
  __shared__ uint64_t s_red[kCntThreads / 32];
  __shared__ uint64_t s_win[kPvMaxTopn];
  __shared__ uint32_t s_rows[kPvMaxTopn];
  const uint32_t r = blockIdx.x;
  if (r >= tot[0]) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ql = find_segment(q_ridx0, nq, r);
  const PvQuery m = pq[ql];
  const uint32_t p = r - m.ridx0;
  uint64_t *__restrict__ out = cand + (int64_t)r * topn;
  const uint32_t n = fill[r];
  if (n == 0 || qover[q_lo + ql]) {
    if (tid < topn) { out[tid] = 0ull; if (ROWS) cand_rows[(int64_t)r * topn + tid] = 0u; }
    return;
  }
  uint32_t S = 256;
  while (S < 2 * n && S < (uint32_t)kSlots) S <<= 1;
  const uint32_t mask = S - 1;
  for (uint32_t i = tid; i < S / 2; i += kCntThreads) reinterpret_cast<ulonglong2 *>(tab)[i] = make_ulonglong2(0ull, 0ull);
  if (tid == 0) s_ndup = 0;
  if (tid < kPvMaxTopn) s_rows[tid] = 0;
  __syncthreads();
  const uint64_t *__restrict__ reg = regions + m.reg_off + (int64_t)p * m.cap;
  uint32_t fresh = 0;
  for (uint32_t i0 = tid; i0 < n; i0 += 4 * kCntThreads) {
    uint64_t t[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { const uint32_t i = i0 + u * kCntThreads; t[u] = i < n ? __ldcs(reg + i) : 0ull; }
#pragma unroll
    for (int u = 0; u < 4; ++u) if (i0 + u * kCntThreads < n) pv_insert(tab, mask, t[u] >> 1, s_dup, &s_ndup, fresh);
  }
  if (n_bins) {
#pragma unroll
    for (int d = 16; d; d >>= 1) fresh += __shfl_xor_sync(0xffffffffu, fresh, d);
    if (lane == 0 && fresh) atomicAdd(n_bins, (unsigned long long)fresh);
  }
  __syncthreads();
  const uint32_t ndup_raw = s_ndup;
  const uint32_t nd = min(ndup_raw, (uint32_t)kDup);
  bool full = ndup_raw > (uint32_t)kDup;
  int nres = 0;
  while (nres < topn) {
    uint64_t best = 0;
    const uint32_t lim = full ? S : nd;
    for (uint32_t i = tid; i < lim; i += kCntThreads) {
      uint64_t c = pv_rank(tab[full ? i : (uint32_t)s_dup[i]]);
      if (c > best) {
        const uint64_t song = (c >> kDiffBits) & kM24;
        for (int w = 0; w < nres; ++w) if (((s_win[w] >> kDiffBits) & kM24) == song) c = 0;
        if (c > best) best = c;
      }
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) { const uint64_t o = __shfl_xor_sync(0xffffffffu, best, d); if (o > best) best = o; }
    if (lane == 0) s_red[warp] = best;
    __syncthreads();
    best = 0;
#pragma unroll
    for (int w = 0; w < kCntThreads / 32; ++w) { const uint64_t o = s_red[w]; if (o > best) best = o; }
    if (best == 0ull) {
      if (full) break;
      full = true;                          // fewer than topn songs with a repeated bin: the bins of count 1 decide
      __syncthreads();
      continue;
    }
    if (tid == 0) s_win[nres] = best;
    ++nres;
    __syncthreads();
  }
  if (ROWS && nres > 0) {                   // dedup_hashes of the region's winners: head tuples of their songs
    for (uint32_t i = tid; i < n; i += kCntThreads) {
      const uint64_t t = reg[i];
      if (t & 1ull) {
        const uint64_t isong = kM24 - (t >> (kDiffBits + 1));
        for (int w = 0; w < nres; ++w) if (((s_win[w] >> kDiffBits) & kM24) == isong) atomicAdd(&s_rows[w], 1u);
      }
    }
    __syncthreads();
  }
  if (tid < topn) {
    out[tid] = tid < nres ? s_win[tid] : 0ull;
    if (ROWS) cand_rows[(int64_t)r * topn + tid] = tid < nres ? s_rows[tid] : 0u;
  }
}

// ---- merge ----------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pv_merge_kernel(const uint64_t *__restrict__ cand, const uint32_t *__restrict__ cand_rows, const PvQuery *__restrict__ pq, int nq,
                int q_lo, int qid_base, const uint32_t *__restrict__ qover, int topn, PvOut out, uint32_t *__restrict__ over_count) {
  const int lane = threadIdx.x & 31;
  const int ql = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (ql >= nq) return;
  const int q = q_lo + ql;
  if (qover[q]) { if (lane == 0 && over_count) atomicAdd(over_count, 1u); return; }
  const PvQuery m = pq[ql];
  if (m.np == 0) return;                    // no tuples: the outputs are already zero
  const uint64_t *__restrict__ c = cand + (int64_t)m.ridx0 * topn;
  const uint32_t total = m.np * (uint32_t)topn;
  uint64_t prev = ~0ull;
  int nres = 0;
  for (int r = 0; r < topn; ++r) {
    uint64_t best = 0;
    uint32_t at = 0;
    for (uint32_t i = lane; i < total; i += 32) { const uint64_t v = c[i]; if (v < prev && v > best) { best = v; at = i; } }
#pragma unroll
    for (int d = 16; d; d >>= 1) {
      const uint64_t o = __shfl_xor_sync(0xffffffffu, best, d);
      const uint32_t oa = __shfl_xor_sync(0xffffffffu, at, d);
      if (o > best) { best = o; at = oa; }
    }
    if (best == 0ull) break;
    if (lane == 0) {
      const int64_t o = ((int64_t)q + qid_base) * topn + r;
      out.song[o] = (int32_t)(kM24 - ((best >> kDiffBits) & kM24));
      out.count[o] = (int32_t)(best >> 49);
      out.diff[o] = (int32_t)(kDiffMask - (best & kDiffMask)) - SIA_DIFF_BIAS;
      out.rows[o] = cand_rows ? (int32_t)cand_rows[(int64_t)m.ridx0 * topn + at] : 0;
    }
    prev = best;
    ++nres;
  }
  if (lane == 0) out.nres[q + qid_base] = nres;
}

struct PvScratch {
  int64_t *seg_lo; uint32_t *seg_cnt, *seg_blk0, *q_ridx0, *tot, *fill, *cand_rows;
  PvQuery *pq; uint64_t *regions, *cand;
};

size_t pv_scratch_bytes(int64_t n_seg, int64_t nq, int64_t regions, int64_t region_tuples, int topn) {
  return (size_t)n_seg * 16 + (size_t)(n_seg + nq + 2) * 4 + (size_t)nq * sizeof(PvQuery) + (size_t)regions * (4 + 12 * (size_t)topn) +
         (size_t)region_tuples * 8 + 16 * 256 + 64;
}

bool pv_take(Arena &ar, PvScratch &S, int64_t n_seg, int64_t nq, int64_t regions, int64_t region_tuples, int topn) {
  S.seg_lo = ar.take<int64_t>(n_seg);
  S.seg_cnt = ar.take<uint32_t>(n_seg);
  S.seg_blk0 = ar.take<uint32_t>(n_seg + 1);
  S.q_ridx0 = ar.take<uint32_t>(nq + 1);
  S.tot = ar.take<uint32_t>(4);
  S.pq = ar.take<PvQuery>(nq);
  S.fill = ar.take<uint32_t>(regions);
  S.cand = ar.take<uint64_t>(regions * topn);
  S.cand_rows = ar.take<uint32_t>(regions * topn);
  S.regions = ar.take<uint64_t>(region_tuples);
  return S.seg_lo && S.seg_cnt && S.seg_blk0 && S.q_ridx0 && S.tot && S.pq && S.fill && S.cand && S.cand_rows && S.regions;
}

int pv_attrs() {
  static bool done[64] = {false};
  int dev = 0;
  SIA_CUDA(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && done[dev]) return SIA_OK;
  SIA_CUDA(cudaFuncSetAttribute(pv_scatter_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kScatterSmem));
  SIA_CUDA(cudaFuncSetAttribute(pv_scatter_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kScatterSmem));
  SIA_CUDA(cudaFuncSetAttribute(pv_count_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSlots * 8));
  SIA_CUDA(cudaFuncSetAttribute(pv_count_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSlots * 8));
  if (dev >= 0 && dev < 64) done[dev] = true;
  return SIA_OK;
}

}  // namespace

namespace sia {

// worst case over all splits of `tuples` into nq queries: a query that fits one region takes its own size (rounded to 2),
// a larger one ceil(t / avg) <= t / avg + 1 regions of cap slots — and there are at most min(nq, tuples / cap) of those
static void pv_bounds(int64_t tuples, int64_t nq, const PvTune &t, int64_t &regions, int64_t &region_tuples) {
  regions = tuples / t.avg + nq + 1;
  region_tuples = (tuples / t.avg + 1) * (int64_t)t.cap + std::min<int64_t>(nq * (int64_t)t.cap, tuples) + 2 * nq + t.cap;
}

size_t pvote_bytes(int64_t tuples, int64_t nq, int n_src, int topn) {
  int64_t regions, region_tuples;
  pv_bounds(tuples, nq, pv_tune(), regions, region_tuples);
  return pv_scratch_bytes(nq * n_src, nq, regions, region_tuples, topn);
}

int pvote_entries(Arena &ar, const Lookup &L, const uint64_t *post, const int64_t *d_qs, int64_t i0, const int64_t *d_goff,
                  const int64_t *h_goff, int qa, int qb, int qid_base, int topn, const PvOut &out, uint32_t *d_qover,
                  unsigned long long *d_nbins, cudaStream_t s) {
  const int nq = qb - qa;
  if (nq <= 0) return SIA_OK;
  int rc = pv_attrs();
  if (rc) return rc;
  const PvTune tune = pv_tune();
  int64_t regions = 0, region_tuples = 0, blocks = 0;
  for (int q = qa; q < qb; ++q) {
    const uint64_t t = (uint64_t)(h_goff[q + 1] - h_goff[q]);
    if (t == 0) continue;
    const uint32_t np = pv_parts(t, tune.cap, tune.avg);
    regions += np;
    region_tuples += (int64_t)np * pv_region_cap(t, tune.cap);
    blocks += (int64_t)ceil_div((int64_t)t, kBlk);
  }
  if (blocks == 0) return SIA_OK;
  SIA_REQUIRE(blocks < (1ll << 31) && regions < (1ll << 31), SIA_E_UNSUPPORTED, "vote: group too large");
  PvScratch S;
  SIA_REQUIRE(pv_take(ar, S, nq, nq, regions, region_tuples, topn), SIA_E_NOMEM, "index scratch arena too small (partitioned vote)");
  SIA_CUDA(cudaMemsetAsync(S.fill, 0, sizeof(uint32_t) * regions, s));
  pv_segs_entries_kernel<<<grid_for(nq), 256, 0, s>>>(d_goff, qa, nq, S.seg_lo, S.seg_cnt);
  pv_layout_kernel<<<1, 1024, 0, s>>>(S.seg_cnt, nq, 1, tune, S.pq, S.q_ridx0, S.seg_blk0, S.tot);
  ScatterArgs a{};
  a.seg_blk0 = S.seg_blk0; a.seg_lo = S.seg_lo; a.seg_cnt = S.seg_cnt; a.n_seg = nq; a.G = 1;
  a.pq = S.pq; a.tot = S.tot; a.regions = S.regions; a.fill = S.fill; a.qover = d_qover; a.q_lo = qa;
  a.ent = L.ent; a.first = L.first; a.off = L.off_all; a.cnt_head = L.cnt_head; a.post = post; a.q_ent = d_qs; a.i0 = i0;
  pv_scatter_kernel<0><<<(unsigned)blocks, kScThreads, kScatterSmem, s>>>(a);
  pv_count_kernel<false><<<(unsigned)regions, kCntThreads, kSlots * 8, s>>>(S.regions, S.fill, S.pq, S.q_ridx0, nq, S.tot, qa, d_qover,
                                                                            topn, S.cand, nullptr, d_nbins);
  pv_merge_kernel<<<(unsigned)ceil_div((int64_t)nq * 32, 256), 256, 0, s>>>(S.cand, nullptr, S.pq, nq, qa, qid_base, d_qover, topn, out,
                                                                            nullptr);
  SIA_CHECK_LAUNCH();
  return SIA_OK;
}

int pvote_key_slots(Arena &ar, const uint64_t *d_keys, int n_slots, int64_t cap, const int64_t *d_counts, int nq, int topn,
                    const PvOut &out, uint32_t *d_qover, uint32_t *d_flags2, cudaStream_t s) {
  if (nq <= 0) return SIA_OK;
  int rc = pv_attrs();
  if (rc) return rc;
  const PvTune tune = pv_tune();
  const int64_t T = (int64_t)n_slots * cap;                 // upper bound of the keys
  int64_t regions, region_tuples;
  pv_bounds(T, nq, tune, regions, region_tuples);
  const int64_t blocks = ceil_div(T, kBlk) + (int64_t)nq * n_slots;
  SIA_REQUIRE(blocks < (1ll << 31) && regions < (1ll << 31), SIA_E_UNSUPPORTED, "vote: too many keys in one call");
  const int64_t n_seg = (int64_t)nq * n_slots;
  PvScratch S;
  SIA_REQUIRE(pv_take(ar, S, n_seg, nq, regions, region_tuples, topn), SIA_E_NOMEM, "vote scratch too small (partitioned vote)");
  SIA_CUDA(cudaMemsetAsync(S.fill, 0, sizeof(uint32_t) * regions, s));
  pv_segs_keys_kernel<<<grid_for(n_seg), 256, 0, s>>>(d_keys, cap, d_counts, n_slots, nq, S.seg_lo, S.seg_cnt, d_flags2);
  pv_layout_kernel<<<1, 1024, 0, s>>>(S.seg_cnt, nq, n_slots, tune, S.pq, S.q_ridx0, S.seg_blk0, S.tot);
  ScatterArgs a{};
  a.seg_blk0 = S.seg_blk0; a.seg_lo = S.seg_lo; a.seg_cnt = S.seg_cnt; a.n_seg = (int)n_seg; a.G = n_slots;
  a.pq = S.pq; a.tot = S.tot; a.regions = S.regions; a.fill = S.fill; a.qover = d_qover; a.q_lo = 0;
  a.keys = d_keys; a.key_cap = cap; a.counts = d_counts; a.unsorted = d_flags2;
  pv_scatter_kernel<1><<<(unsigned)blocks, kScThreads, kScatterSmem, s>>>(a);
  pv_count_kernel<true><<<(unsigned)regions, kCntThreads, kSlots * 8, s>>>(S.regions, S.fill, S.pq, S.q_ridx0, nq, S.tot, 0, d_qover, topn,
                                                                           S.cand, S.cand_rows, nullptr);
  pv_merge_kernel<<<(unsigned)ceil_div((int64_t)nq * 32, 256), 256, 0, s>>>(S.cand, S.cand_rows, S.pq, nq, 0, 0, d_qover, topn, out,
                                                                            d_flags2 + 1);
  SIA_CHECK_LAUNCH();
  return SIA_OK;
}

}  // namespace sia
