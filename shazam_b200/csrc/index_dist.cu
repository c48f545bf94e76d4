// K4, hash-prefix sharding — the device steps on either side of the two NCCL all-to-alls of a query pass
// (shazam_b200/distributed.py drives them):
//   sia_route_entries      the rank that OWNS the queries packs their (query, hash, offset) entries into one fixed-size
//                          slot per destination shard, owner = floor(prefix16(hash) * world / 65536);
//   [all-to-all #1: slots of 16-byte entries]
//   sia_index_expand_slots the shard that owns the hashes sorts what it received, looks the hashes up and expands the
//                          posting runs into vote keys, written straight into one slot per query-owning rank;
//   [all-to-all #2: slots of 8-byte vote keys]
//   sia_vote_key_slots     the query's owner votes the keys of all shards with the same passes as the single-GPU path
//                          (a bin's true count is the SUM over shards, so the keys — not local winners — travel).
// Slots have a fixed capacity, so both exchanges are equal-split all-to-alls with no host-side size negotiation; element 0
// of a slot is its count.  A slot that would overflow sets a flag (and reports the size it needed) and the pass is redone
// with larger slots.
#include "index.cuh"

#include <cstring>

using namespace sia;

namespace {

constexpr int kVoteTuples = 4096;

__global__ void init_entry_headers_kernel(ulonglong2 *__restrict__ slots, int world, int64_t cap) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d < world) slots[(int64_t)d * cap] = make_ulonglong2(0ull, 0ull);
}

__global__ void __launch_bounds__(256)
route_entries_kernel(const uint8_t *__restrict__ hash, const int32_t *__restrict__ qoff, const int64_t *__restrict__ query_starts,
                     int n_queries, int64_t n, int qid_base, int world, int64_t cap, ulonglong2 *__restrict__ slots,
                     int32_t *__restrict__ status) {
  const int lane = threadIdx.x & 31;
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x; i0 < n; i0 += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = i0 + threadIdx.x;
    const bool valid = i < n;
    uint64_t hi = 0; uint32_t lo16 = 0;
    int32_t o = 0, q = 0;
    if (valid) {
      load_digest(hash + i * SIA_HASH_BYTES, hi, lo16);
      o = qoff[i];
      q = find_segment(query_starts, n_queries, i) + qid_base;
      if (o < 0 || o > (int32_t)kM24 || q < 0 || q >= (1 << 24) - 1) atomicOr(status, 2);
    }
    const uint32_t owner = (uint32_t)(((hi >> 48) * (uint64_t)world) >> 16);
    const uint32_t active = __ballot_sync(0xffffffffu, valid);
    if (!valid) continue;
    const uint32_t peers = __match_any_sync(active, owner);
    const int leader = __ffs(peers) - 1;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(&slots[(int64_t)owner * cap].x, (unsigned long long)__popc(peers));
    base = __shfl_sync(peers, base, leader);
    const int64_t pos = (int64_t)base + __popc(peers & ((1u << lane) - 1u)) + 1;
    if (pos < cap) slots[(int64_t)owner * cap + pos] = make_entry((uint32_t)q, hi, lo16, (uint32_t)o);
  }
}

// received slots -> one dense array (padding = all-ones entries, which sort last and match nothing)
__global__ void __launch_bounds__(256)
gather_entries_kernel(const ulonglong2 *__restrict__ slots, int world, int64_t cap, ulonglong2 *__restrict__ out,
                      int64_t *__restrict__ info) {
  const int64_t per = cap - 1, total = (int64_t)world * per;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t s = i / per, k = i - s * per;
    const int64_t c = (int64_t)slots[s * cap].x;
    if (k == 0 && c > per) { atomicOr((unsigned long long *)info, 1ull); atomicMax((unsigned long long *)info + 2, (unsigned long long)c + 1); }
    out[i] = k < min(c, per) ? slots[s * cap + 1 + k] : make_ulonglong2(~0ull, ~0ull);
  }
}

// tuple offsets at which the keys of destination rank d start (entries are sorted by query id, rank d owns the ids
// [d * qp, (d + 1) * qp)); writes the slot headers
__global__ void cuts_kernel(const ulonglong2 *__restrict__ ent, int64_t n, const int64_t *__restrict__ off_all, int world, int qp,
                            int64_t cap_out, int64_t *__restrict__ cut, uint64_t *__restrict__ key_slots,
                            int64_t *__restrict__ info) {
  __shared__ int64_t s_cut[1025];
  const int d = threadIdx.x;
  if (d <= world) {
    const int64_t qlim = (int64_t)d * qp;
    int64_t lo = 0, hi = n;
    while (lo < hi) { const int64_t mid = lo + ((hi - lo) >> 1); if ((int64_t)(ent[mid].y >> 40) < qlim) lo = mid + 1; else hi = mid; }
    s_cut[d] = off_all[lo];
    cut[d] = s_cut[d];
  }
  __syncthreads();
  if (d < world) {
    const int64_t c = s_cut[d + 1] - s_cut[d];
    key_slots[(int64_t)d * cap_out] = (uint64_t)c;
    if (c > cap_out - 1) { atomicOr((unsigned long long *)info, 2ull); }
    atomicMax((unsigned long long *)info + 1, (unsigned long long)c + 1);
  }
}

// posting runs -> vote keys in the destination rank's slot (the entry walk of entries_pass_kernel, index_query.cu)
__global__ void __launch_bounds__(256)
expand_slots_kernel(const ulonglong2 *__restrict__ ent, int64_t n, const int64_t *__restrict__ first,
                    const int64_t *__restrict__ off, const uint32_t *__restrict__ cnt_head, const uint64_t *__restrict__ post,
                    const int64_t *__restrict__ cut, int qp, int64_t cap_out, uint64_t *__restrict__ key_slots) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int per_warp = kVoteTuples >> 3;
  const int64_t j_lo = (int64_t)blockIdx.x * kVoteTuples + (int64_t)warp * per_warp;
  const int64_t j_hi = min(off[n], j_lo + per_warp);
  if (j_lo >= j_hi) return;
  int64_t ei = 0, hi = n;
  while (hi - ei > 1) { const int64_t mid = ei + ((hi - ei) >> 1); if (off[mid] <= j_lo) ei = mid; else hi = mid; }
  for (; ei < n; ++ei) {
    const int64_t o_this = off[ei], o_next = off[ei + 1];
    if (o_this >= j_hi) break;
    if (o_next == o_this) continue;
    const ulonglong2 e = ent[ei];
    const uint32_t q = (uint32_t)(e.y >> 40);
    const uint32_t d = q / (uint32_t)qp;
    const uint64_t khead = ((uint64_t)(cnt_head[ei] != 0) << 63) |
                           ((uint64_t)(q - d * (uint32_t)qp) << (SIA_KEY_SONG_BITS + SIA_KEY_DIFF_BITS));
    const uint32_t qoff = (uint32_t)(e.x & kM24);
    const uint64_t *__restrict__ run = post + first[ei];
    uint64_t *__restrict__ dst = key_slots + (int64_t)d * cap_out + 1 - cut[d];     // key of tuple j goes to dst[j]
    const int64_t lim = cut[d] + cap_out - 1;                                        // first tuple that does not fit
    const int64_t a = max(j_lo, o_this), b = min(j_hi, o_next);
    for (int64_t j = a + lane; j < b; j += 32) {
      const uint64_t r = run[j - o_this];
      const uint64_t song = (r >> 24) & kM24;
      const uint64_t dbits = (uint64_t)((uint32_t)(r & kM24) - qoff + SIA_DIFF_BIAS);
      if (j < lim) dst[j] = khead | (song << SIA_KEY_DIFF_BITS) | dbits;
    }
  }
}

// entries are sorted by global query id: entry / tuple offsets at which every query starts, and its tuple count here
__global__ void query_bounds_kernel(const ulonglong2 *__restrict__ ent, int64_t n, const int64_t *__restrict__ off_all, int nq,
                                    int64_t *__restrict__ q_ent, int64_t *__restrict__ goff, int64_t *__restrict__ tuples) {
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q <= nq; q += gridDim.x * blockDim.x) {
    int64_t b[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      int64_t lo = 0, hi = n;
      while (lo < hi) { const int64_t mid = lo + ((hi - lo) >> 1); if ((int64_t)(ent[mid].y >> 40) < (int64_t)q + h) lo = mid + 1; else hi = mid; }
      b[h] = lo;
    }
    q_ent[q] = b[0];
    goff[q] = off_all[b[0]];
    if (q < nq) tuples[q] = off_all[b[1]] - off_all[b[0]];
  }
}

}  // namespace

extern "C" {

int sia_route_entries(int device, const uint8_t *d_hash, const int32_t *d_qoff, const int64_t *d_query_starts,
                      int32_t n_queries, int64_t n, int32_t qid_base, int32_t world, int64_t slot_cap, void *d_slots,
                      int32_t *d_status, void *stream) {
  SIA_REQUIRE(world >= 1 && world <= 1023 && slot_cap >= 2 && n >= 0 && n_queries >= 0, SIA_E_INVALID, "route: bad sizes");
  SIA_REQUIRE(d_slots && d_status && (n == 0 || (d_hash && d_qoff && d_query_starts)), SIA_E_INVALID, "NULL argument");
  SIA_CUDA(cudaSetDevice(device));
  cudaStream_t s = (cudaStream_t)stream;
  ulonglong2 *slots = static_cast<ulonglong2 *>(d_slots);
  init_entry_headers_kernel<<<ceil_div(world, 256), 256, 0, s>>>(slots, world, slot_cap);
  if (n)
    route_entries_kernel<<<grid_for(n), 256, 0, s>>>(d_hash, d_qoff, d_query_starts, n_queries, n, qid_base, world, slot_cap,
                                                   slots, d_status);
  SIA_CHECK_LAUNCH();
  return SIA_OK;
}

int sia_index_expand_slots(sia_index *ix, const void *d_entry_slots, int32_t world, int64_t entry_cap, int32_t queries_per_rank,
                           uint64_t *d_key_slots, int64_t key_cap, int64_t *d_info, void *stream) {
  SIA_REQUIRE(ix && d_entry_slots && d_key_slots && d_info, SIA_E_INVALID, "NULL argument");
  SIA_REQUIRE(ix->n_pending == 0, SIA_E_INVALID, "index has pending rows: call sia_index_finalize first");
  SIA_REQUIRE(world >= 1 && world <= 1023 && entry_cap >= 2 && key_cap >= 2 && queries_per_rank >= 1 &&
              queries_per_rank <= (1 << SIA_KEY_QID_BITS) && (int64_t)world * queries_per_rank < (1 << 24) - 1,
              SIA_E_INVALID, "expand_slots: bad sizes");
  SIA_CUDA(cudaSetDevice(ix->device));
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t n = (int64_t)world * (entry_cap - 1);
  int rc = ix->arena.reserve((size_t)n * 32 + lookup_bytes(n) + (size_t)(world + 1) * 8 + (1 << 20));
  if (rc) return rc;
  ulonglong2 *a = ix->arena.take<ulonglong2>(n), *b = ix->arena.take<ulonglong2>(n);
  int64_t *cut = ix->arena.take<int64_t>(world + 1);
  SIA_REQUIRE(a && b && cut, SIA_E_NOMEM, "index scratch arena too small (expand_slots)");
  gather_entries_kernel<<<grid_for(n), 256, 0, s>>>(static_cast<const ulonglong2 *>(d_entry_slots), world, entry_cap, a, d_info);
  SIA_CHECK_LAUNCH();
  Lookup L;
  if ((rc = lookup_sorted(ix, ix->arena, a, b, n, nullptr, 0, 0, 0, L, s))) return rc;
  cuts_kernel<<<1, 1024, 0, s>>>(L.ent, n, L.off_all, world, queries_per_rank, key_cap, cut, d_key_slots, d_info);
  if (L.tuples)
    expand_slots_kernel<<<(unsigned)ceil_div(L.tuples, kVoteTuples), 256, 0, s>>>(L.ent, n, L.first, L.off_all, L.cnt_head,
                                                                                  ix->post, cut, queries_per_rank, key_cap,
                                                                                  d_key_slots);
  SIA_CHECK_LAUNCH();
  SIA_CUDA(cudaStreamSynchronize(s));        // the lookup scratch is reused by the next call
  return SIA_OK;
}

int sia_vote_key_slots(int device, const uint64_t *d_key_slots, int32_t n_slots, int64_t key_cap, int32_t n_queries,
                       int32_t topn, int32_t max_song, int32_t *d_out_song, int32_t *d_out_diff, int32_t *d_out_count,
                       int32_t *d_out_rows, int32_t *d_out_nres, int32_t defer, void *stream) {
  SIA_REQUIRE(d_key_slots && key_cap >= 2, SIA_E_INVALID, "vote_key_slots: bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  SIA_CUDA(cudaSetDevice(device));
  if (n_queries > 0 && d_out_nres) {
    SIA_CUDA(cudaMemsetAsync(d_out_nres, 0, sizeof(int32_t) * n_queries, s));
    for (int32_t *o : {d_out_song, d_out_diff, d_out_count, d_out_rows})
      if (o) SIA_CUDA(cudaMemsetAsync(o, 0, sizeof(int32_t) * (size_t)n_queries * topn, s));
  }
  return vote_key_slots(device, d_key_slots, n_slots, key_cap, nullptr, n_queries, topn, max_song, d_out_song, d_out_diff,
                        d_out_count, d_out_rows, d_out_nres, s, defer);
}

int sia_vote_finish(int device) { return vote_key_slots_finish(device); }

// ---- hash-prefix sharding over peer memory (NVLink P2P) -----------------------------------------------------------
int sia_peer_alloc(int device, int64_t bytes, void **d_ptr, uint8_t *h_handle64) {
  SIA_REQUIRE(d_ptr && h_handle64 && bytes > 0, SIA_E_INVALID, "peer_alloc: bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  SIA_CUDA(cudaSetDevice(device));
  SIA_CUDA(cudaMalloc(d_ptr, (size_t)bytes));
  cudaIpcMemHandle_t h;
  SIA_CUDA(cudaIpcGetMemHandle(&h, *d_ptr));
  memcpy(h_handle64, &h, 64);
  return SIA_OK;
}

int sia_peer_open(int device, const uint8_t *h_handle64, void **d_ptr) {
  SIA_REQUIRE(d_ptr && h_handle64, SIA_E_INVALID, "peer_open: bad argument");
  SIA_CUDA(cudaSetDevice(device));
  cudaIpcMemHandle_t h;
  memcpy(&h, h_handle64, 64);
  SIA_CUDA(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return SIA_OK;
}

int sia_peer_close(int device, void *d_ptr) {
  if (!d_ptr) return SIA_OK;
  SIA_CUDA(cudaSetDevice(device));
  SIA_CUDA(cudaIpcCloseMemHandle(d_ptr));
  return SIA_OK;
}

int sia_peer_free(int device, void *d_ptr) {
  if (!d_ptr) return SIA_OK;
  SIA_CUDA(cudaSetDevice(device));
  SIA_CUDA(cudaFree(d_ptr));
  return SIA_OK;
}

int sia_index_lookup_slots(sia_index *ix, const void *d_entry_slots, int32_t world, int64_t entry_cap, int32_t queries_per_rank,
                           int64_t *d_tuples, int64_t *d_info, void *stream) {
  SIA_REQUIRE(ix && d_entry_slots && d_tuples && d_info, SIA_E_INVALID, "NULL argument");
  SIA_REQUIRE(ix->n_pending == 0, SIA_E_INVALID, "index has pending rows: call sia_index_finalize first");
  SIA_REQUIRE(world >= 1 && world <= kPvMaxPeers && entry_cap >= 2 && queries_per_rank >= 1 &&
              (int64_t)world * queries_per_rank < (1 << 24) - 1, SIA_E_INVALID, "lookup_slots: bad sizes");
  SIA_CUDA(cudaSetDevice(ix->device));
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t n = (int64_t)world * (entry_cap - 1);
  const int nq = world * queries_per_rank;
  int rc = ix->arena.reserve((size_t)n * (32 + 20) + lookup_bytes(n) + (size_t)(nq + 1) * 16 + (1 << 20));
  if (rc) return rc;
  ulonglong2 *a = ix->arena.take<ulonglong2>(n), *b = ix->arena.take<ulonglong2>(n);
  ix->dist_q_ent = ix->arena.take<int64_t>(nq + 1);
  ix->dist_goff = ix->arena.take<int64_t>(nq + 1);
  ix->dist_einfo = ix->arena.take<longlong2>(n);
  ix->dist_qh = ix->arena.take<uint32_t>(n);
  ix->dist_nq = 0;
  SIA_REQUIRE(a && b && ix->dist_q_ent && ix->dist_goff && ix->dist_einfo && ix->dist_qh, SIA_E_NOMEM,
              "index scratch arena too small (lookup_slots)");
  gather_entries_kernel<<<grid_for(n), 256, 0, s>>>(static_cast<const ulonglong2 *>(d_entry_slots), world, entry_cap, a, d_info);
  SIA_CHECK_LAUNCH();
  if ((rc = lookup_sorted(ix, ix->arena, a, b, n, nullptr, 0, 0, 0, ix->dist_L, s))) return rc;
  query_bounds_kernel<<<grid_for(nq + 1), 256, 0, s>>>(ix->dist_L.ent, n, ix->dist_L.off_all, nq, ix->dist_q_ent, ix->dist_goff, d_tuples);
  SIA_CHECK_LAUNCH();
  if ((rc = pvote_entry_info(ix->dist_L, ix->dist_einfo, ix->dist_qh, s))) return rc;
  ix->dist_nq = nq;
  return SIA_OK;
}

int sia_index_scatter_peers(sia_index *ix, int32_t world, int32_t rank, int32_t queries_per_rank, const int64_t *d_tuples_total,
                            void *const *h_peer_regions, void *const *h_peer_fill, void *const *h_peer_qover, int64_t region_cap,
                            int64_t fill_cap, int64_t *d_info, void *stream) {
  SIA_REQUIRE(ix && d_tuples_total && h_peer_regions && h_peer_fill && h_peer_qover && d_info, SIA_E_INVALID, "NULL argument");
  SIA_REQUIRE(ix->dist_nq == world * queries_per_rank && ix->dist_nq > 0, SIA_E_INVALID,
              "scatter_peers: call sia_index_lookup_slots for this pass first");
  SIA_CUDA(cudaSetDevice(ix->device));
  cudaStream_t s = (cudaStream_t)stream;
  const int nq = ix->dist_nq;
  const int64_t blocks = ceil_div(std::max<int64_t>(ix->dist_L.tuples, 1), 8192) + nq;
  int rc = ix->arena3.reserve((size_t)nq * 80 + (size_t)blocks * 4 + (1 << 16));
  if (rc) return rc;
  rc = pvote_scatter_peers(ix->arena3, ix->dist_L, ix->dist_einfo, ix->dist_qh, ix->post, ix->dist_q_ent, ix->dist_goff, world, rank,
                           queries_per_rank, d_tuples_total, h_peer_regions, h_peer_fill, h_peer_qover, region_cap, fill_cap, d_info, s);
  ix->dist_nq = 0;
  return rc;
}

int sia_vote_count_regions(int device, const int64_t *d_tuples_total, int32_t n_queries, int32_t topn, uint64_t *d_regions,
                           uint32_t *d_fill, uint32_t *d_qover, int64_t region_cap, int64_t fill_cap, int32_t *d_out_song,
                           int32_t *d_out_diff, int32_t *d_out_count, int32_t *d_out_rows, int32_t *d_out_nres, int64_t *d_info,
                           void *stream) {
  SIA_REQUIRE(d_tuples_total && d_regions && d_fill && d_qover && d_info, SIA_E_INVALID, "NULL argument");
  SIA_REQUIRE(n_queries >= 0 && topn >= 1 && topn <= kPvMaxTopn && fill_cap >= 1 && fill_cap < (1ll << 31), SIA_E_INVALID,
              "count_regions: bad sizes (topn <= 32)");
  if (n_queries == 0) return SIA_OK;
  SIA_REQUIRE(d_out_song && d_out_diff && d_out_count && d_out_rows && d_out_nres, SIA_E_INVALID, "NULL output");
  SIA_REQUIRE(device >= 0 && device < 64, SIA_E_INVALID, "device index");
  SIA_CUDA(cudaSetDevice(device));
  cudaStream_t s = (cudaStream_t)stream;
  SIA_CUDA(cudaMemsetAsync(d_out_nres, 0, sizeof(int32_t) * n_queries, s));
  for (int32_t *o : {d_out_song, d_out_diff, d_out_count, d_out_rows})
    SIA_CUDA(cudaMemsetAsync(o, 0, sizeof(int32_t) * (size_t)n_queries * topn, s));
  int rc = vote_scratch_finish_and_reserve(device, (size_t)n_queries * 64 + (size_t)fill_cap * (4 + 8 * (size_t)topn) + (1 << 16));
  if (rc) return rc;
  const PvOut po{d_out_song, d_out_diff, d_out_count, d_out_rows, d_out_nres};
  return pvote_count_regions(vote_arena(device), d_tuples_total, n_queries, topn, d_regions, d_fill, d_qover, region_cap, fill_cap, po,
                             d_info, reinterpret_cast<uint32_t *>(d_info + 1), s);
}

}  // extern "C"
