// LSD radix sort of fixed-size records (8 or 16 bytes) on selected key bytes — the one
// ordering primitive behind the index build (rows by (hash, song, offset)), the query
// grouping ((query, hash, offset)) and the vote ((query, song, diff) bins).
#pragma once
#include "sia_common.cuh"

namespace sia {

// bytes of scratch for n records (histograms + scanned offsets + scan temp)
size_t radix_sort_tmp_bytes(int64_t n);

// Sort n records of rec_bytes (8: uint64_t, 16: ulonglong2 with .x = low 64 bits) by the
// little-endian key bytes [byte_lo, byte_hi).  d_a holds the input; d_b is a same-size
// alternate buffer.  *result_in_b tells where the sorted data ended up.  Bytes that are
// identical in every record are skipped (one small D2H sync per call).  Stable.
int radix_sort(void *d_a, void *d_b, int64_t n, int rec_bytes, int byte_lo, int byte_hi, void *d_tmp,
               cudaStream_t s, bool *result_in_b);

}  // namespace sia
