// K4 placeholder — replaced by the real index in the next milestone.
#include "sia_common.cuh"
using namespace sia;
#define NOTYET() do { set_error("index: not implemented yet"); return SIA_E_UNSUPPORTED; } while (0)
extern "C" {
int sia_index_create(int, int64_t, sia_index **) { NOTYET(); }
int sia_index_destroy(sia_index *) { return SIA_OK; }
int sia_index_insert(sia_index *, int32_t, const uint8_t *, const int32_t *, int64_t, void *) { NOTYET(); }
int sia_index_insert_rows(sia_index *, const int32_t *, const uint8_t *, const int32_t *, int64_t, void *) { NOTYET(); }
int sia_index_insert_host(sia_index *, int32_t, const uint8_t *, const int32_t *, int64_t) { NOTYET(); }
int sia_index_finalize(sia_index *, int64_t *) { NOTYET(); }
int64_t sia_index_rows(const sia_index *) { return 0; }
int sia_index_select_host(sia_index *, const uint8_t *, int64_t, int32_t *, int32_t *, int32_t *, int64_t, int64_t *) { NOTYET(); }
int sia_index_query_batch(sia_index *, const uint8_t *, const int32_t *, const int64_t *, int32_t, int32_t, int32_t *,
                          int32_t *, int32_t *, int32_t *, int32_t *, int64_t *, void *) { NOTYET(); }
int sia_index_query_partial(sia_index *, const uint8_t *, const int32_t *, const int32_t *, int64_t, uint64_t *, int32_t *,
                            int64_t, int64_t *, uint64_t *, int32_t *, int64_t, int64_t *, void *) { NOTYET(); }
int sia_vote_bins(int, const uint64_t *, const int32_t *, int64_t, const uint64_t *, const int32_t *, int64_t, int32_t,
                  int32_t, int32_t *, int32_t *, int32_t *, int32_t *, int32_t *, void *) { NOTYET(); }
}
