// flk_dist.cu — multi-GPU strip decomposition (flockers_mpi protocol).  Placeholder until the NCCL path lands.
#include "../../include/flk.h"
#include "flk_internal.cuh"

namespace flk {
void dist_free(flk_field *) {}
int dist_set_objects(flk_field *, uint64_t, const uint32_t *, const float *, const float *) { return fail(FLK_E_UNSUPPORTED, "dist not built"); }
int dist_run(flk_field *, uint64_t, uint64_t, uint64_t) { return fail(FLK_E_UNSUPPORTED, "dist not built"); }
int dist_refresh_counts(flk_field *) { return fail(FLK_E_UNSUPPORTED, "dist not built"); }
}  // namespace flk

extern "C" {
int flk_dist_unique_id(uint8_t *) { return flk::fail(FLK_E_UNSUPPORTED, "dist not built"); }
int flk_dist_create(float, float, float, uint64_t, int, int, int, const uint8_t *, flk_field **) { return flk::fail(FLK_E_UNSUPPORTED, "dist not built"); }
int flk_dist_owner(const flk_field *, float, float, int32_t *) { return flk::fail(FLK_E_UNSUPPORTED, "dist not built"); }
int flk_dist_strip(const flk_field *, int32_t *, int32_t *) { return flk::fail(FLK_E_UNSUPPORTED, "dist not built"); }
int flk_dist_step(flk_field *, uint64_t, uint64_t, const float *, uint64_t) { return flk::fail(FLK_E_UNSUPPORTED, "dist not built"); }
int flk_dist_num_owned(flk_field *, uint64_t *) { return flk::fail(FLK_E_UNSUPPORTED, "dist not built"); }
}
