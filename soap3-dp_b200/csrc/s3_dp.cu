// s3_dp.cu -- placeholder until the DP kernels land (next commit)
#include "s3_common.cuh"
#include "../../include/soap3dp_b200.h"
extern "C" int s3_dp_create(uint32_t, uint32_t, uint32_t, s3_dp_scores, int, s3_dp **) { s3_set_error("DP not built yet"); return S3_EINVAL; }
extern "C" void s3_dp_free(s3_dp *) {}
extern "C" void *s3_dp_stream(const s3_dp *) { return NULL; }
extern "C" uint32_t s3_dp_pattern_length(const s3_dp *) { return 0; }
extern "C" int s3_dp_align(s3_dp *, const uint32_t *, const uint32_t *, const uint32_t *, const uint32_t *, const int32_t *, int32_t *, uint32_t *, uint32_t *, uint8_t *, uint32_t, const uint32_t *, uint32_t *, const uint32_t *, const uint32_t *) { return S3_EINVAL; }
extern "C" int s3_dp_align_device(s3_dp *, const uint32_t *, const uint32_t *, const uint32_t *, const uint32_t *, const int32_t *, int32_t *, uint32_t *, uint32_t *, uint8_t *, uint32_t, const uint32_t *, uint32_t *, const uint32_t *, const uint32_t *) { return S3_EINVAL; }
