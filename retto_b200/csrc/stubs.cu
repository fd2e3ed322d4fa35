// stubs.cu — entry points not implemented yet return ERR_UNSUPPORTED (removed as stages land)
#include "common.cuh"
#define STUB(name, ...) extern "C" retto_b200_status name(__VA_ARGS__) { return RETTO_B200_ERR_UNSUPPORTED; }
STUB(retto_b200_crop_boxes, retto_b200_ctx*, const retto_b200_crop_job*, int32_t, retto_b200_crop_info*)
STUB(retto_b200_crop_fetch, retto_b200_ctx*, int32_t, uint8_t*)
STUB(retto_b200_plan_batches, const retto_b200_config*, int32_t, const retto_b200_crop_info*, int32_t, int32_t*, retto_b200_batch*, int32_t*, uint64_t*)
STUB(retto_b200_build_batches, retto_b200_ctx*, int32_t, const retto_b200_line_job*, int32_t, uint64_t, float**)
STUB(retto_b200_cls_postprocess, retto_b200_ctx*, const float*, int32_t, const int32_t*, retto_b200_cls_result*)
STUB(retto_b200_run_pages, retto_b200_ctx*, const retto_b200_page*, int32_t, retto_b200_forward_fn, void*, retto_b200_results*)
