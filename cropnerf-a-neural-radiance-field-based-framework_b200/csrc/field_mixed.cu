// FruitField on tensor cores (mixed precision) -- placeholder until the fused kernel lands.
#include "field_common.cuh"

bool cnb_field_mixed_supported(const cnb_field*) { return false; }

int cnb_field_mixed_fwd(const cnb_field*, const cnb_samples*, float*, float*, float*, float*, float*, cudaStream_t) {
  cnb_set_error("field mixed: not built");
  return CNB_ERR_UNSUPPORTED;
}

int cnb_field_mixed_bwd(const cnb_field*, const cnb_samples*, const float*, const float*, const float*, cudaStream_t) {
  cnb_set_error("field mixed: not built");
  return CNB_ERR_UNSUPPORTED;
}
