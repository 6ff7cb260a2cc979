// placeholder until craft.cu / trocr.cu define the model teardown
#include "common.cuh"
void mb_free_models(mb_ctx* ctx) { (void)ctx; }
