// placeholder: TrOCR encoder/decoder (filled in next)
#include "common.cuh"
void mb_free_trocr(mb_ctx* ctx) { (void)ctx; }
