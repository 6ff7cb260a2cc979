// Model teardown shared by mb_free (api.cu).
#include "common.cuh"
void mb_free_craft(mb_ctx* ctx);
void mb_free_trocr(mb_ctx* ctx);
void mb_free_refine(mb_ctx* ctx);
void mb_free_models(mb_ctx* ctx) {
    mb_free_craft(ctx);
    mb_free_trocr(ctx);
    mb_free_refine(ctx);
}
