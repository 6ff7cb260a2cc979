/*
 * marie_b200.h — C ABI of libmarie_b200.so, the B200 (sm_100a) implementation of Marie-AI's OCR hot
 * path: CRAFT text-box detection (marie.boxes) feeding TrOCR word/line recognition (marie.document).
 *
 * The reference has no FFI on this path (pure Python over torch / OpenCV / PIL / fairseq); the drop-in
 * boundary is the Python plugin API (BoxProcessor / OcrProcessor).  These entry points are
 * what our two plugin classes bind through ctypes; each one cites the reference code it replaces
 * (paths relative to the reference repo root).
 *
 * Conventions
 *   - plain C: pointers + sizes only.  Pointers named *_dev are device pointers (caller-owned, e.g.
 *     torch tensors); *_host are host pointers.  `stream` is a cudaStream_t passed as void*.
 *   - every function returns 0 on success or a negative MB_ERR_* code; mb_last_error(ctx) gives text.
 *     Nothing throws across the ABI.  There is NO CPU fallback: without a CUDA device mb_init fails.
 *   - one mb_ctx per (process, device); a context is not re-entrant.
 *   - all work is enqueued on `stream`; functions that return host-visible counts synchronise it.
 */
#ifndef MARIE_B200_H
#define MARIE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mb_ctx mb_ctx;

enum {
    MB_OK = 0,
    MB_ERR_ARG = -1,      /* bad argument / unsupported shape */
    MB_ERR_CUDA = -2,     /* CUDA runtime / driver error */
    MB_ERR_NO_DEVICE = -3,/* no sm_100 device: the product path refuses to run */
    MB_ERR_STATE = -4,    /* model not loaded, workspace too small, ... */
    MB_ERR_OOM = -5
};

/* ---- context -------------------------------------------------------------------------------- */
const char* mb_version(void);
/* Creates the per-device context (replaces the implicit `.cuda()` device state the reference keeps in
 * BoxProcessorCraft.__load, marie/boxes/craft_box_processor.py:260-315, and TrOcrProcessor.__init__,
 * marie/document/trocr_ocr_processor.py:191-246). */
int mb_init(int device, mb_ctx** out);
void mb_free(mb_ctx* ctx);
const char* mb_last_error(const mb_ctx* ctx);
/* Number of kernels this context has launched so far (bench.py "gpu_launches"). */
unsigned long long mb_launch_count(const mb_ctx* ctx);

/* ---- tensor-core building block --------------------------------------------------------------
 * out[M,N] = act(A[M,K] @ W[N,K]^T + bias) (+ residual); bf16 operands, fp32 accumulation in TMEM.
 * Replaces the cuBLAS calls behind nn.Linear in timm's ViT blocks (marie/models/unilm/trocr/deit.py:105-146)
 * and fairseq's TransformerDecoder (built at marie/models/unilm/trocr/trocr_models.py:142-147).
 * K must be a multiple of 64; lda/ldw/out_ld in elements. act: 0 none, 1 relu, 2 gelu(erf).
 * out_mode: 0 bf16, 1 fp32. */
int mb_gemm_bf16(mb_ctx* ctx, const void* a_dev, long long lda, const void* w_dev, int n_rows_w,
                 int M, int N, int K, const float* bias_dev, int act, const void* residual_dev,
                 long long res_ld, void* out_dev, long long out_ld, int out_mode, void* stream);

/* NHWC convolution as implicit GEMM (taps = 1: 1x1; taps = 9: 3x3 with padding = dilation).  Input is the
 * channel-concatenation of up to two NHWC tensors (the U-Net `torch.cat`, marie/models/craft/craft.py:64-77).
 * Weights are [n_rows_w, taps*(c0+c1)] with k = (ky*3+kx)*(c0+c1) + c.  Replaces cuDNN behind nn.Conv2d in
 * marie/models/craft/basenet/vgg16_bn.py:33-47 and marie/models/craft/craft.py:14-51.
 * out_mode 2 writes fp32 channel planes `out_plane` elements apart (the score maps). */
int mb_conv_bf16(mb_ctx* ctx, const void* a0_dev, int c0, int a0_ld, const void* a1_dev, int c1,
                 int a1_ld, int n, int h, int w, int taps, int dil, const void* w_dev, int n_rows_w,
                 int n_out, const float* bias_dev, int act, void* out_dev, long long out_ld,
                 int out_mode, long long out_plane, void* stream);

/* ---- K5-K7: score-map post-processing ---------------------------------------------------------
 * Replaces getDetBoxes_core / getDetBoxes / adjustResultCoordinates (marie/models/craft/craft_utils.py:25-98,
 * 257-274) and the box -> int rect conversion of BoxProcessorCraft.extract_bounding_boxes
 * (marie/boxes/craft_box_processor.py:499-521) for a batch of n_img score-map pairs [n_img, h, w] fp32.
 *   labels   [n_img, h, w]  i32   component ids in cv2.connectedComponentsWithStats order (0 = background)
 *   n_labels [n_img]        i32   number of labels including background
 *   stats    [n_img, max_labels, 5] i32  cv2 layout (left, top, width, height, area); row 0 is zeroed
 *   det      [n_img, max_boxes, 4, 2] f32  boxes in heat-map coordinates (getDetBoxes_core `det`)
 *   adj      [n_img, max_boxes, 4, 2] f32  boxes * (ratio_w*2, ratio_h*2) (adjustResultCoordinates)
 *   rects    [n_img, max_boxes, 4]  i32   (x, y, w, h) after int32 truncation, boundingRect, +4 px, clamps
 *   mapper   [n_img, max_boxes]     i32   label id of each box;  n_boxes [n_img] i32
 * ratios_dev: [n_img, 2] f64 = (ratio_w*2, ratio_h*2) or NULL (= 1); page_hw_dev: [n_img, 2] i32 page (H, W) used
 * for the clamps, or NULL.  Synchronises `stream` (reports capacity overflow as MB_ERR_STATE). */
int mb_craft_post(mb_ctx* ctx, const float* text_dev, const float* link_dev, int n_img, int h, int w,
                  float text_threshold, float link_threshold, float low_text, const double* ratios_dev,
                  const int32_t* page_hw_dev, int32_t* labels_dev, int32_t* n_labels_dev, int32_t* stats_dev,
                  int max_labels, float* det_dev, float* adj_dev, int32_t* rects_dev, int32_t* mapper_dev,
                  int32_t* n_boxes_dev, int max_boxes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MARIE_B200_H */
