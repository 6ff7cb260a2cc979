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
/* 16-bit element type of every activation / weight buffer that crosses this ABI ("16-bit" below): fp16 (default)
 * or bf16; accumulation is always fp32.  fp16 is the reference's own TrOCR dtype (`.half()`,
 * marie/document/trocr_ocr_processor.py:75-76) and keeps the CRAFT score maps within 1e-2 of the fp32 reference
 * (bf16's 8-bit mantissa does not: DESIGN.md §precision).  Changing the dtype unloads any loaded weights. */
enum { MB_DTYPE_BF16 = 0, MB_DTYPE_F16 = 1 };
int mb_set_dtype(mb_ctx* ctx, int dtype);
int mb_get_dtype(const mb_ctx* ctx);
/* Live CUDA-event timing of the tensor-core GEMM launches on their launch stream (bench.py roofline):
 * mb_profile_enable(ctx, 1) resets and starts, mb_profile_read drains: out3 = {sum of launch durations in ms,
 * sum of algorithmic FLOPs, number of launches}. */
int mb_profile_enable(mb_ctx* ctx, int enable);
int mb_profile_read(mb_ctx* ctx, double* out3_host);
/* Number of kernels this context has launched so far (bench.py "gpu_launches"). */
unsigned long long mb_launch_count(const mb_ctx* ctx);

/* ---- tensor-core building block --------------------------------------------------------------
 * out[M,N] = act(A[M,K] @ W[N,K]^T + bias) (+ residual); 16-bit operands, fp32 accumulation in TMEM.
 * Replaces the cuBLAS calls behind nn.Linear in timm's ViT blocks (marie/models/unilm/trocr/deit.py:105-146)
 * and fairseq's TransformerDecoder (built at marie/models/unilm/trocr/trocr_models.py:142-147).
 * K must be a multiple of 64; lda/ldw/out_ld in elements. act: 0 none, 1 relu, 2 gelu(erf).
 * out_mode: 0 16-bit, 1 fp32. */
int mb_gemm16(mb_ctx* ctx, const void* a_dev, long long lda, const void* w_dev, int n_rows_w,
                 int M, int N, int K, const float* bias_dev, int act, const void* residual_dev,
                 long long res_ld, void* out_dev, long long out_ld, int out_mode, void* stream);

/* Block-diagonal variant: batch b uses A columns [b*a_col_stride, +K), weight rows [b*w_row_stride, +N) and writes
 * output / bias columns [b*out_col_stride, +N) — the per-head projections of the decoder's cross-attention. */
int mb_gemm16_batched(mb_ctx* ctx, const void* a_dev, long long lda, const void* w_dev, int n_rows_w, int M, int N, int K,
                      int batches, int a_col_stride, int w_row_stride, int out_col_stride, const float* bias_dev, int act,
                      void* out_dev, long long out_ld, void* stream);

/* NHWC convolution as implicit GEMM (taps = 1: 1x1; taps = 9: 3x3 with padding = dilation).  Input is the
 * channel-concatenation of up to two NHWC tensors (the U-Net `torch.cat`, marie/models/craft/craft.py:64-77).
 * Weights are [n_rows_w, taps*(c0+c1)] with k = (ky*3+kx)*(c0+c1) + c.  Replaces cuDNN behind nn.Conv2d in
 * marie/models/craft/basenet/vgg16_bn.py:33-47 and marie/models/craft/craft.py:14-51.
 * out_mode 2 writes fp32 channel planes `out_plane` elements apart (the score maps). */
int mb_conv16(mb_ctx* ctx, const void* a0_dev, int c0, int a0_ld, const void* a1_dev, int c1,
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

/* ---- K1: page preprocessing -------------------------------------------------------------------
 * Replaces imgproc.resize_aspect_ratio + normalizeMeanVariance (marie/models/craft/imgproc.py:45-73,26-32) and
 * the HWC->CHW / H2D pre-amble of get_prediction (marie/boxes/craft_box_processor.py:96-106).
 * pages: [n_pages, page_h, page_w, 3] u8 BGR on the device.  The image is resampled to (target_h, target_w) with
 * cv2's INTER_LINEAR u8 fixed-point arithmetic, pasted on a zero canvas (out_h, out_w) (multiples of 32),
 * normalised (v-127.5)/127.5 and written as NHWC 16-bit with C padded to 4: out [n_pages, out_h, out_w, 4]. */
int mb_page_preprocess(mb_ctx* ctx, const uint8_t* pages_dev, int n_pages, int page_h, int page_w,
                       int target_h, int target_w, int out_h, int out_w, void* out_dev, void* stream);

/* ---- K9: crop -> 384x384 network input ----------------------------------------------------------
 * Replaces crop_poly_low on the expanded rect (marie/boxes/craft_box_processor.py:42-73,524),
 * MemoryDataset.__getitem__ (BGR->RGB; marie/models/icr/memory_dataset.py:43-53) and preprocess_image /
 * preprocess_samples (PIL bicubic 384x384, ToTensor, Normalize(0.5,0.5), stack;
 * marie/document/trocr_ocr_processor.py:95-101,116-139).
 * mb_pack_crops: crop i = pages[page_idx[i]][y:y+h+1, x:x+w+1] for rects[i] = (x,y,w,h) (numpy clipping).
 * mb_pack_fragments: crop i = the [h_i, w_i, 3] u8 BGR image at buf + offsets[i] (hw = [n,2] (h,w)).
 * layout 0: out [n, 3, 384, 384] 16-bit (RGB planes); layout 1: out [n*576, 768] 16-bit patch rows
 * (k = c*256 + py*16 + px), the A operand of the ViT patch embedding.  Both synchronise `stream`. */
int mb_pack_crops(mb_ctx* ctx, const uint8_t* pages_dev, int page_h, int page_w, const int32_t* rects_dev,
                  const int32_t* page_idx_dev, int n_crops, void* out_dev, int layout, void* stream);
int mb_pack_fragments(mb_ctx* ctx, const uint8_t* buf_dev, const long long* offsets_dev, const int32_t* hw_dev,
                      int n_crops, void* out_dev, int layout, void* stream);

/* ---- K2-K4: CRAFT network ------------------------------------------------------------------------
 * mb_load_craft: takes the flat blob produced by marie-icr_b200/weights.py:pack_craft from the reference's own
 * state dict (replaces CRAFT() + load_state_dict + .cuda() in BoxProcessorCraft.__load,
 * marie/boxes/craft_box_processor.py:260-285; DataParallel is dropped: SURVEY.md C1).
 * mb_craft_forward: replaces CRAFT.forward (marie/models/craft/craft.py:59-81).  x: [n, h, w, 4] 16-bit NHWC
 * (mb_page_preprocess output), h and w multiples of 32.  scores: [2, n, h/2, w/2] fp32 — plane 0 = text/region,
 * plane 1 = link/affinity (y[..., 0] / y[..., 1] of the reference).  feature_dev (optional, may be NULL):
 * [n, h/2, w/2, 64] 16-bit, channels 0..31 = the reference's `feature`, 32..63 zero. */
int mb_load_craft(mb_ctx* ctx, const void* blob_host, size_t nbytes);
int mb_craft_forward(mb_ctx* ctx, const void* x_dev, int n, int h, int w, float* scores_dev, void* feature_dev,
                     void* stream);

/* ---- section 8f rank 2: the link refiner and its line branch ------------------------------------------------------
 * mb_load_refine: flat blob of marie-icr_b200/weights.py:pack_refine from the reference's RefineNet state dict
 * (marie/models/craft/refinenet.py:15-55; loading: marie/boxes/craft_box_processor.py:287-312).
 * mb_refine_forward: replaces RefineNet.forward (refinenet.py:57-66) as called by get_prediction
 * (craft_box_processor.py:116-120, 150-156).  feature_dev: [n, h, w, 64] 16-bit as written by mb_craft_forward — its
 * unused channels 32 / 33 are OVERWRITTEN with the two score maps (the concatenated 34-channel input is never
 * materialised).  scores_dev: [2, n, h, w] fp32 (mb_craft_forward output).  link_out_dev: [n, h, w] fp32 =
 * y_refiner[..., 0].
 * mb_line_components: replaces the threshold / MORPH_CLOSE / connectedComponentsWithStats block of the line branch
 * (craft_box_processor.py:161-205): link > link_threshold (strict, float32), 3x3 closing, 4-connected labelling.
 * labels_dev [n, h, w] i32 (optional, may be NULL); n_labels_dev [n] (background included); stats_dev
 * [n, max_labels, 5] i32 = cv2's (left, top, width, height, area), rows >= n_labels and row 0 zero.  The boxes
 * (left, top, width, height) of labels 1.. feed mb_line_merge; scaling to page coordinates is the caller's
 * int(v * ratio * 2) (craft_box_processor.py:207-217).  Synchronises `stream`. */
int mb_load_refine(mb_ctx* ctx, const void* blob_host, size_t nbytes);
int mb_refine_forward(mb_ctx* ctx, void* feature_dev, const float* scores_dev, int n, int h, int w, float* link_out_dev,
                      void* stream);
int mb_line_components(mb_ctx* ctx, const float* link_dev, int n, int h, int w, float link_threshold,
                       int32_t* labels_dev, int32_t* n_labels_dev, int32_t* stats_dev, int max_labels, void* stream);

/* ---- K8: line grouping (host side: a few thousand boxes, O(n^2) integer work) -----------------------------------
 * mb_line_merge replaces line_merge / __line_merge (marie/boxes/line_processor.py:48-171): boxes [n,4] i32 (x,y,w,h) ->
 * lines [<= n, 4] i32 sorted by y.  mb_find_line_numbers replaces find_line_number (:15-45): 1-based line id per box,
 * -1 when there are no lines.  Both over find_overlap_vertical (marie/utils/overlap.py:42-103).  Host pointers. */
int mb_line_merge(mb_ctx* ctx, const int32_t* boxes_host, int n, int32_t* lines_out_host, int* n_lines);
int mb_find_line_numbers(mb_ctx* ctx, const int32_t* lines_host, int n_lines, const int32_t* boxes_host, int n_boxes,
                         int32_t* ids_out_host);

/* ---- K10-K12: TrOCR recogniser ----------------------------------------------------------------------
 * mb_load_trocr: flat blob from marie-icr_b200/weights.py:pack_trocr built from the fairseq state dict
 * (`encoder.deit.*`, `decoder.*`) — replaces checkpoint_utils.load_model_ensemble_and_task + .half().to(device) in
 * init() (marie/document/trocr_ocr_processor.py:36-113).  mb_trocr_dims: {enc_dim, dec_dim, vocab, tokens}.
 * mb_trocr_encode: replaces TrOCREncoder.forward / forward_features (marie/models/unilm/trocr/trocr_models.py:508-524,
 * deit.py:105-146).  patches: [n*576, 768] 16-bit patch rows (mb_pack_* layout 1); enc_out: [n*577, enc_dim] 16-bit.
 * mb_trocr_decode: replaces TextRecognitionGenerator._generate (marie/models/unilm/trocr/generator.py:11-374) for
 * beam in 1..8 (1 = greedy): tokens_out [n, out_ld] i32 = best hypothesis including its final EOS (id 2), padded with
 * 1; lengths [n]; scores [n] = sum of log-probs / length.  max_len_b as in task.py:266 (200).  steps_run (host,
 * optional) receives the number of decoder steps executed.  Synchronises `stream`.
 * mb_trocr_forced_logits: parity hook — teacher-forced decoder, logits_out [L, n, vocab] fp32.
 * mb_trocr_recognize: encode + decode over n crops in chunks (0 = default 512 crops per chunk). */
int mb_load_trocr(mb_ctx* ctx, const void* blob_host, size_t nbytes);
/* Test hook: row-wise LayerNorm (fp32 statistics) of a [rows, D] 16-bit matrix, D % 8 == 0, D <= 1024. */
int mb_layernorm16(mb_ctx* ctx, const void* in_dev, void* out_dev, const float* gamma_dev, const float* beta_dev,
                   long long rows, int D, float eps, void* stream);
/* Test hook for the two attention kernels: softmax(Q K^T * scale) V over a packed qkv buffer [n*T, 3*D] (heads of 64)
 * -> out [n*T, D].  mode 0 = tcgen05/TMEM kernel (encoder), 1 = mma.sync flash kernel (decoder cross-attention). */
int mb_attention16(mb_ctx* ctx, const void* qkv_dev, void* out_dev, int n, int T, int D, float scale, int mode,
                   void* stream);
/* Test hook for the residual GEMM that leaves the next LayerNorm's row statistics behind (encoder proj / fc2 in front of
 * the folded LayerNorms, marie/models/unilm/trocr/deit.py:105-146): out = a W^T + bias + residual, stats_out [M, 2] fp32 =
 * (-mean, rstd) of the written rows.  part_ws: M * 4 * ceil(N / 256) floats of scratch. */
int mb_gemm16_res_stats(mb_ctx* ctx, const void* a_dev, const void* w_dev, const float* bias_dev, const void* residual_dev,
                        void* out_dev, long long M, int N, int K, float eps, float* part_ws_dev, float* stats_out_dev,
                        void* stream);
/* Test hook for the greedy cross-attention core (the per-step attention of fairseq's decoder layers over the encoder
 * states, marie/models/unilm/trocr/trocr_models.py:142-147, with the K / V projections hoisted out): qp [rows, heads*E]
 * per-head projected queries, enc [rows*T, E] -> out [rows, heads*E].  mode 0 = tcgen05 / TMA kernel, 1 = mma.sync
 * kernel.  rows = crops; beam > 1 (mode 0 only): qp / out hold rows * beam hypotheses ([crop][beam]) that share the crop's
 * encoder states in one pass.  finished (or null): crops to skip; live_ws: rows + 1 ints of scratch (mode 0 with a mask). */
int mb_cross_enc16(mb_ctx* ctx, const void* qp_dev, const void* enc_dev, void* out_dev, int rows, int beam, int T, int heads,
                   int E, const unsigned char* finished_dev, int32_t* live_ws_dev, int mode, void* stream);
int mb_trocr_dims(mb_ctx* ctx, int* dims4_host);
/* cumulative search statistics: {mb_trocr_decode calls, decoder steps executed, rows (crops * beam) decoded} */
int mb_trocr_stats(mb_ctx* ctx, unsigned long long* out3_host);
int mb_trocr_encode(mb_ctx* ctx, const void* patches_dev, int n, void* enc_out_dev, void* stream);
int mb_trocr_decode(mb_ctx* ctx, const void* enc_out_dev, int n, int beam, int max_len_b, int32_t* tokens_out_dev,
                    int out_ld, int32_t* lengths_dev, float* scores_dev, int* steps_run, void* stream);
int mb_trocr_forced_logits(mb_ctx* ctx, const void* enc_out_dev, int n, const int32_t* forced_dev, int L,
                           float* logits_out_dev, void* stream);
int mb_trocr_recognize(mb_ctx* ctx, const void* patches_dev, int n, int beam, int max_len_b, int chunk,
                       int32_t* tokens_out_dev, int out_ld, int32_t* lengths_dev, float* scores_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MARIE_B200_H */
