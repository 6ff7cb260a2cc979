// K10-K12: TrOCR recogniser — ViT encoder, transformer decoder with a device-resident KV cache, and fairseq's
// beam / greedy search run entirely on the device.
//
// Reference (paths relative to the reference repo):
//   encoder  AdaptedVisionTransformer.forward_features  marie/models/unilm/trocr/deit.py:105-146 (timm 0.6.12 blocks:
//            pre-LN, qkv without bias, softmax(QK^T d^-0.5)V, erf-GELU MLP, LN eps 1e-6), TrOCREncoder.forward
//            marie/models/unilm/trocr/trocr_models.py:508-524
//   decoder  fairseq TransformerDecoder built at trocr_models.py:142-147 with the arch table :423-447 (post-LN, ReLU,
//            sinusoidal positions, embed scale sqrt(d), bias-free output projection, cross-attn kdim = encoder dim)
//   search   TextRecognitionGenerator._generate  marie/models/unilm/trocr/generator.py:11-374 (fairseq BeamSearch.step,
//            finalize_hypos): log-softmax in fp32, pad/min-len/max-len masking, top 2*beam candidates, EOS finalisation,
//            length-normalised scores, beam reorder of the incremental state
//
// All dense layers run on the tcgen05 tap-GEMM (gemm_tc.cu) with fused bias / GELU / ReLU / residual epilogues.  The
// kernels in this file are the glue that is not a GEMM: LayerNorm, token assembly, the 577-token encoder attention,
// the decode-step attentions over the KV caches, the fused log-softmax + top-k and the search bookkeeping.  The beam
// reorder never moves KV data: an ancestor table (step x row) redirects each hypothesis to the rows that hold its
// history (generator.py:139 reorder_incremental_state copies the whole cache instead).
#include "common.cuh"
#include "blob.cuh"
#include <math.h>
#include <stdlib.h>

int mb_attention_tc(mb_ctx* ctx, const bf16* qkv, bf16* out, int n, int T, int D, int heads, float scale_log2e,
                    cudaStream_t stream);
// xattn_tc.cu: greedy cross-attention over the encoder states on tcgen05 / TMA
bool mb_cross_enc_tc_supported(int E, int heads);
int mb_cross_enc_tc_groups(int E, int beam);
int mb_live_list(mb_ctx* ctx, const unsigned char* finished, int n, int* live_ws, cudaStream_t s);
int mb_cross_enc_tc(mb_ctx* ctx, const bf16* qp, const bf16* enc, bf16* out, int n, int beam, int T, int heads, int E,
                    const int* live_ws, cudaStream_t s);

namespace {

constexpr int DH = 64;            // head dim of every TrOCR variant (768/12, 1024/16)
constexpr int MAX_BEAM = 8;
constexpr int TOK_PAD = 1, TOK_EOS = 2;

struct EncLayer {
    const float *ln1_w, *ln1_b, *ln2_w, *ln2_b, *proj_b, *fc1_b, *fc2_b;
    const bf16 *qkv_w, *proj_w, *fc1_w, *fc2_w;
    // LayerNorm folded into the following GEMM (weights.py pack_trocr): W * gamma, its row sums, b + W beta
    const bf16 *qkv_wf = nullptr, *fc1_wf = nullptr;
    const float *qkv_c = nullptr, *qkv_bf = nullptr, *fc1_c = nullptr, *fc1_bf = nullptr;
};
struct DecLayer {
    const bf16 *sqkv_w, *sout_w, *cq_w, *ckv_w, *ckT_w, *cout_w, *fc1_w, *fc2_w;
    const float *sqkv_b, *sout_b, *cq_b, *ckv_b, *cout_b, *fc1_b, *fc2_b;
    const float *ln1_w, *ln1_b, *ln2_w, *ln2_b, *ln3_w, *ln3_b;
};

}  // namespace

struct TrocrModel {
    WeightBlob blob;
    int enc_dim = 0, enc_layers = 0, enc_heads = 0, enc_ffn = 0;
    int dec_dim = 0, dec_layers = 0, dec_heads = 0, dec_ffn = 0;
    int vocab = 0, tokens = 0, max_pos = 0;
    const bf16* patch_w = nullptr; const float* patch_b = nullptr; const float* cls_pos = nullptr;
    const float *norm_w = nullptr, *norm_b = nullptr;
    std::vector<EncLayer> enc;
    std::vector<DecLayer> dec;
    const bf16* embed = nullptr; const float* pe = nullptr; const bf16* out_w = nullptr;
    void* arena = nullptr; size_t arena_bytes = 0;
    // self-attention K / V caches of the decode in flight: their own allocation, sized for `kv_cap` steps and grown on demand
    // (a cache for max_len = 200 steps is 81 GB at 8192 rows; hypotheses are a handful of tokens long)
    void* kv_arena = nullptr; size_t kv_bytes = 0;
    void* rec_enc = nullptr; size_t rec_enc_bytes = 0;      // mb_trocr_recognize: encoder states of one chunk (kept between calls)
    unsigned long long decode_calls = 0, decode_steps = 0, decode_rows = 0;
    bool ln_fold = true;          // encoder LayerNorms folded into qkv / fc1 when the blob carries the folded tensors (MB_LNFOLD=0 disables)
};

namespace {

// ------------------------------------------------------------------------------------------------ LayerNorm
// One warp per row (grid-stride over rows); D % 8 == 0, D <= 1024.  in/out 16-bit (may alias), gamma/beta fp32, statistics in fp32
// (two-pass: mean, then centred variance — the order torch's CPU kernel uses up to summation order).
constexpr int LN_MAXV = 4;      // D <= 1024
// gamma / beta live in shared memory (fp32), each warp strides over rows and keeps the NEXT row's 16-byte loads in
// flight while it reduces the current one.  History: re-reading the affine parameters through L1 for every row and one
// row in flight per warp left the kernel latency-bound at 2.7 TB/s (9 % of a bench step, profiles/r01_launches_bench_summary.md).
__global__ void __launch_bounds__(256) layernorm_kernel(const bf16* __restrict__ in, bf16* __restrict__ out,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        long long rows, int D, float eps, int f16) {
    __shared__ __align__(16) float sg[1024], sb[1024];
    for (int i = threadIdx.x; i < D; i += blockDim.x) { sg[i] = gamma[i]; sb[i] = beta[i]; }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int nv = D >> 3;
    const long long wstride = (long long)gridDim.x * (blockDim.x >> 5);
    long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    uint4 nxt[LN_MAXV];
    if (row < rows) {
        const uint4* src = reinterpret_cast<const uint4*>(in + row * D);
#pragma unroll
        for (int i = 0; i < LN_MAXV; ++i)
            if (lane + 32 * i < nv) nxt[i] = src[lane + 32 * i];
    }
    for (; row < rows; row += wstride) {
        uint4 cur[LN_MAXV];
#pragma unroll
        for (int i = 0; i < LN_MAXV; ++i) cur[i] = nxt[i];
        if (row + wstride < rows) {
            const uint4* src = reinterpret_cast<const uint4*>(in + (row + wstride) * D);
#pragma unroll
            for (int i = 0; i < LN_MAXV; ++i)
                if (lane + 32 * i < nv) nxt[i] = src[lane + 32 * i];
        }
        float v[LN_MAXV][8];
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < LN_MAXV; ++i) {
            if (lane + 32 * i < nv) {
                const uint32_t w[4] = {cur[i].x, cur[i].y, cur[i].z, cur[i].w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float2 f = unpack2(w[k], f16);
                    v[i][2 * k] = f.x; v[i][2 * k + 1] = f.y;
                    sum += f.x + f.y;
                }
            }
        }
        const float mean = warp_sum(sum) / (float)D;
        float sq = 0.f;
#pragma unroll
        for (int i = 0; i < LN_MAXV; ++i) {
            if (lane + 32 * i < nv) {
#pragma unroll
                for (int k = 0; k < 8; ++k) { const float d = v[i][k] - mean; sq += d * d; }
            }
        }
        const float rstd = rsqrtf(warp_sum(sq) / (float)D + eps);
        uint4* dst = reinterpret_cast<uint4*>(out + row * D);
#pragma unroll
        for (int i = 0; i < LN_MAXV; ++i) {
            const int vi = lane + 32 * i;
            if (vi < nv) {
                const float4 g0 = *reinterpret_cast<const float4*>(sg + vi * 8), g1 = *reinterpret_cast<const float4*>(sg + vi * 8 + 4);
                const float4 b0 = *reinterpret_cast<const float4*>(sb + vi * 8), b1 = *reinterpret_cast<const float4*>(sb + vi * 8 + 4);
                uint4 o;
                o.x = pack2((v[i][0] - mean) * rstd * g0.x + b0.x, (v[i][1] - mean) * rstd * g0.y + b0.y, f16);
                o.y = pack2((v[i][2] - mean) * rstd * g0.z + b0.z, (v[i][3] - mean) * rstd * g0.w + b0.w, f16);
                o.z = pack2((v[i][4] - mean) * rstd * g1.x + b1.x, (v[i][5] - mean) * rstd * g1.y + b1.y, f16);
                o.w = pack2((v[i][6] - mean) * rstd * g1.z + b1.z, (v[i][7] - mean) * rstd * g1.w + b1.w, f16);
                dst[vi] = o;
            }
        }
    }
}

// x[n, 0] = cls + pos[0];  x[n, 1+p] = patch_out[n, p] + pos[1+p]   (deit.py:121-143; cls folded into cls_pos[0])
__global__ void assemble_tokens_kernel(const bf16* __restrict__ patch_out, const float* __restrict__ cls_pos,
                                       bf16* __restrict__ x, long long total8, int T, int D, int f16) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const int d8 = D >> 3;
    for (; i < total8; i += stride) {
        const int c = (int)(i % d8);
        const long long r = i / d8;
        const int t = (int)(r % T);
        const long long n = r / T;
        float f[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (t > 0) {
            const uint4 u = __ldg(reinterpret_cast<const uint4*>(patch_out + ((n * (T - 1) + t - 1) * D)) + c);
            const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) { const float2 p = unpack2(w[k], f16); f[2 * k] = p.x; f[2 * k + 1] = p.y; }
        }
        const float4 p0 = __ldg(reinterpret_cast<const float4*>(cls_pos + (long long)t * D) + 2 * c);
        const float4 p1 = __ldg(reinterpret_cast<const float4*>(cls_pos + (long long)t * D) + 2 * c + 1);
        uint4 o;
        o.x = pack2(f[0] + p0.x, f[1] + p0.y, f16); o.y = pack2(f[2] + p0.z, f[3] + p0.w, f16);
        o.z = pack2(f[4] + p1.x, f[5] + p1.y, f16); o.w = pack2(f[6] + p1.z, f[7] + p1.w, f16);
        reinterpret_cast<uint4*>(x + r * D)[c] = o;
    }
}

// ------------------------------------------------------------------------------------------------ encoder attention
// softmax(Q K^T * scale) V, flash-style (online softmax), one CTA = 64 query rows of one
// (image, head); 4 warps x 16 rows; K/V streamed in 64-key tiles through double-buffered cp.async.  Tensor math is
// mma.sync m16n8k16 (legacy HMMA path) — 11% of the encoder FLOPs; the tcgen05 version is future work.
// qkv: [n*T, 3*D] (q | k | v, head h at columns h*64); out: [n*T, D].
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
template <bool F16>
__device__ __forceinline__ void mma16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
    if (F16)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// tile [64 rows][64 x 16-bit] with the 16-byte chunk index XOR-swizzled by (row & 7)
__device__ __forceinline__ int swz(int row, int chunk) { return row * 128 + ((chunk ^ (row & 7)) << 4); }

// Generic form: batch item b has Tq query rows (q + (b*Tq + t)*q_ld + head*64) and Tk key/value rows
// (k|v + (b*Tk + t)*kv_ld + head*64); out + (b*Tq + t)*o_ld + head*64.  The encoder uses it with q/k/v inside one
// qkv buffer (Tq = Tk = 577); the decoder's cross-attention uses Tq = beam (<= 8) query rows per crop against the
// crop's 577 cached encoder keys/values, which turns the step into a streaming read of the K/V cache with cp.async
// keeping ~32 KB in flight per CTA (HBM-bound: 2.36 MB per crop and layer).
constexpr int ATT_STAGES = 3;                       // K/V tiles in flight per CTA (cp.async groups)
constexpr int ATT_SMEM = (1 + 2 * ATT_STAGES) * 64 * 128;   // Q tile + STAGES x (K tile + V tile) = 56 KB
template <bool F16>
__global__ void __launch_bounds__(128, 4) attention_kernel(const bf16* __restrict__ qptr, long long q_ld,
                                                        const bf16* __restrict__ kptr, const bf16* __restrict__ vptr,
                                                        long long kv_ld, bf16* __restrict__ out, long long o_ld, int Tq,
                                                        int T, float scale_log2e) {
    extern __shared__ __align__(128) unsigned char att_smem[];
    unsigned char* sQ = att_smem;
    unsigned char* sKbase = att_smem + 64 * 128;
    unsigned char* sVbase = sKbase + ATT_STAGES * 64 * 128;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q0 = blockIdx.x * 64, head = blockIdx.y;
    const long long img = blockIdx.z;
    const bf16* qbase = qptr + img * Tq * q_ld + head * DH;
    const bf16* kbase = kptr + img * T * kv_ld + head * DH;
    const bf16* vbase = vptr + img * T * kv_ld + head * DH;
    const int n_tiles = (T + 63) >> 6;
    const bool warp_active = q0 + warp * 16 < Tq;

    auto load_tile = [&](unsigned char* dst, const bf16* src, long long ld, int row0, int limit) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int idx = tid + 128 * i;            // 512 chunks of 16 B
            const int r = idx >> 3, c = idx & 7;
            const bool ok = row0 + r < limit;
            cp_async16(smem_addr(dst + swz(r, c)), src + (long long)(ok ? row0 + r : 0) * ld + c * 8, ok);
        }
    };
    load_tile(sQ, qbase, q_ld, q0, Tq);
#pragma unroll
    for (int st = 0; st < ATT_STAGES - 1; ++st) {     // prologue: tiles 0 .. STAGES-2 (one commit group per tile)
        if (st < n_tiles) {
            load_tile(sKbase + st * 64 * 128, kbase, kv_ld, st * 64, T);
            load_tile(sVbase + st * 64 * 128, vbase, kv_ld, st * 64, T);
        }
        cp_async_commit();
    }

    uint32_t aq[4][4];
    float o[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

    for (int j = 0; j < n_tiles; ++j) {
        const int buf = j % ATT_STAGES;
        unsigned char* sKb = sKbase + buf * 64 * 128;
        unsigned char* sVb = sVbase + buf * 64 * 128;
        {   // prefetch tile j + STAGES-1 into the buffer freed by tile j-1 (all warps passed the trailing barrier)
            const int jn = j + ATT_STAGES - 1;
            if (jn < n_tiles) {
                const int bn = jn % ATT_STAGES;
                load_tile(sKbase + bn * 64 * 128, kbase, kv_ld, jn * 64, T);
                load_tile(sVbase + bn * 64 * 128, vbase, kv_ld, jn * 64, T);
            }
            cp_async_commit();                         // always commit: keeps the group count uniform
        }
        cp_async_wait<ATT_STAGES - 1>();               // tile j (and Q) have landed
        __syncthreads();
        // warps whose 16 query rows are all beyond Tq (decoder cross-attention: Tq = beam <= 8, so warps 1..3) only
        // help streaming K/V; their HMMAs on zero rows used to cap the kernel at the legacy tensor pipe's rate
        if (warp_active) {
        if (j == 0) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                const int r = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                const int c = ks * 2 + (lane >> 4);
                ldsm_x4(smem_addr(sQ + swz(r, c)), aq[ks][0], aq[ks][1], aq[ks][2], aq[ks][3]);
            }
        }
        // S = Q K^T  (16 x 64 per warp)
        float s[8][4];
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
            s[nb][0] = s[nb][1] = s[nb][2] = s[nb][3] = 0.f;
#pragma unroll
            for (int kp = 0; kp < 2; ++kp) {          // two k-steps per ldmatrix.x4
                uint32_t b0, b1, b2, b3;
                const int r = nb * 8 + (lane & 7);
                const int c = kp * 4 + (lane >> 3);
                ldsm_x4(smem_addr(sKb + swz(r, c)), b0, b1, b2, b3);
                mma16816<F16>(s[nb], aq[kp * 2], b0, b1);
                mma16816<F16>(s[nb], aq[kp * 2 + 1], b2, b3);
            }
        }
        // online softmax (rows g and g+8 of this warp's 16) on the raw scores: the scale is folded into the exp2
        // argument (one FFMA + one MUFU.EX2 per element), keys are masked only in the ragged last tile, and the
        // accumulator is rescaled only when some row's running maximum actually moved
        if (j == n_tiles - 1) {
            const int key0 = j * 64 + (lane & 3) * 2;
#pragma unroll
            for (int nb = 0; nb < 8; ++nb) {
                const int k = key0 + nb * 8;
                if (k >= T) { s[nb][0] = -INFINITY; s[nb][2] = -INFINITY; }
                if (k + 1 >= T) { s[nb][1] = -INFINITY; s[nb][3] = -INFINITY; }
            }
        }
        float mx0 = fmaxf(s[0][0], s[0][1]), mx1 = fmaxf(s[0][2], s[0][3]);
#pragma unroll
        for (int nb = 1; nb < 8; ++nb) {
            mx0 = fmaxf(mx0, fmaxf(s[nb][0], s[nb][1]));
            mx1 = fmaxf(mx1, fmaxf(s[nb][2], s[nb][3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);   // raw-score maxima; finite: a tile has a valid key
        if (__any_sync(0xffffffffu, mn0 != m0 || mn1 != m1)) {
            const float c0 = fast_exp2((m0 - mn0) * scale_log2e), c1 = fast_exp2((m1 - mn1) * scale_log2e);
            l0 *= c0; l1 *= c1;
#pragma unroll
            for (int nb = 0; nb < 8; ++nb) { o[nb][0] *= c0; o[nb][1] *= c0; o[nb][2] *= c1; o[nb][3] *= c1; }
            m0 = mn0; m1 = mn1;
        }
        const float ms0 = -mn0 * scale_log2e, ms1 = -mn1 * scale_log2e;
        uint32_t ap[4][4];
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
            const float p0 = fast_exp2(fmaf(s[nb][0], scale_log2e, ms0)), p1 = fast_exp2(fmaf(s[nb][1], scale_log2e, ms0));
            const float p2 = fast_exp2(fmaf(s[nb][2], scale_log2e, ms1)), p3 = fast_exp2(fmaf(s[nb][3], scale_log2e, ms1));
            l0 += p0 + p1; l1 += p2 + p3;
            ap[nb >> 1][(nb & 1) * 2] = pack2(p0, p1, F16);
            ap[nb >> 1][(nb & 1) * 2 + 1] = pack2(p2, p3, F16);
        }
        // O += P V
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
            for (int dp = 0; dp < 4; ++dp) {          // two 8-wide dh blocks per ldmatrix.x4.trans
                uint32_t b0, b1, b2, b3;
                const int r = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                const int c = dp * 2 + (lane >> 4);
                ldsm_x4_t(smem_addr(sVb + swz(r, c)), b0, b1, b2, b3);
                mma16816<F16>(o[dp * 2], ap[kk], b0, b1);
                mma16816<F16>(o[dp * 2 + 1], ap[kk], b2, b3);
            }
        }
        }   // warp_active
        __syncthreads();   // all warps done with this K/V buffer before a later prefetch overwrites it
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = 1.f / l0, i1 = 1.f / l1;
    // stage the 64x64 output tile in sQ (each warp only touches its own 16 rows), then coalesced 16 B stores
    const int g = lane >> 2, tg = lane & 3;
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
        const int r0 = warp * 16 + g, r1 = r0 + 8;
        *reinterpret_cast<uint32_t*>(sQ + swz(r0, nb) + tg * 4) = pack2(o[nb][0] * i0, o[nb][1] * i0, F16);
        *reinterpret_cast<uint32_t*>(sQ + swz(r1, nb) + tg * 4) = pack2(o[nb][2] * i1, o[nb][3] * i1, F16);
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int idx = lane + 32 * i;                // 128 chunks per warp (16 rows x 8)
        const int r = warp * 16 + (idx >> 3), c = idx & 7;
        if (q0 + r < Tq)
            *reinterpret_cast<uint4*>(out + (img * Tq + q0 + r) * o_ld + head * DH + c * 8) =
                *reinterpret_cast<const uint4*>(sQ + swz(r, c));
    }
}

// ------------------------------------------------------------------------------------------------ greedy cross-attention
// Cross-attention of a greedy (beam 1) decode step WITHOUT a K/V cache.  The cached form reads K and V = 2 x 577 x 1024
// 16-bit values per crop and layer every step (2.36 MB; HBM-bound, ~60 % of a decode step).  Since
//   score_t^h = q^h . (Wk^h e_t + bk^h) = (Wk^h^T q^h) . e_t + const      (the constant cancels in the softmax)
//   out^h     = sum_t p_t (Wv^h e_t + bv^h) = Wv^h (sum_t p_t e_t) + bv^h
// the step can attend over the ENCODER STATES e_t themselves (577 x E, 0.89 MB for E = 768) with per-head projected
// queries q'^h = Wk^h^T q^h in R^E: 2.7x fewer bytes per layer, no per-crop K/V precompute (12 GEMMs per chunk) and no
// 29 GB cache.  The two per-head projections are block-diagonal tap-GEMMs (gemm_tc.cu, `batches`).
// This kernel: one CTA per crop, the 16 heads are the 16 rows of an m16n8k16 tile; 4 warps split the E dimension (8 warps measured slower: the redundant per-warp softmax and the larger partial-score reduction outweigh the extra latency hiding)
// (scores: partial sums reduced through shared memory; context: each warp owns E/4 output columns); encoder rows are
// streamed in 32-key tiles through a cp.async ring.  qp, ctx: [rows, 16*E]; enc: [rows*T, E].
template <bool F16, int ESLICE, int WARPS, int STAGES, int XE_KEYS = 32>
__global__ void __launch_bounds__(WARPS * 32, (STAGES * XE_KEYS * ESLICE * WARPS * 2 + WARPS * 64 * XE_KEYS <= 56 * 1024) ? 4 : ((STAGES * XE_KEYS * ESLICE * WARPS * 2 + WARPS * 64 * XE_KEYS <= 112 * 1024) ? 2 : 1)) dec_cross_enc_kernel(const bf16* __restrict__ qp, const bf16* __restrict__ enc,
                                                                bf16* __restrict__ ctxo, int T, int heads,
                                                                const unsigned char* __restrict__ finished) {
    constexpr int E = ESLICE * WARPS;
    constexpr int NT = WARPS * 32;
    constexpr int ROWB = E * 2;                       // bytes per smem row
    constexpr int NB = ESLICE / 8;                    // 8-wide output blocks per warp
    extern __shared__ __align__(128) unsigned char xe_smem[];
    unsigned char* sE = xe_smem;                      // STAGES x [32][E]; the projected queries are staged through stage 0
    unsigned char* sQ = sE;                           // [16][E], only until their fragments sit in registers
    float* sS = reinterpret_cast<float*>(sE + STAGES * XE_KEYS * ROWB);   // [WARPS][16][XE_KEYS] partial scores
    constexpr int NKB = XE_KEYS / 8;                   // 8-key blocks per tile
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long crop = blockIdx.x;
    // a crop whose hypothesis has ended (EOS chosen in an earlier step) is dropped from the batch by the reference's
    // generator (generator.py:266-303); here its row stays in place and simply stops reading its encoder states — the
    // 0.9 MB per crop and layer that bound this kernel.  Its stale context row is never looked at again.
    if (finished != nullptr && finished[crop]) return;
    const bf16* qbase = qp + crop * heads * E;     // heads <= 16 rows; the rest of the m16 tile is zero
    const bf16* ebase = enc + crop * T * E;
    const int n_tiles = (T + XE_KEYS - 1) / XE_KEYS;
    auto swz_w = [](int row, int chunk) { return row * ROWB + ((chunk ^ (row & 7)) << 4); };
    auto load_rows = [&](unsigned char* dst, const bf16* src, int rows, int row0, int limit) {
        const int chunks = rows * (E / 8);
        for (int idx = tid; idx < chunks; idx += NT) {
            const int r = idx / (E / 8), c = idx - r * (E / 8);
            const bool ok = row0 + r < limit;
            cp_async16(smem_addr(dst + swz_w(r, c)), src + (long long)(ok ? row0 + r : 0) * E + c * 8, ok);
        }
    };
    const int cbase = warp * (ESLICE / 8);            // first 16-byte chunk of this warp's E slice
    load_rows(sQ, qbase, 16, 0, heads);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    uint32_t aq[ESLICE / 16][4];                      // this warp's A fragments of q' (16 heads x ESLICE), kept for all tiles
#pragma unroll
    for (int ks = 0; ks < ESLICE / 16; ++ks)
        ldsm_x4(smem_addr(sQ + swz_w((lane & 7) + ((lane >> 3) & 1) * 8, cbase + ks * 2 + (lane >> 4))), aq[ks][0], aq[ks][1],
                aq[ks][2], aq[ks][3]);
    __syncthreads();
#pragma unroll
    for (int st = 0; st < STAGES - 1; ++st) {
        if (st < n_tiles) load_rows(sE + st * XE_KEYS * ROWB, ebase, XE_KEYS, st * XE_KEYS, T);
        cp_async_commit();
    }
    float o[NB][4];
#pragma unroll
    for (int i = 0; i < NB; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    const int g = lane >> 2, tg = lane & 3;
    for (int j = 0; j < n_tiles; ++j) {
        unsigned char* tile = sE + (j % STAGES) * XE_KEYS * ROWB;
        cp_async_wait<STAGES - 2>();      // tile j has landed (tiles j+1 .. j+STAGES-2 may still be in flight)
        __syncthreads();                  // ... for every thread; and every warp is done with iteration j-1 (its tile, sS)
        {
            // refill the stage that iteration j-1 just released: STAGES-1 tiles are in flight while tile j is processed
            const int jn = j + STAGES - 1;
            if (jn < n_tiles) load_rows(sE + (jn % STAGES) * XE_KEYS * ROWB, ebase, XE_KEYS, jn * XE_KEYS, T);
            cp_async_commit();
        }
        // partial scores over this warp's slice of E: S[16 heads x 32 keys]
        float s[NKB][4];
#pragma unroll
        for (int nb = 0; nb < NKB; ++nb) { s[nb][0] = s[nb][1] = s[nb][2] = s[nb][3] = 0.f; }
#pragma unroll
        for (int kp = 0; kp < ESLICE / 32; ++kp) {
#pragma unroll
            for (int nb = 0; nb < NKB; ++nb) {
                uint32_t b0, b1, b2, b3;
                ldsm_x4(smem_addr(tile + swz_w(nb * 8 + (lane & 7), cbase + kp * 4 + (lane >> 3))), b0, b1, b2, b3);
                mma16816<F16>(s[nb], aq[kp * 2], b0, b1);
                mma16816<F16>(s[nb], aq[kp * 2 + 1], b2, b3);
            }
        }
        float* mine = sS + warp * 16 * XE_KEYS;
#pragma unroll
        for (int nb = 0; nb < NKB; ++nb) {
            *reinterpret_cast<float2*>(mine + g * XE_KEYS + nb * 8 + tg * 2) = make_float2(s[nb][0], s[nb][1]);
            *reinterpret_cast<float2*>(mine + (g + 8) * XE_KEYS + nb * 8 + tg * 2) = make_float2(s[nb][2], s[nb][3]);
        }
        __syncthreads();
        const int key0 = j * XE_KEYS + tg * 2;
#pragma unroll
        for (int nb = 0; nb < NKB; ++nb) {
            float2 t0 = make_float2(0.f, 0.f), t1 = make_float2(0.f, 0.f);
#pragma unroll
            for (int w = 0; w < WARPS; ++w) {
                const float2 u0 = *reinterpret_cast<const float2*>(sS + w * 16 * XE_KEYS + g * XE_KEYS + nb * 8 + tg * 2);
                const float2 u1 = *reinterpret_cast<const float2*>(sS + w * 16 * XE_KEYS + (g + 8) * XE_KEYS + nb * 8 + tg * 2);
                t0.x += u0.x; t0.y += u0.y; t1.x += u1.x; t1.y += u1.y;
            }
            const int k = key0 + nb * 8;
            s[nb][0] = k < T ? t0.x : -INFINITY; s[nb][1] = k + 1 < T ? t0.y : -INFINITY;
            s[nb][2] = k < T ? t1.x : -INFINITY; s[nb][3] = k + 1 < T ? t1.y : -INFINITY;
        }
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int nb = 0; nb < NKB; ++nb) {
            mx0 = fmaxf(mx0, fmaxf(s[nb][0], s[nb][1]));
            mx1 = fmaxf(mx1, fmaxf(s[nb][2], s[nb][3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
        const float L2E = 1.4426950408889634f;
        if (__any_sync(0xffffffffu, mn0 != m0 || mn1 != m1)) {
            const float c0 = fast_exp2((m0 - mn0) * L2E), c1 = fast_exp2((m1 - mn1) * L2E);
            l0 *= c0; l1 *= c1;
#pragma unroll
            for (int i = 0; i < NB; ++i) { o[i][0] *= c0; o[i][1] *= c0; o[i][2] *= c1; o[i][3] *= c1; }
            m0 = mn0; m1 = mn1;
        }
        uint32_t ap[NKB / 2][4];
#pragma unroll
        for (int nb = 0; nb < NKB; ++nb) {
            const float p0 = fast_exp2((s[nb][0] - mn0) * L2E), p1 = fast_exp2((s[nb][1] - mn0) * L2E);
            const float p2 = fast_exp2((s[nb][2] - mn1) * L2E), p3 = fast_exp2((s[nb][3] - mn1) * L2E);
            l0 += p0 + p1; l1 += p2 + p3;
            ap[nb >> 1][(nb & 1) * 2] = pack2(p0, p1, F16);
            ap[nb >> 1][(nb & 1) * 2 + 1] = pack2(p2, p3, F16);
        }
        // context slice: O[16 x ESLICE] += P[16 x 32] . E[32 x ESLICE]
#pragma unroll
        for (int kk = 0; kk < NKB / 2; ++kk) {
#pragma unroll
            for (int dp = 0; dp < NB / 2; ++dp) {
                uint32_t b0, b1, b2, b3;
                const int r = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                ldsm_x4_t(smem_addr(tile + swz_w(r, cbase + dp * 2 + (lane >> 4))), b0, b1, b2, b3);
                mma16816<F16>(o[dp * 2], ap[kk], b0, b1);
                mma16816<F16>(o[dp * 2 + 1], ap[kk], b2, b3);
            }
        }
        // no barrier here: the next iteration's first barrier orders this iteration's reads of `tile` / sS before their reuse
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = 1.f / l0, i1 = 1.f / l1;
    bf16* obase = ctxo + crop * heads * E + warp * ESLICE;
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) {
        if (g < heads)
            *reinterpret_cast<uint32_t*>(obase + (long long)g * E + nb * 8 + tg * 2) = pack2(o[nb][0] * i0, o[nb][1] * i0, F16);
        if (g + 8 < heads)
            *reinterpret_cast<uint32_t*>(obase + (long long)(g + 8) * E + nb * 8 + tg * 2) = pack2(o[nb][2] * i1, o[nb][3] * i1, F16);
    }
}

// ------------------------------------------------------------------------------------------------ decoder glue
// x[r] = sqrt(H) * E[tok[r]] + PE[pos]   (fairseq: embed_scale * embed_tokens + sinusoidal positions)
__global__ void dec_embed_kernel(const int* __restrict__ tokens, int tok_ld, int step, const bf16* __restrict__ embed,
                                 const float* __restrict__ pe, bf16* __restrict__ x, int rows, int H, float scale,
                                 int f16) {
    const int r = blockIdx.x;
    if (r >= rows) return;
    const int tok = tokens[(long long)r * tok_ld + step];
    const float* p = pe + (long long)(TOK_PAD + 1 + step) * H;
    for (int c = threadIdx.x * 2; c < H; c += blockDim.x * 2) {
        const float2 e = unpack2(*reinterpret_cast<const uint32_t*>(embed + (long long)tok * H + c), f16);
        *reinterpret_cast<uint32_t*>(x + (long long)r * H + c) = pack2(scale * e.x + p[c], scale * e.y + p[c + 1], f16);
    }
}

// Self-attention of one new token over the cached history.  qkv: [R, 3H] (q pre-scaled | k | v) of this step.
// Appends k, v to the caches at (step, row) and attends over steps 0..step through the ancestor table.
// grid (heads, R), one warp (2 dims of the head per lane).
__global__ void __launch_bounds__(32) dec_self_attn_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ kcache,
                                                           bf16* __restrict__ vcache, const int* __restrict__ anc,
                                                           bf16* __restrict__ out, int R, int H, int step, int f16) {
    const int head = blockIdx.x, r = blockIdx.y, lane = threadIdx.x;
    const long long col = head * DH + lane * 2;
    const uint32_t qraw = *reinterpret_cast<const uint32_t*>(qkv + (long long)r * 3 * H + col);
    const uint32_t kraw = *reinterpret_cast<const uint32_t*>(qkv + (long long)r * 3 * H + H + col);
    const uint32_t vraw = *reinterpret_cast<const uint32_t*>(qkv + (long long)r * 3 * H + 2 * H + col);
    *reinterpret_cast<uint32_t*>(kcache + ((long long)step * R + r) * H + col) = kraw;
    *reinterpret_cast<uint32_t*>(vcache + ((long long)step * R + r) * H + col) = vraw;
    const float2 q = unpack2(qraw, f16);
    float m = -INFINITY, l = 0.f, ox = 0.f, oy = 0.f;
    for (int t = 0; t <= step; ++t) {
        uint32_t kr, vr;
        if (t == step) { kr = kraw; vr = vraw; }
        else {
            const long long src = ((long long)t * R + anc[(long long)t * R + r]) * H + col;
            kr = *reinterpret_cast<const uint32_t*>(kcache + src);
            vr = *reinterpret_cast<const uint32_t*>(vcache + src);
        }
        const float2 k = unpack2(kr, f16), v = unpack2(vr, f16);
        const float s = warp_sum(q.x * k.x + q.y * k.y);
        const float mn = fmaxf(m, s);
        const float c = __expf(m - mn), p = __expf(s - mn);
        l = l * c + p;
        ox = ox * c + p * v.x;
        oy = oy * c + p * v.y;
        m = mn;
    }
    *reinterpret_cast<uint32_t*>(out + (long long)r * H + col) = pack2(ox / l, oy / l, f16);
}

// ------------------------------------------------------------------------------------------------ search
// Per row: log-softmax statistics of the fp32 logits and the top `cand` (= 2*beam) masked log-probs
// (generator.py:153-177: pad never, EOS banned below min_len, only EOS at max_len; NaN -> -inf).
// Ties resolve to the lower token id.
constexpr int TOPK_THREADS = 256;
constexpr int MAX_CAND = 2 * MAX_BEAM;
__device__ __forceinline__ bool cand_better(float v, int i, float bv, int bi) { return v > bv || (v == bv && i < bi); }

__global__ void __launch_bounds__(TOPK_THREADS) logits_topk_kernel(const float* __restrict__ logits, int V, int ld,
                                                                   int cand, int eos_banned, int only_eos,
                                                                   float* __restrict__ cand_val, int* __restrict__ cand_idx,
                                                                   const unsigned char* __restrict__ finished, int beam) {
    __shared__ float red_m[TOPK_THREADS / 32], red_s[TOPK_THREADS / 32];
    __shared__ float s_stat[2];
    __shared__ float sv[TOPK_THREADS / 32][MAX_CAND];
    __shared__ int si[TOPK_THREADS / 32][MAX_CAND];
    const int r = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // a crop whose sentence is done takes no further candidates (search_step_kernel returns on it): its rows are skipped
    if (finished != nullptr && finished[r / beam]) return;
    const float* x = logits + (long long)r * ld;
    // ONE pass over the row's 200 KB: running (max, sum of exp) per thread — every vocabulary entry takes part in the
    // normaliser, masked or not — and the per-thread top list of the unmasked logits.  (The first version read the row
    // twice: a max pass, then the sums.)
    float tv[MAX_CAND];
    int ti[MAX_CAND];
#pragma unroll
    for (int k = 0; k < MAX_CAND; ++k) { tv[k] = -INFINITY; ti[k] = 0x7fffffff; }
    float mx = -INFINITY, sum = 0.f;
    for (int i = tid; i < V; i += TOPK_THREADS) {
        float v = x[i];
        if (v == v) {
            if (v > mx) { sum = sum * __expf(mx - v) + 1.f; mx = v; }       // first entry: 0 * exp(-inf) + 1
            else sum += __expf(v - mx);
        } else {
            v = -INFINITY;
        }
        if (i == TOK_PAD || (eos_banned && i == TOK_EOS) || (only_eos && i != TOK_EOS)) v = -INFINITY;
        if (cand_better(v, i, tv[cand - 1], ti[cand - 1])) {
            int k = cand - 1;
            while (k > 0 && cand_better(v, i, tv[k - 1], ti[k - 1])) { tv[k] = tv[k - 1]; ti[k] = ti[k - 1]; --k; }
            tv[k] = v; ti[k] = i;
        }
    }
    // merge the (max, sum) pairs: warp, then block
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float m2 = __shfl_xor_sync(0xffffffffu, mx, o), s2 = __shfl_xor_sync(0xffffffffu, sum, o);
        const float M = fmaxf(mx, m2);
        sum = (M == -INFINITY) ? 0.f : sum * __expf(mx - M) + s2 * __expf(m2 - M);
        mx = M;
    }
    if (lane == 0) { red_m[warp] = mx; red_s[warp] = sum; }
    __syncthreads();
    if (tid == 0) {
        float M = red_m[0];
        for (int w = 1; w < TOPK_THREADS / 32; ++w) M = fmaxf(M, red_m[w]);
        float S = 0.f;
        for (int w = 0; w < TOPK_THREADS / 32; ++w) S += (red_m[w] == -INFINITY) ? 0.f : red_s[w] * __expf(red_m[w] - M);
        s_stat[0] = M;
        s_stat[1] = logf(S);
    }
    // warp-level merge: repeatedly take the best head among the 32 sorted lists
    int head = 0;
    for (int k = 0; k < cand; ++k) {
        float bv = head < cand ? tv[head] : -INFINITY;
        int bi = head < cand ? ti[head] : 0x7fffffff;
        float wv = bv; int wi = bi;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, wv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, wi, o);
            if (cand_better(ov, oi, wv, wi)) { wv = ov; wi = oi; }
        }
        if (wi == bi && wv == bv && head < cand) ++head;      // token ids are unique: exactly one lane advances
        if (lane == 0) { sv[warp][k] = wv; si[warp][k] = wi; }
    }
    __syncthreads();
    if (warp == 0) {
        // merge the 8 warp lists (lane w < 8 owns list w)
        int h = 0;
        const float lse = s_stat[0] + s_stat[1];
        for (int k = 0; k < cand; ++k) {
            float bv = (lane < TOPK_THREADS / 32 && h < cand) ? sv[lane][h] : -INFINITY;
            int bi = (lane < TOPK_THREADS / 32 && h < cand) ? si[lane][h] : 0x7fffffff;
            float wv = bv; int wi = bi;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, wv, o);
                const int oi = __shfl_xor_sync(0xffffffffu, wi, o);
                if (cand_better(ov, oi, wv, wi)) { wv = ov; wi = oi; }
            }
            if (wi == bi && wv == bv && lane < TOPK_THREADS / 32 && h < cand) ++h;
            if (lane == 0) {
                cand_val[(long long)r * cand + k] = (wv == -INFINITY) ? -INFINITY : wv - lse;
                cand_idx[(long long)r * cand + k] = wi;
            }
        }
    }
}

struct SearchState {
    int* tokens;        // [R, max_len + 2]
    float* scores;      // [R, max_len + 1] cumulative log-probs
    int* anc;           // [max_len + 1, R] ancestor table for the self-attention caches
    int* tokens_tmp; float* scores_tmp; int* anc_tmp;
    unsigned char* ignore;   // [n, beam]
    int* fin_count;     // [n]
    int* fin_tokens;    // [n, beam, max_len + 1]
    int* fin_len;       // [n, beam]
    float* fin_score;   // [n, beam]
    unsigned char* finished;   // [n]
    int* n_unfinished;  // [1]
};

// One thread per sentence: fairseq BeamSearch.step over the per-row candidate lists + generator.py:196-362.
__global__ void search_step_kernel(SearchState st, const float* __restrict__ cand_val, const int* __restrict__ cand_idx,
                                   int n, int beam, int step, int max_len) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const int R = n * beam, cand = 2 * beam, tl = max_len + 2, sl = max_len + 1;
    // rows keep their history unless a live sentence reorders them
    if (st.finished[s]) {
        for (int b = 0; b < beam; ++b) st.anc[(long long)step * R + s * beam + b] = s * beam + b;
        return;
    }
    // candidates: step 0 uses beam 0 only; later steps add the beam's cumulative score
    float cv[MAX_CAND]; int cb[MAX_CAND], ct[MAX_CAND];
    int heads[MAX_BEAM];
    for (int b = 0; b < beam; ++b) heads[b] = 0;
    const int nb = step == 0 ? 1 : beam;
    for (int k = 0; k < cand; ++k) {
        float bv = -INFINITY; int bb = -1, bt = 0x7fffffff;
        for (int b = 0; b < nb; ++b) {
            if (heads[b] >= cand) continue;
            const long long row = (long long)(s * beam + b);
            float v = cand_val[row * cand + heads[b]];
            const int t = cand_idx[row * cand + heads[b]];
            if (step > 0) v += st.scores[row * sl + step - 1];
            // torch.topk over the flattened [beam * V] row: ties by flat index (beam-major)
            if (bb < 0 || v > bv) { bv = v; bb = b; bt = t; }
        }
        cv[k] = bv; cb[k] = bb; ct[k] = bt;
        heads[bb]++;
    }
    // finalise EOS hypotheses among the top `beam` candidates
    bool eos[MAX_CAND];
    for (int k = 0; k < cand; ++k) eos[k] = (ct[k] == TOK_EOS) && (cv[k] != -INFINITY);
    for (int k = 0; k < beam; ++k) eos[k] = eos[k] && !st.ignore[s * beam + k];
    int fc = st.fin_count[s];
    for (int k = 0; k < beam; ++k) {
        if (!eos[k] || fc >= beam) continue;
        const long long src = (long long)(s * beam + cb[k]);
        int* ft = st.fin_tokens + ((long long)s * beam + fc) * sl;
        for (int t = 0; t < step; ++t) ft[t] = st.tokens[src * tl + 1 + t];
        ft[step] = TOK_EOS;
        st.fin_len[s * beam + fc] = step + 1;
        st.fin_score[s * beam + fc] = cv[k] / (float)(step + 1);
        ++fc;
    }
    st.fin_count[s] = fc;
    if (fc == beam || step == max_len) {
        st.finished[s] = 1;
        atomicSub(st.n_unfinished, 1);
        for (int b = 0; b < beam; ++b) st.anc[(long long)step * R + s * beam + b] = s * beam + b;
        return;
    }
    // pick the `beam` best non-EOS candidates (smallest of eos*cand + k, ascending)
    for (int k = 0; k < beam; ++k) eos[k] = eos[k] || st.ignore[s * beam + k];
    int hyp[MAX_BEAM]; bool ign[MAX_BEAM];
    int j = 0;
    for (int k = 0; k < cand && j < beam; ++k) if (!eos[k]) { hyp[j] = k; ign[j] = false; ++j; }
    for (int k = 0; k < cand && j < beam; ++k) if (eos[k]) { hyp[j] = k; ign[j] = true; ++j; }
    for (int b = 0; b < beam; ++b) {
        const int k = hyp[b];
        const long long src = (long long)(s * beam + cb[k]), dst = (long long)(s * beam + b);
        st.ignore[s * beam + b] = ign[b];
        for (int t = 0; t <= step; ++t) st.tokens_tmp[dst * tl + t] = st.tokens[src * tl + t];
        st.tokens_tmp[dst * tl + step + 1] = ct[k];
        for (int t = 0; t < step; ++t) st.scores_tmp[dst * sl + t] = st.scores[src * sl + t];
        st.scores_tmp[dst * sl + step] = cv[k];
        for (int t = 0; t < step; ++t) st.anc_tmp[(long long)t * R + dst] = st.anc[(long long)t * R + src];
        st.anc_tmp[(long long)step * R + dst] = (int)src;
    }
}
// second phase: publish the reordered rows of live sentences (tmp -> main), after every thread has read the old state
__global__ void search_commit_kernel(SearchState st, int n, int beam, int step, int max_len) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    const int R = n * beam, tl = max_len + 2, sl = max_len + 1;
    if (row >= R) return;
    if (st.finished[row / beam]) return;
    for (int t = 0; t <= step + 1; ++t) st.tokens[(long long)row * tl + t] = st.tokens_tmp[(long long)row * tl + t];
    for (int t = 0; t <= step; ++t) st.scores[(long long)row * sl + t] = st.scores_tmp[(long long)row * sl + t];
    for (int t = 0; t <= step; ++t) st.anc[(long long)t * R + row] = st.anc_tmp[(long long)t * R + row];
}

// teacher forcing (parity hook): next token = forced[r, step]; cumulative score of that token
__global__ void forced_step_kernel(SearchState st, const float* __restrict__ logits, int V, int ld,
                                   const int* __restrict__ forced, int forced_ld, int R, int step, int max_len) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    st.tokens[(long long)r * (max_len + 2) + step + 1] = forced[(long long)r * forced_ld + step];
    st.anc[(long long)step * R + r] = r;
}

__global__ void search_init_kernel(SearchState st, int n, int beam, int max_len) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int R = n * beam;
    if (i < R) {
        for (int t = 0; t < max_len + 2; ++t) st.tokens[(long long)i * (max_len + 2) + t] = t == 0 ? TOK_EOS : TOK_PAD;
        st.ignore[i] = 0;
        st.fin_len[i] = 0;
        st.fin_score[i] = -INFINITY;
    }
    if (i < n) { st.fin_count[i] = 0; st.finished[i] = 0; }
    if (i == 0) *st.n_unfinished = n;
}

// best finalised hypothesis per sentence (generator.py:364-373 sorts by score, get_text takes [0])
__global__ void search_pick_kernel(SearchState st, int n, int beam, int max_len, int* __restrict__ out_tokens,
                                   int out_ld, int* __restrict__ out_len, float* __restrict__ out_score) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    int best = -1;
    for (int b = 0; b < st.fin_count[s]; ++b)
        if (best < 0 || st.fin_score[s * beam + b] > st.fin_score[s * beam + best]) best = b;
    if (best < 0) { out_len[s] = 0; out_score[s] = -INFINITY; return; }
    const int len = st.fin_len[s * beam + best];
    for (int t = 0; t < len && t < out_ld; ++t)
        out_tokens[(long long)s * out_ld + t] = st.fin_tokens[((long long)s * beam + best) * (max_len + 1) + t];
    for (int t = len; t < out_ld; ++t) out_tokens[(long long)s * out_ld + t] = TOK_PAD;
    out_len[s] = len;
    out_score[s] = st.fin_score[s * beam + best];
}

// ------------------------------------------------------------------------------------------------ host helpers
struct Arena {
    unsigned char* base; size_t off = 0, cap;
    template <typename T> T* take(size_t n) {
        off = mb_align_up(off, 256);
        T* p = base ? (T*)(base + off) : nullptr;
        off += n * sizeof(T);
        return p;
    }
};

int ensure_arena(mb_ctx* ctx, TrocrModel* m, size_t bytes) {
    if (bytes <= m->arena_bytes) return 0;
    if (m->arena) cudaFree(m->arena);
    m->arena = nullptr; m->arena_bytes = 0;
    if (cudaMalloc(&m->arena, bytes) != cudaSuccess) {
        cudaGetLastError();
        return mb_set_err(ctx, MB_ERR_OOM, "trocr: workspace of %zu bytes failed", bytes);
    }
    m->arena_bytes = bytes;
    return 0;
}

// MB_ATTN_LEGACY=1 routes the encoder through the mma.sync flash kernel (A/B testing of the tcgen05 kernel)
bool attention_legacy() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("MB_ATTN_LEGACY"); v = (e && e[0] == '1') ? 1 : 0; }
    return v == 1;
}

// Row statistics only: (-mean, rstd) per row, same two-pass arithmetic as layernorm_kernel.  With the LayerNorm folded
// into the next GEMM (TapGemm::ln_stats) the normalised activations are never written: 2 B/element read instead of
// 2 B read + 2 B written + 2 B read again by the GEMM.
__global__ void __launch_bounds__(256) ln_stats_kernel(const bf16* __restrict__ in, float2* __restrict__ stats,
                                                       long long rows, int D, float eps, int f16) {
    const int lane = threadIdx.x & 31;
    const int nv = D >> 3;
    const long long wstride = (long long)gridDim.x * (blockDim.x >> 5);
    long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    uint4 nxt[LN_MAXV];
    if (row < rows) {
        const uint4* src = reinterpret_cast<const uint4*>(in + row * D);
#pragma unroll
        for (int i = 0; i < LN_MAXV; ++i)
            if (lane + 32 * i < nv) nxt[i] = src[lane + 32 * i];
    }
    for (; row < rows; row += wstride) {
        uint4 cur[LN_MAXV];
#pragma unroll
        for (int i = 0; i < LN_MAXV; ++i) cur[i] = nxt[i];
        if (row + wstride < rows) {
            const uint4* src = reinterpret_cast<const uint4*>(in + (row + wstride) * D);
#pragma unroll
            for (int i = 0; i < LN_MAXV; ++i)
                if (lane + 32 * i < nv) nxt[i] = src[lane + 32 * i];
        }
        float v[LN_MAXV][8];
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < LN_MAXV; ++i) {
            if (lane + 32 * i < nv) {
                const uint32_t w[4] = {cur[i].x, cur[i].y, cur[i].z, cur[i].w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float2 f = unpack2(w[k], f16);
                    v[i][2 * k] = f.x; v[i][2 * k + 1] = f.y;
                    sum += f.x + f.y;
                }
            }
        }
        const float mean = warp_sum(sum) / (float)D;
        float sq = 0.f;
#pragma unroll
        for (int i = 0; i < LN_MAXV; ++i) {
            if (lane + 32 * i < nv) {
#pragma unroll
                for (int k = 0; k < 8; ++k) { const float d = v[i][k] - mean; sq += d * d; }
            }
        }
        const float rstd = rsqrtf(warp_sum(sq) / (float)D + eps);
        if (lane == 0) stats[row] = make_float2(-mean, rstd);
    }
}

// (sum, sum of squares) partials written by the residual GEMM's epilogue (TapGemm::stat_out) -> (-mean, rstd) per row.
// var = E[x^2] - mean^2 in fp32: the stream's row means are small against its spread, the cancellation costs ~1e-6.
__global__ void __launch_bounds__(256) ln_stats_finish_kernel(const float2* __restrict__ part, float2* __restrict__ stats,
                                                              long long rows, int slots, float inv_d, float eps) {
    for (long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x; row < rows; row += (long long)gridDim.x * blockDim.x) {
        float s = 0.f, q = 0.f;
        for (int i = 0; i < slots; ++i) {
            const float2 v = part[row * slots + i];
            s += v.x;
            q += v.y;
        }
        const float mean = s * inv_d;
        const float var = fmaxf(fmaf(-mean, mean, q * inv_d), 0.f);
        stats[row] = make_float2(-mean, rsqrtf(var + eps));
    }
}

int attention_setup(mb_ctx* ctx) {
    static bool done = false;
    if (done) return 0;
    MB_CUDA(ctx, cudaFuncSetAttribute(attention_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
    MB_CUDA(ctx, cudaFuncSetAttribute(attention_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
    MB_CUDA(ctx, cudaFuncSetAttribute(attention_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    MB_CUDA(ctx, cudaFuncSetAttribute(attention_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    done = true;
    return 0;
}

int gemm(mb_ctx* ctx, const bf16* a, int K, const bf16* w, int rows_w, long long M, int N, const float* bias, int act,
         const bf16* residual, void* out, int out_mode, cudaStream_t s) {
    TapGemm g;
    g.a0 = a; g.c0 = K; g.a0_ld = K;
    g.n = 1; g.h = 1; g.w = (int)M;
    g.wgt = w; g.n_rows_w = rows_w; g.n_out = N;
    g.bias = bias; g.act = act;
    g.residual = residual; g.res_ld = N;
    g.out = out; g.out_ld = N; g.out_mode = out_mode;
    return mb_tap_gemm(ctx, g, s);
}

// residual GEMM (out = a W^T + b + residual) that also leaves the LayerNorm statistics of its output rows in `stats`
int gemm_res_stats(mb_ctx* ctx, const bf16* a, int K, const bf16* w, long long M, int N, const float* bias, const bf16* residual,
                   bf16* out, float* part, float* stats, float eps, cudaStream_t s) {
    TapGemm g;
    g.a0 = a; g.c0 = K; g.a0_ld = K;
    g.n = 1; g.h = 1; g.w = (int)M;
    g.wgt = w; g.n_rows_w = N; g.n_out = N;
    g.bias = bias; g.act = MB_ACT_NONE;
    g.residual = residual; g.res_ld = N;
    g.out = out; g.out_ld = N; g.out_mode = MB_OUT_BF16;
    g.block_n = 256;
    g.stat_out = part;
    int rc = mb_tap_gemm(ctx, g, s);
    if (rc) return rc;
    const int slots = 2 * mb_cdiv(N, 256);
    const long long blocks = (M + 255) / 256;
    const long long cap = (long long)ctx->num_sms * 8;
    ln_stats_finish_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, s>>>(reinterpret_cast<const float2*>(part),
                                                                                  reinterpret_cast<float2*>(stats), M, slots, 1.0f / (float)N, eps);
    MB_LAUNCH_CHECK(ctx);
    return 0;
}

int layernorm(mb_ctx* ctx, const bf16* in, bf16* out, const float* g, const float* b, long long rows, int D, float eps,
              cudaStream_t s) {
    if (D % 8 != 0 || D > 256 * LN_MAXV) return mb_set_err(ctx, MB_ERR_ARG, "layernorm: unsupported width %d", D);
    const long long blocks = (rows + 7) / 8;
    const long long cap = (long long)ctx->num_sms * 8;
    layernorm_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, s>>>(in, out, g, b, rows, D, eps, ctx->f16);
    MB_LAUNCH_CHECK(ctx);
    return 0;
}

int ln_stats(mb_ctx* ctx, const bf16* in, float* stats, long long rows, int D, float eps, cudaStream_t s) {
    if (D % 8 != 0 || D > 256 * LN_MAXV) return mb_set_err(ctx, MB_ERR_ARG, "ln_stats: unsupported width %d", D);
    const long long blocks = (rows + 7) / 8;
    const long long cap = (long long)ctx->num_sms * 8;
    ln_stats_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, s>>>(in, reinterpret_cast<float2*>(stats), rows, D, eps,
                                                                            ctx->f16);
    MB_LAUNCH_CHECK(ctx);
    return 0;
}

// GEMM over the raw residual stream with the LayerNorm folded in (see TapGemm::ln_stats)
int gemm_ln(mb_ctx* ctx, const bf16* a, int K, const bf16* wf, long long M, int N, const float* bias_f, const float* c,
            const float* stats, int act, void* out, cudaStream_t s) {
    TapGemm g;
    g.a0 = a; g.c0 = K; g.a0_ld = K;
    g.n = 1; g.h = 1; g.w = (int)M;
    g.wgt = wf; g.n_rows_w = N; g.n_out = N;
    g.bias = bias_f; g.act = act;
    g.out = out; g.out_ld = N; g.out_mode = MB_OUT_BF16;
    g.ln_stats = stats; g.ln_c = c;
    return mb_tap_gemm(ctx, g, s);
}

int grid1d(mb_ctx* ctx, long long total, int threads) {
    long long g = (total + threads - 1) / threads;
    const long long cap = (long long)ctx->num_sms * 16;
    return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

#define RC(expr) do { int _rc = (expr); if (_rc) return _rc; } while (0)

// patches [n*576, 768] -> enc_out [n*577, D]; ws must hold x, y [n*577*D] and big [n*577*max(3D, ffn)]
int encode(mb_ctx* ctx, TrocrModel* m, const bf16* patches, int n, bf16* enc_out, bf16* x, bf16* y, bf16* big,
           float* stats, float* stat_part, cudaStream_t s) {
    const int D = m->enc_dim, T = m->tokens, F = m->enc_ffn;
    const long long M = (long long)n * T;
    // patch embedding (Conv2d k=s=16 == GEMM over patch rows) into `big`, then cls/pos assembly into x
    RC(gemm(ctx, patches, 768, m->patch_w, D, (long long)n * (T - 1), D, m->patch_b, MB_ACT_NONE, nullptr, big, MB_OUT_BF16, s));
    {
        const long long total8 = M * (D / 8);
        assemble_tokens_kernel<<<grid1d(ctx, total8, 256), 256, 0, s>>>(big, m->cls_pos, x, total8, T, D, ctx->f16);
        MB_LAUNCH_CHECK(ctx);
    }
    const float scale_log2e = 0.125f * 1.4426950408889634f;
    // timing aid (tools/gpu_probe_encoder.py): MB_PROBE_SKIP=attn|ln|qkv|proj|fc1|fc2 leaves that launch class out, so
    // its in-situ cost (sustained clocks, warm L2) is the difference of two runs.  Results are garbage when set.
    const char* skip_env = getenv("MB_PROBE_SKIP");
    const std::string skip = skip_env ? skip_env : "";
    // the residual GEMMs (proj, fc2) leave the row statistics of the stream they write for the LayerNorm fold of the next
    // GEMM: the separate 2 B / element pass of ln_stats_kernel only runs in front of the first layer.  MB_LNSTAT_FUSE=0
    // (and the two-CTA GEMM mode) keep the separate pass.
    static int fuse_env = -1;
    if (fuse_env < 0) {
        const char* e = getenv("MB_LNSTAT_FUSE");
        const char* g2 = getenv("MB_GEMM2");
        const char* et = getenv("MB_EPI_TMA");
        fuse_env = ((e && e[0] == '0') || (g2 && g2[0] == '1') || (et && et[0] == '0')) ? 0 : 1;
    }
    const bool fuse = fuse_env == 1 && stat_part != nullptr && D % 32 == 0 && skip.empty();
    bool have_stats = false;                       // `stats` already describes x (written by the previous residual GEMM)
    for (int l = 0; l < m->enc_layers; ++l) {
        const EncLayer& L = m->enc[l];
        const bool fold = m->ln_fold && L.qkv_wf && L.fc1_wf;
        const bool next_fold = l + 1 < m->enc_layers && m->ln_fold && m->enc[l + 1].qkv_wf && m->enc[l + 1].fc1_wf;
        if (fold) {
            if (skip != "ln" && !have_stats) RC(ln_stats(ctx, x, stats, M, D, 1e-6f, s));
            if (skip != "qkv") RC(gemm_ln(ctx, x, D, L.qkv_wf, M, 3 * D, L.qkv_bf, L.qkv_c, stats, MB_ACT_NONE, big, s));
        } else {
            RC(layernorm(ctx, x, y, L.ln1_w, L.ln1_b, M, D, 1e-6f, s));
            RC(gemm(ctx, y, D, L.qkv_w, 3 * D, M, 3 * D, nullptr, MB_ACT_NONE, nullptr, big, MB_OUT_BF16, s));
        }
        if (skip == "attn") {
        } else if (!attention_legacy()) {
            RC(mb_attention_tc(ctx, big, y, n, T, D, m->enc_heads, scale_log2e, s));
        } else {
            dim3 grid((T + 63) / 64, m->enc_heads, n);
            RC(attention_setup(ctx));
            if (ctx->f16) attention_kernel<true><<<grid, 128, ATT_SMEM, s>>>(big, 3LL * D, big + D, big + 2 * D, 3LL * D, y, D, T, T, scale_log2e);
            else attention_kernel<false><<<grid, 128, ATT_SMEM, s>>>(big, 3LL * D, big + D, big + 2 * D, 3LL * D, y, D, T, T, scale_log2e);
            MB_LAUNCH_CHECK(ctx);
        }
        have_stats = false;
        if (fold && fuse) {
            RC(gemm_res_stats(ctx, y, D, L.proj_w, M, D, L.proj_b, x, x, stat_part, stats, 1e-6f, s));
            have_stats = true;
        } else if (skip != "proj") RC(gemm(ctx, y, D, L.proj_w, D, M, D, L.proj_b, MB_ACT_NONE, x, x, MB_OUT_BF16, s));
        if (fold) {
            if (skip != "ln" && !have_stats) RC(ln_stats(ctx, x, stats, M, D, 1e-6f, s));
            if (skip != "fc1") RC(gemm_ln(ctx, x, D, L.fc1_wf, M, F, L.fc1_bf, L.fc1_c, stats, MB_ACT_GELU, big, s));
        } else {
            RC(layernorm(ctx, x, y, L.ln2_w, L.ln2_b, M, D, 1e-6f, s));
            RC(gemm(ctx, y, D, L.fc1_w, F, M, F, L.fc1_b, MB_ACT_GELU, nullptr, big, MB_OUT_BF16, s));
        }
        have_stats = false;
        if (fuse && next_fold) {
            RC(gemm_res_stats(ctx, big, F, L.fc2_w, M, D, L.fc2_b, x, x, stat_part, stats, 1e-6f, s));
            have_stats = true;
        } else if (skip != "fc2") RC(gemm(ctx, big, F, L.fc2_w, D, M, D, L.fc2_b, MB_ACT_NONE, x, x, MB_OUT_BF16, s));
    }
    RC(layernorm(ctx, x, enc_out, m->norm_w, m->norm_b, M, D, 1e-6f, s));
    return 0;
}

struct DecodeWs {
    int* live = nullptr;          // greedy mode, tcgen05 cross-attention: compacted list of open crops + its counter (n + 1)
    bf16 *cross_kv, *kcache, *vcache, *x, *qkv, *att, *tmp, *ffn;
    int kv_cap = 0;               // steps the K / V caches currently hold per layer
    bf16 *qp, *ctxe;              // greedy mode: per-head projected queries / attended encoder states [R, heads*E]
    bool greedy;
    float* logits; float* cand_val; int* cand_idx;
    SearchState st;
};

// the tcgen05 / TMA cross-attention (xattn_tc.cu) is the default greedy kernel; MB_XE_TC=0 selects the mma.sync kernel
bool cross_tc(const TrocrModel* m) {
    static int on = -1;
    if (on < 0) { const char* e = getenv("MB_XE_TC"); on = (e && e[0] == '0') ? 0 : 1; }
    return on == 1 && mb_cross_enc_tc_supported(m->enc_dim, m->dec_heads);
}

// beam 1 attends over the encoder states directly (dec_cross_enc_kernel); MB_CROSS_CACHED=1 forces the K/V-cache path
bool cross_uncached(const TrocrModel* m, int beam) {
    static int forced = -1;
    if (forced < 0) { const char* e = getenv("MB_CROSS_CACHED"); forced = (e && e[0] == '1') ? 1 : 0; }
    // beam >= 2: the tcgen05 kernel shares a pass over a crop's encoder states between up to three of its hypotheses (two
    // for E = 1024); wider beams take ceil(beam / 3) passes, the second one mostly out of L2.  Measured on B200 against the
    // K/V-cache path (which also pays 12 projection GEMMs per crop and caps the decode batch at 2048 crops: 29 GB of cache
    // per 1024 crops): TrOCR-base beam 3 9.7 vs 8.2 pages/s, beam 5 decode 3.21 vs 3.47 s per 64 pages, TrOCR-large beam 3
    // decode 1.22 vs 1.47 s per 16 dense pages.  MB_CROSS_CACHED=1 keeps the cache (A/B, and the only path without the
    // tcgen05 kernel).
    const bool beams_ok = beam == 1 || (cross_tc(m) && beam <= MAX_BEAM);
    return beams_ok && !forced && m->dec_heads <= 16 && (m->enc_dim == 128 || m->enc_dim == 768 || m->enc_dim == 1024);
}

size_t plan_decode(TrocrModel* m, int n, int beam, int max_len, unsigned char* base, DecodeWs* w) {
    Arena a{base, 0, 0};
    const int H = m->dec_dim, T = m->tokens, L = m->dec_layers, V = m->vocab;
    const long long R = (long long)n * beam;
    const int cand = 2 * beam;
    w->greedy = cross_uncached(m, beam);
    if (w->greedy) {
        w->cross_kv = nullptr;
        w->qp = a.take<bf16>((size_t)R * m->dec_heads * m->enc_dim);
        w->ctxe = a.take<bf16>((size_t)R * m->dec_heads * m->enc_dim);
        w->live = a.take<int>((size_t)n + 1);
    } else {
        w->qp = w->ctxe = nullptr;
        w->cross_kv = a.take<bf16>((size_t)L * n * T * 2 * H);
    }
    w->x = a.take<bf16>(R * H);
    w->qkv = a.take<bf16>(R * 3 * H);
    w->att = a.take<bf16>(R * H);
    w->tmp = a.take<bf16>(R * H);
    w->ffn = a.take<bf16>(R * m->dec_ffn);
    w->logits = a.take<float>(R * V);
    w->cand_val = a.take<float>(R * cand);
    w->cand_idx = a.take<int>(R * cand);
    w->st.tokens = a.take<int>(R * (max_len + 2));
    w->st.tokens_tmp = a.take<int>(R * (max_len + 2));
    w->st.scores = a.take<float>(R * (max_len + 1));
    w->st.scores_tmp = a.take<float>(R * (max_len + 1));
    w->st.anc = a.take<int>((size_t)(max_len + 1) * R);
    w->st.anc_tmp = a.take<int>((size_t)(max_len + 1) * R);
    w->st.ignore = a.take<unsigned char>(R);
    w->st.fin_count = a.take<int>(n);
    w->st.fin_tokens = a.take<int>(R * (max_len + 1));
    w->st.fin_len = a.take<int>(R);
    w->st.fin_score = a.take<float>(R);
    w->st.finished = a.take<unsigned char>(n);
    w->st.n_unfinished = a.take<int>(1);
    return mb_align_up(a.off, 256);
}

// one decoder step for all R rows: tokens[:, step] -> logits [R, V]
template <bool F16, int ESLICE, int WARPS, int STAGES, int XE_KEYS = 32>
int launch_cross_enc(mb_ctx* ctx, const bf16* qp, const bf16* enc, bf16* ctxe, int n, int T, int heads,
                     const unsigned char* finished, cudaStream_t s) {
    const size_t smem = (size_t)(STAGES * XE_KEYS) * ESLICE * WARPS * 2 + WARPS * 16 * XE_KEYS * sizeof(float);
    static bool done = false;
    if (!done) {
        MB_CUDA(ctx, cudaFuncSetAttribute(dec_cross_enc_kernel<F16, ESLICE, WARPS, STAGES, XE_KEYS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        done = true;
    }
    dec_cross_enc_kernel<F16, ESLICE, WARPS, STAGES, XE_KEYS><<<n, WARPS * 32, smem, s>>>(qp, enc, ctxe, T, heads, finished);
    MB_LAUNCH_CHECK(ctx);
    return 0;
}

// greedy cross-attention of one layer: q [R, H] (in w.qkv) -> w.att [R, H]
int cross_enc_attention(mb_ctx* ctx, TrocrModel* m, DecodeWs& w, const DecLayer& L, const bf16* enc_out, int n, int beam,
                        cudaStream_t s) {
    const int H = m->dec_dim, E = m->enc_dim, heads = m->dec_heads, T = m->tokens;
    const int R = n * beam;
    TapGemm g;                                    // q'^h = Wk^h^T q^h  : [R, heads*E]
    g.a0 = w.qkv; g.c0 = DH; g.a0_ld = H; g.n = 1; g.h = 1; g.w = R;
    g.wgt = L.ckT_w; g.n_rows_w = heads * E; g.n_out = E;
    g.out = w.qp; g.out_ld = (long long)heads * E; g.out_mode = MB_OUT_BF16;
    g.batches = heads; g.a_col_stride = DH; g.w_row_stride = E; g.out_col_stride = E;
    RC(mb_tap_gemm(ctx, g, s));
    int rc;
    // E = 768: the kernel is bound by how many encoder rows are in flight per SM, not by its arithmetic: one CTA per SM with
    // a four-stage ring (three 48 KB tiles in flight) beats two CTAs with two stages each (MB_XE_STAGES=2 / 3 for A/B)
    // The kernel is bound by the latency of its per-tile chain (cp.async wait -> partial scores -> smem reduction -> softmax
    // -> P E, three barriers), not by arithmetic: 16-key tiles through a four-stage ring (three tiles = 72 KB in flight per
    // CTA, two CTAs / SM) measured 8.70 ms per decode step at 2048 live crops against 9.66 for 32-key tiles x 2 stages
    // (tools/gpu_probe_decode.py; MB_XE_MODE = 0: 32 x 2, 1: 16 x 2 (4 CTAs / SM, 11.8 ms), 2: 16 x 3 (8.8), 3: 16 x 4, 4: 8 warps (12.2)).
    if (cross_tc(m)) {
        rc = mb_cross_enc_tc(ctx, w.qp, enc_out, w.ctxe, n, beam, T, heads, E, w.live, s);
    } else {
    if (beam != 1) return mb_set_err(ctx, MB_ERR_STATE, "cross_enc_attention: the mma.sync kernel is greedy-only");
    static int xe_mode = -1;
    if (xe_mode < 0) { const char* e = getenv("MB_XE_MODE"); xe_mode = e ? atoi(e) : 3; if (xe_mode < 0 || xe_mode > 4) xe_mode = 3; }
    if (E == 768 && xe_mode == 1) rc = ctx->f16 ? launch_cross_enc<true, 192, 4, 2, 16>(ctx, w.qp, enc_out, w.ctxe, n, T, heads, w.st.finished, s) : launch_cross_enc<false, 192, 4, 2, 16>(ctx, w.qp, enc_out, w.ctxe, n, T, heads, w.st.finished, s);
    else if (E == 768 && xe_mode == 2) rc = ctx->f16 ? launch_cross_enc<true, 192, 4, 3, 16>(ctx, w.qp, enc_out, w.ctxe, n, T, heads, w.st.finished, s) : launch_cross_enc<false, 192, 4, 3, 16>(ctx, w.qp, enc_out, w.ctxe, n, T, heads, w.st.finished, s);
    else if (E == 768 && xe_mode == 3) rc = ctx->f16 ? launch_cross_enc<true, 192, 4, 4, 16>(ctx, w.qp, enc_out, w.ctxe, n, T, heads, w.st.finished, s) : launch_cross_enc<false, 192, 4, 4, 16>(ctx, w.qp, enc_out, w.ctxe, n, T, heads, w.st.finished, s);
    else if (E == 768 && xe_mode == 4) rc = ctx->f16 ? launch_cross_enc<true, 96, 8, 3, 16>(ctx, w.qp, enc_out, w.ctxe, n, T, heads, w.st.finished, s) : launch_cross_enc<false, 96, 8, 3, 16>(ctx, w.qp, enc_out, w.ctxe, n, T, heads, w.st.finished, s);
    else if (E == 768) rc = ctx->f16 ? launch_cross_enc<true, 192, 4, 2>(ctx, w.qp, enc_out, w.ctxe, n, T, heads, w.st.finished, s) : launch_cross_enc<false, 192, 4, 2>(ctx, w.qp, enc_out, w.ctxe, n, T, heads, w.st.finished, s);
    else if (E == 1024) rc = ctx->f16 ? launch_cross_enc<true, 256, 4, 2>(ctx, w.qp, enc_out, w.ctxe, n, T, heads, w.st.finished, s) : launch_cross_enc<false, 256, 4, 2>(ctx, w.qp, enc_out, w.ctxe, n, T, heads, w.st.finished, s);
    else if (E == 128) rc = ctx->f16 ? launch_cross_enc<true, 32, 4, 3>(ctx, w.qp, enc_out, w.ctxe, n, T, heads, w.st.finished, s) : launch_cross_enc<false, 32, 4, 3>(ctx, w.qp, enc_out, w.ctxe, n, T, heads, w.st.finished, s);
    else return mb_set_err(ctx, MB_ERR_STATE, "cross_enc_attention: unsupported encoder width %d", E);
    }
    if (rc) return rc;
    TapGemm v;                                    // att^h = Wv^h ctx^h + bv^h : [R, H]
    v.a0 = w.ctxe; v.c0 = E; v.a0_ld = heads * E; v.n = 1; v.h = 1; v.w = R;
    v.wgt = L.ckv_w + (size_t)H * E; v.n_rows_w = H; v.n_out = DH;
    v.bias = L.ckv_b + H;
    v.out = w.att; v.out_ld = H; v.out_mode = MB_OUT_BF16;
    v.batches = heads; v.a_col_stride = E; v.w_row_stride = DH; v.out_col_stride = DH;
    return mb_tap_gemm(ctx, v, s);
}

// K / V caches [L][kv_cap][R][H] x 2 in their own allocation.  need_steps: steps the caches must be able to hold;
// keep_steps: steps already written that must survive a growth (copied layer by layer).
int ensure_kv(mb_ctx* ctx, TrocrModel* m, DecodeWs& w, long long R, int need_steps, int keep_steps, int max_steps,
              cudaStream_t s) {
    if (need_steps <= w.kv_cap) return 0;
    static int first_cap = 0;
    if (!first_cap) { const char* e = getenv("MB_KV_STEPS"); first_cap = e ? atoi(e) : 32; if (first_cap < 1) first_cap = 32; }
    int cap = w.kv_cap ? w.kv_cap * 2 : first_cap;
    while (cap < need_steps) cap *= 2;
    if (cap > max_steps) cap = max_steps;
    const int L = m->dec_layers, H = m->dec_dim;
    const size_t per = (size_t)L * cap * R * H * sizeof(bf16);
    const size_t bytes = 2 * mb_align_up(per, 256);
    void* fresh = nullptr;
    const bool reuse = keep_steps == 0 && bytes <= m->kv_bytes;      // a new decode fits the allocation of the last one
    if (reuse) {
        fresh = m->kv_arena;
    } else {
        if (keep_steps == 0 && m->kv_arena) { cudaFree(m->kv_arena); m->kv_arena = nullptr; m->kv_bytes = 0; }
        if (cudaMalloc(&fresh, bytes) != cudaSuccess) {
            cudaGetLastError();
            return mb_set_err(ctx, MB_ERR_OOM, "trocr: K/V cache of %zu bytes failed", bytes);
        }
    }
    bf16* nk = (bf16*)fresh;
    bf16* nv = (bf16*)((unsigned char*)fresh + mb_align_up(per, 256));
    if (keep_steps > 0) {
        for (int l = 0; l < L; ++l) {
            const size_t n_el = (size_t)keep_steps * R * H;
            MB_CUDA(ctx, cudaMemcpyAsync(nk + (size_t)l * cap * R * H, w.kcache + (size_t)l * w.kv_cap * R * H, n_el * sizeof(bf16),
                                         cudaMemcpyDeviceToDevice, s));
            MB_CUDA(ctx, cudaMemcpyAsync(nv + (size_t)l * cap * R * H, w.vcache + (size_t)l * w.kv_cap * R * H, n_el * sizeof(bf16),
                                         cudaMemcpyDeviceToDevice, s));
        }
        MB_CUDA(ctx, cudaStreamSynchronize(s));
        cudaFree(m->kv_arena);
    }
    if (!reuse) { m->kv_arena = fresh; m->kv_bytes = bytes; }
    w.kcache = nk; w.vcache = nv; w.kv_cap = cap;
    return 0;
}

int decoder_step(mb_ctx* ctx, TrocrModel* m, DecodeWs& w, const bf16* enc_out, int n, int beam, int step, int max_len, cudaStream_t s) {
    const int H = m->dec_dim, T = m->tokens, F = m->dec_ffn, V = m->vocab;
    const int R = n * beam;
    RC(ensure_kv(ctx, m, w, R, step + 1, step, max_len + 1, s));
    dec_embed_kernel<<<R, 128, 0, s>>>(w.st.tokens, max_len + 2, step, m->embed, m->pe, w.x, R, H, sqrtf((float)H), ctx->f16);
    MB_LAUNCH_CHECK(ctx);
    if (w.greedy && cross_tc(m)) RC(mb_live_list(ctx, w.st.finished, n, w.live, s));   // open crops of this step, all layers
    for (int l = 0; l < m->dec_layers; ++l) {
        const DecLayer& L = m->dec[l];
        const size_t cache_off = (size_t)l * w.kv_cap * R * H;
        RC(gemm(ctx, w.x, H, L.sqkv_w, 3 * H, R, 3 * H, L.sqkv_b, MB_ACT_NONE, nullptr, w.qkv, MB_OUT_BF16, s));
        dec_self_attn_kernel<<<dim3(m->dec_heads, R), 32, 0, s>>>(w.qkv, w.kcache + cache_off, w.vcache + cache_off, w.st.anc,
                                                                  w.att, R, H, step, ctx->f16);
        MB_LAUNCH_CHECK(ctx);
        RC(gemm(ctx, w.att, H, L.sout_w, H, R, H, L.sout_b, MB_ACT_NONE, w.x, w.tmp, MB_OUT_BF16, s));
        RC(layernorm(ctx, w.tmp, w.x, L.ln1_w, L.ln1_b, R, H, 1e-5f, s));
        RC(gemm(ctx, w.x, H, L.cq_w, H, R, H, L.cq_b, MB_ACT_NONE, nullptr, w.qkv, MB_OUT_BF16, s));
        if (w.greedy) {
            RC(cross_enc_attention(ctx, m, w, L, enc_out, n, beam, s));
        } else {
            // all beams of a crop share one pass over the crop's cached K/V (q is pre-scaled: weights carry d^-0.5)
            const bf16* kvl = w.cross_kv + (size_t)l * n * T * 2 * H;
            dim3 grid(1, m->dec_heads, n);
            RC(attention_setup(ctx));
            if (ctx->f16) attention_kernel<true><<<grid, 128, ATT_SMEM, s>>>(w.qkv, H, kvl, kvl + H, 2LL * H, w.att, H, beam, T, 1.4426950408889634f);
            else attention_kernel<false><<<grid, 128, ATT_SMEM, s>>>(w.qkv, H, kvl, kvl + H, 2LL * H, w.att, H, beam, T, 1.4426950408889634f);
            MB_LAUNCH_CHECK(ctx);
        }
        RC(gemm(ctx, w.att, H, L.cout_w, H, R, H, L.cout_b, MB_ACT_NONE, w.x, w.tmp, MB_OUT_BF16, s));
        RC(layernorm(ctx, w.tmp, w.x, L.ln2_w, L.ln2_b, R, H, 1e-5f, s));
        RC(gemm(ctx, w.x, H, L.fc1_w, F, R, F, L.fc1_b, MB_ACT_RELU, nullptr, w.ffn, MB_OUT_BF16, s));
        RC(gemm(ctx, w.ffn, F, L.fc2_w, H, R, H, L.fc2_b, MB_ACT_NONE, w.x, w.tmp, MB_OUT_BF16, s));
        RC(layernorm(ctx, w.tmp, w.x, L.ln3_w, L.ln3_b, R, H, 1e-5f, s));
    }
    RC(gemm(ctx, w.x, H, m->out_w, V, R, V, nullptr, MB_ACT_NONE, nullptr, w.logits, MB_OUT_F32, s));
    return 0;
}

int decode_prepare(mb_ctx* ctx, TrocrModel* m, DecodeWs& w, const bf16* enc_out, int n, int beam, int max_len,
                   cudaStream_t s) {
    const int H = m->dec_dim, T = m->tokens, D = m->enc_dim;
    // static cross-attention K/V, once per crop (not per beam): [n*T, 2H] per layer — not needed in greedy mode
    for (int l = 0; l < m->dec_layers && !w.greedy; ++l)
        RC(gemm(ctx, enc_out, D, m->dec[l].ckv_w, 2 * H, (long long)n * T, 2 * H, m->dec[l].ckv_b, MB_ACT_NONE, nullptr,
                w.cross_kv + (size_t)l * n * T * 2 * H, MB_OUT_BF16, s));
    const int R = n * beam;
    search_init_kernel<<<mb_cdiv(R, 128), 128, 0, s>>>(w.st, n, beam, max_len);
    MB_LAUNCH_CHECK(ctx);
    return 0;
}

}  // namespace

void mb_free_trocr(mb_ctx* ctx) {
    if (!ctx->trocr) return;
    ctx->trocr->blob.release();
    if (ctx->trocr->arena) cudaFree(ctx->trocr->arena);
    if (ctx->trocr->kv_arena) cudaFree(ctx->trocr->kv_arena);
    if (ctx->trocr->rec_enc) cudaFree(ctx->trocr->rec_enc);
    delete ctx->trocr;
    ctx->trocr = nullptr;
}

extern "C" int mb_load_trocr(mb_ctx* ctx, const void* blob_host, size_t nbytes) {
    MbDeviceGuard _mb_guard(ctx);
    if (!ctx) return MB_ERR_ARG;
    mb_free_trocr(ctx);
    TrocrModel* m = new TrocrModel();
    ctx->trocr = m;
    int rc = m->blob.load(ctx, blob_host, nbytes);
    if (rc) { mb_free_trocr(ctx); return rc; }
    const int wdt = ctx->f16 ? 3 : 1;
    bool ok = true;
    std::string missing;
    auto W = [&](const std::string& name) -> const bf16* {
        const BlobTensor* t = m->blob.get(name);
        if (!t || t->dtype != wdt) { ok = false; missing = name; return nullptr; }
        return (const bf16*)t->dev;
    };
    auto Fp = [&](const std::string& name) -> const float* {
        const BlobTensor* t = m->blob.get(name);
        if (!t || t->dtype != 0) { ok = false; missing = name; return nullptr; }
        return (const float*)t->dev;
    };
    const BlobTensor* cfg = m->blob.get("config");
    if (!cfg || cfg->dtype != 2 || cfg->nbytes < 11 * 4) {
        mb_free_trocr(ctx);
        return mb_set_err(ctx, MB_ERR_ARG, "trocr blob: config missing");
    }
    int c[11];
    cudaMemcpy(c, cfg->dev, sizeof(c), cudaMemcpyDeviceToHost);
    m->enc_dim = c[0]; m->enc_layers = c[1]; m->enc_heads = c[2]; m->enc_ffn = c[3];
    m->dec_dim = c[4]; m->dec_layers = c[5]; m->dec_heads = c[6]; m->dec_ffn = c[7];
    m->vocab = c[8]; m->tokens = c[9]; m->max_pos = c[10];
    if (m->enc_dim != m->enc_heads * DH || m->dec_dim != m->dec_heads * DH || m->enc_dim % 64 || m->dec_dim % 64 ||
        m->enc_ffn % 64 || m->dec_ffn % 64) {
        mb_free_trocr(ctx);
        return mb_set_err(ctx, MB_ERR_ARG, "trocr blob: unsupported geometry (head dim must be 64, widths multiples of 64)");
    }
    m->patch_w = W("enc.patch.w"); m->patch_b = Fp("enc.patch.b"); m->cls_pos = Fp("enc.cls_pos");
    m->norm_w = Fp("enc.norm.w"); m->norm_b = Fp("enc.norm.b");
    m->enc.resize(m->enc_layers);
    for (int i = 0; i < m->enc_layers; ++i) {
        const std::string p = "enc.L" + std::to_string(i) + ".";
        EncLayer& L = m->enc[i];
        L.ln1_w = Fp(p + "ln1.w"); L.ln1_b = Fp(p + "ln1.b"); L.ln2_w = Fp(p + "ln2.w"); L.ln2_b = Fp(p + "ln2.b");
        L.qkv_w = W(p + "qkv.w"); L.proj_w = W(p + "proj.w"); L.proj_b = Fp(p + "proj.b");
        L.fc1_w = W(p + "fc1.w"); L.fc1_b = Fp(p + "fc1.b"); L.fc2_w = W(p + "fc2.w"); L.fc2_b = Fp(p + "fc2.b");
        if (m->blob.get(p + "qkv.wf") && m->blob.get(p + "fc1.wf")) {      // optional: LayerNorm-folded set
            L.qkv_wf = W(p + "qkv.wf"); L.qkv_c = Fp(p + "qkv.c"); L.qkv_bf = Fp(p + "qkv.bf");
            L.fc1_wf = W(p + "fc1.wf"); L.fc1_c = Fp(p + "fc1.c"); L.fc1_bf = Fp(p + "fc1.bf");
        }
    }
    {
        const char* e = getenv("MB_LNFOLD");
        m->ln_fold = !(e && e[0] == '0');
    }
    m->embed = W("dec.embed"); m->pe = Fp("dec.pe"); m->out_w = W("dec.out.w");
    m->dec.resize(m->dec_layers);
    for (int i = 0; i < m->dec_layers; ++i) {
        const std::string p = "dec.L" + std::to_string(i) + ".";
        DecLayer& L = m->dec[i];
        L.sqkv_w = W(p + "self.qkv.w"); L.sqkv_b = Fp(p + "self.qkv.b"); L.sout_w = W(p + "self.out.w"); L.sout_b = Fp(p + "self.out.b");
        L.cq_w = W(p + "cross.q.w"); L.cq_b = Fp(p + "cross.q.b"); L.ckv_w = W(p + "cross.kv.w"); L.ckv_b = Fp(p + "cross.kv.b"); L.ckT_w = W(p + "cross.kT.w");
        L.cout_w = W(p + "cross.out.w"); L.cout_b = Fp(p + "cross.out.b");
        L.fc1_w = W(p + "fc1.w"); L.fc1_b = Fp(p + "fc1.b"); L.fc2_w = W(p + "fc2.w"); L.fc2_b = Fp(p + "fc2.b");
        L.ln1_w = Fp(p + "ln1.w"); L.ln1_b = Fp(p + "ln1.b"); L.ln2_w = Fp(p + "ln2.w"); L.ln2_b = Fp(p + "ln2.b");
        L.ln3_w = Fp(p + "ln3.w"); L.ln3_b = Fp(p + "ln3.b");
    }
    if (!ok) {
        mb_free_trocr(ctx);
        return mb_set_err(ctx, MB_ERR_ARG, "trocr blob: tensor %s missing or not packed for the context dtype (%s)",
                          missing.c_str(), ctx->f16 ? "fp16" : "bf16");
    }
    return 0;
}

extern "C" int mb_trocr_dims(mb_ctx* ctx, int* dims) {
    MbDeviceGuard _mb_guard(ctx);
    if (!ctx || !ctx->trocr) return mb_set_err(ctx, MB_ERR_STATE, "trocr: weights not loaded");
    TrocrModel* m = ctx->trocr;
    dims[0] = m->enc_dim; dims[1] = m->dec_dim; dims[2] = m->vocab; dims[3] = m->tokens;
    return 0;
}

// Test hook: row-wise LayerNorm of a [rows, D] 16-bit matrix with fp32 gamma / beta.
extern "C" int mb_layernorm16(mb_ctx* ctx, const void* in_dev, void* out_dev, const float* gamma_dev, const float* beta_dev,
                              long long rows, int D, float eps, void* stream) {
    MbDeviceGuard _mb_guard(ctx);
    if (!ctx) return MB_ERR_ARG;
    return layernorm(ctx, (const bf16*)in_dev, (bf16*)out_dev, gamma_dev, beta_dev, rows, D, eps, (cudaStream_t)stream);
}

// Test hook: softmax(Q K^T * scale) V over a packed qkv buffer [n*T, 3*D] (heads of 64) -> out [n*T, D].
// mode 0: tcgen05 kernel (attn_tc.cu); mode 1: mma.sync flash kernel (the one the decoder's cross-attention uses).
extern "C" int mb_attention16(mb_ctx* ctx, const void* qkv_dev, void* out_dev, int n, int T, int D, float scale,
                              int mode, void* stream) {
    MbDeviceGuard _mb_guard(ctx);
    if (!ctx) return MB_ERR_ARG;
    MB_REQUIRE(ctx, n > 0 && T > 0 && D > 0 && D % DH == 0, "attention16: bad geometry");
    cudaStream_t s = (cudaStream_t)stream;
    const float scale_log2e = scale * 1.4426950408889634f;
    const bf16* q = (const bf16*)qkv_dev;
    if (mode == 0) return mb_attention_tc(ctx, q, (bf16*)out_dev, n, T, D, D / DH, scale_log2e, s);
    RC(attention_setup(ctx));
    dim3 grid((T + 63) / 64, D / DH, n);
    if (ctx->f16) attention_kernel<true><<<grid, 128, ATT_SMEM, s>>>(q, 3LL * D, q + D, q + 2 * D, 3LL * D, (bf16*)out_dev, D, T, T, scale_log2e);
    else attention_kernel<false><<<grid, 128, ATT_SMEM, s>>>(q, 3LL * D, q + D, q + 2 * D, 3LL * D, (bf16*)out_dev, D, T, T, scale_log2e);
    MB_LAUNCH_CHECK(ctx);
    return 0;
}

// Test hook: out = a W^T + bias + residual (16-bit) through the residual GEMM whose epilogue also produces the LayerNorm
// statistics of the rows it writes: stats_out [M, 2] fp32 = (-mean, rstd).  part_ws: M * 4 * ceil(N / 256) floats.
extern "C" int mb_gemm16_res_stats(mb_ctx* ctx, const void* a_dev, const void* w_dev, const float* bias_dev,
                                   const void* residual_dev, void* out_dev, long long M, int N, int K, float eps,
                                   float* part_ws_dev, float* stats_out_dev, void* stream) {
    MbDeviceGuard _mb_guard(ctx);
    if (!ctx) return MB_ERR_ARG;
    MB_REQUIRE(ctx, M > 0 && N > 0 && K > 0 && part_ws_dev && stats_out_dev, "gemm16_res_stats: bad arguments");
    return gemm_res_stats(ctx, (const bf16*)a_dev, K, (const bf16*)w_dev, M, N, bias_dev, (const bf16*)residual_dev,
                          (bf16*)out_dev, part_ws_dev, stats_out_dev, eps, (cudaStream_t)stream);
}

// Test hook: the greedy cross-attention core on its own.  qp [rows, heads*E] (per-head projected queries), enc [rows*T, E]
// -> out [rows, heads*E] = softmax_t(qp^h . e_t) . e.  mode 0: tcgen05 / TMA kernel (xattn_tc.cu), 1: mma.sync kernel.
// rows = crops; with beam > 1 (mode 0) qp / out hold rows * beam hypotheses, [crop][beam], sharing the crop's states.
// finished (or null): crops to skip; live_ws: rows + 1 ints of scratch for mode 0's compacted crop list.
extern "C" int mb_cross_enc16(mb_ctx* ctx, const void* qp_dev, const void* enc_dev, void* out_dev, int rows, int beam, int T,
                              int heads, int E, const unsigned char* finished_dev, int32_t* live_ws_dev, int mode,
                              void* stream) {
    MbDeviceGuard _mb_guard(ctx);
    if (!ctx) return MB_ERR_ARG;
    MB_REQUIRE(ctx, rows > 0 && beam >= 1 && T > 0 && heads > 0 && heads <= 16, "cross_enc16: bad geometry");
    MB_REQUIRE(ctx, mode == 0 || beam == 1, "cross_enc16: the mma.sync kernel takes one hypothesis per crop");
    cudaStream_t s = (cudaStream_t)stream;
    const bf16* qp = (const bf16*)qp_dev;
    const bf16* enc = (const bf16*)enc_dev;
    bf16* out = (bf16*)out_dev;
    if (mode == 0) {
        if (finished_dev) {
            MB_REQUIRE(ctx, live_ws_dev != nullptr, "cross_enc16: live_ws is required with a finished mask");
            RC(mb_live_list(ctx, finished_dev, rows, live_ws_dev, s));
        }
        return mb_cross_enc_tc(ctx, qp, enc, out, rows, beam, T, heads, E, finished_dev ? live_ws_dev : nullptr, s);
    }
    if (E == 768) return ctx->f16 ? launch_cross_enc<true, 192, 4, 4, 16>(ctx, qp, enc, out, rows, T, heads, finished_dev, s)
                                  : launch_cross_enc<false, 192, 4, 4, 16>(ctx, qp, enc, out, rows, T, heads, finished_dev, s);
    if (E == 1024) return ctx->f16 ? launch_cross_enc<true, 256, 4, 2>(ctx, qp, enc, out, rows, T, heads, finished_dev, s)
                                   : launch_cross_enc<false, 256, 4, 2>(ctx, qp, enc, out, rows, T, heads, finished_dev, s);
    if (E == 128) return ctx->f16 ? launch_cross_enc<true, 32, 4, 3>(ctx, qp, enc, out, rows, T, heads, finished_dev, s)
                                  : launch_cross_enc<false, 32, 4, 3>(ctx, qp, enc, out, rows, T, heads, finished_dev, s);
    return mb_set_err(ctx, MB_ERR_ARG, "cross_enc16: unsupported encoder width %d", E);
}

// cumulative search statistics: {mb_trocr_decode calls, decoder steps executed, rows (crops * beam) decoded}
extern "C" int mb_trocr_stats(mb_ctx* ctx, unsigned long long* out3) {
    MbDeviceGuard _mb_guard(ctx);
    if (!ctx || !ctx->trocr || !out3) return MB_ERR_STATE;
    out3[0] = ctx->trocr->decode_calls; out3[1] = ctx->trocr->decode_steps; out3[2] = ctx->trocr->decode_rows;
    return 0;
}

extern "C" int mb_trocr_encode(mb_ctx* ctx, const void* patches_dev, int n, void* enc_out_dev, void* stream) {
    MbDeviceGuard _mb_guard(ctx);
    if (!ctx) return MB_ERR_ARG;
    TrocrModel* m = ctx->trocr;
    if (!m) return mb_set_err(ctx, MB_ERR_STATE, "trocr: weights not loaded (mb_load_trocr)");
    MB_REQUIRE(ctx, n > 0, "trocr_encode: empty batch");
    const int D = m->enc_dim, T = m->tokens;
    const size_t wide = (size_t)(3 * D > m->enc_ffn ? 3 * D : m->enc_ffn);
    const size_t M = (size_t)n * T;
    const size_t part_floats = M * 4 * (size_t)mb_cdiv(D, 256);           // (sum, sum of squares) x 2 halves x N tiles per row
    const size_t bytes = (2 * M * D + M * wide) * 2 + M * 8 + part_floats * 4 + 4096;
    RC(ensure_arena(ctx, m, bytes));
    Arena a{(unsigned char*)m->arena, 0, 0};
    bf16* x = a.take<bf16>(M * D);
    bf16* y = a.take<bf16>(M * D);
    bf16* big = a.take<bf16>(M * wide);
    float* stats = a.take<float>(M * 2);
    float* stat_part = a.take<float>(part_floats);
    return encode(ctx, m, (const bf16*)patches_dev, n, (bf16*)enc_out_dev, x, y, big, stats, stat_part, (cudaStream_t)stream);
}

// Greedy (beam 1) / beam search over encoder states.  tokens_out [n, out_ld] i32 (hypothesis incl. the final EOS,
// padded with 1), lengths [n], scores [n] (length-normalised sum of log-probs, generator.py finalize_hypos).
extern "C" int mb_trocr_decode(mb_ctx* ctx, const void* enc_out_dev, int n, int beam, int max_len_b,
                               int32_t* tokens_out_dev, int out_ld, int32_t* lengths_dev, float* scores_dev,
                               int* steps_run, void* stream) {
    MbDeviceGuard _mb_guard(ctx);
    if (!ctx) return MB_ERR_ARG;
    TrocrModel* m = ctx->trocr;
    if (!m) return mb_set_err(ctx, MB_ERR_STATE, "trocr: weights not loaded (mb_load_trocr)");
    MB_REQUIRE(ctx, n > 0 && beam >= 1 && beam <= MAX_BEAM, "trocr_decode: beam must be in 1..%d", MAX_BEAM);
    MB_REQUIRE(ctx, 2 * beam < m->vocab, "trocr_decode: vocabulary smaller than the candidate list");
    cudaStream_t s = (cudaStream_t)stream;
    const int max_len = max_len_b < m->max_pos - 1 ? max_len_b : m->max_pos - 1;   // generator.py:57-61
    MB_REQUIRE(ctx, max_len >= 1, "trocr_decode: max_len must be >= min_len (1)");
    DecodeWs w;
    const size_t bytes = plan_decode(m, n, beam, max_len, nullptr, &w);
    RC(ensure_arena(ctx, m, bytes));
    plan_decode(m, n, beam, max_len, (unsigned char*)m->arena, &w);
    RC(decode_prepare(ctx, m, w, (const bf16*)enc_out_dev, n, beam, max_len, s));
    const int R = n * beam, cand = 2 * beam;
    int step = 0;
    for (; step <= max_len; ++step) {
        RC(decoder_step(ctx, m, w, (const bf16*)enc_out_dev, n, beam, step, max_len, s));
        logits_topk_kernel<<<R, TOPK_THREADS, 0, s>>>(w.logits, m->vocab, m->vocab, cand, step < 1, step >= max_len,
                                                     w.cand_val, w.cand_idx, w.st.finished, beam);
        MB_LAUNCH_CHECK(ctx);
        search_step_kernel<<<mb_cdiv(n, 64), 64, 0, s>>>(w.st, w.cand_val, w.cand_idx, n, beam, step, max_len);
        MB_LAUNCH_CHECK(ctx);
        search_commit_kernel<<<mb_cdiv(R, 128), 128, 0, s>>>(w.st, n, beam, step, max_len);
        MB_LAUNCH_CHECK(ctx);
        if (step >= 3 || (step & 1) == 1 || step == max_len) {   // poll the "all finished" counter (a 4-byte D2H + sync, ~1 % of a step)
            int left = 0;
            MB_CUDA(ctx, cudaMemcpyAsync(&left, w.st.n_unfinished, sizeof(int), cudaMemcpyDeviceToHost, s));
            MB_CUDA(ctx, cudaStreamSynchronize(s));
            if (left == 0) { ++step; break; }
        }
    }
    if (steps_run) *steps_run = step;
    m->decode_calls++; m->decode_steps += step; m->decode_rows += (unsigned long long)n * beam;
    search_pick_kernel<<<mb_cdiv(n, 64), 64, 0, s>>>(w.st, n, beam, max_len, tokens_out_dev, out_ld, lengths_dev, scores_dev);
    MB_LAUNCH_CHECK(ctx);
    MB_CUDA(ctx, cudaStreamSynchronize(s));
    return 0;
}

// Parity hook: teacher-forced decoding (beam 1).  forced [n, L] i32 = the tokens chosen at steps 0..L-1;
// logits_out [L, n, V] fp32 receives the raw decoder outputs of every step (before log-softmax / masking).
extern "C" int mb_trocr_forced_logits(mb_ctx* ctx, const void* enc_out_dev, int n, const int32_t* forced_dev, int L,
                                      float* logits_out_dev, void* stream) {
    MbDeviceGuard _mb_guard(ctx);
    if (!ctx) return MB_ERR_ARG;
    TrocrModel* m = ctx->trocr;
    if (!m) return mb_set_err(ctx, MB_ERR_STATE, "trocr: weights not loaded (mb_load_trocr)");
    MB_REQUIRE(ctx, n > 0 && L >= 1 && L < m->max_pos, "trocr_forced_logits: bad arguments");
    cudaStream_t s = (cudaStream_t)stream;
    DecodeWs w;
    const size_t bytes = plan_decode(m, n, 1, L, nullptr, &w);
    RC(ensure_arena(ctx, m, bytes));
    plan_decode(m, n, 1, L, (unsigned char*)m->arena, &w);
    RC(decode_prepare(ctx, m, w, (const bf16*)enc_out_dev, n, 1, L, s));
    for (int step = 0; step < L; ++step) {
        RC(decoder_step(ctx, m, w, (const bf16*)enc_out_dev, n, 1, step, L, s));
        MB_CUDA(ctx, cudaMemcpyAsync(logits_out_dev + (size_t)step * n * m->vocab, w.logits, (size_t)n * m->vocab * 4,
                                     cudaMemcpyDeviceToDevice, s));
        forced_step_kernel<<<mb_cdiv(n, 128), 128, 0, s>>>(w.st, w.logits, m->vocab, m->vocab, forced_dev, L, n, step, L);
        MB_LAUNCH_CHECK(ctx);
    }
    MB_CUDA(ctx, cudaStreamSynchronize(s));
    return 0;
}

// Full recogniser on packed patch rows (mb_pack_crops / mb_pack_fragments layout 1): encoder + search, processed in
// chunks of `chunk` crops (0 = pick from free memory).  Replaces task.inference_step(generator, ...) inside get_text
// (marie/document/trocr_ocr_processor.py:142-149).
extern "C" int mb_trocr_recognize(mb_ctx* ctx, const void* patches_dev, int n, int beam, int max_len_b, int chunk,
                                  int32_t* tokens_out_dev, int out_ld, int32_t* lengths_dev, float* scores_dev,
                                  void* stream) {
    MbDeviceGuard _mb_guard(ctx);
    if (!ctx) return MB_ERR_ARG;
    TrocrModel* m = ctx->trocr;
    if (!m) return mb_set_err(ctx, MB_ERR_STATE, "trocr: weights not loaded (mb_load_trocr)");
    if (n == 0) return 0;
    MB_REQUIRE(ctx, n > 0 && beam >= 1 && beam <= MAX_BEAM, "trocr_recognize: bad arguments");
    if (chunk <= 0) chunk = 512;
    const int D = m->enc_dim, T = m->tokens;
    const int cmax = n < chunk ? n : chunk;
    const size_t need = (size_t)cmax * T * D * 2;
    if (need > m->rec_enc_bytes) {                       // grown on demand, never per call
        if (m->rec_enc) cudaFree(m->rec_enc);
        m->rec_enc = nullptr; m->rec_enc_bytes = 0;
        if (cudaMalloc(&m->rec_enc, need) != cudaSuccess) {
            cudaGetLastError();
            return mb_set_err(ctx, MB_ERR_OOM, "trocr_recognize: encoder output buffer");
        }
        m->rec_enc_bytes = need;
    }
    void* enc_out = m->rec_enc;
    int rc = 0;
    for (int i0 = 0; i0 < n && !rc; i0 += chunk) {
        const int c = n - i0 < chunk ? n - i0 : chunk;
        rc = mb_trocr_encode(ctx, (const bf16*)patches_dev + (size_t)i0 * (T - 1) * 768, c, enc_out, stream);
        if (!rc)
            rc = mb_trocr_decode(ctx, enc_out, c, beam, max_len_b, tokens_out_dev + (size_t)i0 * out_ld, out_ld,
                                 lengths_dev + i0, scores_dev + i0, nullptr, stream);
    }
    return rc;
}
