// Encoder self-attention on the 5th-generation tensor cores: softmax(Q K^T / 8) V for T = 577 tokens, head dim 64.
// Reference: timm Attention inside AdaptedVisionTransformer.forward_features (marie/models/unilm/trocr/deit.py:35-48,
// 105-146).  Replaces the mma.sync flash kernel (trocr.cu attention_kernel) for the encoder: ncu showed that kernel
// pinned at the legacy HMMA pipe's rate (~255 TFLOP/s with the pipe 92 % busy, profiles/r01_attention_decode_cross.md).
//
// One CTA = 128 query rows of one (image, head); two CTAs per SM (112 KB smem, 256 of 512 TMEM columns each), so one
// CTA's softmax overlaps the other's MMAs.  Per 128-key tile:
//   warp 0      TMA: Q once; K and V tiles (two stages) straight out of the qkv activation buffer [rows, 3D]
//   warp 1      tcgen05.mma  S[128x128] = Q K^T  (A, B K-major);  O[128x64] += P V  (B = V tile as loaded, MN-major)
//   warps 2..5  one thread per query row (TMEM lane): P = exp2((S - ref) / 8 * log2 e) -> 16-bit -> shared memory in
//               the swizzled K-major operand layout, one pass over S per tile (fixed per-row reference, re-based
//               with an in-TMEM rescale of O only on fp16 head-room overflow); after the last tile O / l -> global
// Rows / keys beyond the image's 577 tokens are whatever follows in the buffer (finite) or TMA zero fill; keys are
// masked in the last tile, rows are simply not stored.
#include "common.cuh"
#include "tc_ptx.cuh"

int mb_encode_2d_map(mb_ctx* ctx, CUtensorMap* m, const void* base, long long cols, long long rows, long long ld,
                     int box_cols, int box_rows);

namespace {

constexpr int AT_THREADS = 192;
constexpr int TILE = 128;                 // query rows per CTA
constexpr int KT = 64;                    // keys per step: 64 -> 65 KB smem, 128 TMEM columns, three CTAs per SM
constexpr int HD = 64;                    // head dim
constexpr int TILE_BYTES = TILE * HD * 2; // 16 KB: one [128 x 64] 16-bit operand tile (128-byte rows, SW128)
constexpr int KV_BYTES = KT * HD * 2;     // 8 KB: one K or V tile
constexpr int P_BYTES = TILE * KT * 2;    // 16 KB: P as KT/64 [128 x 64] K-major atoms
constexpr int AT_SMEM = 1024 + TILE_BYTES + 4 * KV_BYTES + P_BYTES;
constexpr int TMEM_COLS_AT = (KT + HD) <= 128 ? 128 : 256;   // S: columns 0..KT-1, O: columns KT..KT+63
constexpr int O_COL = KT;
constexpr int AT_CTAS = KT == 64 ? 3 : 2;

struct AttnParams {
    int T, D, heads;
    float scale_log2e;
    bf16* out;
    int f16;
    unsigned int* diag;
};

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <bool F16>
__global__ void __launch_bounds__(AT_THREADS, AT_CTAS)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmKV, const AttnParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // layout: [mbarriers, 1 KB] [Q] [K x2] [V x2] [P]; the dynamic segment must itself be 1024-byte aligned (128-byte
    // swizzle) — two CTAs of 114 KB fill the SM exactly, there is no room for an alignment pad
    if ((smem_u32(smem_raw) & 1023u) != 0) {
        if (threadIdx.x == 0 && p.diag) atomicExch(p.diag, 0xA11C0000u);
        __trap();
    }
    uint8_t* smem = smem_raw + 1024;
    uint8_t* sQ = smem;
    uint8_t* sK = smem + TILE_BYTES;              // 2 stages
    uint8_t* sV = sK + 2 * KV_BYTES;              // 2 stages
    uint8_t* sP = sV + 2 * KV_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
    uint64_t* q_full = bars;
    uint64_t* kv_full = bars + 1;                 // [2]
    uint64_t* kv_empty = bars + 3;                // [2]
    uint64_t* s_full = bars + 5;
    uint64_t* p_full = bars + 6;
    uint64_t* o_full = bars + 7;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * TILE, head = blockIdx.y;
    const long long img = blockIdx.z;
    const int T = p.T;
    const int n_tiles = (T + KT - 1) / KT;
    const int row_base = (int)(img * T);

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQKV) : "memory");
        mbar_init(smem_u32(q_full), 1);
        for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&kv_full[i]), 1); mbar_init(smem_u32(&kv_empty[i]), 1); }
        mbar_init(smem_u32(s_full), 1);
        mbar_init(smem_u32(p_full), 128);
        mbar_init(smem_u32(o_full), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                     "r"(TMEM_COLS_AT) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_ptr_smem);

    if (warp == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(smem_u32(q_full), TILE_BYTES);
            tma_load_2d(smem_u32(sQ), &tmQKV, smem_u32(q_full), head * HD, row_base + q0);
            for (int j = 0; j < n_tiles; ++j) {
                const int st = j & 1;
                mbar_wait(smem_u32(&kv_empty[st]), ((j >> 1) & 1) ^ 1, p.diag, 11);
                const uint32_t fb = smem_u32(&kv_full[st]);
                mbar_arrive_expect_tx(fb, 2 * KV_BYTES);
                tma_load_2d(smem_u32(sK + st * KV_BYTES), &tmKV, fb, p.D + head * HD, row_base + j * KT);
                tma_load_2d(smem_u32(sV + st * KV_BYTES), &tmKV, fb, 2 * p.D + head * HD, row_base + j * KT);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t ab_fmt = F16 ? 0u : ((1u << 7) | (1u << 10));
            const uint32_t idesc_qk = (1u << 4) | ab_fmt | ((uint32_t)(KT >> 3) << 17) | ((uint32_t)(TILE >> 4) << 24);
            // P V: B operand (V tile: keys x 64 contiguous head dims) is MN-major — bit 16
            const uint32_t idesc_pv = (1u << 4) | ab_fmt | (1u << 16) | ((uint32_t)(HD >> 3) << 17) | ((uint32_t)(TILE >> 4) << 24);
            mbar_wait(smem_u32(q_full), 0, p.diag, 12);
            for (int j = 0; j < n_tiles; ++j) {
                const int st = j & 1;
                mbar_wait(smem_u32(&kv_full[st]), (j >> 1) & 1, p.diag, 13);
                tcgen05_fence_after();
                const uint64_t qd = make_smem_desc(smem_u32(sQ));
                const uint64_t kd = make_smem_desc(smem_u32(sK + st * KV_BYTES));
#pragma unroll
                for (int k = 0; k < HD / 16; ++k)
                    umma_bf16(tmem_base, qd + (uint64_t)(2 * k), kd + (uint64_t)(2 * k), idesc_qk, k > 0 ? 1u : 0u);
                tcgen05_commit(smem_u32(s_full));
                mbar_wait(smem_u32(p_full), j & 1, p.diag, 14);
                tcgen05_fence_after();
                const uint64_t vd = make_smem_desc(smem_u32(sV + st * KV_BYTES));
#pragma unroll
                for (int k = 0; k < KT / 16; ++k) {
                    // A = P: KT/64 K-major atoms of 64 keys (16 KB apart), 32 B per 16-key step inside an atom
                    const uint64_t pd = make_smem_desc(smem_u32(sP + (k >> 2) * TILE_BYTES)) + (uint64_t)(2 * (k & 3));
                    // B = V tile, MN-major: 16 keys = two 8-row groups of 1024 B
                    umma_bf16(tmem_base + O_COL, pd, vd + (uint64_t)((k * 2048) >> 4), idesc_pv, (j > 0 || k > 0) ? 1u : 0u);
                }
                tcgen05_commit(smem_u32(&kv_empty[st]));
                tcgen05_commit(smem_u32(o_full));
            }
        }
    } else {
        // ------------------------------------------------------------ softmax / correction / epilogue: one row per thread
        const int quarter = warp & 3;                         // TMEM lane quarter this warp may access
        const int row = quarter * 32 + lane;                  // query row inside the tile == TMEM lane
        const uint32_t t_s = tmem_base + ((uint32_t)(quarter * 32) << 16);
        const uint32_t t_o = t_s + O_COL;
        float l_run = 0.f;
        float ms = 0.f;                                       // -(reference score) * scale * log2(e) of this row
        const float sc = p.scale_log2e;
        uint32_t va[32], vb[32];
        // The softmax is shift invariant, so instead of the running row maximum (which costs a second pass over S in
        // TMEM per tile) the row keeps ONE reference: the exact maximum of its first key tile.  Later tiles are read
        // once; probabilities may exceed 1 and only when one passes 2^11 (fp16 operand head-room) the row is re-based:
        // O and l are scaled in place and the tile's P is recomputed.  The reference never exceeds the true maximum,
        // so the largest probability of a row is >= 1 and nothing underflows as a whole.
        // TMEM loads are software-pipelined: chunk c+1 is in flight while chunk c is processed.
        auto pass_p = [&](int valid, float ms_, float& pmax) -> float {
            float lsum = 0.f;
            float m0 = 0.f, m1 = 0.f, m2 = 0.f, m3 = 0.f;
            tmem_ld32(t_s, va);
#pragma unroll
            for (int ci = 0; ci < KT / 32; ++ci) {
                uint32_t* cur = (ci & 1) ? vb : va;
                uint32_t* nxt = (ci & 1) ? va : vb;
                tmem_wait_ld();
                if (ci + 1 < KT / 32) tmem_ld32(t_s + (uint32_t)((ci + 1) * 32), nxt);
                const int c = ci * 32;
                float pr[32];
                float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    pr[i] = ex2_approx(fmaf(__uint_as_float(cur[i]), sc, ms_));
                    pr[i + 1] = ex2_approx(fmaf(__uint_as_float(cur[i + 1]), sc, ms_));
                    pr[i + 2] = ex2_approx(fmaf(__uint_as_float(cur[i + 2]), sc, ms_));
                    pr[i + 3] = ex2_approx(fmaf(__uint_as_float(cur[i + 3]), sc, ms_));
                    if (c + 32 > valid) {                      // ragged last tile only
                        if (c + i >= valid) pr[i] = 0.f;
                        if (c + i + 1 >= valid) pr[i + 1] = 0.f;
                        if (c + i + 2 >= valid) pr[i + 2] = 0.f;
                        if (c + i + 3 >= valid) pr[i + 3] = 0.f;
                    }
                    s0 += pr[i]; s1 += pr[i + 1]; s2 += pr[i + 2]; s3 += pr[i + 3];
                    m0 = fmaxf(m0, pr[i]); m1 = fmaxf(m1, pr[i + 1]); m2 = fmaxf(m2, pr[i + 2]); m3 = fmaxf(m3, pr[i + 3]);
                }
                lsum += (s0 + s1) + (s2 + s3);
                uint8_t* prow = sP + (c >> 6) * TILE_BYTES + row * 128;
                const int chunk0 = (c & 63) >> 3;             // first 16-byte chunk of these 32 columns inside the atom row
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    uint4 w;
                    w.x = pack2(pr[8 * q + 0], pr[8 * q + 1], F16 ? 1 : 0);
                    w.y = pack2(pr[8 * q + 2], pr[8 * q + 3], F16 ? 1 : 0);
                    w.z = pack2(pr[8 * q + 4], pr[8 * q + 5], F16 ? 1 : 0);
                    w.w = pack2(pr[8 * q + 6], pr[8 * q + 7], F16 ? 1 : 0);
                    *reinterpret_cast<uint4*>(prow + (((chunk0 + q) ^ (row & 7)) << 4)) = w;
                }
            }
            pmax = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
            return lsum;
        };
        for (int j = 0; j < n_tiles; ++j) {
            mbar_wait(smem_u32(s_full), j & 1, p.diag, 15);
            tcgen05_fence_after();
            const int valid = T - j * KT;                     // keys >= valid are beyond this image
            if (j == 0) {
                // exact row maximum of the first tile = the row's reference
                float mx = -INFINITY;
                tmem_ld32(t_s, va);
#pragma unroll
                for (int ci = 0; ci < KT / 32; ++ci) {
                    uint32_t* cur = (ci & 1) ? vb : va;
                    uint32_t* nxt = (ci & 1) ? va : vb;
                    tmem_wait_ld();
                    if (ci + 1 < KT / 32) tmem_ld32(t_s + (uint32_t)((ci + 1) * 32), nxt);
                    const int c = ci * 32;
                    float a0 = -INFINITY, a1 = -INFINITY, a2 = -INFINITY, a3 = -INFINITY;
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        if (c + i < valid) a0 = fmaxf(a0, __uint_as_float(cur[i]));
                        if (c + i + 1 < valid) a1 = fmaxf(a1, __uint_as_float(cur[i + 1]));
                        if (c + i + 2 < valid) a2 = fmaxf(a2, __uint_as_float(cur[i + 2]));
                        if (c + i + 3 < valid) a3 = fmaxf(a3, __uint_as_float(cur[i + 3]));
                    }
                    mx = fmaxf(mx, fmaxf(fmaxf(a0, a1), fmaxf(a2, a3)));
                }
                ms = -mx * sc;
            }
            float pmax;
            float lt = pass_p(valid, ms, pmax);
            if (__any_sync(0xffffffffu, pmax > 2048.f)) {
                // re-base the rows that grew: shift by log2(pmax) so their largest probability becomes 1
                const float lg = pmax > 1.f ? log2f(pmax) : 0.f;
                const float f = ex2_approx(-lg);
                ms -= lg;
                l_run *= f;
                if (j > 0) {                                  // s_full(j) was committed after P V (j-1): O is complete
#pragma unroll 1
                    for (int c = 0; c < HD; c += 32) {
                        tmem_ld32(t_o + (uint32_t)c, vb);
                        tmem_wait_ld();
#pragma unroll
                        for (int i = 0; i < 32; ++i) vb[i] = __float_as_uint(__uint_as_float(vb[i]) * f);
                        tmem_st32(t_o + (uint32_t)c, vb);
                    }
                    tmem_wait_st();
                }
                lt = pass_p(valid, ms, pmax);
            }
            l_run += lt;
            fence_proxy_async_smem();                         // generic-proxy writes of P -> visible to the MMA (async proxy)
            tcgen05_fence_before();
            mbar_arrive(smem_u32(p_full));
        }
        // epilogue: O / l -> 16-bit -> global (one 128-byte row per thread)
        mbar_wait(smem_u32(o_full), (n_tiles - 1) & 1, p.diag, 16);
        tcgen05_fence_after();
        const float inv = 1.0f / l_run;
        const bool store = q0 + row < T;
        bf16* orow = p.out + ((long long)row_base + q0 + row) * p.D + head * HD;
#pragma unroll 1
        for (int c = 0; c < HD; c += 32) {
            uint32_t v[32];
            tmem_ld32(t_o + (uint32_t)c, v);
            tmem_wait_ld();
            if (store) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    uint4 w;
                    w.x = pack2(__uint_as_float(v[8 * q + 0]) * inv, __uint_as_float(v[8 * q + 1]) * inv, F16 ? 1 : 0);
                    w.y = pack2(__uint_as_float(v[8 * q + 2]) * inv, __uint_as_float(v[8 * q + 3]) * inv, F16 ? 1 : 0);
                    w.z = pack2(__uint_as_float(v[8 * q + 4]) * inv, __uint_as_float(v[8 * q + 5]) * inv, F16 ? 1 : 0);
                    w.w = pack2(__uint_as_float(v[8 * q + 6]) * inv, __uint_as_float(v[8 * q + 7]) * inv, F16 ? 1 : 0);
                    *reinterpret_cast<uint4*>(orow + c + 8 * q) = w;
                }
            }
        }
        tcgen05_fence_before();
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS_AT) : "memory");
    }
}

}  // namespace

// qkv: [n*T, 3*D] 16-bit (q | k | v, head h at columns h*64) -> out [n*T, D]
int mb_attention_tc(mb_ctx* ctx, const bf16* qkv, bf16* out, int n, int T, int D, int heads, float scale_log2e,
                    cudaStream_t stream) {
    MB_REQUIRE(ctx, D == heads * HD && n > 0 && T > 0, "attention_tc: bad geometry");
    CUtensorMap tm, tmkv;
    int rc = mb_encode_2d_map(ctx, &tm, qkv, 3LL * D, (long long)n * T, 3LL * D, HD, TILE);
    if (rc) return rc;
    rc = mb_encode_2d_map(ctx, &tmkv, qkv, 3LL * D, (long long)n * T, 3LL * D, HD, KT);
    if (rc) return rc;
    static bool attr_set = false;
    if (!attr_set) {
        MB_CUDA(ctx, cudaFuncSetAttribute(attn_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM));
        MB_CUDA(ctx, cudaFuncSetAttribute(attn_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM));
        attr_set = true;
    }
    AttnParams p;
    p.T = T; p.D = D; p.heads = heads; p.scale_log2e = scale_log2e; p.out = out; p.f16 = ctx->f16; p.diag = ctx->dev_diag;
    dim3 grid((T + TILE - 1) / TILE, heads, n);
    if (ctx->f16) attn_tc_kernel<true><<<grid, AT_THREADS, AT_SMEM, stream>>>(tm, tmkv, p);
    else attn_tc_kernel<false><<<grid, AT_THREADS, AT_SMEM, stream>>>(tm, tmkv, p);
    MB_LAUNCH_CHECK(ctx);
    return 0;
}
