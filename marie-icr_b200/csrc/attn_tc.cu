// Encoder self-attention on the 5th-generation tensor cores: softmax(Q K^T / 8) V for T = 577 tokens, head dim 64.
// Reference: timm Attention inside AdaptedVisionTransformer.forward_features (marie/models/unilm/trocr/deit.py:35-48,
// 105-146).  Replaces the mma.sync flash kernel (trocr.cu attention_kernel) for the encoder: ncu showed that kernel
// pinned at the legacy HMMA pipe's rate (~255 TFLOP/s with the pipe 92 % busy, profiles/r01_attention_decode_cross.md).
//
// One CTA = 128 query rows of one (image, head); two CTAs per SM (98 KB smem, 256 of 512 TMEM columns each).
// The kernel is bound by the MUFU pipe (one ex2 per score), so the schedule keeps the softmax warps fed: S is DOUBLE
// BUFFERED in TMEM and P in shared memory, and the MMA warp issues S(j+1) = Q K(j+1)^T before it waits for P(j) — while
// the softmax warps work on tile j the tensor pipe already produces tile j+1, and P V of tile j runs under the softmax
// of tile j+1.  (The first version serialised S -> softmax -> P V inside a CTA and relied on three resident CTAs to
// overlap: ncu 331 TFLOP/s, MUFU a third busy.)  Per 64-key tile:
//   warp 0      TMA: Q once; K and V tiles (three stages) straight out of the qkv activation buffer [rows, 3D]
//   warp 1      tcgen05.mma  S[128x64] = Q K^T  (A, B K-major);  O[128x64] += P V  (B = V tile as loaded, MN-major)
//   warps 2..5  one thread per query row (TMEM lane): P = exp2((S - ref) / 8 * log2 e) -> 16-bit -> shared memory in
//               the swizzled K-major operand layout, one pass over S per tile (fixed per-row reference, re-based
//               with an in-TMEM rescale of O only on fp16 head-room overflow); after the last tile O / l -> global
// Rows / keys beyond the image's 577 tokens are whatever follows in the buffer (finite) or TMA zero fill; keys are
// masked in the last tile, rows are simply not stored.
#include "common.cuh"
#include "tc_ptx.cuh"
#include <stdlib.h>

int mb_encode_2d_map(mb_ctx* ctx, CUtensorMap* m, const void* base, long long cols, long long rows, long long ld,
                     int box_cols, int box_rows);

namespace {

constexpr int AT_THREADS = 192;
constexpr int TILE = 128;                 // query rows per CTA
constexpr int KT = 64;                    // keys per step
constexpr int HD = 64;                    // head dim
constexpr int KV_STAGES = 3;
constexpr int TILE_BYTES = TILE * HD * 2; // 16 KB: one [128 x 64] 16-bit operand tile (128-byte rows, SW128)
constexpr int KV_BYTES = KT * HD * 2;     // 8 KB: one K or V tile
constexpr int P_BYTES = TILE * KT * 2;    // 16 KB: P as one [128 x 64] K-major atom
constexpr int AT_SMEM = 1024 + TILE_BYTES + 2 * KV_STAGES * KV_BYTES + 2 * P_BYTES;
constexpr int TMEM_COLS_AT = 256;         // S0: columns 0..63, S1: 64..127, O: 128..191
constexpr int O_COL = 2 * KT;
constexpr int AT_CTAS = 2;
constexpr int MAX_TAIL = 2;                 // tail K / V rows are staged in the barrier page: MAX_TAIL * 256 B at +512
static_assert(KT == 64, "P is one 64-key K-major atom per buffer");

struct AttnParams {
    int T, D, heads;
    float scale_log2e;
    const bf16* qkv;
    bf16* out;
    int f16;
    unsigned int* diag;
    int poly;              // exp2 of every second score pair on the FMA pipe (packed polynomial) instead of MUFU
};

// exp2 of a PAIR of values on the FMA pipe: x = n + f with n = round(x), |f| <= 0.5 (magic-number rounding), a degree-4
// polynomial for 2^f (max relative error 3.1e-6, far below the 16-bit rounding of P) in packed fp32, and the exponent
// inserted with one integer shift-add.  x is clamped at -125, so tiny probabilities stay normal numbers.  The MUFU pipe
// (16 ex2 / clk / SM) is what bounds the softmax warps; half of the scores take this route.
__device__ __forceinline__ void exp2_poly2(float x0, float x1, float& y0, float& y1) {
    const float MAGIC = 12582912.0f;                    // 1.5 * 2^23: x + MAGIC rounds x to the nearest integer
    const f32x2 x = pk2(fmaxf(x0, -125.0f), fmaxf(x1, -125.0f));
    const f32x2 t = add2(x, pk2(MAGIC, MAGIC));
    const f32x2 n = add2(t, pk2(-MAGIC, -MAGIC));
    const f32x2 f = fma2(n, pk2(-1.0f, -1.0f), x);
    f32x2 q = fma2(pk2(0.009600395517050074f, 0.009600395517050074f), f, pk2(0.05591689382504809f, 0.05591689382504809f));
    q = fma2(q, f, pk2(0.24023718463104982f, 0.24023718463104982f));
    q = fma2(q, f, pk2(0.6931219948398218f, 0.6931219948398218f));
    q = fma2(q, f, pk2(1.0f, 1.0f));
    float q0, q1, t0, t1;
    upk2(q, q0, q1);
    upk2(t, t0, t1);
    y0 = __int_as_float(__float_as_int(q0) + (__float_as_int(t0) << 23));
    y1 = __int_as_float(__float_as_int(q1) + (__float_as_int(t1) << 23));
}

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// 16-bit pack WITHOUT saturation: P of a tile that overflows the fp16 range is never consumed (the row is re-based and
// the tile recomputed before the MMA is released), and two FMNMX per score are what made the loop issue-bound.
template <bool F16> __device__ __forceinline__ uint32_t pack2_raw(float lo, float hi) {
    if (F16) {
        __half2 h = __floats2half2_rn(lo, hi);
        return *reinterpret_cast<uint32_t*>(&h);
    }
    return pack_bf16x2(lo, hi);
}

// One pass over a 64-key S tile of this thread's row: P = exp2(S * sc + ms) -> 16-bit -> shared memory (swizzled K-major
// atom); returns the row's sum of P.  RAGGED masks keys >= valid (a partial last tile only).  TMEM loads are
// software-pipelined: chunk c+1 is in flight while chunk c is processed.  Per score: FFMA, MUFU.EX2, FADD, half a pack.
template <bool F16, bool RAGGED, bool POLY = false>
__device__ __forceinline__ float attn_pass_p(uint32_t t_s, uint8_t* pbuf, int row, int valid, float ms_, float sc,
                                             uint32_t* va, uint32_t* vb) {
    float lsum = 0.f;
    tmem_ld32(t_s, va);
#pragma unroll
    for (int ci = 0; ci < KT / 32; ++ci) {
        uint32_t* cur = (ci & 1) ? vb : va;
        uint32_t* nxt = (ci & 1) ? va : vb;
        tmem_wait_ld();
        if (ci + 1 < KT / 32) tmem_ld32(t_s + (uint32_t)((ci + 1) * 32), nxt);
        const int c = ci * 32;
        // packed fp32 arithmetic (FFMA2 / FADD2): half an instruction per score for the affine map and for the row sum
        float pr[32];
        const f32x2 sc2 = pk2(sc, sc), ms2 = pk2(ms_, ms_);
        f32x2 acc[4] = {0ull, 0ull, 0ull, 0ull};
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
            float x0, x1, x2, x3;
            upk2(fma2(pk2(cur[i], cur[i + 1]), sc2, ms2), x0, x1);
            upk2(fma2(pk2(cur[i + 2], cur[i + 3]), sc2, ms2), x2, x3);
            pr[i] = ex2_approx(x0);
            pr[i + 1] = ex2_approx(x1);
            if (POLY) {
                exp2_poly2(x2, x3, pr[i + 2], pr[i + 3]);
            } else {
                pr[i + 2] = ex2_approx(x2);
                pr[i + 3] = ex2_approx(x3);
            }
            if (RAGGED) {
                if (c + i >= valid) pr[i] = 0.f;
                if (c + i + 1 >= valid) pr[i + 1] = 0.f;
                if (c + i + 2 >= valid) pr[i + 2] = 0.f;
                if (c + i + 3 >= valid) pr[i + 3] = 0.f;
            }
            acc[(i >> 2) & 1] = add2(acc[(i >> 2) & 1], pk2(pr[i], pr[i + 1]));
            acc[2 + ((i >> 2) & 1)] = add2(acc[2 + ((i >> 2) & 1)], pk2(pr[i + 2], pr[i + 3]));
        }
        {
            float a0, a1;
            upk2(add2(add2(acc[0], acc[1]), add2(acc[2], acc[3])), a0, a1);
            lsum += a0 + a1;
        }
        uint8_t* prow = pbuf + row * 128;
        const int chunk0 = c >> 3;                    // first 16-byte chunk of these 32 columns inside the atom row
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint4 w;
            w.x = pack2_raw<F16>(pr[8 * q + 0], pr[8 * q + 1]);
            w.y = pack2_raw<F16>(pr[8 * q + 2], pr[8 * q + 3]);
            w.z = pack2_raw<F16>(pr[8 * q + 4], pr[8 * q + 5]);
            w.w = pack2_raw<F16>(pr[8 * q + 6], pr[8 * q + 7]);
            *reinterpret_cast<uint4*>(prow + (((chunk0 + q) ^ (row & 7)) << 4)) = w;
        }
    }
    return lsum;
}

template <bool F16>
__global__ void __launch_bounds__(AT_THREADS, AT_CTAS)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmKV, const AttnParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // layout: [mbarriers, 1 KB] [Q] [K x3] [V x3] [P x2]; the dynamic segment must itself be 1024-byte aligned (128-byte
    // swizzle)
    if ((smem_u32(smem_raw) & 1023u) != 0) {
        if (threadIdx.x == 0 && p.diag) atomicExch(p.diag, 0xA11C0000u);
        __trap();
    }
    uint8_t* smem = smem_raw + 1024;
    uint8_t* sQ = smem;
    uint8_t* sK = smem + TILE_BYTES;
    uint8_t* sV = sK + KV_STAGES * KV_BYTES;
    uint8_t* sP = sV + KV_STAGES * KV_BYTES;      // 2 buffers
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
    uint64_t* q_full = bars;
    uint64_t* kv_full = bars + 1;                 // [3]  TMA -> MMA
    uint64_t* kv_empty = bars + 4;                // [3]  P V (j) retired -> TMA
    uint64_t* s_full = bars + 7;                  // [2]  S(j) complete -> softmax
    uint64_t* p_full = bars + 9;                  // [2]  P(j) written and S(j) consumed -> MMA (128 arrivals)
    uint64_t* pv_done = bars + 11;                // [2]  P V (j) retired: P buffer reusable, O up to tile j complete
    uint64_t* tail_full = bars + 13;              //      tail K / V rows landed in shared memory (cp.async, warp 0)
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 14);
    uint8_t* sTail = smem_raw + 512;              // [MAX_TAIL][K row 128 B | V row 128 B]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * TILE, head = blockIdx.y;
    const long long img = blockIdx.z;
    const int T = p.T;
    // T = 577 is nine 64-key tiles plus ONE key: a remainder of up to MAX_TAIL keys is folded into the epilogue on the
    // CUDA cores (two 64-long dot products per key and row) instead of costing a whole masked tile of MUFU / MMA work
    const int tail = (T >= KT && (T % KT) <= MAX_TAIL) ? (T % KT) : 0;
    const int n_tiles = tail ? T / KT : (T + KT - 1) / KT;
    const int row_base = (int)(img * T);

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQKV) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmKV) : "memory");
        mbar_init(smem_u32(q_full), 1);
        for (int i = 0; i < KV_STAGES; ++i) { mbar_init(smem_u32(&kv_full[i]), 1); mbar_init(smem_u32(&kv_empty[i]), 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(smem_u32(&s_full[i]), 1);
            mbar_init(smem_u32(&p_full[i]), 128);
            mbar_init(smem_u32(&pv_done[i]), 1);
        }
        mbar_init(smem_u32(tail_full), 32);
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
        // tail K / V rows (16 B per lane) go to shared memory asynchronously; the epilogue picks them up a few
        // microseconds later without a global-memory round trip on its critical path
        if (lane < 16 * tail) {
            const int t = lane >> 4, part = lane & 15;        // part 0..7: K row, 8..15: V row
            const bf16* g = p.qkv + ((long long)row_base + n_tiles * KT + t) * (3LL * p.D) + (part < 8 ? p.D : 2 * p.D) +
                            head * HD + (part & 7) * 8;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(sTail + t * 256 + part * 16)), "l"(g)
                         : "memory");
        }
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(tail_full)) : "memory");
        if (lane == 0) {
            mbar_arrive_expect_tx(smem_u32(q_full), TILE_BYTES);
            tma_load_2d(smem_u32(sQ), &tmQKV, smem_u32(q_full), head * HD, row_base + q0);
            int st = 0, use = 0;                  // stage of tile j, number of times the ring wrapped
            for (int j = 0; j < n_tiles; ++j) {
                mbar_wait(smem_u32(&kv_empty[st]), (use & 1) ^ 1, p.diag, 11);
                const uint32_t fb = smem_u32(&kv_full[st]);
                mbar_arrive_expect_tx(fb, 2 * KV_BYTES);
                tma_load_2d(smem_u32(sK + st * KV_BYTES), &tmKV, fb, p.D + head * HD, row_base + j * KT);
                tma_load_2d(smem_u32(sV + st * KV_BYTES), &tmKV, fb, 2 * p.D + head * HD, row_base + j * KT);
                if (++st == KV_STAGES) { st = 0; ++use; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t ab_fmt = F16 ? 0u : ((1u << 7) | (1u << 10));
            const uint32_t idesc_qk = (1u << 4) | ab_fmt | ((uint32_t)(KT >> 3) << 17) | ((uint32_t)(TILE >> 4) << 24);
            // P V: B operand (V tile: keys x 64 contiguous head dims) is MN-major — bit 16
            const uint32_t idesc_pv = (1u << 4) | ab_fmt | (1u << 16) | ((uint32_t)(HD >> 3) << 17) | ((uint32_t)(TILE >> 4) << 24);
            const uint64_t qd = make_smem_desc(smem_u32(sQ));
            auto issue_s = [&](int j, int st, int use) {
                mbar_wait(smem_u32(&kv_full[st]), use & 1, p.diag, 13);
                tcgen05_fence_after();
                const uint64_t kd = make_smem_desc(smem_u32(sK + st * KV_BYTES));
                const uint32_t d = tmem_base + (uint32_t)((j & 1) * KT);
#pragma unroll
                for (int k = 0; k < HD / 16; ++k)
                    umma_bf16(d, qd + (uint64_t)(2 * k), kd + (uint64_t)(2 * k), idesc_qk, k > 0 ? 1u : 0u);
                tcgen05_commit(smem_u32(&s_full[j & 1]));
            };
            mbar_wait(smem_u32(q_full), 0, p.diag, 12);
            issue_s(0, 0, 0);
            int st = 0, use = 0;                  // stage / wrap count of tile j
            for (int j = 0; j < n_tiles; ++j) {
                int st1 = st + 1, use1 = use;
                if (st1 == KV_STAGES) { st1 = 0; ++use1; }
                // S(j+1) goes to the tensor pipe before P(j) is awaited: its buffer was released by p_full(j-1), which
                // this thread observed before issuing P V (j-1)
                if (j + 1 < n_tiles) issue_s(j + 1, st1, use1);
                mbar_wait(smem_u32(&p_full[j & 1]), (j >> 1) & 1, p.diag, 14);
                tcgen05_fence_after();
                const uint64_t vd = make_smem_desc(smem_u32(sV + st * KV_BYTES));
                const uint64_t pd0 = make_smem_desc(smem_u32(sP + (j & 1) * P_BYTES));
#pragma unroll
                for (int k = 0; k < KT / 16; ++k) {
                    // A = P: 32 B per 16-key step inside the K-major atom; B = V tile, MN-major: 16 keys = two 8-row
                    // groups of 1024 B
                    umma_bf16(tmem_base + O_COL, pd0 + (uint64_t)(2 * k), vd + (uint64_t)((k * 2048) >> 4), idesc_pv,
                              (j > 0 || k > 0) ? 1u : 0u);
                }
                tcgen05_commit(smem_u32(&kv_empty[st]));
                tcgen05_commit(smem_u32(&pv_done[j & 1]));
                st = st1; use = use1;
            }
        }
    } else {
        // ------------------------------------------------------------ softmax / correction / epilogue: one row per thread
        const int quarter = warp & 3;                         // TMEM lane quarter this warp may access
        const int row = quarter * 32 + lane;                  // query row inside the tile == TMEM lane
        const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
        const uint32_t t_o = t_lane + O_COL;
        float l_run = 0.f;
        float ms = 0.f;                                       // -(reference score) * scale * log2(e) of this row
        const float sc = p.scale_log2e;
        uint32_t va[32], vb[32];
        // The softmax is shift invariant, so instead of the running row maximum (which costs a second pass over S in
        // TMEM per tile) the row keeps ONE reference: the exact maximum of its first key tile.  Later tiles are read
        // once; probabilities may exceed 1 and only when a row's tile sum passes 2^11 (fp16 operand head-room; the sum
        // bounds every probability) the row is re-based: O and l are scaled in place and the tile's P is recomputed.
        // The reference never exceeds the true maximum, so the largest probability of a row is >= 1 and nothing
        // underflows as a whole.  TMEM loads are software-pipelined: chunk c+1 is in flight while chunk c is processed.
        // exact maximum of the valid scores of a tile (first tile: the reference; re-basing: the new reference)
        auto pass_max = [&](uint32_t t_s, int valid) -> float {
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
            return mx;
        };
        for (int j = 0; j < n_tiles; ++j) {
            const int b = j & 1;
            const uint32_t t_s = t_lane + (uint32_t)(b * KT);
            uint8_t* pbuf = sP + b * P_BYTES;
            mbar_wait(smem_u32(&s_full[b]), (j >> 1) & 1, p.diag, 15);
            // P buffer b was last read by P V (j-2)
            if (j >= 2) mbar_wait(smem_u32(&pv_done[b]), ((j >> 1) - 1) & 1, p.diag, 17);
            tcgen05_fence_after();
            const int valid = T - j * KT;                     // keys >= valid are beyond this image
            const bool ragged = valid < KT;                   // CTA-uniform: a partial last tile only
            if (j == 0) ms = -pass_max(t_s, valid) * sc;      // the row's reference
            float lt = ragged ? attn_pass_p<F16, true>(t_s, pbuf, row, valid, ms, sc, va, vb)
                              : (p.poly ? attn_pass_p<F16, false, true>(t_s, pbuf, row, valid, ms, sc, va, vb)
                                        : attn_pass_p<F16, false>(t_s, pbuf, row, valid, ms, sc, va, vb));
            if (__any_sync(0xffffffffu, !(lt <= 2048.f))) {
                // re-base the rows that grew: the tile's maximum becomes the new reference (their largest probability
                // of this tile becomes 1); everything accumulated so far shrinks by the same factor
                const float ms_new = -pass_max(t_s, valid) * sc;   // all lanes: tcgen05.ld is warp-collective
                float f = 1.f;
                if (!(lt <= 2048.f)) {
                    f = ex2_approx(ms_new - ms);              // 2^-(shift), shift > 0
                    ms = ms_new;
                }
                l_run *= f;
                if (j > 0) {
                    // O must hold tiles 0..j-1 completely and no P V may be in flight: P V (j) is not issued before this
                    // thread's p_full arrival below
                    mbar_wait(smem_u32(&pv_done[(j - 1) & 1]), ((j - 1) >> 1) & 1, p.diag, 18);
                    tcgen05_fence_after();
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
                lt = attn_pass_p<F16, true>(t_s, pbuf, row, valid, ms, sc, va, vb);
            }
            l_run += lt;
            fence_proxy_async_smem();                         // generic-proxy writes of P -> visible to the MMA (async proxy)
            tcgen05_fence_before();
            mbar_arrive(smem_u32(&p_full[b]));
        }
        // tail keys on the CUDA cores: s = q . k, p = exp2(s * sc + ms) (re-basing the row if p would leave the fp32
        // comfort zone), folded into l and O below
        float pt[MAX_TAIL];
        float fo = 1.f;                                       // factor applied to the tensor-core part of O
        if (tail) {
            mbar_wait(smem_u32(tail_full), 0, p.diag, 19);
            uint4 qv[8];
            const uint8_t* qrow = sQ + row * 128;
#pragma unroll
            for (int i = 0; i < 8; ++i) qv[i] = *reinterpret_cast<const uint4*>(qrow + ((i ^ (row & 7)) << 4));
#pragma unroll
            for (int t = 0; t < MAX_TAIL; ++t) {
                pt[t] = 0.f;
                if (t < tail) {
                    const uint4* kg = reinterpret_cast<const uint4*>(sTail + t * 256);
                    float acc = 0.f;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const uint4 kk = kg[i];
                        const uint32_t qa[4] = {qv[i].x, qv[i].y, qv[i].z, qv[i].w};
                        const uint32_t ka[4] = {kk.x, kk.y, kk.z, kk.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float2 a = unpack2(qa[e], F16 ? 1 : 0), b = unpack2(ka[e], F16 ? 1 : 0);
                            acc = fmaf(a.x, b.x, acc);
                            acc = fmaf(a.y, b.y, acc);
                        }
                    }
                    float x = fmaf(acc, sc, ms);
                    if (x > 16.f) {                           // re-base: this key becomes the row's reference
                        const float f = ex2_approx(-x);
                        l_run *= f; fo *= f; ms -= x;
#pragma unroll
                        for (int u = 0; u < MAX_TAIL; ++u)
                            if (u < t) pt[u] *= f;
                        x = 0.f;
                    }
                    pt[t] = ex2_approx(x);
                    l_run += pt[t];
                }
            }
        }
        // epilogue: O / l -> 16-bit -> global (one 128-byte row per thread)
        mbar_wait(smem_u32(&pv_done[(n_tiles - 1) & 1]), ((n_tiles - 1) >> 1) & 1, p.diag, 16);
        tcgen05_fence_after();
        const float inv = 1.0f / l_run;
        const bool store = q0 + row < T;
        bf16* orow = p.out + ((long long)row_base + q0 + row) * p.D + head * HD;
#pragma unroll 1
        for (int c = 0; c < HD; c += 32) {
            uint32_t v[32];
            tmem_ld32(t_o + (uint32_t)c, v);
            tmem_wait_ld();
            float o[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __uint_as_float(v[i]) * fo;
            if (tail) {
#pragma unroll
                for (int t = 0; t < MAX_TAIL; ++t) {
                    if (t < tail) {
                        const uint4* vg = reinterpret_cast<const uint4*>(sTail + t * 256 + 128 + c * 2);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const uint4 vv = vg[i];
                            const uint32_t a4[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float2 a = unpack2(a4[e], F16 ? 1 : 0);
                                o[8 * i + 2 * e] = fmaf(pt[t], a.x, o[8 * i + 2 * e]);
                                o[8 * i + 2 * e + 1] = fmaf(pt[t], a.y, o[8 * i + 2 * e + 1]);
                            }
                        }
                    }
                }
            }
            if (store) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    uint4 w;
                    w.x = pack2(o[8 * q + 0] * inv, o[8 * q + 1] * inv, F16 ? 1 : 0);
                    w.y = pack2(o[8 * q + 2] * inv, o[8 * q + 3] * inv, F16 ? 1 : 0);
                    w.z = pack2(o[8 * q + 4] * inv, o[8 * q + 5] * inv, F16 ? 1 : 0);
                    w.w = pack2(o[8 * q + 6] * inv, o[8 * q + 7] * inv, F16 ? 1 : 0);
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


// ------------------------------------------------------------------------------------------------ persistent variant
// Same arithmetic, one CTA per SM-half for the whole launch.  ncu on the per-item kernel above: 9 key tiles of ~1.3 K
// cycles per CTA, but ~8 K more cycles per CTA in launch, barrier / TMEM set-up, the first (HBM-latency) Q / K / V loads
// and the epilogue.  Here every CTA walks a static list of (image, head, query tile) items; the TMA warp runs ahead
// into the next item (Q and the tail rows are double buffered, K three and V two stages), the MMA warp issues the next
// item's first S while the softmax warps are still in the previous item's epilogue, and set-up happens once.
constexpr int PK_STAGES = 3, PV_STAGES = 2;
constexpr int ATP_SMEM = 1024 + 2 * TILE_BYTES + (PK_STAGES + PV_STAGES) * KV_BYTES + 2 * P_BYTES + 2 * MAX_TAIL * 256;

template <bool F16>
__global__ void __launch_bounds__(AT_THREADS, AT_CTAS)
attn_tc_persist_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmKV,
                       const AttnParams p, const int n_items, const int q_tiles) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    if ((smem_u32(smem_raw) & 1023u) != 0) {
        if (threadIdx.x == 0 && p.diag) atomicExch(p.diag, 0xA11C0001u);
        __trap();
    }
    uint8_t* smem = smem_raw + 1024;
    uint8_t* sQ = smem;                                   // 2 buffers
    uint8_t* sK = sQ + 2 * TILE_BYTES;                    // PK_STAGES
    uint8_t* sV = sK + PK_STAGES * KV_BYTES;              // PV_STAGES
    uint8_t* sP = sV + PV_STAGES * KV_BYTES;              // 2 buffers
    uint8_t* sTail = sP + 2 * P_BYTES;                    // 2 x [MAX_TAIL][K row 128 B | V row 128 B]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
    uint64_t* q_full = bars;            // [2] TMA -> MMA
    uint64_t* q_empty = bars + 2;       // [2] softmax threads (epilogue) -> TMA, 128 arrivals
    uint64_t* k_full = bars + 4;        // [3]
    uint64_t* k_empty = bars + 7;       // [3] S(j) retired
    uint64_t* v_full = bars + 10;       // [2]
    uint64_t* v_empty = bars + 12;      // [2] P V (j) retired
    uint64_t* s_full = bars + 14;       // [2]
    uint64_t* p_full = bars + 16;       // [2] 128 arrivals
    uint64_t* pv_done = bars + 18;      // [2]
    uint64_t* o_free = bars + 20;       //     epilogue has read O (128 arrivals)
    uint64_t* tail_full = bars + 21;    // [2] 32 cp.async arrivals (warp 0)
    uint64_t* tail_empty = bars + 23;   // [2] 128 arrivals
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 25);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int T = p.T;
    const int tail = (T >= KT && (T % KT) <= MAX_TAIL) ? (T % KT) : 0;
    const int n_tiles = tail ? T / KT : (T + KT - 1) / KT;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQKV) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmKV) : "memory");
        for (int i = 0; i < 2; ++i) {
            mbar_init(smem_u32(&q_full[i]), 1); mbar_init(smem_u32(&q_empty[i]), 128);
            mbar_init(smem_u32(&v_full[i]), 1); mbar_init(smem_u32(&v_empty[i]), 1);
            mbar_init(smem_u32(&s_full[i]), 1); mbar_init(smem_u32(&p_full[i]), 128); mbar_init(smem_u32(&pv_done[i]), 1);
            mbar_init(smem_u32(&tail_full[i]), 32); mbar_init(smem_u32(&tail_empty[i]), 128);
        }
        for (int i = 0; i < PK_STAGES; ++i) { mbar_init(smem_u32(&k_full[i]), 1); mbar_init(smem_u32(&k_empty[i]), 1); }
        mbar_init(smem_u32(o_free), 128);
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

    // item -> (query tile, head, image): query tiles of one (image, head) are neighbours in the item order, so the CTAs
    // that run at the same time share its K / V in L2
    auto decode = [&](int item, int& q0, int& head, int& row_base) {
        const int qt = item % q_tiles;
        const int r = item / q_tiles;
        head = r % p.heads;
        row_base = (r / p.heads) * T;
        q0 = qt * TILE;
    };

    if (warp == 0) {
        int ks = 0, kuse = 0, vs = 0, vuse = 0;           // ring positions / wrap counts, running across items
        int it = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
            int q0, head, row_base;
            decode(item, q0, head, row_base);
            const int qb = it & 1;
            const uint32_t qpar = (uint32_t)(((it >> 1) & 1) ^ 1);
            // the whole warp walks the schedule (waits included); one elected lane issues — coordinates / addresses then live
            // in uniform registers instead of going through an R2UR waterfall per TMA / MMA instruction
            mbar_wait(smem_u32(&tail_empty[qb]), qpar, p.diag, 21);
            mbar_wait(smem_u32(&q_empty[qb]), qpar, p.diag, 22);
            if (elect_one_sync()) {
                mbar_arrive_expect_tx(smem_u32(&q_full[qb]), TILE_BYTES);
                tma_load_2d(smem_u32(sQ + qb * TILE_BYTES), &tmQKV, smem_u32(&q_full[qb]), head * HD, row_base + q0);
            }
            __syncwarp();
            if (lane < 16 * tail) {
                const int t = lane >> 4, part = lane & 15;    // part 0..7: K row, 8..15: V row
                const bf16* g = p.qkv + ((long long)row_base + n_tiles * KT + t) * (3LL * p.D) + (part < 8 ? p.D : 2 * p.D) +
                                head * HD + (part & 7) * 8;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(sTail + qb * MAX_TAIL * 256 + t * 256 + part * 16)),
                             "l"(g) : "memory");
            }
            asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&tail_full[qb])) : "memory");
            {
                auto load_k = [&](int j) {
                    mbar_wait(smem_u32(&k_empty[ks]), (kuse & 1) ^ 1, p.diag, 23);
                    if (elect_one_sync()) {
                        mbar_arrive_expect_tx(smem_u32(&k_full[ks]), KV_BYTES);
                        tma_load_2d(smem_u32(sK + ks * KV_BYTES), &tmKV, smem_u32(&k_full[ks]), p.D + head * HD, row_base + j * KT);
                    }
                    __syncwarp();
                    if (++ks == PK_STAGES) { ks = 0; ++kuse; }
                };
                load_k(0);
                for (int j = 0; j < n_tiles; ++j) {
                    if (j + 1 < n_tiles) load_k(j + 1);       // K runs one tile ahead of V: S(j+1) is issued before P V (j)
                    mbar_wait(smem_u32(&v_empty[vs]), (vuse & 1) ^ 1, p.diag, 24);
                    if (elect_one_sync()) {
                        mbar_arrive_expect_tx(smem_u32(&v_full[vs]), KV_BYTES);
                        tma_load_2d(smem_u32(sV + vs * KV_BYTES), &tmKV, smem_u32(&v_full[vs]), 2 * p.D + head * HD, row_base + j * KT);
                    }
                    __syncwarp();
                    if (++vs == PV_STAGES) { vs = 0; ++vuse; }
                }
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        {
            const uint32_t ab_fmt = F16 ? 0u : ((1u << 7) | (1u << 10));
            const uint32_t idesc_qk = (1u << 4) | ab_fmt | ((uint32_t)(KT >> 3) << 17) | ((uint32_t)(TILE >> 4) << 24);
            const uint32_t idesc_pv = (1u << 4) | ab_fmt | (1u << 16) | ((uint32_t)(HD >> 3) << 17) | ((uint32_t)(TILE >> 4) << 24);
            int ks = 0, kuse = 0, vs = 0, vuse = 0;
            int jt = 0;                                   // running tile index: S / P buffer = jt & 1
            int it = 0;
            uint64_t qd = 0;
            auto issue_s = [&](int jj) {                  // S(jj) = Q K^T into S buffer jj & 1
                mbar_wait(smem_u32(&k_full[ks]), kuse & 1, p.diag, 25);
                tcgen05_fence_after();
                const uint64_t kd = make_smem_desc(smem_u32(sK + ks * KV_BYTES));
                const uint32_t d = tmem_base + (uint32_t)((jj & 1) * KT);
                if (elect_one_sync()) {
#pragma unroll
                    for (int k = 0; k < HD / 16; ++k)
                        umma_bf16(d, qd + (uint64_t)(2 * k), kd + (uint64_t)(2 * k), idesc_qk, k > 0 ? 1u : 0u);
                    tcgen05_commit(smem_u32(&k_empty[ks]));
                    tcgen05_commit(smem_u32(&s_full[jj & 1]));
                }
                __syncwarp();
                if (++ks == PK_STAGES) { ks = 0; ++kuse; }
            };
            for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
                const int qb = it & 1;
                mbar_wait(smem_u32(&q_full[qb]), (it >> 1) & 1, p.diag, 26);
                tcgen05_fence_after();
                qd = make_smem_desc(smem_u32(sQ + qb * TILE_BYTES));
                issue_s(jt);
                for (int j = 0; j < n_tiles; ++j, ++jt) {
                    if (j + 1 < n_tiles) issue_s(jt + 1);
                    mbar_wait(smem_u32(&p_full[jt & 1]), (jt >> 1) & 1, p.diag, 27);
                    if (j == 0 && it > 0) mbar_wait(smem_u32(o_free), (it - 1) & 1, p.diag, 28);   // previous epilogue read O
                    mbar_wait(smem_u32(&v_full[vs]), vuse & 1, p.diag, 29);
                    tcgen05_fence_after();
                    const uint64_t vd = make_smem_desc(smem_u32(sV + vs * KV_BYTES));
                    const uint64_t pd0 = make_smem_desc(smem_u32(sP + (jt & 1) * P_BYTES));
                    if (elect_one_sync()) {
#pragma unroll
                        for (int k = 0; k < KT / 16; ++k)
                            umma_bf16(tmem_base + O_COL, pd0 + (uint64_t)(2 * k), vd + (uint64_t)((k * 2048) >> 4), idesc_pv,
                                      (j > 0 || k > 0) ? 1u : 0u);
                        tcgen05_commit(smem_u32(&v_empty[vs]));
                        tcgen05_commit(smem_u32(&pv_done[jt & 1]));
                    }
                    __syncwarp();
                    if (++vs == PV_STAGES) { vs = 0; ++vuse; }
                }
            }
        }
    } else {
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
        const uint32_t t_o = t_lane + O_COL;
        const float sc = p.scale_log2e;
        uint32_t va[32], vb[32];
        auto pass_max = [&](uint32_t t_s, int valid) -> float {
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
            return mx;
        };
        int jt = 0, it = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
            int q0, head, row_base;
            decode(item, q0, head, row_base);
            const int qb = it & 1;
            float l_run = 0.f, ms = 0.f;
            for (int j = 0; j < n_tiles; ++j, ++jt) {
                const int b = jt & 1;
                const uint32_t t_s = t_lane + (uint32_t)(b * KT);
                uint8_t* pbuf = sP + b * P_BYTES;
                mbar_wait(smem_u32(&s_full[b]), (jt >> 1) & 1, p.diag, 15);
                if (jt >= 2) mbar_wait(smem_u32(&pv_done[b]), ((jt >> 1) - 1) & 1, p.diag, 17);   // P buffer b: read by P V (jt-2)
                tcgen05_fence_after();
                const int valid = T - j * KT;
                const bool ragged = valid < KT;
                if (j == 0) ms = -pass_max(t_s, valid) * sc;
                float lt = ragged ? attn_pass_p<F16, true>(t_s, pbuf, row, valid, ms, sc, va, vb)
                                  : (p.poly ? attn_pass_p<F16, false, true>(t_s, pbuf, row, valid, ms, sc, va, vb)
                                            : attn_pass_p<F16, false>(t_s, pbuf, row, valid, ms, sc, va, vb));
                if (__any_sync(0xffffffffu, !(lt <= 2048.f))) {
                    const float ms_new = -pass_max(t_s, valid) * sc;
                    float f = 1.f;
                    if (!(lt <= 2048.f)) {
                        f = ex2_approx(ms_new - ms);
                        ms = ms_new;
                    }
                    l_run *= f;
                    if (j > 0) {
                        mbar_wait(smem_u32(&pv_done[(jt - 1) & 1]), ((jt - 1) >> 1) & 1, p.diag, 18);
                        tcgen05_fence_after();
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
                    lt = attn_pass_p<F16, true>(t_s, pbuf, row, valid, ms, sc, va, vb);
                }
                l_run += lt;
                fence_proxy_async_smem();
                tcgen05_fence_before();
                mbar_arrive(smem_u32(&p_full[b]));
            }
            // ---- item epilogue: tail keys on the CUDA cores, then O / l -> global
            float pt[MAX_TAIL];
            float fo = 1.f;
            if (tail) {
                mbar_wait(smem_u32(&tail_full[qb]), (it >> 1) & 1, p.diag, 19);
                uint4 qv[8];
                const uint8_t* qrow = sQ + qb * TILE_BYTES + row * 128;
#pragma unroll
                for (int i = 0; i < 8; ++i) qv[i] = *reinterpret_cast<const uint4*>(qrow + ((i ^ (row & 7)) << 4));
#pragma unroll
                for (int t = 0; t < MAX_TAIL; ++t) {
                    pt[t] = 0.f;
                    if (t < tail) {
                        const uint4* kg = reinterpret_cast<const uint4*>(sTail + qb * MAX_TAIL * 256 + t * 256);
                        float acc = 0.f;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const uint4 kk = kg[i];
                            const uint32_t qa[4] = {qv[i].x, qv[i].y, qv[i].z, qv[i].w};
                            const uint32_t ka[4] = {kk.x, kk.y, kk.z, kk.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float2 a = unpack2(qa[e], F16 ? 1 : 0), bq = unpack2(ka[e], F16 ? 1 : 0);
                                acc = fmaf(a.x, bq.x, acc);
                                acc = fmaf(a.y, bq.y, acc);
                            }
                        }
                        float x = fmaf(acc, sc, ms);
                        if (x > 16.f) {
                            const float f = ex2_approx(-x);
                            l_run *= f; fo *= f; ms -= x;
#pragma unroll
                            for (int u = 0; u < MAX_TAIL; ++u)
                                if (u < t) pt[u] *= f;
                            x = 0.f;
                        }
                        pt[t] = ex2_approx(x);
                        l_run += pt[t];
                    }
                }
            }
            mbar_arrive(smem_u32(&q_empty[qb]));          // this thread is done with the item's Q tile
            mbar_wait(smem_u32(&pv_done[(jt - 1) & 1]), ((jt - 1) >> 1) & 1, p.diag, 16);
            tcgen05_fence_after();
            const float inv = 1.0f / l_run;
            const bool store = q0 + row < T;
            bf16* orow = p.out + ((long long)row_base + q0 + row) * p.D + head * HD;
            uint32_t v0[32], v1[32];
            tmem_ld32(t_o, v0);
            tmem_ld32(t_o + 32u, v1);
            tmem_wait_ld();
            tcgen05_fence_before();
            mbar_arrive(smem_u32(o_free));                // O is in registers: the next item's first P V may overwrite it
#pragma unroll
            for (int hb = 0; hb < 2; ++hb) {
                const uint32_t* v = hb ? v1 : v0;
                const int c = hb * 32;
                float o[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) o[i] = __uint_as_float(v[i]) * fo;
                if (tail) {
#pragma unroll
                    for (int t = 0; t < MAX_TAIL; ++t) {
                        if (t < tail) {
                            const uint4* vg = reinterpret_cast<const uint4*>(sTail + qb * MAX_TAIL * 256 + t * 256 + 128 + c * 2);
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const uint4 vv = vg[i];
                                const uint32_t a4[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const float2 a = unpack2(a4[e], F16 ? 1 : 0);
                                    o[8 * i + 2 * e] = fmaf(pt[t], a.x, o[8 * i + 2 * e]);
                                    o[8 * i + 2 * e + 1] = fmaf(pt[t], a.y, o[8 * i + 2 * e + 1]);
                                }
                            }
                        }
                    }
                }
                if (store) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint4 w;
                        w.x = pack2(o[8 * q + 0] * inv, o[8 * q + 1] * inv, F16 ? 1 : 0);
                        w.y = pack2(o[8 * q + 2] * inv, o[8 * q + 3] * inv, F16 ? 1 : 0);
                        w.z = pack2(o[8 * q + 4] * inv, o[8 * q + 5] * inv, F16 ? 1 : 0);
                        w.w = pack2(o[8 * q + 6] * inv, o[8 * q + 7] * inv, F16 ? 1 : 0);
                        *reinterpret_cast<uint4*>(orow + c + 8 * q) = w;
                    }
                }
            }
            mbar_arrive(smem_u32(&tail_empty[qb]));       // tail rows consumed
        }
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
    static int persist = 1;
    if (!attr_set) {
        MB_CUDA(ctx, cudaFuncSetAttribute(attn_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM));
        MB_CUDA(ctx, cudaFuncSetAttribute(attn_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM));
        MB_CUDA(ctx, cudaFuncSetAttribute(attn_tc_persist_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATP_SMEM));
        MB_CUDA(ctx, cudaFuncSetAttribute(attn_tc_persist_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATP_SMEM));
        const char* e = getenv("MB_ATTN_PERSIST");
        persist = !(e && e[0] == '0');
        attr_set = true;
    }
    AttnParams p;
    p.T = T; p.D = D; p.heads = heads; p.scale_log2e = scale_log2e; p.qkv = qkv; p.out = out; p.f16 = ctx->f16; p.diag = ctx->dev_diag;
    { const char* e = getenv("MB_ATTN_POLY"); p.poly = e ? (e[0] != '0') : 0; }
    const int q_tiles = (T + TILE - 1) / TILE;
    const long long items = (long long)q_tiles * heads * n;
    if (persist && items < 0x7fffffffLL) {
        const int grid = (int)(items < (long long)AT_CTAS * ctx->num_sms ? items : (long long)AT_CTAS * ctx->num_sms);
        if (ctx->f16) attn_tc_persist_kernel<true><<<grid, AT_THREADS, ATP_SMEM, stream>>>(tm, tmkv, p, (int)items, q_tiles);
        else attn_tc_persist_kernel<false><<<grid, AT_THREADS, ATP_SMEM, stream>>>(tm, tmkv, p, (int)items, q_tiles);
        MB_LAUNCH_CHECK(ctx);
        return 0;
    }
    dim3 grid(q_tiles, heads, n);
    if (ctx->f16) attn_tc_kernel<true><<<grid, AT_THREADS, AT_SMEM, stream>>>(tm, tmkv, p);
    else attn_tc_kernel<false><<<grid, AT_THREADS, AT_SMEM, stream>>>(tm, tmkv, p);
    MB_LAUNCH_CHECK(ctx);
    return 0;
}
