// Greedy-decode cross-attention over the encoder states on the 5th-generation tensor cores (tcgen05 + TMEM + TMA).
// Same arithmetic as dec_cross_enc_kernel (trocr.cu), which it replaces: per crop and decoder layer, with the per-head
// projected queries q'^h = Wk^h^T q^h in R^E (16 heads) and the crop's encoder states e_t (T x E),
//     s_t^h = q'^h . e_t,   p^h = softmax_t(s^h),   ctx^h = sum_t p_t^h e_t        (marie/models/unilm/trocr: fairseq
// MultiheadAttention inside TransformerDecoderLayer, trocr_models.py:142-147; the K / V projections are hoisted out, see
// the header of dec_cross_enc_kernel).  The kernel is a single pass over the crop's 0.89 MB of encoder states and should
// be bound by HBM; the mma.sync version was bound by the instruction stream of its eight resident warps (ncu: 4.39 TB/s,
// 12 % occupancy at 245 registers, 120 of 625 instructions per warp and tile were cp.async address arithmetic).
//
// Formulation for tcgen05 (M = 128 is the narrow side of the tensor core, so the 16 heads go into N):
//     S^T  [keys x heads] = E_tile [keys x E] . q'^T [E x heads]          A = encoder tile as loaded (K-major), B = q'
//     ctx^T[E    x heads] += E_tile^T [E x keys] . P^T [keys x heads]     A = THE SAME tile read MN-major, B = P^T
//     l    [ *   x heads] += ones [128 x keys]  . P^T                     row sums of the rounded P, by the tensor core
// One persistent CTA per SM walks a list of live crops (rows whose hypothesis has ended are compacted away by
// live_list_kernel).  Warp 0: TMA — the crop's q' (16 x E) and its encoder rows in tiles of KT keys (E / 64 boxes of
// [KT x 64], 128-byte swizzle; keys beyond T are zero-filled by the tensor map, not fetched), two tile buffers.  Warp 1:
// tcgen05.mma, S^T double buffered in TMEM so S^T(j+1) is issued before P(j) is awaited.  Warps 2-5: one thread per key
// (TMEM lane) — p = exp2(s - m^h) against a per-head reference m^h that only moves when a score exceeds it by 2^8 (the
// accumulators are then rescaled in TMEM, all four warps; P stays far inside fp16 range) -> P^T in shared memory; after
// the last tile ctx^T / l -> global.  The S^T MMA runs with M = 128 although a tile has KT <= 64 keys: rows beyond the
// tile read whatever follows in shared memory and land in TMEM lanes nobody reads.
// Beams: the hypotheses of a crop attend over the same encoder states, so up to NB = 3 of them (2 for E = 1024; TMEM holds
// (3 + E / 128) * 16 * NB columns) go through one pass as N = 16 * NB columns of every MMA — q' rows [crop][beam][head] are
// one TMA box, TMEM column c belongs to q' row row0 + c; a wider beam is ceil(beam / NB) work items per crop (the second
// pass over the crop's states mostly hits L2).  No beam width keeps a cross-attention K/V cache.
#include "common.cuh"
#include "tc_ptx.cuh"
#include <stdlib.h>
#include <mutex>

int mb_encode_2d_map(mb_ctx* ctx, CUtensorMap* m, const void* base, long long cols, long long rows, long long ld,
                     int box_cols, int box_rows);
int mb_encode_3d_map(mb_ctx* ctx, CUtensorMap* m, const void* base, long long d0, long long d1, long long d2, long long ld1,
                     long long ld2, int b0, int b1);

namespace {

__device__ __forceinline__ float ex2a(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

constexpr int XT_THREADS = 192;

// NB = hypotheses (beams) of one crop that share a pass over its encoder states: N = 16 * NB columns in every MMA
template <int E, int KT, int NB> struct XtCfg {
    static constexpr int N = 16 * NB;
    static constexpr int NCH = E / 64;                   // 64-column boxes per encoder row
    static constexpr int CHUNK = KT * 128;               // bytes of one [KT x 64] box
    static constexpr int TILE = NCH * CHUNK;             // one tile of KT encoder rows
    static constexpr int MT = E / 128;                   // M tiles of ctx^T
    static constexpr int QCH = N * 128;                  // one [N x 64] box of q'
    static constexpr int Q_BYTES = NCH * QCH;
    static constexpr int P_BYTES = N * 128;              // one P^T buffer, [N x 64 keys] K-major
    static constexpr int OFF_Q = 2 * TILE;
    static constexpr int OFF_P = OFF_Q + Q_BYTES;
    static constexpr int OFF_ONES = OFF_P + 2 * P_BYTES; // [128 x 16] MN-major tile of ones
    static constexpr int OFF_BAR = OFF_ONES + 4096;      // control page: mbarriers, TMEM pointer, flags, tile maxima, refs
    static constexpr int SMEM = OFF_BAR + 2048;
    static constexpr int S_COL = 0;                      // S^T(0), S^T(1): N columns each
    static constexpr int C_COL = 2 * N;                  // ctx^T: MT x N columns
    static constexpr int L_COL = C_COL + MT * N;         // row sums: N columns
    static constexpr int TMEM_COLS = L_COL + N <= 256 ? 256 : 512;
    static_assert(E % 128 == 0 && KT % 16 == 0 && KT <= 64 && NB >= 1 && NB <= 3, "geometry");
    static_assert(L_COL + N <= 512, "TMEM columns");
    static_assert(SMEM <= 227 * 1024, "shared memory");
    static_assert(104 + 32 + 24 + 4 * N * 4 <= 2048, "control page");
    // the M = 128 S^T MMA reads 128 rows from the start of a box: beyond the last box of the second buffer that is the q'
    // / P / ones region, which must cover the overrun
    static_assert((128 - KT) * 128 <= Q_BYTES + 2 * P_BYTES + 4096, "operand overrun stays inside the allocation");
};

struct XtParams {
    int T, heads, n, beam, groups;   // n crops of `beam` hypotheses each; groups = ceil(beam / NB) passes per crop
    const int* live;             // compacted crop indices (or null: all crops)
    const int* n_live;           // device counter (or null: n)
    bf16* out;                   // [n * beam, heads * E]
    unsigned int* diag;
};

// bars (u64 slots): 0,1 tile_full | 2,3 tile_empty | 4 q_full | 5 q_empty | 6,7 s_full | 8,9 p_full | 10,11 ctx_done
template <bool F16, int E, int KT, int NB>
__global__ void __launch_bounds__(XT_THREADS, 1)
dec_cross_tc_kernel(const __grid_constant__ CUtensorMap tmE, const __grid_constant__ CUtensorMap tmQ, const XtParams p) {
    using C = XtCfg<E, KT, NB>;
    constexpr int N = C::N;
    extern __shared__ __align__(1024) uint8_t xt_smem[];
    if ((smem_u32(xt_smem) & 1023u) != 0) {
        if (threadIdx.x == 0 && p.diag) atomicExch(p.diag, 0xDEAD00B1u);
        __trap();
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* sTile = xt_smem;
    uint8_t* sQ = xt_smem + C::OFF_Q;
    uint8_t* sP = xt_smem + C::OFF_P;
    uint8_t* sOnes = xt_smem + C::OFF_ONES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(xt_smem + C::OFF_BAR);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 12);
    volatile uint32_t* flags = reinterpret_cast<volatile uint32_t*>(bars + 13);      // [2 buffers][4 warps]
    // exchanged across the named barrier (its asm carries a memory clobber, so plain loads are re-issued after it)
    float* tmax = reinterpret_cast<float*>(bars + 20);                               // [2 key warps][N]
    float* mref = tmax + 2 * N;                                                      // [N] softmax references (log2 units)
    float* fsc = mref + N;                                                           // [N] rescale factors of the last move
    const uint32_t tile_full = smem_u32(&bars[0]), tile_empty = smem_u32(&bars[2]);
    const uint32_t q_full = smem_u32(&bars[4]), q_empty = smem_u32(&bars[5]);
    const uint32_t s_full = smem_u32(&bars[6]), p_full = smem_u32(&bars[8]), ctx_done = smem_u32(&bars[10]);

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmE) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQ) : "memory");
        for (int i = 0; i < 12; ++i) mbar_init(smem_u32(&bars[i]), (i == 8 || i == 9) ? 4u : 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)),
                     "r"(C::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (warp >= 2) {
        // the constant operand of the row-sum MMA, and the "nothing exceeded the reference" flags of the warps without keys
        const uint32_t one2 = F16 ? 0x3C003C00u : 0x3F803F80u;
        for (int i = threadIdx.x - 64; i < 4096 / 4; i += 128) reinterpret_cast<uint32_t*>(sOnes)[i] = one2;
        if (threadIdx.x - 64 < 8) flags[threadIdx.x - 64] = 0u;
        fence_proxy_async_smem();
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_ptr);

    const int n_items = (p.n_live ? *p.n_live : p.n) * p.groups;
    const int nt = (p.T + KT - 1) / KT;
    // item -> (crop, first q' row of the group): q' / out rows are [crop][beam][head]
    auto item_of = [&](int i, int& crop, int& row0, int& row_end) {
        const int ci = i / p.groups, grp = i - ci * p.groups;
        crop = p.live ? p.live[ci] : ci;
        const int b0 = grp * NB;
        row0 = (crop * p.beam + b0) * p.heads;
        row_end = (crop * p.beam + min(b0 + NB, p.beam)) * p.heads;
    };

    if (warp == 0) {
        // ---------------------------------------------------------------------------------------- TMA producer
        uint32_t g = 0, ci = 0;
        for (int i = blockIdx.x; i < n_items; i += gridDim.x, ++ci) {
            int crop, row0, row_end;
            item_of(i, crop, row0, row_end);
            mbar_wait(q_empty, (ci & 1u) ^ 1u, p.diag, 21);           // every S^T MMA of the previous item has read q'
            if (elect_one_sync()) {
                mbar_arrive_expect_tx(q_full, C::Q_BYTES);
#pragma unroll 1
                for (int c = 0; c < C::NCH; ++c)
                    tma_load_2d(smem_u32(sQ + c * C::QCH), &tmQ, q_full, c * 64, row0);
            }
            __syncwarp();
            for (int j = 0; j < nt; ++j, ++g) {
                const uint32_t b = g & 1u;
                mbar_wait(tile_empty + 8u * b, ((g >> 1) & 1u) ^ 1u, p.diag, 22);
                if (elect_one_sync()) {
                    mbar_arrive_expect_tx(tile_full + 8u * b, C::TILE);
#pragma unroll 1
                    for (int c = 0; c < C::NCH; ++c)
                        tma_load_3d(smem_u32(sTile + b * C::TILE + c * C::CHUNK), &tmE, tile_full + 8u * b, c * 64, j * KT, crop);
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        // ---------------------------------------------------------------------------------------- MMA issuer
        const uint32_t ab_fmt = F16 ? 0u : ((1u << 7) | (1u << 10));
        const uint32_t idesc_s = (1u << 4) | ab_fmt | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t idesc_c = idesc_s | (1u << 15);                 // A (encoder tile / ones) MN-major
        const uint32_t tile_addr = smem_u32(sTile), q_addr = smem_u32(sQ), p_addr = smem_u32(sP);
        const uint64_t ones_desc = make_smem_desc_mn(smem_u32(sOnes), 2048);
        auto issue_s = [&](uint32_t gt, bool last) {
            const uint32_t b = gt & 1u;
            mbar_wait(tile_full + 8u * b, (gt >> 1) & 1u, p.diag, 23);
            tcgen05_fence_after();
            if (elect_one_sync()) {
                const uint32_t d = tmem_base + (uint32_t)C::S_COL + b * (uint32_t)N;
#pragma unroll 1
                for (int c = 0; c < C::NCH; ++c) {
                    const uint64_t ad = make_smem_desc(tile_addr + b * C::TILE + c * C::CHUNK);
                    const uint64_t bd = make_smem_desc(q_addr + c * C::QCH);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16(d, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc_s, (c > 0 || k > 0) ? 1u : 0u);
                }
                tcgen05_commit(s_full + 8u * b);
                if (last) tcgen05_commit(q_empty);
            }
            __syncwarp();
        };
        uint32_t g = 0, ci = 0;
        for (int i = blockIdx.x; i < n_items; i += gridDim.x, ++ci) {
            mbar_wait(q_full, ci & 1u, p.diag, 24);
            tcgen05_fence_after();
            issue_s(g, nt == 1);
            for (int j = 0; j < nt; ++j) {
                const uint32_t gj = g + (uint32_t)j, b = gj & 1u;
                // S^T(j+1) goes to the tensor pipe as soon as its tile has landed — under the softmax of tile j when it is
                // already there, but never in front of ctx(j): with two tile buffers the buffer of tile j+2 is only released
                // by ctx(j), so waiting for tile j+1 first would serialise every load behind the previous tile's softmax
                bool s_next = j + 1 >= nt;
                unsigned int spins = 0;
                while (true) {
                    if (++spins > (1u << 26)) {                      // a lost arrival must fault, never hang the box
                        if (p.diag) atomicExch(p.diag, 0xDEAD0000u | 25u);
                        __trap();
                    }
                    if (!s_next && mbar_test(tile_full + 8u * ((gj + 1u) & 1u), ((gj + 1u) >> 1) & 1u)) {
                        issue_s(gj + 1, j + 2 == nt);
                        s_next = true;
                    }
                    if (mbar_test(p_full + 8u * b, (gj >> 1) & 1u)) break;
                }
                tcgen05_fence_after();
                if (elect_one_sync()) {
                    const uint64_t pd = make_smem_desc(p_addr + b * (uint32_t)C::P_BYTES);
#pragma unroll 1
                    for (int mt = 0; mt < C::MT; ++mt) {
#pragma unroll
                        for (int k = 0; k < KT / 16; ++k) {
                            // A: e columns mt*128 .. +127 = two 64-column boxes CHUNK bytes apart (LBO); 16 keys = two 8-row
                            // groups of 1024 B
                            const uint64_t ad = make_smem_desc_mn(tile_addr + b * C::TILE + 2 * mt * C::CHUNK + k * 2048, C::CHUNK);
                            umma_bf16(tmem_base + (uint32_t)(C::C_COL + mt * N), ad, pd + (uint64_t)(2 * k), idesc_c,
                                      (j > 0 || k > 0) ? 1u : 0u);
                        }
                    }
#pragma unroll
                    for (int k = 0; k < KT / 16; ++k)
                        umma_bf16(tmem_base + (uint32_t)C::L_COL, ones_desc, pd + (uint64_t)(2 * k), idesc_c, (j > 0 || k > 0) ? 1u : 0u);
                    tcgen05_commit(tile_empty + 8u * b);
                    tcgen05_commit(ctx_done + 8u * b);
                }
                __syncwarp();
                if (!s_next) issue_s(gj + 1, j + 2 == nt);
            }
            g += (uint32_t)nt;
        }
    } else {
        // ---------------------------------------------------------------------------------------- softmax / epilogue
        const int q = warp & 3;                                    // TMEM lane quarter this warp may access
        const int tid = threadIdx.x - 64;                          // 0..127 among the four warps
        const bool keywarp = q * 32 < KT;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        const float L2E = 1.4426950408889634f;
        uint32_t g = 0;
        for (int i = blockIdx.x; i < n_items; i += gridDim.x) {
            int crop, row0, row_end;
            item_of(i, crop, row0, row_end);
            if (tid < N) mref[tid] = -INFINITY;
            named_bar_sync(1, 128);
            for (int j = 0; j < nt; ++j) {
                const uint32_t gj = g + (uint32_t)j, b = gj & 1u, par = (gj >> 1) & 1u;
                const uint32_t ts = tmem_base + lane_base + (uint32_t)C::S_COL + b * (uint32_t)N;
                const bool valid = j * KT + q * 32 + lane < p.T;
                uint32_t vk[16];                                  // NB == 1: the tile's scores stay in registers between the passes
                if (keywarp) {
                    mbar_wait(s_full + 8u * b, par, p.diag, 26);
                    tcgen05_fence_after();
                    bool over = false;
#pragma unroll
                    for (int c16 = 0; c16 < NB; ++c16) {
                        uint32_t v[16];
                        tmem_ld16(ts + (uint32_t)(c16 * 16), v);
                        tmem_wait_ld();
                        float rf[16];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float4 r4 = *reinterpret_cast<const float4*>(mref + c16 * 16 + 4 * i);
                            rf[4 * i] = r4.x; rf[4 * i + 1] = r4.y; rf[4 * i + 2] = r4.z; rf[4 * i + 3] = r4.w;
                        }
#pragma unroll
                        for (int h = 0; h < 16; ++h) {
                            over |= valid && (__uint_as_float(v[h]) * L2E > rf[h] + 8.0f);
                            if (NB == 1) vk[h] = v[h];
                        }
                    }
                    const bool any = __any_sync(0xffffffffu, over);
                    if (lane == 0) flags[b * 4 + q] = any ? 1u : 0u;
                }
                named_bar_sync(1, 128);
                const uint32_t ov = flags[b * 4] | (KT > 32 ? flags[b * 4 + 1] : 0u);
                if (ov) {
                    // some score is more than 2^8 above its column's reference (always in an item's first tile): move the
                    // references to the exact tile maxima and rescale what has been accumulated so far
                    if (keywarp) {
#pragma unroll
                        for (int c16 = 0; c16 < NB; ++c16) {
                            uint32_t v[16];
                            if (NB > 1) {
                                tmem_ld16(ts + (uint32_t)(c16 * 16), v);
                                tmem_wait_ld();
                            }
#pragma unroll
                            for (int h = 0; h < 16; ++h) {
                                if (NB == 1) v[h] = vk[h];
                                const float t = warp_max(valid ? __uint_as_float(v[h]) * L2E : -INFINITY);
                                if (lane == 0) tmax[q * N + c16 * 16 + h] = t;
                            }
                        }
                    }
                    named_bar_sync(1, 128);
                    if (tid < N) {
                        float tm = tmax[tid];
                        if (KT > 32) tm = fmaxf(tm, tmax[N + tid]);
                        const float mo = mref[tid], mn = fmaxf(mo, tm);
                        fsc[tid] = ex2a(mo - mn);                    // 0 for the first tile (ref = -inf), 1 when unchanged
                        mref[tid] = mn;
                    }
                    named_bar_sync(1, 128);
                    if (j > 0) {
                        mbar_wait(ctx_done + 8u * ((gj - 1u) & 1u), ((gj - 1u) >> 1) & 1u, p.diag, 27);
                        tcgen05_fence_after();
#pragma unroll 1
                        for (int mt = 0; mt <= C::MT; ++mt) {        // MT context tiles + the row-sum tile
#pragma unroll
                            for (int c16 = 0; c16 < NB; ++c16) {
                                const uint32_t ta = tmem_base + lane_base + (uint32_t)(C::C_COL + mt * N + c16 * 16);
                                uint32_t a[16];
                                tmem_ld16(ta, a);
                                tmem_wait_ld();
#pragma unroll
                                for (int h = 0; h < 16; ++h) a[h] = __float_as_uint(__uint_as_float(a[h]) * fsc[c16 * 16 + h]);
                                tmem_st16(ta, a);
                            }
                        }
                        tmem_wait_st();
                    }
                }
                if (keywarp) {
                    mbar_wait(ctx_done + 8u * b, par ^ 1u, p.diag, 28);   // the MMAs that read P buffer b two tiles ago have retired
                    const int t = q * 32 + lane;
                    uint8_t* pbuf = sP + b * C::P_BYTES + (t & 7) * 2;
#pragma unroll
                    for (int c16 = 0; c16 < NB; ++c16) {
                        uint32_t v[16];
                        if (NB > 1) {
                            tmem_ld16(ts + (uint32_t)(c16 * 16), v);
                            tmem_wait_ld();
                        }
                        float rf[16];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float4 r4 = *reinterpret_cast<const float4*>(mref + c16 * 16 + 4 * i);
                            rf[4 * i] = r4.x; rf[4 * i + 1] = r4.y; rf[4 * i + 2] = r4.z; rf[4 * i + 3] = r4.w;
                        }
#pragma unroll
                        for (int h = 0; h < 16; ++h) {
                            if (NB == 1) v[h] = vk[h];
                            const int col = c16 * 16 + h;
                            const float pv = valid ? ex2a(__uint_as_float(v[h]) * L2E - rf[h]) : 0.f;
                            const uint32_t pk = pack2(pv, 0.f, F16 ? 1 : 0);
                            *reinterpret_cast<unsigned short*>(pbuf + col * 128 + ((((uint32_t)t >> 3) ^ (uint32_t)(col & 7)) << 4)) =
                                (unsigned short)(pk & 0xFFFFu);
                        }
                    }
                    fence_proxy_async_smem();
                }
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(p_full + 8u * b);
            }
            // epilogue: ctx^T / l for this warp's 32 e columns of every M tile; TMEM column c = q' row row0 + c
            const uint32_t gl = g + (uint32_t)nt - 1u;
            mbar_wait(ctx_done + 8u * (gl & 1u), (gl >> 1) & 1u, p.diag, 29);
            tcgen05_fence_after();
#pragma unroll 1
            for (int c16 = 0; c16 < NB; ++c16) {
                float inv[16];
                {
                    uint32_t lv[16];
                    tmem_ld16(tmem_base + lane_base + (uint32_t)(C::L_COL + c16 * 16), lv);
                    tmem_wait_ld();
#pragma unroll
                    for (int h = 0; h < 16; ++h) inv[h] = 1.0f / __uint_as_float(lv[h]);
                }
                bf16* obase = p.out + (long long)(row0 + c16 * 16) * E + q * 32 + lane;
#pragma unroll 1
                for (int mt = 0; mt < C::MT; ++mt) {
                    uint32_t a[16];
                    tmem_ld16(tmem_base + lane_base + (uint32_t)(C::C_COL + mt * N + c16 * 16), a);
                    tmem_wait_ld();
#pragma unroll
                    for (int h = 0; h < 16; ++h)
                        if (row0 + c16 * 16 + h < row_end)
                            store16(obase + (long long)h * E + mt * 128, __uint_as_float(a[h]) * inv[h], F16 ? 1 : 0);
                }
            }
            tcgen05_fence_before();
            g += (uint32_t)nt;
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(C::TMEM_COLS) : "memory");
    }
}

// ascending list of the crops whose hypothesis is still open (one block; n <= a few 10 k)
__global__ void __launch_bounds__(1024) live_list_kernel(const unsigned char* __restrict__ finished, int n, int* __restrict__ live,
                                                          int* __restrict__ n_live) {
    __shared__ int warp_cnt[32];
    __shared__ int base;
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (threadIdx.x == 0) base = 0;
    __syncthreads();
    for (int i0 = 0; i0 < n; i0 += 1024) {
        const int i = i0 + threadIdx.x;
        const bool alive = i < n && !finished[i];
        const unsigned bal = __ballot_sync(0xffffffffu, alive);
        if (l == 0) warp_cnt[w] = __popc(bal);
        __syncthreads();
        int off = base;
        for (int k = 0; k < w; ++k) off += warp_cnt[k];
        if (alive) live[off + __popc(bal & ((1u << l) - 1u))] = i;
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int k = 0; k < 32; ++k) t += warp_cnt[k];
            base += t;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_live = base;
}

template <bool F16, int E, int KT, int NB>
int launch_xt(mb_ctx* ctx, const CUtensorMap& tmE, const CUtensorMap& tmQ, const XtParams& p, int grid, cudaStream_t s) {
    using C = XtCfg<E, KT, NB>;
    static bool done = false;
    if (!done) {
        MB_CUDA(ctx, cudaFuncSetAttribute(dec_cross_tc_kernel<F16, E, KT, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
        done = true;
    }
    dec_cross_tc_kernel<F16, E, KT, NB><<<grid, XT_THREADS, C::SMEM, s>>>(tmE, tmQ, p);
    MB_LAUNCH_CHECK(ctx);
    return 0;
}

// hypotheses of a crop that share one pass (TMEM: (3 + E / 128) * 16 * NB columns of 512)
int xt_group(int E, int beam, int* groups) {
    const int nb_max = E == 1024 ? 2 : 3;
    *groups = (beam + nb_max - 1) / nb_max;
    return (beam + *groups - 1) / *groups;
}

}  // namespace

// passes over a crop's encoder states that `beam` hypotheses need (1: all of them share one pass)
int mb_cross_enc_tc_groups(int E, int beam) {
    int groups = 1;
    xt_group(E, beam, &groups);
    return groups;
}

bool mb_cross_enc_tc_supported(int E, int heads) { return (E == 128 || E == 768 || E == 1024) && heads >= 1 && heads <= 16; }

// live = finished ? compacted list built here into live_ws (n + 1 ints: list, then the counter) : all rows
int mb_live_list(mb_ctx* ctx, const unsigned char* finished, int n, int* live_ws, cudaStream_t s) {
    live_list_kernel<<<1, 1024, 0, s>>>(finished, n, live_ws, live_ws + n);
    MB_LAUNCH_CHECK(ctx);
    return 0;
}

// n crops x beam hypotheses.  qp, out: [n * beam, heads * E]; enc: [n * T, E]; live_ws: the list of mb_live_list (or null:
// every crop).  The hypotheses of a crop share the pass over its encoder states (up to three per pass).
int mb_cross_enc_tc(mb_ctx* ctx, const bf16* qp, const bf16* enc, bf16* out, int n, int beam, int T, int heads, int E,
                    const int* live_ws, cudaStream_t s) {
    MB_REQUIRE(ctx, mb_cross_enc_tc_supported(E, heads) && n > 0 && T > 0 && beam >= 1,
               "cross_enc_tc: unsupported geometry (E=%d heads=%d beam=%d)", E, heads, beam);
    int groups = 1;
    const int nb = xt_group(E, beam, &groups);
    const int KT = (E == 1024 || (E == 768 && nb > 1)) ? 32 : 64;
    const long long rows = (long long)n * beam;
    // descriptors: the same few (buffer, shape) pairs come back every step and layer
    struct Key { const mb_ctx* c; const void* q; const void* e; int n, beam, T, heads, E, f16; };
    static std::mutex mu;
    static Key last = {nullptr, nullptr, nullptr, 0, 0, 0, 0, 0, 0};
    static CUtensorMap last_e, last_q;
    CUtensorMap tmE, tmQ;
    {
        std::lock_guard<std::mutex> lock(mu);
        if (last.c != ctx || last.q != qp || last.e != enc || last.n != n || last.beam != beam || last.T != T ||
            last.heads != heads || last.E != E || last.f16 != ctx->f16) {
            int rc = mb_encode_3d_map(ctx, &last_e, enc, E, T, n, E, (long long)T * E, 64, KT);
            if (rc) { last.c = nullptr; return rc; }
            rc = mb_encode_2d_map(ctx, &last_q, qp, E, rows * heads, E, 64, 16 * nb);
            if (rc) { last.c = nullptr; return rc; }
            last = Key{ctx, qp, enc, n, beam, T, heads, E, ctx->f16};
        }
        tmE = last_e;
        tmQ = last_q;
    }
    XtParams p;
    p.T = T; p.heads = heads; p.n = n; p.beam = beam; p.groups = groups;
    p.live = live_ws; p.n_live = live_ws ? live_ws + n : nullptr;
    p.out = out; p.diag = ctx->dev_diag;
    const long long items = (long long)n * groups;
    const int grid = (int)(items < ctx->num_sms ? items : ctx->num_sms);
    const bool h = ctx->f16 != 0;
#define XT_GO(EE, KK, BB) return h ? launch_xt<true, EE, KK, BB>(ctx, tmE, tmQ, p, grid, s) : launch_xt<false, EE, KK, BB>(ctx, tmE, tmQ, p, grid, s)
    if (E == 768 && nb == 1) XT_GO(768, 64, 1);
    if (E == 768 && nb == 2) XT_GO(768, 32, 2);
    if (E == 768 && nb == 3) XT_GO(768, 32, 3);
    if (E == 1024 && nb == 1) XT_GO(1024, 32, 1);
    if (E == 1024 && nb == 2) XT_GO(1024, 32, 2);
    if (E == 128 && nb == 1) XT_GO(128, 64, 1);
    if (E == 128 && nb == 2) XT_GO(128, 64, 2);
    if (E == 128 && nb == 3) XT_GO(128, 64, 3);
#undef XT_GO
    return mb_set_err(ctx, MB_ERR_ARG, "cross_enc_tc: no kernel for E=%d with %d hypotheses per pass", E, nb);
}
