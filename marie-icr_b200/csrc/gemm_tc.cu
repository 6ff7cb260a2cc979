// tap-GEMM: the one tensor-core kernel of the library (sm_100a: tcgen05.mma + TMEM + TMA).
//
//   D[pixel, n] = epilogue( sum_{tap} sum_{c} A[pixel + offset(tap), c] * W[n, tap*C + c] )
//
// * A is one (or two channel-concatenated) NHWC bf16 tensor(s).  An M tile is 128 consecutive pixels of one
//   image row; the A tile for tap (dy,dx) is the TMA box shifted by (dy,dx)*dilation — out-of-bounds
//   pixels are zero-filled by TMA, which is exactly the convolution's zero padding.  A plain GEMM is the
//   degenerate case n = h = 1, w = M, taps = 1.
// * W is [N, K] K-major; both operands land in shared memory in the 128-byte-swizzled K-major canonical
//   layout that tcgen05.mma consumes through shared-memory descriptors.
// * Warp roles (persistent CTA, one per SM): warp 0 = TMA producer, warp 1 = MMA issuer (single thread),
//   warp 2 = TMEM allocator, warps 4..7 = epilogue (TMEM -> registers -> global).  The accumulator is
//   double buffered in TMEM (2 x 256 columns) so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// Replaces cuDNN/cuBLAS behind nn.Conv2d / nn.Linear on the reference's hot path
// (marie/models/craft/basenet/vgg16_bn.py:33-47, marie/models/craft/craft.py:14-51,
//  marie/models/unilm/trocr/deit.py:105-146, trocr_models.py:142-147).
#include "common.cuh"
#include "tc_ptx.cuh"
#include <mutex>
#include <unordered_map>
#include <stdlib.h>

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;          // 64 bf16 = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 384;          // warps: 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4..11 epilogue
constexpr int EPI_WARPS = 8;
constexpr int STAGE_TILE_BYTES = 32 * 32 * 4;   // per epilogue warp: 32 rows x 32 fp32 columns
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;   // 16 KiB
constexpr int HALO_ROWS = BLOCK_M + 2;                  // 3x3 convolutions: one A tile with a one-pixel halo serves dx = -1, 0, +1
constexpr int HALO_BYTES = HALO_ROWS * BLOCK_K * 2;     // 16640 B delivered by the TMA
constexpr int A_HALO_STAGE_BYTES = 17 * 1024;           // ... in a 1024-byte aligned slot
constexpr int MAX_STAGES = 8;
constexpr int TMEM_COLS = 512;
constexpr int SMEM_LIMIT = 227 * 1024;
constexpr int BAR_BYTES = 512;   // 2*8 ring + 4 accumulator mbarriers, the TMEM pointer slot, 8 per-warp residual mbarriers;
                                 // 512 keeps the staging tiles aligned for the 64-byte TMA swizzle
constexpr int RES_BAR_SLOT = 2 * MAX_STAGES + 5;   // first of the EPI_WARPS residual mbarriers (u64 slots after the pointer)
constexpr int BRES_BAR_SLOT = RES_BAR_SLOT + EPI_WARPS;   // "resident weights have landed"
constexpr int COEF_BAR_SLOT = BRES_BAR_SLOT + 1;          // two "per-column coefficients of this tile are staged" mbarriers,
                                                          // then two "the epilogue warps have read them" (EPI_WARPS arrivals)
constexpr int COEF_RES_BYTES = 2 * 1024;                  // TMA epilogue with residual: two bias tiles behind the staging tiles

struct KParams {
    int n, h, w;
    int w_tiles, m_tiles, n_tiles;
    int taps, dil;
    int kb0, kb1;          // 64-wide k blocks per tap from source 0 / source 1
    int block_n;
    int n_out;
    int stages;
    const float* bias;
    int act;
    const bf16* residual;
    long long res_ld;
    void* out;
    long long out_ld;
    int out_mode;
    long long out_plane;
    int f16;               // operand / 16-bit output element type: 0 = bf16, 1 = fp16
    int batches, a_col_stride, w_row_stride, out_col_stride;   // block-diagonal (per-head) GEMMs
    const float2* ln_stats;   // LNF: per row (-mean, rstd)
    const float* ln_c;        // LNF: per column sum_k W'[n, k]
    float2* stat_out;         // STATS: per row, N tile and epilogue-warp half (sum, sum of squares) of the written values
    unsigned int* diag;
    unsigned backoff;         // ns slept between polls of the long waits (0 = poll continuously)
    // 3x3 (dilation 1) convolutions: a stage holds the A tile of one image row WITH a one-pixel halo (130 pixel rows) and
    // the three dx taps are three MMAs on the same tile, the operand start shifted by one 128-byte row each — a third of
    // the A traffic from L2.  bres: the whole weight matrix (9 taps x kb blocks) stays resident in shared memory.
    int halo, bres;
};

// exact-erf GELU (timm nn.GELU).  erf(z) = 1 - 2^q(z) for z in [0, 3.92] with a degree-6 polynomial q fitted to
// log2(erfc) (max |erf error| 7.6e-7, max |GELU error| 2.8e-7 — far below the 16-bit output rounding; fit script in
// profiles/r01_gelu_erf_fit.md): 6 FFMA + ONE MUFU.EX2.  erff's branchy path (and the earlier A&S 7.1.26 form with
// two MUFU ops) made the fc1 epilogue issue-bound: 539 vs 1114 TFLOP/s for the same GEMM without activation.
__device__ __forceinline__ float gelu_erf(float x) {
    const float z = fminf(fabsf(x) * 0.70710678118654752440f, 3.92f);
    float q = fmaf(0.00022097790497355163f, z, -0.004072441719472408f);
    q = fmaf(q, z, 0.031655170023441315f);
    q = fmaf(q, z, -0.15032003819942474f);
    q = fmaf(q, z, -0.9179494380950928f);
    q = fmaf(q, z, -1.6279493570327759f);
    float e2;                                           // q * z in [-26, 0]: no denormal range handling needed
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e2) : "f"(q * z));
    const float e = 1.0f - e2;
    const float hx = 0.5f * x;
    return fmaf(fabsf(hx), e, hx);                      // hx * sign(x) * e == |hx| * e: the sign rides on the operand modifier
}
// The same GELU on a PAIR of values with packed fp32 arithmetic (FFMA2), re-parametrised so that no sign handling is
// left: with u = min(|x|, 3.92*sqrt(2)) and t(u) = q(u/sqrt2)*(u/sqrt2) - 1 (a degree-6 polynomial in u, zero-free Horner
// form), h = 2^t = erfc(|x|/sqrt2)/2 and GELU(x) = relu(x) - |x|*h.  Per pair: 2 FMNMX + 6 FFMA2 + 2 MUFU.EX2 + 2 FMUL +
// 2 FMNMX + 1 FFMA2 = 7.5 instructions per value (12 in the scalar form); max |error| 3.4e-7.
__device__ __forceinline__ f32x2 gelu_erf2(f32x2 x2) {
    float x0, x1;
    upk2(x2, x0, x1);
    const float U = 5.5437171645f;
    const f32x2 u = pk2(fminf(fabsf(x0), U), fminf(fabsf(x1), U));
    f32x2 t = fma2(pk2(2.7622238121693954e-05f, 2.7622238121693954e-05f), u, pk2(-0.0007199128158390522f, -0.0007199128158390522f));
    t = fma2(t, u, pk2(0.007913792505860329f, 0.007913792505860329f));
    t = fma2(t, u, pk2(-0.053146157413721085f, -0.053146157413721085f));
    t = fma2(t, u, pk2(-0.4589747190475464f, -0.4589747190475464f));
    t = fma2(t, u, pk2(-1.1511340141296387f, -1.1511340141296387f));
    t = fma2(t, u, pk2(-1.0f, -1.0f));
    float t0, t1, h0, h1;
    upk2(t, t0, t1);
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(h0) : "f"(t0));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(h1) : "f"(t1));
    return fma2(pk2(fabsf(x0) * h0, fabsf(x1) * h1), pk2(-1.0f, -1.0f), pk2(fmaxf(x0, 0.0f), fmaxf(x1, 0.0f)));
}
template <int ACT> __device__ __forceinline__ f32x2 apply_act2(f32x2 x) {
    if (ACT == MB_ACT_GELU) return gelu_erf2(x);
    if (ACT == MB_ACT_RELU) {
        float a, b;
        upk2(x, a, b);
        return pk2(fmaxf(a, 0.0f), fmaxf(b, 0.0f));
    }
    return x;
}
template <int ACT> __device__ __forceinline__ float apply_act(float x) {
    if (ACT == MB_ACT_RELU) return fmaxf(x, 0.0f);
    if (ACT == MB_ACT_GELU) return gelu_erf(x);
    return x;
}

// ---------------------------------------------------------------- kernel
template <bool F16> __device__ __forceinline__ float2 unpack2t(uint32_t v) { return unpack2(v, F16 ? 1 : 0); }
template <bool F16> __device__ __forceinline__ uint32_t pack2t(float a, float b) { return pack2(a, b, F16 ? 1 : 0); }

// CG2 = true: clusters of two CTAs (neighbouring SMs) work on two M tiles of the same N tile with ONE shared B tile:
// each CTA loads its own A tile and half of the B tile, the leader CTA issues tcgen05.mma.cta_group::2 (M = 256: rows
// 0..127 accumulate in the leader's TMEM, 128..255 in the peer's), completions are multicast to both CTAs' mbarriers.
// Per SM and k-block that is 32 KB through shared memory instead of 48 KB, and six pipeline stages instead of four.
// TMAOUT = true (16-bit output with TMA-compatible pitch): the epilogue works row-per-lane straight out of TMEM — bias /
// LayerNorm fold / activation / residual with packed fp32 arithmetic — writes the 16-bit tile once into a 64-byte-swizzled
// staging tile and hands it to the TMA (cp.async.bulk.tensor store, ragged edges clipped by the tensor map); the residual
// tile arrives the same way (TMA load, one chunk ahead).  No fp32 staging round trip, no per-thread global address
// arithmetic, no separate ragged path.
template <int ACT, int OUT, bool RES, bool F16, bool CG2, bool LNF = false, bool TMAOUT = false, bool STATS = false>
__global__ void __launch_bounds__(NUM_THREADS, 1)
tap_gemm_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmOut,
                const __grid_constant__ CUtensorMap tmRes, const KParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // 1024 B alignment is required by the 128B swizzle.  The kernel has no static shared memory, so the dynamic window
    // starts at the declared alignment; a build that breaks this must fault, not corrupt operands (the host sizes the
    // allocation without slack: the last kilobyte pays for the coefficient tiles of the TMA epilogue).
    uint8_t* smem = smem_raw;
    if ((smem_u32(smem_raw) & 1023u) != 0) {
        if (threadIdx.x == 0 && p.diag) atomicExch(p.diag, 0xDEAD00A1u);
        __trap();
    }
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int b_slot_bytes = (CG2 ? p.block_n / 2 : p.block_n) * BLOCK_K * 2;
    const int b_stage_bytes = p.halo ? (p.bres ? 0 : 3 * b_slot_bytes) : b_slot_bytes;
    const int a_stage_bytes = p.halo ? A_HALO_STAGE_BYTES : A_STAGE_BYTES;
    const uint32_t crank = CG2 ? cluster_ctarank() : 0u;
    const bool leader = crank == 0;
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + p.stages * a_stage_bytes;
    uint8_t* smem_bres = smem_b + p.stages * b_stage_bytes;            // resident weights (halo mode, small layers)
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_bres + (p.bres ? 9 * (p.kb0 + p.kb1) * b_slot_bytes : 0));
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + MAX_STAGES;
    uint64_t* tmem_full_bar = bars + 2 * MAX_STAGES;
    uint64_t* tmem_empty_bar = bars + 2 * MAX_STAGES + 2;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * MAX_STAGES + 4);
    uint8_t* smem_stage = reinterpret_cast<uint8_t*>(bars) + BAR_BYTES;      // epilogue staging tiles (16 B aligned)

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA0) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA1) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        if (TMAOUT) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmOut) : "memory");
        if (TMAOUT && RES) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmRes) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < p.stages; ++i) {
            mbar_init(smem_u32(&full_bar[i]), 1);
            mbar_init(smem_u32(&empty_bar[i]), 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(smem_u32(&tmem_full_bar[i]), 1);
            mbar_init(smem_u32(&tmem_empty_bar[i]), CG2 ? 2 * EPI_WARPS : EPI_WARPS);   // one arrival per epilogue warp
        }
        if (TMAOUT && RES)
            for (int i = 0; i < EPI_WARPS; ++i) mbar_init(smem_u32(&bars[RES_BAR_SLOT + i]), 1);
        mbar_init(smem_u32(&bars[BRES_BAR_SLOT]), 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(smem_u32(&bars[COEF_BAR_SLOT + i]), 1);
            mbar_init(smem_u32(&bars[COEF_BAR_SLOT + 2 + i]), EPI_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        if (CG2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                         "r"(TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                         "r"(TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (CG2) cluster_sync_all();      // the peer's mbarriers are initialised before anything is signalled remotely
    tcgen05_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_ptr_smem);

    // CG2: a "tile" index addresses a PAIR of M tiles (2*mp, 2*mp+1) of one N tile; this CTA takes M tile 2*mp + rank
    const int m_units = CG2 ? (p.m_tiles + 1) / 2 : p.m_tiles;
    const int tiles_per_batch = m_units * p.n_tiles;
    const int total_tiles = tiles_per_batch * p.batches;
    const int tile0 = CG2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int tile_step = CG2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int kb_per_tap = p.kb0 + p.kb1;
    const int k_iters = (p.halo ? 3 : p.taps) * kb_per_tap;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer (whole warp loops, one lane issues)
        {
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t tx_bytes = p.halo ? HALO_BYTES + b_stage_bytes : A_STAGE_BYTES + b_stage_bytes;
            if (p.halo && p.bres && tile0 < total_tiles) {         // the whole weight matrix, once per CTA
                const uint32_t rb = smem_u32(&bars[BRES_BAR_SLOT]);
                if (elect_one_sync()) {
                    mbar_arrive_expect_tx(rb, 9 * kb_per_tap * b_slot_bytes);
                    for (int i = 0; i < 9 * kb_per_tap; ++i)
                        tma_load_2d(smem_u32(smem_bres + i * b_slot_bytes), &tmB, rb, i * BLOCK_K, 0);
                }
                __syncwarp();
            }
            for (int tile = tile0; tile < total_tiles; tile += tile_step) {
                const int bt = tile / tiles_per_batch;
                const int trem = tile - bt * tiles_per_batch;
                const int nt = trem % p.n_tiles;
                int mt = trem / p.n_tiles;
                if (CG2) mt = min(2 * mt + (int)crank, p.m_tiles - 1);   // a missing odd partner re-loads the last tile
                const int wt = mt % p.w_tiles;
                const int rest = mt / p.w_tiles;
                const int hh = rest % p.h;
                const int nn = rest / p.h;
                const int w0 = wt * BLOCK_M;
                const int acol = bt * p.a_col_stride, wrow = bt * p.w_row_stride + (CG2 ? (int)crank * (p.block_n / 2) : 0);
                if (p.halo) {
                    for (int dyi = 0; dyi < 3; ++dyi) {
                        for (int kb = 0; kb < kb_per_tap; ++kb) {
                            mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1, p.diag, 1, p.backoff);
                            const uint32_t fb = smem_u32(&full_bar[stage]);
                            const uint32_t sa = smem_u32(smem_a + stage * a_stage_bytes);
                            const uint32_t sb = smem_u32(smem_b + stage * b_stage_bytes);
                            if (elect_one_sync()) {
                                mbar_arrive_expect_tx(fb, tx_bytes);
                                if (kb < p.kb0)
                                    tma_load_4d(sa, &tmA0, fb, kb * BLOCK_K + acol, w0 - 1, hh + dyi - 1, nn);
                                else
                                    tma_load_4d(sa, &tmA1, fb, (kb - p.kb0) * BLOCK_K, w0 - 1, hh + dyi - 1, nn);
                                if (!p.bres) {
#pragma unroll
                                    for (int dxi = 0; dxi < 3; ++dxi)
                                        tma_load_2d(sb + dxi * b_slot_bytes, &tmB, fb, ((dyi * 3 + dxi) * kb_per_tap + kb) * BLOCK_K,
                                                    nt * p.block_n + wrow);
                                }
                            }
                            __syncwarp();
                            if (++stage == p.stages) { stage = 0; phase ^= 1; }
                        }
                    }
                    continue;
                }
                for (int tap = 0; tap < p.taps; ++tap) {
                    const int dy = (p.taps == 9) ? (tap / 3 - 1) * p.dil : 0;
                    const int dx = (p.taps == 9) ? (tap % 3 - 1) * p.dil : 0;
                    for (int kb = 0; kb < kb_per_tap; ++kb) {
                        mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1, p.diag, 1, p.backoff);
                        const uint32_t fb = smem_u32(&full_bar[stage]);
                        const uint32_t sa = smem_u32(smem_a + stage * A_STAGE_BYTES);
                        const uint32_t sb = smem_u32(smem_b + stage * b_stage_bytes);
                        if (elect_one_sync()) {
                        if (CG2) {
                            // the leader's barrier collects the bytes of both CTAs
                            if (leader) mbar_arrive_expect_tx(fb, 2 * tx_bytes);
                            if (kb < p.kb0)
                                tma_load_4d_cg2(sa, &tmA0, fb, kb * BLOCK_K + acol, w0 + dx, hh + dy, nn);
                            else
                                tma_load_4d_cg2(sa, &tmA1, fb, (kb - p.kb0) * BLOCK_K, w0 + dx, hh + dy, nn);
                            tma_load_2d_cg2(sb, &tmB, fb, (tap * kb_per_tap + kb) * BLOCK_K, nt * p.block_n + wrow);
                        } else {
                        mbar_arrive_expect_tx(fb, tx_bytes);
                        if (kb < p.kb0)
                            tma_load_4d(sa, &tmA0, fb, kb * BLOCK_K + acol, w0 + dx, hh + dy, nn);
                        else
                            tma_load_4d(sa, &tmA1, fb, (kb - p.kb0) * BLOCK_K, w0 + dx, hh + dy, nn);
                        tma_load_2d(sb, &tmB, fb, (tap * kb_per_tap + kb) * BLOCK_K, nt * p.block_n + wrow);
                        }
                        }
                        __syncwarp();
                        if (++stage == p.stages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer (whole warp loops, one lane issues)
        if (leader) {
            // instruction descriptor: D=f32, A=B=bf16, both K-major, N>>3 at bit 17, M>>4 at bit 24 (M = 256 for a CTA pair)
            const uint32_t ab_fmt = p.f16 ? 0u : ((1u << 7) | (1u << 10));   // a/b format: 0 = f16, 1 = bf16
            const uint32_t idesc = (1u << 4) | ab_fmt |
                                   ((uint32_t)(p.block_n >> 3) << 17) | ((uint32_t)((CG2 ? 2 * BLOCK_M : BLOCK_M) >> 4) << 24);
            int stage = 0;
            uint32_t phase = 0;
            int as = 0;
            uint32_t aphase = 0;
            bool bres_ready = false;
            for (int tile = tile0; tile < total_tiles; tile += tile_step) {
                mbar_wait(smem_u32(&tmem_empty_bar[as]), aphase ^ 1, p.diag, 2);
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * 256);
                for (int it = 0; it < k_iters; ++it) {
                    mbar_wait(smem_u32(&full_bar[stage]), phase, p.diag, 3);
                    tcgen05_fence_after();
                    if (p.halo) {
                        if (p.bres && !bres_ready) {
                            mbar_wait(smem_u32(&bars[BRES_BAR_SLOT]), 0, p.diag, 6);
                            tcgen05_fence_after();
                            bres_ready = true;
                        }
                        const uint32_t sa = smem_u32(smem_a + stage * a_stage_bytes);
                        const int dyi = it / kb_per_tap, kb = it - dyi * kb_per_tap;
                        if (elect_one_sync()) {
#pragma unroll
                        for (int dxi = 0; dxi < 3; ++dxi) {
                            // rows dxi .. dxi+127 of the halo tile: the operand starts one 128-byte row further.  The 128-byte
                            // swizzle is a function of the absolute shared-memory address (TMA wrote it that way), so the
                            // descriptor needs no base offset (measured: setting it breaks the result)
                            const uint64_t adesc = make_smem_desc(sa + (uint32_t)dxi * 128u);
                            const uint32_t sbp = p.bres ? smem_u32(smem_bres + ((dyi * 3 + dxi) * kb_per_tap + kb) * b_slot_bytes)
                                                        : smem_u32(smem_b + stage * b_stage_bytes + dxi * b_slot_bytes);
                            const uint64_t bdesc = make_smem_desc(sbp);
#pragma unroll
                            for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                                umma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc,
                                          (it > 0 || dxi > 0 || k > 0) ? 1u : 0u);
                        }
                        tcgen05_commit(smem_u32(&empty_bar[stage]));
                        if (it == k_iters - 1) tcgen05_commit(smem_u32(&tmem_full_bar[as]));
                        }
                        __syncwarp();
                        if (++stage == p.stages) { stage = 0; phase ^= 1; }
                        continue;
                    }
                    const uint64_t adesc = make_smem_desc(smem_u32(smem_a + stage * A_STAGE_BYTES));
                    const uint64_t bdesc = make_smem_desc(smem_u32(smem_b + stage * b_stage_bytes));
                    if (elect_one_sync()) {
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                        // advance 16 elements = 32 B along K inside the swizzle row: +2 in 16 B units
                        if (CG2)
                            umma_bf16_cg2(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc,
                                          (it > 0 || k > 0) ? 1u : 0u);
                        else
                            umma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc,
                                      (it > 0 || k > 0) ? 1u : 0u);
                    }
                    if (CG2) {
                        tcgen05_commit_cg2_mc(smem_u32(&empty_bar[stage]));   // frees the slot in BOTH CTAs
                        if (it == k_iters - 1) tcgen05_commit_cg2_mc(smem_u32(&tmem_full_bar[as]));
                    } else {
                    tcgen05_commit(smem_u32(&empty_bar[stage]));     // frees the smem slot when MMAs retire
                    if (it == k_iters - 1) tcgen05_commit(smem_u32(&tmem_full_bar[as]));
                    }
                    }
                    __syncwarp();
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
                as ^= 1;
                if (as == 0) aphase ^= 1;
            }
        }
    } else if (warp == 3 && TMAOUT) {
        // ------------------------------------------------------------ coefficient producer (TMA epilogue only)
        // The epilogue works row-per-lane, so every lane needs ALL per-column coefficients of its chunk (bias; LayerNorm
        // fold: c_n as well).  Read through L1 they cost the epilogue warps half their time (ncu, r02: 16 LDG.128 per chunk,
        // every sector missing — the 227 KB carve-out leaves L1 a few KB — and stall_long_scoreboard on each FFMA2 that
        // consumed them: the three K = 768 encoder GEMMs all ran at the same ~10.6 K cycles per tile, epilogue-bound).  This
        // otherwise idle warp stages the tile's <= 256 bias (and c) values in shared memory one accumulator buffer ahead;
        // the epilogue reads them as broadcast LDS.128.  Buffer `as` is free again when this CTA's eight epilogue warps have
        // finished the tile that used it (a CTA-local barrier: in a CTA pair the TMEM release goes to the leader).
        int as = 0;
        uint32_t aphase = 0;
        for (int tile = tile0; tile < total_tiles; tile += tile_step) {
            const int bt = tile / tiles_per_batch;
            const int nt = (tile - bt * tiles_per_batch) % p.n_tiles;
            const int nbase = bt * p.out_col_stride, ntile0 = nt * p.block_n;
            float* cb = RES ? reinterpret_cast<float*>(smem_stage + EPI_WARPS * STAGE_TILE_BYTES + as * (COEF_RES_BYTES / 2))
                            : reinterpret_cast<float*>(smem_stage + as * STAGE_TILE_BYTES + 2048);
            mbar_wait(smem_u32(&bars[COEF_BAR_SLOT + 2 + as]), aphase ^ 1, p.diag, 7, p.backoff);
            for (int j = lane; j < p.block_n; j += 32) {
                const bool in = ntile0 + j < p.n_out;
                cb[j] = (in && p.bias != nullptr) ? __ldg(p.bias + nbase + ntile0 + j) : 0.f;
                if (LNF) cb[256 + j] = in ? __ldg(p.ln_c + nbase + ntile0 + j) : 0.f;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bars[COEF_BAR_SLOT + as]));
            as ^= 1;
            if (as == 0) aphase ^= 1;
        }
    } else if (warp >= 4 && TMAOUT) {
        // ------------------------------------------------------------ epilogue, TMA-store form (8 warps)
        // Warp e owns TMEM lane quarter (warp & 3) = 32 tile rows, one row per lane, and every second 32-column chunk.
        static_assert(!(TMAOUT && RES && LNF), "the TMA epilogue keeps the LayerNorm-fold coefficients in the residual staging area");
        const int ew = warp - 4;
        const int q = warp & 3;
        const int half = ew >> 2;
        uint8_t* stg_out = smem_stage + ew * STAGE_TILE_BYTES;           // 32 rows x 64 B, 64-byte swizzle: TMA store source
        uint8_t* stg_res = stg_out + 2048;                                // same layout: TMA load destination (residual)
        const uint32_t res_bar = smem_u32(&bars[RES_BAR_SLOT + ew]);
        uint32_t rphase = 0;
        const uint32_t sw_row = (uint32_t)lane * 64u, sw_x = ((uint32_t)lane >> 1) & 3u;
        const int n_chunks = (p.block_n + 31) >> 5;
        struct TileC { int ntile0, nbase, wrow, hh, nn, rows_valid; };
        auto coords = [&](int tile) {
            TileC t;
            const int bt = tile / tiles_per_batch;
            const int trem = tile - bt * tiles_per_batch;
            const int nt = trem % p.n_tiles;
            int mt = trem / p.n_tiles;
            bool exists = true;
            if (CG2) { mt = 2 * mt + (int)crank; exists = mt < p.m_tiles; mt = min(mt, p.m_tiles - 1); }
            const int wt = mt % p.w_tiles;
            const int rest = mt / p.w_tiles;
            t.hh = rest % p.h;
            t.nn = rest / p.h;
            t.nbase = bt * p.out_col_stride;
            t.wrow = wt * BLOCK_M + q * 32;                          // w coordinate of this warp's first row
            t.rows_valid = exists ? min(32, p.w - t.wrow) : 0;       // <= 0: ragged last tile / missing partner of a pair
            t.ntile0 = nt * p.block_n;
            return t;
        };
        auto chunk_active = [&](const TileC& t, int ci) { return ci < n_chunks && t.ntile0 + ci * 32 < p.n_out && t.rows_valid > 0; };
        auto issue_res = [&](const TileC& t, int ci) {
            mbar_arrive_expect_tx(res_bar, 2048);
            tma_load_4d(smem_u32(stg_res), &tmRes, res_bar, t.nbase + t.ntile0 + ci * 32, t.wrow, t.hh, t.nn);
        };
        // accumulator buffer back to the MMA warp (the pair leader's in two-CTA mode), coefficient tile back to warp 3
        auto release_tile = [&](int a) {
            if (CG2) mbar_arrive_cluster(smem_u32(&tmem_empty_bar[a]), 0);
            else mbar_arrive(smem_u32(&tmem_empty_bar[a]));
            mbar_arrive(smem_u32(&bars[COEF_BAR_SLOT + 2 + a]));
        };
        int as = 0;
        uint32_t aphase = 0;
        bool res_pending = false;        // this tile's first residual chunk was requested under the previous tile's last chunk
        for (int tile = tile0; tile < total_tiles; tile += tile_step) {
            const TileC t = coords(tile);
            const long long pix0 = ((long long)t.nn * p.h + t.hh) * p.w + t.wrow;
            f32x2 nm2 = 0, rstd2 = 0;
            f32x2 st_s = 0, st_q = 0;                    // STATS: this lane's row, this warp's chunks of the tile
            if (LNF) {
                const float2 st = lane < t.rows_valid ? __ldg(p.ln_stats + pix0 + lane) : make_float2(0.f, 0.f);
                nm2 = pk2(st.x, st.x);
                rstd2 = pk2(st.y, st.y);
            }
            if (RES && !res_pending && chunk_active(t, half) && lane == 0) issue_res(t, half);
            res_pending = false;
            mbar_wait(smem_u32(&tmem_full_bar[as]), aphase, p.diag, 4, p.backoff);
            tcgen05_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * 256);
            uint32_t v[32];
            bool released = false;
            if (chunk_active(t, half)) tmem_ld32(taddr + (uint32_t)(half * 32), v);
            mbar_wait(smem_u32(&bars[COEF_BAR_SLOT + as]), aphase, p.diag, 8);
            const float* cb = RES ? reinterpret_cast<const float*>(smem_stage + EPI_WARPS * STAGE_TILE_BYTES + as * (COEF_RES_BYTES / 2))
                                  : reinterpret_cast<const float*>(smem_stage + as * STAGE_TILE_BYTES + 2048);
#pragma unroll 1
            for (int ci = half; chunk_active(t, ci); ci += 2) {
                const int n0 = t.ntile0 + ci * 32;
                uint32_t rr[16];
                if (RES) {
                    mbar_wait(res_bar, rphase, p.diag, 5);
                    rphase ^= 1;
#pragma unroll
                    for (uint32_t c = 0; c < 4; ++c) {
                        const uint4 x = *reinterpret_cast<const uint4*>(stg_res + sw_row + ((c ^ sw_x) << 4));
                        rr[4 * c] = x.x; rr[4 * c + 1] = x.y; rr[4 * c + 2] = x.z; rr[4 * c + 3] = x.w;
                    }
                    __syncwarp();
                    // the next residual chunk lands under this chunk's arithmetic; after the tile's last chunk that is the
                    // first chunk of the NEXT tile (its latency would otherwise sit in front of that tile's epilogue)
                    if (chunk_active(t, ci + 2)) {
                        if (lane == 0) issue_res(t, ci + 2);
                    } else if (tile + tile_step < total_tiles) {
                        const TileC tn = coords(tile + tile_step);
                        if (chunk_active(tn, half)) {
                            if (lane == 0) issue_res(tn, half);
                            res_pending = true;
                        }
                    }
                }
                tmem_wait_ld();
                uint32_t o[16];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float4 b4 = *reinterpret_cast<const float4*>(cb + ci * 32 + 4 * k);      // broadcast reads
                    f32x2 xa = pk2(v[4 * k], v[4 * k + 1]), xb = pk2(v[4 * k + 2], v[4 * k + 3]);
                    const f32x2 ba = pk2(b4.x, b4.y), bb = pk2(b4.z, b4.w);
                    if (LNF) {       // rstd * (acc - mean * c) + b'
                        const float4 c4 = *reinterpret_cast<const float4*>(cb + 256 + ci * 32 + 4 * k);
                        xa = fma2(fma2(nm2, pk2(c4.x, c4.y), xa), rstd2, ba);
                        xb = fma2(fma2(nm2, pk2(c4.z, c4.w), xb), rstd2, bb);
                    } else {
                        xa = add2(xa, ba);
                        xb = add2(xb, bb);
                    }
                    xa = apply_act2<ACT>(xa);
                    xb = apply_act2<ACT>(xb);
                    if (RES) {
                        const float2 r0 = unpack2t<F16>(rr[2 * k]), r1 = unpack2t<F16>(rr[2 * k + 1]);
                        xa = add2(xa, pk2(r0.x, r0.y));
                        xb = add2(xb, pk2(r1.x, r1.y));
                    }
                    if (STATS) {
                        st_s = add2(st_s, add2(xa, xb));
                        st_q = fma2(xa, xa, st_q);
                        st_q = fma2(xb, xb, st_q);
                    }
                    float f0, f1, f2, f3;
                    upk2(xa, f0, f1);
                    upk2(xb, f2, f3);
                    o[2 * k] = pack2t<F16>(f0, f1);
                    o[2 * k + 1] = pack2t<F16>(f2, f3);
                }
                // the accumulators of this chunk are consumed: next chunk's TMEM load overlaps the store sequence; after
                // the last chunk the accumulator buffer goes back to the MMA warp before the stores are even issued
                if (chunk_active(t, ci + 2)) {
                    tmem_ld32(taddr + (uint32_t)((ci + 2) * 32), v);
                } else {
                    tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) release_tile(as);
                    released = true;
                }
                if (lane == 0) bulk_wait_group_read0();              // this warp's previous store has read the staging tile
                __syncwarp();
#pragma unroll
                for (uint32_t c = 0; c < 4; ++c)
                    *reinterpret_cast<uint4*>(stg_out + sw_row + ((c ^ sw_x) << 4)) =
                        make_uint4(o[4 * c], o[4 * c + 1], o[4 * c + 2], o[4 * c + 3]);
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    tma_store_4d(&tmOut, smem_u32(stg_out), t.nbase + n0, t.wrow, t.hh, t.nn);
                    bulk_commit_group();
                }
            }
            if (!released) {
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) release_tile(as);
            }
            if (STATS) {
                if (lane < t.rows_valid) {
                    float s0, s1, q0, q1;
                    upk2(st_s, s0, s1);
                    upk2(st_q, q0, q1);
                    p.stat_out[(pix0 + lane) * (2 * p.n_tiles) + 2 * (t.ntile0 / p.block_n) + half] = make_float2(s0 + s1, q0 + q1);
                }
            }
            as ^= 1;
            if (as == 0) aphase ^= 1;
        }
        if (lane == 0) bulk_wait_group0();     // the staging tile must outlive the last store's read
    } else if (warp >= 4) {
        // ------------------------------------------------------------ epilogue (8 warps)
        // Warp e = warp-4 owns TMEM lane quarter (warp & 3) — the 32 tile rows it may read — and every second
        // 32-column chunk.  A chunk goes TMEM -> registers (row per lane) -> bias/activation -> fp32 staging tile in
        // shared memory (XOR-swizzled, conflict-free) -> read back with lanes running along the row, so the
        // residual loads and the output stores are coalesced row segments instead of 32 scattered rows.
        // The kernel is specialised on (ACT, OUT, RES, F16): a runtime-generic epilogue unrolled to ~280 KB of SASS
        // and ran out of the instruction cache (ncu: stall_no_inst on every epilogue instruction).
        const int ew = warp - 4;
        const int q = warp & 3;
        const int half = ew >> 2;
        float* stage = reinterpret_cast<float*>(smem_stage + ew * STAGE_TILE_BYTES);
        const int rr = lane >> 3, kk = lane & 7;  // read-back role: row offset within a group of 4, 16-byte chunk
        const bool vec_out = (p.out_ld & 3) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0;
        const bool vec_res = !RES || ((p.res_ld & 3) == 0 && (reinterpret_cast<uintptr_t>(p.residual) & 7) == 0);
        int as = 0;
        uint32_t aphase = 0;
        for (int tile = tile0; tile < total_tiles; tile += tile_step) {
            const int bt = tile / tiles_per_batch;
            const int trem = tile - bt * tiles_per_batch;
            const int nt = trem % p.n_tiles;
            int mt = trem / p.n_tiles;
            bool tile_exists = true;
            if (CG2) { mt = 2 * mt + (int)crank; tile_exists = mt < p.m_tiles; mt = min(mt, p.m_tiles - 1); }
            const int wt = mt % p.w_tiles;
            const int rest = mt / p.w_tiles;
            const int hh = rest % p.h;
            const int nn = rest / p.h;
            const int nbase = bt * p.out_col_stride;                // output / bias / residual column of this batch
            const int row0 = wt * BLOCK_M + q * 32;                 // first tile row (pixel) of this warp
            const long long pix0 = ((long long)nn * p.h + hh) * p.w + row0;
            const int rows_valid = tile_exists ? min(32, p.w - row0) : 0;   // <= 0: ragged last tile / missing partner

            // LNF: the (-mean, rstd) pairs of this lane's eight read-back rows, once per tile
            float2 lst[8];
            if (LNF) {
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    lst[i] = (i * 4 + rr < rows_valid) ? __ldg(p.ln_stats + pix0 + i * 4 + rr) : make_float2(0.f, 0.f);
            }
            const int n_chunks = (p.block_n + 31) >> 5;
            // residual tile rows are prefetched one chunk ahead (the first chunk before the accumulator is even ready)
            uint2 rres[8];
            auto chunk_fast = [&](int ci) {
                const int n0c = nt * p.block_n + ci * 32;
                return ci < n_chunks && n0c + 32 <= p.n_out && rows_valid == 32 && vec_out && vec_res;
            };
            auto load_res = [&](int ci) {
                const int n0c = nt * p.block_n + ci * 32;
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    rres[i] = __ldg(reinterpret_cast<const uint2*>(p.residual + (pix0 + i * 4 + rr) * p.res_ld + nbase + n0c + kk * 4));
            };
            if (RES && chunk_fast(half)) load_res(half);
            mbar_wait(smem_u32(&tmem_full_bar[as]), aphase, p.diag, 4, p.backoff);
            tcgen05_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * 256);
            uint32_t v[32];
            if (half < n_chunks) tmem_ld32(taddr + (uint32_t)(half * 32), v);     // first chunk's accumulators in flight
#pragma unroll 1
            for (int ci = half; ci < n_chunks; ci += 2) {
                const int n0 = nt * p.block_n + ci * 32;
                const bool active = n0 < p.n_out && rows_valid > 0;               // warp-uniform
                const int ncols = min(32, p.n_out - n0);
                const bool fast = ncols == 32 && rows_valid == 32 && vec_out && vec_res;   // warp-uniform
                tmem_wait_ld();
                if (OUT == MB_OUT_F32_PLANAR) {   // channel planes, pixel-contiguous: coalesced across lanes as is
                    if (active && lane < rows_valid) {
                        float* o = reinterpret_cast<float*>(p.out) + pix0 + lane;
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            if (j < ncols) {
                                float x = __uint_as_float(v[j]);
                                if (p.bias != nullptr) x += __ldg(p.bias + nbase + n0 + j);
                                o[(long long)(nbase + n0 + j) * p.out_plane] = apply_act<ACT>(x);
                            }
                        }
                    }
                    __syncwarp();
                    if (ci + 2 < n_chunks) tmem_ld32(taddr + (uint32_t)((ci + 2) * 32), v);
                    continue;
                }
                // phase A: raw accumulators -> staging tile (row per lane, 16-byte chunks XOR-swizzled)
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    *reinterpret_cast<uint4*>(stage + lane * 32 + ((k ^ (lane & 7)) << 2)) =
                        make_uint4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
                __syncwarp();
                // the next chunk's TMEM load overlaps phase B (v is dead until the wait at the top of the loop)
                if (ci + 2 < n_chunks) tmem_ld32(taddr + (uint32_t)((ci + 2) * 32), v);
                if (active) {
                    if (fast) {
                        // phase B: lanes run along the row (8 lanes x 4 columns, 4 rows per pass): bias once per chunk,
                        // activation, residual, pack, coalesced store
                        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (p.bias != nullptr) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + nbase + n0) + kk);
                        float4 c4 = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (LNF) c4 = __ldg(reinterpret_cast<const float4*>(p.ln_c + nbase + n0) + kk);
                        float4 xs[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int r = i * 4 + rr;
                            float4 x = *reinterpret_cast<const float4*>(stage + r * 32 + ((kk ^ (r & 7)) << 2));
                            if (LNF) {   // rstd * (acc - mean * c) + b'
                                x.x = fmaf(fmaf(lst[i].x, c4.x, x.x), lst[i].y, b4.x);
                                x.y = fmaf(fmaf(lst[i].x, c4.y, x.y), lst[i].y, b4.y);
                                x.z = fmaf(fmaf(lst[i].x, c4.z, x.z), lst[i].y, b4.z);
                                x.w = fmaf(fmaf(lst[i].x, c4.w, x.w), lst[i].y, b4.w);
                                x.x = apply_act<ACT>(x.x); x.y = apply_act<ACT>(x.y);
                                x.z = apply_act<ACT>(x.z); x.w = apply_act<ACT>(x.w);
                            } else {
                            x.x = apply_act<ACT>(x.x + b4.x); x.y = apply_act<ACT>(x.y + b4.y);
                            x.z = apply_act<ACT>(x.z + b4.z); x.w = apply_act<ACT>(x.w + b4.w);
                            }
                            if (RES) {
                                const float2 r0 = unpack2t<F16>(rres[i].x), r1 = unpack2t<F16>(rres[i].y);
                                x.x += r0.x; x.y += r0.y; x.z += r1.x; x.w += r1.y;
                            }
                            xs[i] = x;
                        }
                        if (RES && chunk_fast(ci + 2)) load_res(ci + 2);    // next chunk's residual rows: a chunk ahead
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int r = i * 4 + rr;
                            const float4 x = xs[i];
                            if (OUT == MB_OUT_BF16)
                                *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(p.out) + (pix0 + r) * p.out_ld + nbase + n0 + kk * 4) =
                                    make_uint2(pack2t<F16>(x.x, x.y), pack2t<F16>(x.z, x.w));
                            else
                                *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + (pix0 + r) * p.out_ld + nbase + n0 + kk * 4) = x;
                        }
                    } else {
                        // ragged chunk (last rows / last columns / unaligned pitch): one element per lane and row
                        const float bl = (p.bias != nullptr && lane < ncols) ? __ldg(p.bias + nbase + n0 + lane) : 0.f;
                        const float cl = (LNF && lane < ncols) ? __ldg(p.ln_c + nbase + n0 + lane) : 0.f;
#pragma unroll 1
                        for (int r = 0; r < rows_valid; ++r) {
                            if (lane < ncols) {
                                float x = stage[r * 32 + ((((lane >> 2) ^ (r & 7)) << 2) | (lane & 3))];
                                if (LNF) {
                                    const float2 st = __ldg(p.ln_stats + pix0 + r);
                                    x = apply_act<ACT>(fmaf(fmaf(st.x, cl, x), st.y, bl));
                                } else {
                                    x = apply_act<ACT>(x + bl);
                                }
                                if (RES) x += load16(p.residual + (pix0 + r) * p.res_ld + nbase + n0 + lane, F16);
                                if (OUT == MB_OUT_BF16)
                                    store16(reinterpret_cast<bf16*>(p.out) + (pix0 + r) * p.out_ld + nbase + n0 + lane, x, F16);
                                else
                                    reinterpret_cast<float*>(p.out)[(pix0 + r) * p.out_ld + nbase + n0 + lane] = x;
                            }
                        }
                    }
                }
                if (RES && !(active && fast) && chunk_fast(ci + 2)) load_res(ci + 2);   // keep the prefetch chain alive
                __syncwarp();
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (CG2) mbar_arrive_cluster(smem_u32(&tmem_empty_bar[as]), 0);   // the leader's barrier
                else mbar_arrive(smem_u32(&tmem_empty_bar[as]));
            }
            as ^= 1;
            if (as == 0) aphase ^= 1;
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (CG2) cluster_sync_all();      // no CTA leaves (or frees TMEM) while its peer may still signal / read it
    if (warp == 2) {
        tcgen05_fence_after();
        if (CG2)
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
        else
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

// NHWC activation map: dims {C, W, H, N}, box {64, 128, 1, 1}
int encode_act_map(mb_ctx* ctx, CUtensorMap* m, const bf16* base, int c, int ld, int n, int h, int w, int f16, int box_rows = BLOCK_M) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return mb_set_err(ctx, MB_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
    cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)w * ld * 2, (cuuint64_t)h * w * ld * 2};
    cuuint32_t box[4] = {BLOCK_K, (cuuint32_t)box_rows, 1, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = fn(m, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<bf16*>(base), dims, strides, box, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return mb_set_err(ctx, MB_ERR_CUDA, "cuTensorMapEncodeTiled(act c=%d ld=%d n=%d h=%d w=%d) -> %d", c,
                          ld, n, h, w, (int)r);
    return 0;
}

int encode_wgt_map(mb_ctx* ctx, CUtensorMap* m, const bf16* base, int ktot, int rows, int block_n, int f16) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return mb_set_err(ctx, MB_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[2] = {(cuuint64_t)ktot, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ktot * 2};
    cuuint32_t box[2] = {BLOCK_K, (cuuint32_t)block_n};
    cuuint32_t es[2] = {1, 1};
    CUresult r = fn(m, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<bf16*>(base), dims, strides, box, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return mb_set_err(ctx, MB_ERR_CUDA, "cuTensorMapEncodeTiled(wgt k=%d rows=%d bn=%d) -> %d", ktot,
                          rows, block_n, (int)r);
    return 0;
}

// NHWC 16-bit output / residual map for the TMA epilogue: dims {C, W, H, N}, box {32, 32, 1, 1}, 64-byte swizzle
int encode_io_map(mb_ctx* ctx, CUtensorMap* m, const void* base, long long c, long long ld, int n, int h, int w, int f16) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return mb_set_err(ctx, MB_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
    cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)w * ld * 2, (cuuint64_t)h * w * ld * 2};
    cuuint32_t box[4] = {32, 32, 1, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = fn(m, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims,
                    strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return mb_set_err(ctx, MB_ERR_CUDA, "cuTensorMapEncodeTiled(io c=%lld ld=%lld n=%d h=%d w=%d) -> %d", c, ld, n, h, w, (int)r);
    return 0;
}

// Descriptor cache: the same (buffer, shape) pairs come back every layer / step / page batch (arena pointers are stable),
// so a tensor map is encoded once instead of three to five driver calls per launch.
struct MapKey {
    const void* base; long long a, b; int kind, n, h, w, f16;
    bool operator==(const MapKey& o) const {
        return base == o.base && a == o.a && b == o.b && kind == o.kind && n == o.n && h == o.h && w == o.w && f16 == o.f16;
    }
};
struct MapKeyHash {
    size_t operator()(const MapKey& k) const {
        size_t x = reinterpret_cast<size_t>(k.base) * 0x9E3779B97F4A7C15ull;
        auto mix = [&x](long long v) { x ^= (size_t)v + 0x9E3779B97F4A7C15ull + (x << 6) + (x >> 2); };
        mix(k.a); mix(k.b); mix(k.kind); mix(k.n); mix(k.h); mix(k.w); mix(k.f16);
        return x;
    }
};
struct MapCache {
    std::unordered_map<MapKey, CUtensorMap, MapKeyHash> maps;
    std::mutex mu;
};
MapCache& map_cache(mb_ctx* ctx) {
    static std::mutex mu;
    static std::unordered_map<mb_ctx*, MapCache*> all;
    std::lock_guard<std::mutex> g(mu);
    MapCache*& c = all[ctx];
    if (!c) c = new MapCache();
    return *c;
}
// kind 0: activation map (a = channels, b = pitch), 1: weight map (a = ktot, b = rows, n = block_n), 2: output / residual map
template <class F> int cached_map(mb_ctx* ctx, CUtensorMap* out, const MapKey& key, F encode) {
    MapCache& c = map_cache(ctx);
    std::lock_guard<std::mutex> g(c.mu);
    auto it = c.maps.find(key);
    if (it != c.maps.end()) { *out = it->second; return 0; }
    const int rc = encode(out);
    if (rc) return rc;
    if (c.maps.size() > 16384) c.maps.clear();
    c.maps.emplace(key, *out);
    return 0;
}

}  // namespace

void mb_profile_drain(mb_ctx* ctx) {
    if (ctx->prof_events.empty()) return;
    cudaEventSynchronize(ctx->prof_events.back());
    for (size_t i = 0; i + 1 < ctx->prof_events.size(); i += 2) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ctx->prof_events[i], ctx->prof_events[i + 1]) == cudaSuccess) {
            ctx->prof_ms_total += ms;
            ctx->prof_flops_total += ctx->prof_flops[i / 2];
            ctx->prof_launches++;
        }
        cudaEventDestroy(ctx->prof_events[i]);
        cudaEventDestroy(ctx->prof_events[i + 1]);
    }
    ctx->prof_events.clear();
    ctx->prof_flops.clear();
    cudaGetLastError();
}

// 2-D row-major 16-bit tensor [rows, cols] with pitch `ld` elements -> TMA map with a {box_cols, box_rows} box,
// 128-byte swizzle (used by attn_tc.cu).
int mb_encode_2d_map(mb_ctx* ctx, CUtensorMap* m, const void* base, long long cols, long long rows, long long ld,
                     int box_cols, int box_rows) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return mb_set_err(ctx, MB_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = fn(m, ctx->f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                    const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return mb_set_err(ctx, MB_ERR_CUDA, "cuTensorMapEncodeTiled(2d cols=%lld rows=%lld) -> %d", cols, rows, (int)r);
    return 0;
}

// 3-D 16-bit tensor [d2][d1][d0] (d0 contiguous, pitches ld1 / ld2 in elements) -> TMA map with a {b0, b1, 1} box,
// 128-byte swizzle (xattn_tc.cu: encoder states [crops][T][E]; rows beyond d1 are zero-filled, never fetched).
int mb_encode_3d_map(mb_ctx* ctx, CUtensorMap* m, const void* base, long long d0, long long d1, long long d2, long long ld1,
                     long long ld2, int b0, int b1) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return mb_set_err(ctx, MB_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[3] = {(cuuint64_t)d0, (cuuint64_t)d1, (cuuint64_t)d2};
    cuuint64_t strides[2] = {(cuuint64_t)ld1 * 2, (cuuint64_t)ld2 * 2};
    cuuint32_t box[3] = {(cuuint32_t)b0, (cuuint32_t)b1, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = fn(m, ctx->f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3,
                    const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return mb_set_err(ctx, MB_ERR_CUDA, "cuTensorMapEncodeTiled(3d %lld x %lld x %lld) -> %d", d0, d1, d2, (int)r);
    return 0;
}

int mb_tap_gemm(mb_ctx* ctx, const TapGemm& g, cudaStream_t stream) {
    MB_REQUIRE(ctx, g.a0 && g.wgt && g.out, "tap_gemm: null pointer");
    MB_REQUIRE(ctx, g.c0 > 0 && g.c0 % BLOCK_K == 0 && g.c1 >= 0 && g.c1 % BLOCK_K == 0,
               "tap_gemm: channel counts must be multiples of 64 (c0=%d c1=%d)", g.c0, g.c1);
    MB_REQUIRE(ctx, g.taps == 1 || g.taps == 9, "tap_gemm: taps must be 1 or 9");
    MB_REQUIRE(ctx, g.a0_ld % 8 == 0 && (g.c1 == 0 || g.a1_ld % 8 == 0), "tap_gemm: pixel pitch must be a multiple of 8");
    MB_REQUIRE(ctx, (reinterpret_cast<uintptr_t>(g.a0) & 15) == 0 && (reinterpret_cast<uintptr_t>(g.wgt) & 15) == 0 &&
                        (g.c1 == 0 || (reinterpret_cast<uintptr_t>(g.a1) & 15) == 0),
               "tap_gemm: operands must be 16-byte aligned");
    MB_REQUIRE(ctx, g.n_out > 0 && g.n_rows_w >= g.n_out, "tap_gemm: bad n_out/n_rows_w");
    MB_REQUIRE(ctx, g.n > 0 && g.h > 0 && g.w > 0, "tap_gemm: empty problem");

    int block_n = g.block_n;
    if (block_n == 0) {
        if (g.n_out >= 256) block_n = 256;
        else if (g.n_out > 128) block_n = 256;
        else if (g.n_out > 64) block_n = 128;
        else if (g.n_out > 32) block_n = 64;
        else if (g.n_out > 16) block_n = 32;
        else block_n = 16;
    }
    if (g.block_n == 0) {
        // small problems (decoder steps): shrink the N tile until the grid covers the SMs
        const long long m_tiles = (long long)g.n * g.h * mb_cdiv(g.w, BLOCK_M) * (g.batches > 0 ? g.batches : 1);
        while (block_n > 64 && m_tiles * mb_cdiv(g.n_out, block_n) < ctx->num_sms) block_n >>= 1;
    }
    MB_REQUIRE(ctx, block_n % 16 == 0 && block_n >= 16 && block_n <= 256, "tap_gemm: bad block_n %d", block_n);

    KParams p;
    p.n = g.n; p.h = g.h; p.w = g.w;
    p.w_tiles = mb_cdiv(g.w, BLOCK_M);
    p.m_tiles = g.n * g.h * p.w_tiles;
    p.n_tiles = mb_cdiv(g.n_out, block_n);
    p.taps = g.taps; p.dil = g.dil;
    p.kb0 = g.c0 / BLOCK_K; p.kb1 = g.c1 / BLOCK_K;
    p.block_n = block_n;
    p.n_out = g.n_out;
    // TMA-store epilogue: 16-bit row-major output whose pitch / base the TMA can address (and the same for the residual);
    // block-diagonal batches need whole chunks per batch (a box must not spill into the next batch's columns)
    const bool res = g.residual != nullptr;
    const bool lnf = g.ln_stats != nullptr;
    const int nbatch = g.batches > 0 ? g.batches : 1;
    const long long out_cols = (long long)(nbatch - 1) * g.out_col_stride + g.n_out;
    const char* epi_env = getenv("MB_EPI_TMA");
    const bool tma_out = !(epi_env && epi_env[0] == '0') && g.out_mode == MB_OUT_BF16 && g.out_ld % 8 == 0 &&
                         (reinterpret_cast<uintptr_t>(g.out) & 15) == 0 && out_cols >= 32 &&
                         (nbatch == 1 || g.n_out % 32 == 0) &&
                         (!res || (g.res_ld % 8 == 0 && (reinterpret_cast<uintptr_t>(g.residual) & 15) == 0));
    // two-CTA mode (cta_group::2): big 16-bit-output problems with full 256-wide N tiles; MB_GEMM2=0 disables it
    static int gemm2_env = -1;
    if (gemm2_env < 0) { const char* e = getenv("MB_GEMM2"); gemm2_env = (e && e[0] == '1') ? 1 : 0; }
    const bool stats = g.stat_out != nullptr;
    MB_REQUIRE(ctx, !stats || (tma_out && res && !lnf && g.act == MB_ACT_NONE && nbatch == 1 && g.n_out % 32 == 0 && block_n == 256),
               "tap_gemm: row statistics need the TMA epilogue of a residual GEMM with block_n = 256");
    const bool cg2 = gemm2_env == 1 && block_n == 256 && g.out_mode == MB_OUT_BF16 && p.m_tiles >= 2 * ctx->num_sms &&
                     (g.batches <= 1) && (!lnf || tma_out) && !stats;
    // halo mode: 3x3 / dilation 1 convolutions (one A load per image row instead of three); small layers also keep the whole
    // weight matrix in shared memory.  MB_HALO=0 switches it off (A/B timing).
    const char* halo_env = getenv("MB_HALO");
    const int b_slot = (cg2 ? block_n / 2 : block_n) * BLOCK_K * 2;
    const int kb_tap = (g.c0 + g.c1) / BLOCK_K;
    p.halo = (g.taps == 9 && g.dil == 1 && !cg2 && (g.batches <= 1) && !(halo_env && halo_env[0] == '0')) ? 1 : 0;
    const int bres_bytes = 9 * kb_tap * b_slot;
    const char* bres_env = getenv("MB_BRES");
    p.bres = (p.halo && p.n_tiles == 1 && bres_bytes <= 80 * 1024 && !(bres_env && bres_env[0] == '0')) ? 1 : 0;
    const int stage_bytes = p.halo ? A_HALO_STAGE_BYTES + (p.bres ? 0 : 3 * b_slot) : A_STAGE_BYTES + b_slot;
    const int bar_bytes = BAR_BYTES + EPI_WARPS * STAGE_TILE_BYTES;
    // (residual present: the TMA epilogue may need COEF_RES_BYTES behind the staging tiles; reserved whether or not it is taken)
    const int coef_bytes = g.residual != nullptr ? COEF_RES_BYTES : 0;
    int stages = (SMEM_LIMIT - bar_bytes - coef_bytes - (p.bres ? bres_bytes : 0)) / stage_bytes;
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    if (p.halo && stages < 2) {                       // wide N tiles: three B slots per stage do not leave a ring
        p.halo = p.bres = 0;
        stages = (SMEM_LIMIT - bar_bytes - coef_bytes) / (A_STAGE_BYTES + b_slot);
        if (stages > MAX_STAGES) stages = MAX_STAGES;
    }
    const int stage_bytes_final = p.halo ? stage_bytes : A_STAGE_BYTES + b_slot;
    p.stages = stages;
    p.bias = g.bias; p.act = g.act;
    p.residual = g.residual; p.res_ld = g.res_ld;
    p.out = g.out; p.out_ld = g.out_ld; p.out_mode = g.out_mode; p.out_plane = g.out_plane;
    p.f16 = ctx->f16;
    p.batches = g.batches > 0 ? g.batches : 1;
    p.a_col_stride = g.a_col_stride; p.w_row_stride = g.w_row_stride; p.out_col_stride = g.out_col_stride;
    p.ln_stats = reinterpret_cast<const float2*>(g.ln_stats); p.ln_c = g.ln_c;
    p.stat_out = reinterpret_cast<float2*>(g.stat_out);
    p.diag = ctx->dev_diag;
    { const char* e = getenv("MB_SPIN_SLEEP"); p.backoff = e ? (unsigned)atoi(e) : 0u; }
    MB_REQUIRE(ctx, !lnf || (g.ln_c && g.out_mode == MB_OUT_BF16 && !g.residual && g.taps == 1 && g.n == 1 && g.h == 1),
               "tap_gemm: the LayerNorm fold needs a plain 16-bit-output GEMM without residual");

    const bool h = ctx->f16 != 0;
    CUtensorMap tmA0, tmA1, tmB, tmOut, tmRes;
    const int a0c = g.batches > 1 ? g.a0_ld : g.c0;
    const int a_rows = p.halo ? HALO_ROWS : BLOCK_M;
    int rc = cached_map(ctx, &tmA0, MapKey{g.a0, a0c, g.a0_ld, p.halo ? 3 : 0, g.n, g.h, g.w, ctx->f16},
                        [&](CUtensorMap* m) { return encode_act_map(ctx, m, g.a0, a0c, g.a0_ld, g.n, g.h, g.w, ctx->f16, a_rows); });
    if (rc) return rc;
    if (g.c1 > 0) {
        MB_REQUIRE(ctx, g.a1 != nullptr, "tap_gemm: c1>0 but a1 null");
        rc = cached_map(ctx, &tmA1, MapKey{g.a1, g.c1, g.a1_ld, p.halo ? 3 : 0, g.n, g.h, g.w, ctx->f16},
                        [&](CUtensorMap* m) { return encode_act_map(ctx, m, g.a1, g.c1, g.a1_ld, g.n, g.h, g.w, ctx->f16, a_rows); });
        if (rc) return rc;
    } else {
        tmA1 = tmA0;
    }
    const int ktot = g.taps * (g.c0 + g.c1), bn_box = cg2 ? block_n / 2 : block_n;
    rc = cached_map(ctx, &tmB, MapKey{g.wgt, ktot, g.n_rows_w, 1, bn_box, 0, 0, ctx->f16},
                    [&](CUtensorMap* m) { return encode_wgt_map(ctx, m, g.wgt, ktot, g.n_rows_w, bn_box, ctx->f16); });
    if (rc) return rc;
    tmOut = tmB;
    tmRes = tmB;
    if (tma_out) {
        rc = cached_map(ctx, &tmOut, MapKey{g.out, out_cols, g.out_ld, 2, g.n, g.h, g.w, ctx->f16},
                        [&](CUtensorMap* m) { return encode_io_map(ctx, m, g.out, out_cols, g.out_ld, g.n, g.h, g.w, ctx->f16); });
        if (rc) return rc;
        if (res) {
            rc = cached_map(ctx, &tmRes, MapKey{g.residual, out_cols, g.res_ld, 2, g.n, g.h, g.w, ctx->f16}, [&](CUtensorMap* m) {
                return encode_io_map(ctx, m, g.residual, out_cols, g.res_ld, g.n, g.h, g.w, ctx->f16);
            });
            if (rc) return rc;
        }
    }

    const size_t smem = (size_t)stages * stage_bytes_final + (p.bres ? bres_bytes : 0) + bar_bytes + coef_bytes;
    typedef void (*KernelFn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap,
                             const KParams);
    KernelFn fn = nullptr;
#define MB_PICK(A, O, R)                                                                               \
    if (g.act == (A) && g.out_mode == (O) && res == (R))                                               \
        fn = h ? (KernelFn)tap_gemm_kernel<A, O, R, true, false> : (KernelFn)tap_gemm_kernel<A, O, R, false, false>;
#define MB_PICK2(A, R)                                                                                 \
    if (cg2 && g.act == (A) && res == (R))                                                             \
        fn = h ? (KernelFn)tap_gemm_kernel<A, MB_OUT_BF16, R, true, true> : (KernelFn)tap_gemm_kernel<A, MB_OUT_BF16, R, false, true>;
    MB_PICK(MB_ACT_NONE, MB_OUT_BF16, false) MB_PICK(MB_ACT_NONE, MB_OUT_BF16, true)
    MB_PICK(MB_ACT_RELU, MB_OUT_BF16, false) MB_PICK(MB_ACT_RELU, MB_OUT_BF16, true)
    MB_PICK(MB_ACT_GELU, MB_OUT_BF16, false) MB_PICK(MB_ACT_GELU, MB_OUT_BF16, true)
    MB_PICK(MB_ACT_NONE, MB_OUT_F32, false) MB_PICK(MB_ACT_RELU, MB_OUT_F32, false)
    MB_PICK(MB_ACT_NONE, MB_OUT_F32_PLANAR, false) MB_PICK(MB_ACT_RELU, MB_OUT_F32_PLANAR, false)
    MB_PICK2(MB_ACT_NONE, false) MB_PICK2(MB_ACT_NONE, true) MB_PICK2(MB_ACT_RELU, false) MB_PICK2(MB_ACT_RELU, true)
    MB_PICK2(MB_ACT_GELU, false) MB_PICK2(MB_ACT_GELU, true)
#define MB_PICKL(A)                                                                                    \
    if (lnf && g.act == (A))                                                                           \
        fn = h ? (KernelFn)tap_gemm_kernel<A, MB_OUT_BF16, false, true, false, true>                  \
               : (KernelFn)tap_gemm_kernel<A, MB_OUT_BF16, false, false, false, true>;
    MB_PICKL(MB_ACT_NONE) MB_PICKL(MB_ACT_GELU)
    if (lnf && g.act == MB_ACT_RELU) fn = nullptr;
#undef MB_PICKL
    if (tma_out && fn) {
#define MB_PICKT(A, R, L)                                                                                            \
    if (g.act == (A) && res == (R) && lnf == (L))                                                                    \
        fn = cg2 ? (h ? (KernelFn)tap_gemm_kernel<A, MB_OUT_BF16, R, true, true, L, true>                           \
                      : (KernelFn)tap_gemm_kernel<A, MB_OUT_BF16, R, false, true, L, true>)                         \
                 : (h ? (KernelFn)tap_gemm_kernel<A, MB_OUT_BF16, R, true, false, L, true>                          \
                      : (KernelFn)tap_gemm_kernel<A, MB_OUT_BF16, R, false, false, L, true>);
        MB_PICKT(MB_ACT_NONE, false, false) MB_PICKT(MB_ACT_NONE, true, false) MB_PICKT(MB_ACT_RELU, false, false)
        MB_PICKT(MB_ACT_RELU, true, false) MB_PICKT(MB_ACT_GELU, false, false) MB_PICKT(MB_ACT_GELU, true, false)
        MB_PICKT(MB_ACT_NONE, false, true) MB_PICKT(MB_ACT_GELU, false, true)
#undef MB_PICKT
        if (stats)
            fn = h ? (KernelFn)tap_gemm_kernel<MB_ACT_NONE, MB_OUT_BF16, true, true, false, false, true, true>
                   : (KernelFn)tap_gemm_kernel<MB_ACT_NONE, MB_OUT_BF16, true, false, false, false, true, true>;
    }
#undef MB_PICK
#undef MB_PICK2
    if (!fn)
        return mb_set_err(ctx, MB_ERR_ARG, "tap_gemm: unsupported epilogue (act %d, out_mode %d, residual %d)", g.act,
                          g.out_mode, (int)res);
    MB_CUDA(ctx, cudaFuncSetAttribute((const void*)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    const long long total = (long long)p.m_tiles * p.n_tiles * p.batches;
    const int grid = (int)(total < ctx->num_sms ? total : ctx->num_sms);
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if (ctx->profile) {
        cudaEventCreate(&ev0);
        cudaEventCreate(&ev1);
        cudaEventRecord(ev0, stream);
    }
    if (cg2) {
        const long long pairs = (long long)((p.m_tiles + 1) / 2) * p.n_tiles * p.batches;
        const int clusters = (int)(pairs < ctx->num_sms / 2 ? pairs : ctx->num_sms / 2);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * clusters, 1, 1);
        cfg.blockDim = dim3(NUM_THREADS, 1, 1);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        MB_CUDA(ctx, cudaLaunchKernelEx(&cfg, fn, tmA0, tmA1, tmB, tmOut, tmRes, p));
    } else {
        fn<<<grid, NUM_THREADS, smem, stream>>>(tmA0, tmA1, tmB, tmOut, tmRes, p);
    }
    if (ctx->profile) {
        cudaEventRecord(ev1, stream);
        ctx->prof_events.push_back(ev0);
        ctx->prof_events.push_back(ev1);
        // algorithmic work: 2 * pixels * n_out * K (padding rows / zero-filled halo not counted)
        ctx->prof_flops.push_back(2.0 * (double)g.n * g.h * g.w * (double)g.n_out * (double)g.taps * (g.c0 + g.c1));
        if (ctx->prof_events.size() >= 8192) mb_profile_drain(ctx);
    }
    MB_LAUNCH_CHECK(ctx);
    return 0;
}

extern "C" int mb_gemm16(mb_ctx* ctx, const void* a_dev, long long lda, const void* w_dev, int n_rows_w,
                            int M, int N, int K, const float* bias_dev, int act, const void* residual_dev,
                            long long res_ld, void* out_dev, long long out_ld, int out_mode, void* stream) {
    MbDeviceGuard _mb_guard(ctx);
    if (!ctx) return MB_ERR_ARG;
    TapGemm g;
    g.a0 = (const bf16*)a_dev; g.c0 = K; g.a0_ld = (int)lda;
    g.n = 1; g.h = 1; g.w = M;
    g.taps = 1; g.dil = 1;
    g.wgt = (const bf16*)w_dev; g.n_rows_w = n_rows_w; g.n_out = N;
    g.bias = bias_dev; g.act = act;
    g.residual = (const bf16*)residual_dev; g.res_ld = (int)res_ld;
    g.out = out_dev; g.out_ld = out_ld; g.out_mode = out_mode;
    return mb_tap_gemm(ctx, g, (cudaStream_t)stream);
}

// Block-diagonal GEMM: for batch b, out[:, b*out_col_stride + n] = act(A[:, b*a_col_stride : +K] @ W[b*w_row_stride + n, :K]^T
// + bias[b*out_col_stride + n]).  Used for the per-head projections of the decoder's cross-attention (trocr.cu).
extern "C" int mb_gemm16_batched(mb_ctx* ctx, const void* a_dev, long long lda, const void* w_dev, int n_rows_w, int M,
                                 int N, int K, int batches, int a_col_stride, int w_row_stride, int out_col_stride,
                                 const float* bias_dev, int act, void* out_dev, long long out_ld, void* stream) {
    MbDeviceGuard _mb_guard(ctx);
    if (!ctx) return MB_ERR_ARG;
    TapGemm g;
    g.a0 = (const bf16*)a_dev; g.c0 = K; g.a0_ld = (int)lda;
    g.n = 1; g.h = 1; g.w = M;
    g.wgt = (const bf16*)w_dev; g.n_rows_w = n_rows_w; g.n_out = N;
    g.bias = bias_dev; g.act = act;
    g.out = out_dev; g.out_ld = out_ld; g.out_mode = MB_OUT_BF16;
    g.batches = batches; g.a_col_stride = a_col_stride; g.w_row_stride = w_row_stride; g.out_col_stride = out_col_stride;
    return mb_tap_gemm(ctx, g, (cudaStream_t)stream);
}

extern "C" int mb_conv16(mb_ctx* ctx, const void* a0_dev, int c0, int a0_ld, const void* a1_dev, int c1,
                            int a1_ld, int n, int h, int w, int taps, int dil, const void* w_dev,
                            int n_rows_w, int n_out, const float* bias_dev, int act, void* out_dev,
                            long long out_ld, int out_mode, long long out_plane, void* stream) {
    MbDeviceGuard _mb_guard(ctx);
    if (!ctx) return MB_ERR_ARG;
    TapGemm g;
    g.a0 = (const bf16*)a0_dev; g.c0 = c0; g.a0_ld = a0_ld;
    g.a1 = (const bf16*)a1_dev; g.c1 = c1; g.a1_ld = a1_ld;
    g.n = n; g.h = h; g.w = w;
    g.taps = taps; g.dil = dil;
    g.wgt = (const bf16*)w_dev; g.n_rows_w = n_rows_w; g.n_out = n_out;
    g.bias = bias_dev; g.act = act;
    g.out = out_dev; g.out_ld = out_ld; g.out_mode = out_mode; g.out_plane = out_plane;
    return mb_tap_gemm(ctx, g, (cudaStream_t)stream);
}
