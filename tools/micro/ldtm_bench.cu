// tcgen05.ld throughput per SM: W warps (quarter = warp % 4) each issue `iters` 32x32b.x32 loads (4 KB each) back to back.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ldtm_bench ldtm_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),
          "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]),
          "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]),
          "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
}
__global__ void k(uint32_t* out, long long* cyc, int iters, int inflight) {
    __shared__ uint32_t tptr;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&tptr)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = tptr + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t a[32], b[32], acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        ld32(base + (uint32_t)((it & 7) * 32), a);
        if (inflight == 2) ld32(base + (uint32_t)(((it + 3) & 7) * 32), b);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        acc += a[0] ^ a[31];
        if (inflight == 2) acc += b[0] ^ b[31];
    }
    const long long t1 = clock64();
    __syncthreads();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tptr) : "memory");
}
int main() {
    int sms = 148, iters = 2048;
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, 4 * sms * 512); cudaMalloc(&cyc, 8 * sms);
    for (int infl : {1, 2}) for (int warps : {1, 4, 8, 16}) {
        k<<<sms, 32 * warps>>>(out, cyc, iters, infl); cudaDeviceSynchronize();
        k<<<sms, 32 * warps>>>(out, cyc, iters, infl);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        double mx = 0; for (int i = 0; i < sms; ++i) mx = h[i] > mx ? h[i] : mx;
        double bytes = (double)warps * iters * infl * 4096;
        printf("LDTM x32: %2d warps/CTA, %d in flight: %.1f B/clk/SM (%.0f clk per load per warp) %s\n", warps, infl, bytes / mx, mx / iters / infl, cudaGetErrorString(e));
    }
    return 0;
}
