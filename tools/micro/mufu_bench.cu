// MUFU.EX2 / FFMA2 / mixed throughput per SM on this GPU (cycles from clock64 inside the kernel).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_bench mufu_bench.cu ; run: ./mufu_bench
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
    float a[8];
    for (int i = 0; i < 8; ++i) a[i] = -0.001f * (threadIdx.x + i);
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) a[i] = ex2(a[i]);
            if (MODE == 1) { a[i] = ex2(a[i]); a[i] = fmaf(a[i], -0.5f, -0.25f); }
            if (MODE == 2) a[i] = fmaf(a[i], 0.999f, -0.25f);
        }
    }
    const long long t1 = clock64();
    float s = 0; for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE> void run(const char* name, int warps_per_sm) {
    int sms = 148, threads = 32 * warps_per_sm, iters = 4096;
    float* out; long long* cyc;
    cudaMalloc(&out, sizeof(float) * sms * threads); cudaMalloc(&cyc, sizeof(long long) * sms);
    k<MODE><<<sms, threads>>>(out, cyc, iters); cudaDeviceSynchronize();
    k<MODE><<<sms, threads>>>(out, cyc, iters); cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double mx = 0; for (int i = 0; i < sms; ++i) mx = h[i] > mx ? h[i] : mx;
    double ops = (double)threads * iters * 8;
    printf("%-28s warps/SM %2d: %.2f ops/clk/SM\n", name, warps_per_sm, ops / mx);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    for (int w : {4, 8, 16, 32}) run<0>("MUFU.EX2", w);
    for (int w : {8, 16, 32}) run<1>("MUFU.EX2 + FFMA (pairs)", w);
    for (int w : {8, 16, 32}) run<2>("FFMA", w);
    return 0;
}
