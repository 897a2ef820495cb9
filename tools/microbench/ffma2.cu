// Microbenchmark: issue rate of the packed fp32 instructions of sm_100 (FFMA2 / FMUL2: two fp32 operations per
// instruction) against scalar FFMA, per SM.  Sizes the exact lazy-Adam replay of bpr_step_kernel (instruction-bound).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2 ffma2.cu ; run: ./ffma2
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long pack(float a, float b) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}

template <int MODE>
__global__ void __launch_bounds__(1024, 1) k(float* out, int iters, long long* cyc) {
    float a[8], b = 1.0001f + threadIdx.x * 1e-9f, c = 1e-7f;
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = 1.0f + i + threadIdx.x * 1e-6f;
    unsigned long long pa[4], pb = pack(b, b), pc = pack(c, c);
#pragma unroll
    for (int i = 0; i < 4; ++i) pa[i] = pack(a[2 * i], a[2 * i + 1]);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], b, c);              // 8 scalar FFMA = 8 fp32 fma
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(pa[i]) : "l"(pb), "l"(pc));   // 4 FFMA2 = 8 fp32 fma
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
#pragma unroll
    for (int i = 0; i < 4; ++i) s += (float)(pa[i] & 0xffff);
    if (s == 12345.678f) out[threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
static void run(const char* name, int threads) {
    float* out; long long* cyc;
    cudaMalloc(&out, 4096 * 4); cudaMalloc(&cyc, 148 * 8);
    const int iters = 20000;
    k<MODE><<<148, threads>>>(out, iters, cyc);
    cudaDeviceSynchronize();
    k<MODE><<<148, threads>>>(out, iters, cyc);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
    const double fma_per_sm = (double)threads * iters * 8;
    printf("%-8s threads/SM=%4d  cycles=%.0f  fp32 fma per cycle per SM = %.1f  (warp-instr per cycle per SM = %.2f)\n", name, threads, c,
           fma_per_sm / c, fma_per_sm / c / 32 / (MODE == 0 ? 1 : 2));
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int t : {128, 256, 512, 1024}) run<0>("FFMA", t);
    for (int t : {128, 256, 512, 1024}) run<1>("FFMA2", t);
    return 0;
}
