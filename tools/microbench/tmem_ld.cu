// Microbenchmark: tcgen05.ld (TMEM -> registers) throughput per SM, and the ALU cost of a max3 reduction over the loaded
// values, as a function of the number of reading warps.  Sizes the epilogue of pda_eval_tc.cu's sweep kernel.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_ld tmem_ld.cu ; run: ./tmem_ld
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// MODE 0: loads only (double-buffered, one wait per load); MODE 1: + max3 over the 32 values and a rare-hit branch;
// MODE 2: + FSET/FFMA hit mask (the r01 epilogue)
template <int MODE>
__global__ void __launch_bounds__(512, 1) k_ld(int iters, float thr, float* out, long long* cycles) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot + ((uint32_t)(32 * (warp & 3)) << 16);
    float acc = -1e30f;
    int hits = 0;
    uint32_t va[32], vb[32];
    const long long t0 = clock64();
    tmem_ld32(base, va);
    for (int i = 0; i < iters; ++i) {
        uint32_t* v = (i & 1) ? vb : va;
        tmem_ld_wait();
        tmem_ld32(base + (uint32_t)(((i + 1) * 32) & 511), (i & 1) ? va : vb);
        if (MODE == 1) {
            float m = -1e30f;
#pragma unroll
            for (int c = 0; c < 32; c += 2) m = fmaxf(m, fmaxf(__uint_as_float(v[c]), __uint_as_float(v[c + 1])));
            if (m >= thr) { hits += __popc(__float_as_uint(m)); }
            acc = fmaxf(acc, m);
        } else if (MODE == 2) {
            float hlo = 0.f, hhi = 0.f;
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                float f;
                asm("set.ge.f32.f32 %0, %1, %2;" : "=f"(f) : "f"(__uint_as_float(v[c])), "f"(thr));
                if (c < 16) hlo = fmaf(f, (float)(1u << (c & 15)), hlo);
                else hhi = fmaf(f, (float)(1u << (c & 15)), hhi);
            }
            hits += (int)hlo + (int)hhi;
        } else {
            acc = fmaxf(acc, __uint_as_float(v[lane & 31 ? 0 : 1]));
        }
    }
    tmem_ld_wait();
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (acc == 12345.f || hits == -7) out[threadIdx.x] = acc + hits;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512) : "memory");
}

template <int MODE>
static void run(const char* name, int warps, int iters) {
    float* out; long long* cyc;
    cudaMalloc(&out, 4096); cudaMalloc(&cyc, 148 * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_ld<MODE><<<148, warps * 32>>>(iters, 3e38f, out, cyc);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k_ld<MODE><<<148, warps * 32>>>(iters, 3e38f, out, cyc);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
    const double bytes_sm = (double)warps * iters * 4096.0;
    printf("%-10s warps=%2d iters=%d  %.3f ms  cycles/CTA=%.0f  TMEM read = %.1f B/cyc/SM  (%.1f cyc per warp-ld; %.2f cyc per 128x256 fp32 tile)  %s\n",
           name, warps, iters, ms, c, bytes_sm / c, c / iters, 131072.0 / (bytes_sm / c), cudaGetErrorString(err));
    cudaFree(out); cudaFree(cyc);
}

int main() {
    const int it = 20000;
    for (int w : {1, 4, 8, 16}) run<0>("ld-only", w, it);
    for (int w : {4, 8, 16}) run<1>("ld+max3", w, it);
    for (int w : {4, 8, 16}) run<2>("ld+fset", w, it);
    return 0;
}
