// Memory skeleton of the fused BPR step (no arithmetic): per triple, read the user row's (w, m, v) and two item rows,
// write (w, m, v) back and reduce two rows into the item-gradient accumulator -- the traffic of bpr_step_pipe_kernel.
// Layout A: w, m, v in three arrays (512 B each);  layout B: one interleaved user array [n_users][3][128] (1536 B).
// Answers: what is the memory-only floor of the step kernel, and does a 3x longer contiguous user record help DRAM?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_bw gather_bw.cu && ./gather_bw
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred P1;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, 0x989680;\n\t@P1 bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, 0xffffffff;\n\t@px mov.s32 %0, 1;\n\t}" : "+r"(pred));
    return pred != 0;
}
__device__ __forceinline__ float4 lds_f4(uint32_t a) { float4 v; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a)); return v; }
__device__ __forceinline__ void red4(float* p, float4 v) { asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory"); }

// INTER = 0: U, M, V separate;  1: interleaved (U points at the [n][3][128] array).  MODE bit 0: do the stores, bit 1: do the reds
template <int INTER, int D, int NW>
__global__ void __launch_bounds__(NW * 32) skel(float* U, float* M, float* V, const float* I, float* GI, const int* iu, const int* ip,
                                                  const int* in, int64_t B, int mode, int hot) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int STAGE = 5 * 512;
    const uint32_t rows0 = smem_u32(smem) + warp * D * STAGE, bars0 = smem_u32(smem) + NW * D * STAGE + warp * D * 8;
    if (lane == 0) for (int s = 0; s < D; ++s) mbar_init(bars0 + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const int64_t nw = (int64_t)gridDim.x * NW, w = (int64_t)blockIdx.x * NW + warp;
    const int64_t per = (B + nw - 1) / nw, t0 = w * per, t1 = t0 + per < B ? t0 + per : B;
    auto issue = [&](int64_t t, int s) {
        if (t >= t1) return;
        const int64_t u = iu[t], p = ip[t], n = in[t];
        if (elect_one()) {
            const uint32_t bar = bars0 + 8 * s, dst = rows0 + s * STAGE;
            mbar_expect_tx(bar, STAGE);
            if (INTER) bulk(dst, U + u * 384, 1536, bar);
            else { bulk(dst, U + u * 128, 512, bar); bulk(dst + 512, M + u * 128, 512, bar); bulk(dst + 1024, V + u * 128, 512, bar); }
            bulk(dst + 1536, I + p * 128, 512, bar); bulk(dst + 2048, I + n * 128, 512, bar);
        }
    };
    for (int i = 0; i < D; ++i) issue(t0 + i, i);
    int s = 0; uint32_t par = 0;
    float acc = 0.f;
    for (int64_t t = t0; t < t1; ++t) {
        mbar_wait(bars0 + 8 * s, par);
        const uint32_t st = rows0 + s * STAGE + lane * 16;
        float4 a = lds_f4(st), b = lds_f4(st + 512), c = lds_f4(st + 1024), p = lds_f4(st + 1536), n = lds_f4(st + 2048);
        __syncwarp();
        issue(t + D, s);
        if (++s == D) { s = 0; par ^= 1; }
        const int64_t u = iu[t], pi = ip[t], ni = in[t];
        a.x += p.x; b.x += n.x; c.x += 1.0f; acc += a.y + b.y + c.y;
        if (mode & 1) {
            if (INTER) {
                float* r = U + u * 384 + lane * 4;
                *reinterpret_cast<float4*>(r) = a; *reinterpret_cast<float4*>(r + 128) = b; *reinterpret_cast<float4*>(r + 256) = c;
            } else {
                *reinterpret_cast<float4*>(U + u * 128 + lane * 4) = a; *reinterpret_cast<float4*>(M + u * 128 + lane * 4) = b;
                *reinterpret_cast<float4*>(V + u * 128 + lane * 4) = c;
            }
        }
        if ((mode & 2) && !(mode & 4)) { if (pi >= hot) red4(GI + pi * 128 + lane * 4, p); red4(GI + ni * 128 + lane * 4, n); }   // hot: the `hot` most popular items' reductions are dropped (= an ideal privatised accumulator)
        if (mode & (4 | 8)) {
            // results back into this warp's out-staging rows (a second 2560 B area per warp), then bulk copies / bulk reduces
            const uint32_t ob = smem_u32(smem) + NW * D * STAGE + NW * D * 8 + 64 + warp * STAGE;
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");      // the previous triple's bulk ops have read the area
            __syncwarp();
            auto sts = [&](uint32_t ad, float4 v) { asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(ad), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory"); };
            sts(ob + lane * 16, a); sts(ob + 512 + lane * 16, b); sts(ob + 1024 + lane * 16, c); sts(ob + 1536 + lane * 16, p); sts(ob + 2048 + lane * 16, n);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (elect_one()) {
                if (mode & 8) {
                    if (INTER) asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 1536;" ::"l"(U + u * 384), "r"(ob) : "memory");
                    else {
                        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 512;" ::"l"(U + u * 128), "r"(ob) : "memory");
                        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 512;" ::"l"(M + u * 128), "r"(ob + 512) : "memory");
                        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 512;" ::"l"(V + u * 128), "r"(ob + 1024) : "memory");
                    }
                }
                if (mode & 4) {
                    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], 512;" ::"l"(GI + pi * 128), "r"(ob + 1536) : "memory");
                    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], 512;" ::"l"(GI + ni * 128), "r"(ob + 2048) : "memory");
                }
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    if (acc == 123.456f) U[0] = acc;
}

static uint64_t rng = 88172645463325252ull;
static uint32_t xr() { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return (uint32_t)(rng >> 16); }

template <int INTER, int D, int NW>
float run(float* U, float* M, float* V, float* I, float* GI, int* iu, int* ip, int* in, int64_t B, int mode, int ctas_per_sm, int hot = 0) {
    const size_t smem = (size_t)NW * D * 2560 + NW * D * 8 + 64 + (size_t)NW * 2560;
    cudaFuncSetAttribute(skel<INTER, D, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int it = 0; it < 4; ++it) {
        cudaEventRecord(e0);
        skel<INTER, D, NW><<<148 * ctas_per_sm, NW * 32, smem>>>(U, M, V, I, GI, iu, ip, in, B, mode, hot);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (it && ms < best) best = ms;
    }
    if (cudaGetLastError() != cudaSuccess) printf("CUDA error\n");
    return best;
}

int main(int argc, char** argv) {
    const int64_t nU = argc > 1 ? atoll(argv[1]) : 10000000, nI = 1000000, B = 1 << 20;
    float *U, *I, *GI; int *iu, *ip, *in;
    cudaMalloc(&U, (size_t)nU * 384 * 4); cudaMalloc(&I, (size_t)nI * 512); cudaMalloc(&GI, (size_t)nI * 512);
    cudaMemset(U, 0, (size_t)nU * 384 * 4); cudaMemset(I, 0, (size_t)nI * 512); cudaMemset(GI, 0, (size_t)nI * 512);
    float* M = U + (size_t)nU * 128; float* V = U + (size_t)nU * 256;
    int *hu = (int*)malloc(B * 4), *hp = (int*)malloc(B * 4), *hn = (int*)malloc(B * 4);
    // distinct users (stride walk), Zipf-like items (log-uniform rank) like the bench set
    for (int64_t i = 0; i < B; ++i) {
        hu[i] = (int)((i * 7919ull * 1013ull + xr() % 9) % nU);
        double r = (xr() % 1000000) / 1e6; hp[i] = (int)(exp(r * log((double)nI + 1)) - 1) % nI; hn[i] = xr() % nI;
    }
    if (argc > 2 && atoi(argv[2]) == 1) {      // batch sorted by pos item (the order of triples inside a batch is free)
        struct T { int u, p, n; };
        T* t = (T*)malloc(B * sizeof(T));
        for (int64_t i = 0; i < B; ++i) t[i] = {hu[i], hp[i], hn[i]};
        qsort(t, B, sizeof(T), [](const void* a, const void* b) { return ((const T*)a)->p - ((const T*)b)->p; });
        for (int64_t i = 0; i < B; ++i) { hu[i] = t[i].u; hp[i] = t[i].p; hn[i] = t[i].n; }
        printf("batch sorted by pos item\n");
    }
    const int dist = argc > 2 ? atoi(argv[2]) : 0;
    if (dist == 2) { for (int64_t i = 0; i < B; ++i) hp[i] = xr() % nI; printf("uniform positives\n"); }
    cudaMalloc(&iu, B * 4); cudaMalloc(&ip, B * 4); cudaMalloc(&in, B * 4);
    cudaMemcpy(iu, hu, B * 4, cudaMemcpyHostToDevice); cudaMemcpy(ip, hp, B * 4, cudaMemcpyHostToDevice); cudaMemcpy(in, hn, B * 4, cudaMemcpyHostToDevice);
    const double gb = B * (40.0 * 128 + 20) / 1e9;
    printf("B = 2^20 triples, %lld users, 1M items; algorithmic %.2f GB per launch (40d+20 per triple)\n", (long long)nU, gb);
    if (dist >= 2) {   // how much of the reduction cost is same-row contention on the popular items?
        const int hots[5] = {0, 32, 256, 4096, 1 << 30};
        for (int k = 0; k < 5; ++k) {
            long long kept = 0;
            for (int64_t i = 0; i < B; ++i) kept += hp[i] >= hots[k];
            printf("reds of positives with rank >= %d only (%lld of 2^20 kept): reads+reds %.3f ms, reads+stores+reds %.3f ms\n", hots[k], kept,
                   run<0, 3, 8>(U, M, V, I, GI, iu, ip, in, B, 2, 3, hots[k]), run<0, 3, 8>(U, M, V, I, GI, iu, ip, in, B, 3, 3, hots[k]));
        }
        return 0;
    }
    {
        // bulk variants (D = 2 stages + the out-staging area, 3 CTAs per SM): 1|4 = plain stores + bulk reduces, 8|2 = bulk stores + red.v4, 8|4 = both bulk
        const int modes[4] = {4, 1 | 4, 8 | 2, 8 | 4};
        for (int k = 0; k < 4; ++k) {
            float a = run<0, 2, 8>(U, M, V, I, GI, iu, ip, in, B, modes[k], 3), b = run<1, 2, 8>(U, M, V, I, GI, iu, ip, in, B, modes[k], 3);
            printf("bulk mode %2d (%s): separate %.3f ms  interleaved %.3f ms\n", modes[k],
                   k == 0 ? "bulk reduces only" : k == 1 ? "plain stores + bulk reduces" : k == 2 ? "bulk stores + red.v4" : "bulk stores + bulk reduces", a, b);
        }
    }
    for (int mode = 0; mode < 4; ++mode) {
        const double g = B * ((mode & 1 ? 40.0 : 28.0) - (mode & 2 ? 0 : 8.0)) * 128 / 1e9;
        float a = run<0, 3, 8>(U, M, V, I, GI, iu, ip, in, B, mode, 3), b = run<1, 3, 8>(U, M, V, I, GI, iu, ip, in, B, mode, 3);
        float c = run<0, 4, 8>(U, M, V, I, GI, iu, ip, in, B, mode, 2), d = run<1, 4, 8>(U, M, V, I, GI, iu, ip, in, B, mode, 2);
        float e = run<0, 2, 8>(U, M, V, I, GI, iu, ip, in, B, mode, 4), f = run<1, 2, 8>(U, M, V, I, GI, iu, ip, in, B, mode, 4);
        printf("mode %d (stores %d, reds %d) bytes %.2f GB | D3x3: separate %.3f ms (%.0f GB/s)  interleaved %.3f ms (%.0f GB/s) | D4x2: %.3f / %.3f | D2x4: %.3f / %.3f\n",
               mode, mode & 1, (mode >> 1) & 1, g, a, g / a * 1e3, b, g / b * 1e3, c, d, e, f);
    }
    return 0;
}
