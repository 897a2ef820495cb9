// Data-parallel item-gradient exchange fused with the optimizer, over NVLink multicast (NVLS):
//
//   reduce-scatter  +  TF1 Adam sweep of the rank's row slice  +  all-gather      in ONE kernel.
//
// The item-gradient accumulator G and the item table W of every rank live in symmetric memory bound to one multicast
// object (the caller sets that up: torch.distributed._symmetric_memory in pda_b200/parallel.py).  For its own row slice a
// rank reads the SUM over all ranks' accumulators with multimem.ld_reduce (the NVSwitch adds in flight: the rank receives
// one reduced copy, 1/world of the table), applies the Adam update of MF/model_api.py:83 (AdamOptimizer._apply_sparse_
// shared: the dense form, same separately rounded fp32 operations as adam_dense_kernel) to its slice of W, m, v, and
// writes the new rows to EVERY replica with multimem.st (one store, multicast by the switch).  Per rank and step
// (n_items x d fp32 = S bytes): sent (7/8 + 1/8) S, received (1/8 + 7/8) S, both directions busy at the same time --
// against (7/8 + 7/8) S each way for a reduce-scatter followed by an all-gather.
// Every element is reduced once, by its owner, and broadcast: the replicas stay bit-identical.
// Ordering: the caller brackets the kernel with two cross-rank barriers on the same stream (all step kernels done before
// the first read; all replicas written and all accumulators read before anything overwrites them).
#include <stdlib.h>

#include "../../include/pda_b200.h"
#include "pda_kernels.h"

namespace pda {

__device__ __forceinline__ float4 multimem_ld_reduce_add(const float* mc) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
    return v;
}
__device__ __forceinline__ void multimem_st(float* mc, const float4& v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__device__ __forceinline__ void multimem_st_weak(float* mc, const float4& v) {
    asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 multimem_ld_reduce_add_weak(const float* mc) {
    float4 v;
    asm volatile("multimem.ld_reduce.weak.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
    return v;
}

// ---- cross-rank barriers INSIDE the exchange kernel (flags in symmetric memory; replaces two barrier launches) ----
// flags of a rank: [0..7] start flags written by rank r, [8..15] end flags written by rank r, [16] grid completion counter
// (local), [17] time-out marker.  `epoch` increases by one per launch on every rank.
struct DpBarrier {
    uint32_t* local;        // this rank's flags (nullptr: the caller brackets the kernel with its own barriers)
    uint32_t* peer[8];      // every rank's flags as mapped here, self included
    uint32_t epoch;
    int world, rank;
};
__device__ __forceinline__ void st_release_sys_u32(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t ld_relaxed_sys_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// wait until flag >= epoch (epochs are monotonic, compared as a signed distance); ~2 s time-out marks flags[17] instead of
// hanging.  The polling loads are RELAXED (an acquire load per iteration invalidates the SM's L1 every time round and
// starved the sampler kernel running beside the exchange: 1.38 instead of 0.55 ms); one acquire load ends the wait.
__device__ __forceinline__ void dp_spin(const uint32_t* flag, uint32_t epoch, uint32_t* timeout_mark) {
    const long long t0 = clock64();
    while ((int32_t)(ld_relaxed_sys_u32(flag) - epoch) < 0) {
        if (clock64() - t0 > 4000000000ll) { *timeout_mark = 1u; break; }
        __nanosleep(256);
    }
    (void)ld_acquire_sys_u32(flag);
}
// every rank's previous kernels on this stream (the step kernel) are complete before any CTA here goes on
__device__ __forceinline__ void dp_barrier_enter(const DpBarrier& b) {
    if (!b.local) return;
    if (blockIdx.x == 0 && threadIdx.x < b.world) st_release_sys_u32(b.peer[threadIdx.x] + b.rank, b.epoch);
    if (threadIdx.x == 0)
        for (int r = 0; r < b.world; ++r) dp_spin(b.local + r, b.epoch, b.local + 17);
    __syncthreads();
}
// the LAST CTA of the grid (all others have fenced and left) tells every rank that this rank's writes have landed and its
// reads are done, and leaves only when every rank has said the same -- the kernel's completion is the barrier
__device__ __forceinline__ void dp_barrier_leave(const DpBarrier& b) {
    if (!b.local) return;
    __shared__ int s_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        s_last = atomicAdd(b.local + 16, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    if (threadIdx.x == 0) { b.local[16] = 0u; __threadfence_system(); }
    if (threadIdx.x < b.world) {
        st_release_sys_u32(b.peer[threadIdx.x] + 8 + b.rank, b.epoch);
        dp_spin(b.local + 8 + threadIdx.x, b.epoch, b.local + 17);
    }
}

__device__ __forceinline__ void adam_elem_x(float& w, float& m, float& v, float g, float lr_t) {
    const float omb1 = fsub(1.0f, 0.9f), omb2 = fsub(1.0f, 0.999f);
    m = fadd(fmul(m, 0.9f), fmul(g, omb1));
    v = fadd(fmul(v, 0.999f), fmul(fmul(g, g), omb2));
    w = fsub(w, fdiv(fmul(lr_t, m), fadd(fsqrt(v), 1e-8f)));
}

// mcG / mcW: multicast addresses of the first element of the rank's slice; W, M, V: the local slice; n4 float4s.
// U float4s per thread and iteration: all multimem loads of an iteration are in flight before the first is consumed.
// dbg (tuning only, results are wrong unless 3): bit 0 = read the accumulator through the multicast address (else the
// local replica), bit 1 = write the table through the multicast address (else the local replica).
template <int U>
__global__ void __launch_bounds__(256) dp_exchange_adam_kernel(const float* __restrict__ mcG, float* __restrict__ mcW,
                                                               const float* __restrict__ Gl, float* __restrict__ W, float* __restrict__ M,
                                                               float* __restrict__ V, int64_t n4, const float* __restrict__ pw, float lr,
                                                               int dbg, DpBarrier bar) {
    dp_barrier_enter(bar);
    const float lr_t = fdiv(fmul(lr, fsqrt(fsub(1.0f, pw[1]))), fsub(1.0f, pw[0]));
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * U;
    for (int64_t base = (int64_t)blockIdx.x * blockDim.x * U + threadIdx.x; base < n4; base += stride) {
        float4 g[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t e = base + (int64_t)u * blockDim.x;
            if (e < n4) g[u] = (dbg & 1) ? ((dbg & 8) ? multimem_ld_reduce_add_weak(mcG + 4 * e) : multimem_ld_reduce_add(mcG + 4 * e))
                                         : reinterpret_cast<const float4*>(Gl)[e];
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t e = base + (int64_t)u * blockDim.x;
            if (e >= n4) continue;
            float4 w = reinterpret_cast<const float4*>(W)[e], m = reinterpret_cast<float4*>(M)[e], v = reinterpret_cast<float4*>(V)[e];
            adam_elem_x(w.x, m.x, v.x, g[u].x, lr_t);
            adam_elem_x(w.y, m.y, v.y, g[u].y, lr_t);
            adam_elem_x(w.z, m.z, v.z, g[u].z, lr_t);
            adam_elem_x(w.w, m.w, v.w, g[u].w, lr_t);
            reinterpret_cast<float4*>(M)[e] = m;
            reinterpret_cast<float4*>(V)[e] = v;
            if (dbg & 4) multimem_st_weak(mcW + 4 * e, w);
            else if (dbg & 2) multimem_st(mcW + 4 * e, w);     // every replica, this rank's included
            else reinterpret_cast<float4*>(W)[e] = w;
        }
    }
    dp_barrier_leave(bar);
}

// The same fused exchange over UNICAST peer pointers (symmetric memory, no multicast): the owner of a slice loads the
// partial gradients of all ranks (fixed order r = 0..world-1: deterministic), applies Adam, and stores the new rows into
// every replica.  Same wire volume as the multicast form at 2 ranks (and there the in-switch reduction has nothing to
// add), more from 4 ranks on.
struct PeerPtrs { const float* G[8]; float* W[8]; };

__device__ __forceinline__ float4 ld_sys_f4(const float* p) {
    float4 v;
    asm volatile("ld.global.relaxed.sys.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_sys_f4(float* p, const float4& v) {
    asm volatile("st.global.relaxed.sys.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__device__ __forceinline__ float4 ld_cg_f4(const float* p) {
    float4 v;
    asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_weak_f4(float* p, const float4& v) {
    asm volatile("st.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// mode bit 0: ld.global.cg instead of ld.relaxed.sys; bit 1: plain (weak) stores instead of st.relaxed.sys
template <int U>
__global__ void __launch_bounds__(256) dp_exchange_adam_p2p_kernel(PeerPtrs pp, int world, int self, float* __restrict__ M,
                                                                   float* __restrict__ V, int64_t n4, const float* __restrict__ pw,
                                                                   float lr, int mode, DpBarrier bar) {
    dp_barrier_enter(bar);
    const float lr_t = fdiv(fmul(lr, fsqrt(fsub(1.0f, pw[1]))), fsub(1.0f, pw[0]));
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * U;
    for (int64_t base = (int64_t)blockIdx.x * blockDim.x * U + threadIdx.x; base < n4; base += stride) {
        float4 g[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t e = base + (int64_t)u * blockDim.x;
            g[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (e < n4) {
                g[u] = (mode & 1) ? ld_cg_f4(pp.G[0] + 4 * e) : ld_sys_f4(pp.G[0] + 4 * e);   // 0 + g0 == g0: start from rank 0's partial
                for (int r = 1; r < world; ++r) {
                    const float4 t = (mode & 1) ? ld_cg_f4(pp.G[r] + 4 * e) : ld_sys_f4(pp.G[r] + 4 * e);
                    g[u].x = fadd(g[u].x, t.x); g[u].y = fadd(g[u].y, t.y); g[u].z = fadd(g[u].z, t.z); g[u].w = fadd(g[u].w, t.w);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t e = base + (int64_t)u * blockDim.x;
            if (e >= n4) continue;
            float4 w = reinterpret_cast<const float4*>(pp.W[self])[e], m = reinterpret_cast<float4*>(M)[e], v = reinterpret_cast<float4*>(V)[e];
            adam_elem_x(w.x, m.x, v.x, g[u].x, lr_t);
            adam_elem_x(w.y, m.y, v.y, g[u].y, lr_t);
            adam_elem_x(w.z, m.z, v.z, g[u].z, lr_t);
            adam_elem_x(w.w, m.w, v.w, g[u].w, lr_t);
            reinterpret_cast<float4*>(M)[e] = m;
            reinterpret_cast<float4*>(V)[e] = v;
            for (int r = 0; r < world; ++r) { if (mode & 2) st_weak_f4(pp.W[r] + 4 * e, w); else st_sys_f4(pp.W[r] + 4 * e, w); }
        }
    }
    dp_barrier_leave(bar);
}

static int env_i(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

static DpBarrier make_bar(const DpSync* sy) {
    DpBarrier b;
    b.local = nullptr; b.epoch = 0; b.world = 0; b.rank = 0;
    for (int r = 0; r < 8; ++r) b.peer[r] = nullptr;
    if (sy && sy->local) {
        b.local = sy->local; b.epoch = sy->epoch; b.world = sy->world; b.rank = sy->rank;
        for (int r = 0; r < sy->world && r < 8; ++r) b.peer[r] = sy->peer[r];
    }
    return b;
}

void launch_dp_exchange_adam(const float* mcG, float* mcW, const float* Gl, float* W, float* M, float* V, int64_t n4,
                             const float* pw, float lr, const DpSync* sy, cudaStream_t st) {
    const DpBarrier bar = make_bar(sy);
    const int U = env_i("PDA_DPX_UNROLL", 2), dbg = env_i("PDA_DPX_DBG", 3);
    int64_t blocks = (n4 + 256 * U - 1) / (256 * U);
    const int cap = env_i("PDA_DPX_BLOCKS", 148 * 2);     // measured at 8 GPUs: 148-296 CTAs x unroll 2 -> 1.08 ms, 592 x 4 -> 1.13 ms
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    if (U == 1) dp_exchange_adam_kernel<1><<<(int)blocks, 256, 0, st>>>(mcG, mcW, Gl, W, M, V, n4, pw, lr, dbg, bar);
    else if (U == 2) dp_exchange_adam_kernel<2><<<(int)blocks, 256, 0, st>>>(mcG, mcW, Gl, W, M, V, n4, pw, lr, dbg, bar);
    else if (U == 8) dp_exchange_adam_kernel<8><<<(int)blocks, 256, 0, st>>>(mcG, mcW, Gl, W, M, V, n4, pw, lr, dbg, bar);
    else dp_exchange_adam_kernel<4><<<(int)blocks, 256, 0, st>>>(mcG, mcW, Gl, W, M, V, n4, pw, lr, dbg, bar);
}

void launch_dp_exchange_adam_p2p(const float* const* G, float* const* W, int world, int self, int64_t off, float* M, float* V,
                                 int64_t n4, const float* pw, float lr, const DpSync* sy, cudaStream_t st) {
    const DpBarrier bar = make_bar(sy);
    PeerPtrs pp;
    for (int r = 0; r < 8; ++r) { pp.G[r] = r < world ? G[r] + off : nullptr; pp.W[r] = r < world ? W[r] + off : nullptr; }
    const int U = env_i("PDA_DPX_UNROLL", 2), mode = env_i("PDA_DPX_P2P_MODE", 0);
    int64_t blocks = (n4 + 256 * U - 1) / (256 * U);
    // measured at 2 GPUs (256 MB slice): 148 CTAs 1.63 ms, 592 -> 0.83-0.88, 1184 x unroll 2 -> 0.785; ld/st flavours equal
    const int cap = env_i("PDA_DPX_BLOCKS", 148 * 8);
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    if (U == 1) dp_exchange_adam_p2p_kernel<1><<<(int)blocks, 256, 0, st>>>(pp, world, self, M, V, n4, pw, lr, mode, bar);
    else if (U == 4) dp_exchange_adam_p2p_kernel<4><<<(int)blocks, 256, 0, st>>>(pp, world, self, M, V, n4, pw, lr, mode, bar);
    else if (U == 8) dp_exchange_adam_p2p_kernel<8><<<(int)blocks, 256, 0, st>>>(pp, world, self, M, V, n4, pw, lr, mode, bar);
    else dp_exchange_adam_p2p_kernel<2><<<(int)blocks, 256, 0, st>>>(pp, world, self, M, V, n4, pw, lr, mode, bar);
}

}  // namespace pda
