// Diagnostics behind the bit-exactness claims of the exact Adam replay (pda_common.cuh): the straight-line
// MUFU + FFMA refinements (sqrt_rn_inrange, div_rn_inrange, the packed zero_grad_pair / lazy_grad_step4) must
// return the bits of __fsqrt_rn / __fdiv_rn -- the operations TF1's Adam sweep (MF/model_api.py:83 ->
// AdamOptimizer._apply_sparse_shared) is restated with -- for EVERY operand inside the guarded ranges.
// These entry points sweep / sample those ranges on the device and count mismatches; tests/test_gpu_numerics.py
// drives them.  Not on the product path.
#include "../../include/pda_b200.h"
#include "pda_kernels.h"

using namespace pda;

namespace {

struct DbgOut {
    unsigned long long checked, mismatches;
    uint32_t first_a, first_b, first_c, pad;
};

__device__ __forceinline__ void report(DbgOut* o, unsigned long long n, unsigned long long bad, uint32_t a, uint32_t b, uint32_t c) {
    // warp-aggregated
    for (int off = 16; off >= 1; off >>= 1) {
        n += __shfl_xor_sync(0xffffffffu, n, off);
        bad += __shfl_xor_sync(0xffffffffu, bad, off);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&o->checked, n);
        if (bad) atomicAdd(&o->mismatches, bad);
    }
    (void)a; (void)b; (void)c;
}

__device__ __forceinline__ void note_first(DbgOut* o, uint32_t a, uint32_t b, uint32_t c) {
    if (atomicCAS(&o->pad, 0u, 1u) == 0u) { o->first_a = a; o->first_b = b; o->first_c = c; }
}

// every fp32 bit pattern in [lo, hi]
__global__ void __launch_bounds__(256) sqrt_sweep_kernel(uint32_t lo, uint32_t hi, DbgOut* o) {
    unsigned long long n = 0, bad = 0;
    const uint64_t total = (uint64_t)hi - lo + 1;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t bits = lo + (uint32_t)i;
        const float x = __uint_as_float(bits);
        const float got = sqrt_rn_inrange(x), want = __fsqrt_rn(x);
        ++n;
        if (__float_as_uint(got) != __float_as_uint(want)) { ++bad; note_first(o, bits, __float_as_uint(got), __float_as_uint(want)); }
    }
    report(o, n, bad, 0, 0, 0);
}

// a float with a uniformly random mantissa and an exponent uniform in [e_lo, e_hi] (unbiased), random sign if `sgn`
__device__ __forceinline__ float rnd_float(uint32_t r_m, uint32_t r_e, int e_lo, int e_hi, bool sgn) {
    const int e = e_lo + (int)mulhi32(r_e, (uint32_t)(e_hi - e_lo + 1));
    uint32_t b = ((uint32_t)(e + 127) << 23) | (r_m & 0x7fffffu);
    if (sgn && (r_m & 0x800000u)) b |= 0x80000000u;
    return __uint_as_float(b);
}

// kind 0: div_rn_inrange(a, b) vs __fdiv_rn, |a| in [2^-100, 2^60], b in [2^-27, 2^21] (the divisor sqrt(v) + eps)
// kind 1: one zero-gradient step of four elements, unguarded packed form vs the generic intrinsics (w, m, v, lr in range)
// kind 2: one gradient step of four elements, lazy_grad_step4 vs lazy_grad_step
// kind 4: three zero-gradient steps on the negated-v state (zero_grad_step4_nv, as lazy_replay4_warp runs them) vs the generic intrinsics
__global__ void __launch_bounds__(256) random_check_kernel(int kind, uint32_t seed, uint64_t per_thread, DbgOut* o) {
    unsigned long long n = 0, bad = 0;
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    for (uint64_t it = 0; it < per_thread; ++it) {
        const u32x4 r0 = philox4x32(tid, (uint32_t)it, (uint32_t)(it >> 32), 0u, seed, 0xD1A60001u);
        const u32x4 r1 = philox4x32(tid, (uint32_t)it, (uint32_t)(it >> 32), 1u, seed, 0xD1A60001u);
        if (kind == 0) {
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const float a = rnd_float(r0.w[2 * k], r1.w[2 * k], -100, 59, true);
                const float b = rnd_float(r0.w[2 * k + 1], r1.w[2 * k + 1], -27, 20, false);
                const float got = div_rn_inrange(a, b), want = __fdiv_rn(a, b);
                ++n;
                if (__float_as_uint(got) != __float_as_uint(want)) { ++bad; note_first(o, __float_as_uint(a), __float_as_uint(b), __float_as_uint(got)); }
            }
        } else {
            const u32x4 r2 = philox4x32(tid, (uint32_t)it, (uint32_t)(it >> 32), 2u, seed, 0xD1A60001u);
            const u32x4 r3 = philox4x32(tid, (uint32_t)it, (uint32_t)(it >> 32), 3u, seed, 0xD1A60001u);
            // one exponent window per float4 (like a real row: neighbouring elements have similar scale), random offsets inside
            const int ev = -98 + (int)mulhi32(r3.w[0], 130u), em = -59 + (int)mulhi32(r3.w[1], 90u);
            float4 w, m, v, g;
            w.x = rnd_float(r0.w[0], r3.w[2], -20, 4, true); w.y = rnd_float(r0.w[1], r3.w[2] * 3u, -20, 4, true);
            w.z = rnd_float(r0.w[2], r3.w[2] * 5u, -20, 4, true); w.w = rnd_float(r0.w[3], r3.w[2] * 7u, -20, 4, true);
            m.x = rnd_float(r1.w[0], r3.w[3], em, em + 6, true); m.y = rnd_float(r1.w[1], r3.w[3] * 3u, em, em + 6, true);
            m.z = rnd_float(r1.w[2], r3.w[3] * 5u, em, em + 6, true); m.w = rnd_float(r1.w[3], r3.w[3] * 7u, em, em + 6, true);
            v.x = rnd_float(r2.w[0], r3.w[1] * 3u, ev, ev + 6, false); v.y = rnd_float(r2.w[1], r3.w[1] * 5u, ev, ev + 6, false);
            v.z = rnd_float(r2.w[2], r3.w[1] * 7u, ev, ev + 6, false); v.w = rnd_float(r2.w[3], r3.w[1] * 9u, ev, ev + 6, false);
            const float lr = rnd_float(r3.w[0] * 11u, r3.w[1] * 13u, -28, 8, false);
            float4 w2 = w, m2 = m, v2 = v;
            if (kind == 4) {
                if (!replay_block_in_range(m, v)) continue;
                Row4 r = row4_pack(w, m, v);
                for (int k = 0; k < 3; ++k) {
                    zero_grad_step4_nv(r, lr);
                    lazy_zero_grad_step(w2.x, m2.x, v2.x, lr); lazy_zero_grad_step(w2.y, m2.y, v2.y, lr);
                    lazy_zero_grad_step(w2.z, m2.z, v2.z, lr); lazy_zero_grad_step(w2.w, m2.w, v2.w, lr);
                }
                row4_unpack(r, w, m, v);
            } else if (kind == 1) {
                if (!replay_block_in_range(m, v)) continue;
                zero_grad_step4_unguarded(w, m, v, lr);
                lazy_zero_grad_step(w2.x, m2.x, v2.x, lr); lazy_zero_grad_step(w2.y, m2.y, v2.y, lr);
                lazy_zero_grad_step(w2.z, m2.z, v2.z, lr); lazy_zero_grad_step(w2.w, m2.w, v2.w, lr);
            } else {
                g.x = rnd_float(r2.w[0] * 3u, r3.w[2] * 11u, em - 4, em + 8, true); g.y = rnd_float(r2.w[1] * 3u, r3.w[2] * 13u, em - 4, em + 8, true);
                g.z = rnd_float(r2.w[2] * 3u, r3.w[2] * 17u, em - 4, em + 8, true); g.w = rnd_float(r2.w[3] * 3u, r3.w[2] * 19u, em - 4, em + 8, true);
                lazy_grad_step4(w, m, v, g, lr);
                lazy_grad_step(w2.x, m2.x, v2.x, g.x, lr); lazy_grad_step(w2.y, m2.y, v2.y, g.y, lr);
                lazy_grad_step(w2.z, m2.z, v2.z, g.z, lr); lazy_grad_step(w2.w, m2.w, v2.w, g.w, lr);
            }
            const float a4[12] = {w.x, w.y, w.z, w.w, m.x, m.y, m.z, m.w, v.x, v.y, v.z, v.w};
            const float b4[12] = {w2.x, w2.y, w2.z, w2.w, m2.x, m2.y, m2.z, m2.w, v2.x, v2.y, v2.z, v2.w};
            n += 4;
#pragma unroll
            for (int k = 0; k < 12; ++k)
                if (__float_as_uint(a4[k]) != __float_as_uint(b4[k])) { ++bad; note_first(o, (uint32_t)k, __float_as_uint(a4[k]), __float_as_uint(b4[k])); }
        }
    }
    report(o, n, bad, 0, 0, 0);
}

}  // namespace

extern "C" int pda_debug_numerics(int kind, uint32_t lo_or_seed, uint32_t hi, uint64_t per_thread, uint64_t* out5) {
    if (!out5) return PDA_ERR_ARG;
    DbgOut* d = nullptr;
    if (cudaMalloc((void**)&d, sizeof(DbgOut)) != cudaSuccess) return PDA_ERR_CUDA;
    cudaMemset(d, 0, sizeof(DbgOut));
    if (kind == 3) sqrt_sweep_kernel<<<148 * 8, 256>>>(lo_or_seed, hi, d);
    else if ((kind >= 0 && kind <= 2) || kind == 4) random_check_kernel<<<148 * 8, 256>>>(kind, lo_or_seed, per_thread, d);
    else { cudaFree(d); return PDA_ERR_ARG; }
    DbgOut h;
    cudaError_t e = cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return PDA_ERR_CUDA;
    out5[0] = h.checked; out5[1] = h.mismatches; out5[2] = h.first_a; out5[3] = h.first_b; out5[4] = h.first_c;
    return PDA_OK;
}
