// Tensor-core (tcgen05 + TMEM + TMA) candidate filter for the all-items recommender, with exact rescoring.
//
// Reference call sites: MF/train_new_api.py:594-640 (U_b I^T, (elu+1)*pop, -inf mask, top_k(50)),
// MF/model_api.py:62,113.
//
// bf16 tensor-core scores cannot give fp32-exact top-K ids by themselves, so this path is a FILTER with a
// certificate; every id and score that leaves the library is computed by the exact fp32 spec
// (sequential-k accumulate, spec_expf), identical to pda_eval_exact.cu and the oracle.
//
//   prep     I, U[users] -> bf16 copies (rows padded to the tile) + row norms; per-tile max item norm
//   pass A   tcgen05 sweep over every `se`-th item tile: per (row, 32-column chunk) the maximum of a LOWER bound of
//            the transformed score, with the column packed into the low mantissa bits        -> cmax[chunk][row]
//   select   per row: drop chunk maxima that are train items (mask), tau = K-th largest of the rest.  K distinct
//            unmasked items have exact score >= tau, so the exact K-th best is >= tau.
//   pass B   tcgen05 sweep over ALL item tiles: items whose UPPER bound reaches tau -> cand[row][...]
//   rescore  per row: exact fp32 score of every candidate, transform, mask, sorted top-K.  The row is CERTIFIED when
//            the candidate buffer did not overflow and at least K unmasked candidates have exact score >= tau
//            (then every member of the exact top-K has upper bound >= its score >= K-th best >= tau, i.e. is a candidate).
//   fallback rows that are not certified are recomputed by recommend_exact_kernel (count read on the device, no host sync).
//
// Bound: s = sum u_k i_k, s_lp = fp32-accumulated sum bf16(u_k) bf16(i_k):  |s - s_lp| <= c * |u| * |i|,
// c = 1.02 * 2^-8 + d * 2^-21 (two RN roundings to 8 significant bits per product, accumulation slack, and the
// distance between the sequential-k fp32 spec and the real-number dot).  f(x) = elu(x)+1 is increasing with
// max(x+1, 0) <= f(x) <= max(x+1, 1), so no exp is needed in the sweeps.
//
// Sweep kernel: one CTA = 128 users (UMMA M) x a range of 256-item tiles (UMMA N).  Warp 0 lane 0 issues TMA
// (128B-swizzled K-major tiles: A once, B through a ring), warp 1 lane 0 issues tcgen05.mma (kind::f16, bf16 -> fp32)
// into two 256-column TMEM accumulator stages, warps 4-11 are the epilogue: tcgen05.ld 32x32b.x32 (a thread = one
// user row x 32 items), bound, compare / max.  The score matrix never leaves TMEM.
#include <cuda.h>
#include <cuda_bf16.h>

#include "pda_kernels.h"

namespace pda {

namespace tc {

constexpr int TM = 128, TN = 256, KB = 64;          // UMMA tile, K block (64 bf16 = one 128 B swizzle row)
constexpr int MR = 2;                               // user tiles resident per CTA: every B tile from L2 serves MR * 128 users
constexpr int EPI_G = 2;                            // epilogue warp groups: each owns TN / EPI_G columns of every tile (4 measured no faster)
constexpr int EPI_COLS = TN / EPI_G;                // columns per thread and tile, in chunks of 32
constexpr int EPI_CH = EPI_COLS / 32;
constexpr int NT = 128 + 128 * EPI_G;               // warps: 0 TMA, 1 MMA, 2-3 idle, then 4 * EPI_G epilogue warps
constexpr int CHUNK = 32;

// ------------------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, 0x989680;\n\t"
        "@P1 bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
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
__device__ __forceinline__ float4 lds_f4(uint32_t saddr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
    return v;
}
// 1.0f if a >= b else 0.0f (FSET): the compare result as a float, so the hit mask can be accumulated with FFMAs on
// the FMA pipe instead of SEL / IADD3 on the half-rate ALU pipe
__device__ __forceinline__ float fset_ge(float a, float b) {
    float d;
    asm("set.ge.f32.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b));
    return d;
}

// K-major, 128B-swizzled operand tile: rows of 128 B, 8-row groups 1024 B apart (SBO), version 1, layout type 2.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}
// kind::f16: D = F32 (bit 4), A = B = BF16 (bits 7, 10), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);

}  // namespace tc

// ------------------------------------------------------------------------------------------------------------
// prep kernels
// ------------------------------------------------------------------------------------------------------------
// one warp per row: fp32 row -> bf16 row (+ L2 norm, rounded up); rows >= n_rows are zero padding.
// src_rows == nullptr: row r of src; else row src_rows[r] (gather of the eval users).
__global__ void __launch_bounds__(256) tc_convert_rows_kernel(const float* __restrict__ src, const int32_t* __restrict__ src_rows,
                                                              int64_t n_rows, int64_t n_pad, int d,
                                                              __nv_bfloat16* __restrict__ dst, float* __restrict__ norm) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (r >= n_pad) return;
    float sq = 0.f;
    if (r < n_rows) {
        const float* s = src + (src_rows ? (int64_t)src_rows[r] : r) * d;
        for (int k = lane * 2; k < d; k += 64) {
            const float2 v = *reinterpret_cast<const float2*>(s + k);
            sq = fmaf(v.x, v.x, fmaf(v.y, v.y, sq));
            *reinterpret_cast<__nv_bfloat162*>(dst + r * d + k) = __floats2bfloat162_rn(v.x, v.y);
        }
    } else {
        for (int k = lane * 2; k < d; k += 64) *reinterpret_cast<__nv_bfloat162*>(dst + r * d + k) = __floats2bfloat162_rn(0.f, 0.f);
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, off);
    if (lane == 0) norm[r] = sqrtf(sq) * 1.00001f;
}

// per 256-item tile: max of v[0..n) (values beyond n count as 0)
__global__ void __launch_bounds__(256) tc_tile_norm_kernel(const float* __restrict__ inorm, int64_t n_tiles, float* __restrict__ tmax,
                                                           int64_t n) {
    const int lane = threadIdx.x & 31;
    const int64_t t = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (t >= n_tiles) return;
    float m = 0.f;
    for (int c = lane; c < tc::TN; c += 32) { const int64_t j = t * tc::TN + c; if (j < n) m = fmaxf(m, inorm[j]); }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
    if (lane == 0) tmax[t] = m;
}

// ------------------------------------------------------------------------------------------------------------
// the sweep
// ------------------------------------------------------------------------------------------------------------
struct SweepArgs {
    int64_t M, N;              // real rows / items
    int64_t M_pad;             // rows of the bf16 user copy (multiple of 128)
    int d;
    int n_tiles;               // item tiles of 256
    int tiles_per_split;       // each CTA of blockIdx.y sweeps [y * tiles_per_split, ...)
    int se;                    // pass A: every se-th tile is sampled
    float c_err;               // error-bound coefficient
    const float* unorm;        // [M_pad]
    const float* tile_inorm;   // [n_tiles]
    const float* col;          // mode 1: pop [N]; mode 0: col_bias [N] or nullptr
    const float* tile_colmax;  // mode 1: max of pop over each 256-item tile
    // pass A
    float* cmax; int n_c;      // [M_pad][n_c] chunk maxima, row-major
    // pass B
    const float* tau;          // [M_pad]
    int32_t* cand;             // [M_pad][n_seg][seg_cap] item ids, ascending inside a segment;
                               // segment = (item split, column half): exactly one owner thread, no atomics
    int32_t* cnt;              // [M_pad][n_seg] entries written (> seg_cap = overflow)
    int n_seg, seg_cap;
};

template <int MODE, int PASS>
__global__ void __launch_bounds__(tc::NT, 1) tc_sweep_kernel(const __grid_constant__ CUtensorMap tmA,
                                                              const __grid_constant__ CUtensorMap tmB, SweepArgs a,
                                                              int n_stages) {
    using namespace tc;
    extern __shared__ unsigned char smem_raw[];
    // 1024 B alignment for the 128B-swizzle atoms
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int kblocks = a.d / KB;
    const uint32_t a1_bytes = (uint32_t)TM * 128u * kblocks;     // one A tile: kblocks sub-tiles of [128 rows x 128 B]
    const uint32_t a_bytes = a1_bytes * MR;                      // MR user tiles stay resident for the whole sweep
    const uint32_t b_bytes = (uint32_t)TN * 128u * kblocks;      // B stage: kblocks sub-tiles of [256 rows x 128 B]
    unsigned char* sA = smem;
    unsigned char* sB = smem + a_bytes;
    unsigned char* tail = sB + (size_t)n_stages * b_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(tail);          // full[8] empty[8] tfull[2] tempty[2] afull[1]
    float* scol = reinterpret_cast<float*>(tail + 256);          // [2][256] column values (pop / bias), double-buffered by tile
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tail + 256 + 2048);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + 8), bar_tfull = smem_u32(bars + 16),
                   bar_tempty = smem_u32(bars + 18), bar_afull = smem_u32(bars + 20);

    const int m_blk = blockIdx.x;                                // MR consecutive 128-row user tiles
    const int t_begin = blockIdx.y * a.tiles_per_split;
    const int t_end = min(a.n_tiles, t_begin + a.tiles_per_split);
    // tiles this CTA visits: pass A -> multiples of se; pass B -> all
    const int step = PASS == 0 ? a.se : 1;
    const int t_first = PASS == 0 ? ((t_begin + a.se - 1) / a.se) * a.se : t_begin;
    const int n_my = t_first < t_end ? (t_end - t_first + step - 1) / step : 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < n_stages; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(bar_tfull + 8 * s, 1); mbar_init(bar_tempty + 8 * s, 4 * EPI_G); }
        mbar_init(bar_afull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512);
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // Every B tile is multiplied with the MR resident user tiles in turn: sub-step sub = i * MR + mr uses accumulator
    // stage sub & 1, so the epilogue of one user tile overlaps the MMAs of the next, and each B tile fetched from L2
    // serves MR * 128 users.
    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0 && n_my > 0) {
            mbar_expect_tx(bar_afull, a_bytes);
            for (int mr = 0; mr < MR; ++mr)
                for (int kb = 0; kb < kblocks; ++kb)
                    tma_load_2d(smem_u32(sA + (size_t)mr * a1_bytes + (size_t)kb * TM * 128), &tmA, bar_afull, kb * KB,
                                (m_blk * MR + mr) * TM);
            for (int i = 0; i < n_my; ++i) {
                const int s = i % n_stages, ph = (i / n_stages) & 1;
                mbar_wait(bar_empty + 8 * s, ph ^ 1);
                mbar_expect_tx(bar_full + 8 * s, b_bytes);
                const int t = t_first + i * step;
                for (int kb = 0; kb < kblocks; ++kb)
                    tma_load_2d(smem_u32(sB + (size_t)s * b_bytes + (size_t)kb * TN * 128), &tmB, bar_full + 8 * s, kb * KB, t * TN);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0 && n_my > 0) {
            mbar_wait(bar_afull, 0);
            for (int i = 0; i < n_my; ++i) {
                const int s = i % n_stages, ph = (i / n_stages) & 1;
                mbar_wait(bar_full + 8 * s, ph);
                for (int mr = 0; mr < MR; ++mr) {
                    const int sub = i * MR + mr, acc = sub & 1, aph = (sub >> 1) & 1;
                    mbar_wait(bar_tempty + 8 * acc, aph ^ 1);
                    fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)acc * TN;
                    for (int kb = 0; kb < kblocks; ++kb) {
                        const uint64_t ad = make_desc(smem_u32(sA + (size_t)mr * a1_bytes + (size_t)kb * TM * 128));
                        const uint64_t bd = make_desc(smem_u32(sB + (size_t)s * b_bytes + (size_t)kb * TN * 128));
#pragma unroll
                        for (int k = 0; k < KB / 16; ++k)    // UMMA K = 16 bf16 = 32 B inside the 128 B swizzle row
                            umma_bf16(d_tmem, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), IDESC, (kb | k) ? 1u : 0u);
                    }
                    if (mr == MR - 1) umma_commit(bar_empty + 8 * s);   // B stage may be refilled once these MMAs retire
                    umma_commit(bar_tfull + 8 * acc);                   // accumulator ready for the epilogue
                }
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue: thread = (row 32q + lane of each resident user tile, column group h of EPI_COLS columns) =====
        const int q = warp & 3, h = (warp - 4) >> 2;
        const int e = (warp - 4) * 32 + lane;                    // epilogue thread index; the first 256 stage scol
        const bool use_col = MODE == 1 || a.col != nullptr;
        // pass B: this thread owns segment (split, h) of each of its rows' candidate lists -> no atomics
        const int seg = blockIdx.y * EPI_G + h;
        int64_t row[MR];
        float un[MR], tau[MR];
        int32_t* my_cand[MR];
        int n_local[MR];
#pragma unroll
        for (int mr = 0; mr < MR; ++mr) {
            row[mr] = ((int64_t)m_blk * MR + mr) * TM + 32 * q + lane;
            un[mr] = a.unorm[row[mr]] * a.c_err;
            tau[mr] = 0.f;
            if (PASS == 1) tau[mr] = row[mr] < a.M ? a.tau[row[mr]] : INFINITY;
            my_cand[mr] = PASS == 1 ? a.cand + (row[mr] * a.n_seg + seg) * a.seg_cap : nullptr;
            n_local[mr] = 0;
        }
        const uint32_t lane_base = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(h * EPI_COLS);
        for (int i = 0; i < n_my; ++i) {
            const int t = t_first + i * step;
            const int64_t j0 = (int64_t)t * TN;
            const float* sc = scol + (i & 1) * 256 + h * EPI_COLS;
            float pmax_t = INFINITY;
            if (use_col) {
                if (e < TN) {
                    const int64_t j = j0 + e;
                    scol[(i & 1) * 256 + e] = j < a.N ? __ldg(a.col + j) : 0.f;
                }
                asm volatile("bar.sync 1, %0;" ::"n"(128 * EPI_G) : "memory");
                if (MODE == 1 && PASS == 1) pmax_t = __ldg(a.tile_colmax + t);
            }
            const float tn = __ldg(a.tile_inorm + t);
#pragma unroll
            for (int mr = 0; mr < MR; ++mr) {
                const int sub = i * MR + mr, acc = sub & 1, aph = (sub >> 1) & 1;
                const float er = un[mr] * tn;                     // |s - s_lp| <= er for every item of this tile
                mbar_wait(bar_tfull + 8 * acc, aph);
                fence_after();
                uint32_t va[32], vb[32];
                float bests[EPI_CH];
                const uint32_t tbase = lane_base + (uint32_t)(acc * TN);
                tmem_ld32(tbase, va);
#pragma unroll
                for (int cc = 0; cc < EPI_CH; ++cc) {
                    uint32_t* v = (cc & 1) ? vb : va;
                    tmem_ld_wait();
                    if (cc < EPI_CH - 1) tmem_ld32(tbase + (uint32_t)((cc + 1) * 32), (cc & 1) ? va : vb);   // next chunk in flight
                    else {
                        // all chunks are in registers: hand the accumulator stage back to the MMA warp
                        fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
                    }
                    const int cb = cc * 32;                          // first column of this chunk inside this thread's group
                    const int64_t jb = j0 + h * EPI_COLS + cb;
                    if (PASS == 0) {
                        float best = -INFINITY;
                        if (jb + 32 <= a.N) {
                            if (MODE == 1) {
                                // lower bound of (elu(s)+1)*pop: f(x) >= x + 1 (also when x + 1 < 0: the product then is <= 0 <= y)
                                const float k1 = 1.0f - er;
#pragma unroll
                                for (int c = 0; c < 32; c += 4) {
                                    const float4 p4 = lds_f4(smem_u32(sc + cb + c));
                                    best = fmaxf(best, (__uint_as_float(v[c + 0]) + k1) * p4.x);
                                    best = fmaxf(best, (__uint_as_float(v[c + 1]) + k1) * p4.y);
                                    best = fmaxf(best, (__uint_as_float(v[c + 2]) + k1) * p4.z);
                                    best = fmaxf(best, (__uint_as_float(v[c + 3]) + k1) * p4.w);
                                }
                            } else if (use_col) {
                                const float k1 = -er;
#pragma unroll
                                for (int c = 0; c < 32; c += 4) {
                                    const float4 p4 = lds_f4(smem_u32(sc + cb + c));
                                    best = fmaxf(best, (__uint_as_float(v[c + 0]) + k1) + p4.x);
                                    best = fmaxf(best, (__uint_as_float(v[c + 1]) + k1) + p4.y);
                                    best = fmaxf(best, (__uint_as_float(v[c + 2]) + k1) + p4.z);
                                    best = fmaxf(best, (__uint_as_float(v[c + 3]) + k1) + p4.w);
                                }
                            } else {
                                // max first, bound after: s - er is monotone in s
#pragma unroll
                                for (int c = 0; c < 32; ++c) best = fmaxf(best, __uint_as_float(v[c]));
                                best -= er;
                            }
                        } else {
                            // the last tile: padded columns (zero rows) must not produce a bound
                            for (int c = 0; c < 32; ++c) {
                                if (jb + c >= a.N) break;
                                const float s = __uint_as_float(v[c]);
                                float y;
                                if (MODE == 1) y = (s + (1.0f - er)) * sc[cb + c];
                                else y = use_col ? (s - er) + sc[cb + c] : s - er;
                                best = fmaxf(best, y);
                            }
                        }
                        bests[cc] = best;
                    } else {
                        // hit mask, accumulated as two exact fp32 sums of distinct powers of two (columns 0-15, 16-31)
                        float hlo = 0.f, hhi = 0.f;
                        const uint32_t sc_s = smem_u32(sc + cb);
                        if (MODE == 1) {
                            // upper bound: f(x) <= max(x + 1, 1), and max(a, 1) * p = max(a * p, p) for p >= 0.  When tau
                            // exceeds every pop of the tile (the usual case: pop <= 1 < tau) only a * p >= tau can fire.
                            const float k1 = 1.0f + er;
                            const bool simple = __all_sync(0xffffffffu, tau[mr] > pmax_t);
                            if (simple) {
#pragma unroll
                                for (int c = 0; c < 32; c += 4) {
                                    const float4 p4 = lds_f4(sc_s + 4 * c);
                                    const float pc[4] = {p4.x, p4.y, p4.z, p4.w};
#pragma unroll
                                    for (int z = 0; z < 4; ++z) {
                                        const float f = fset_ge((__uint_as_float(v[c + z]) + k1) * pc[z], tau[mr]);
                                        if (c + z < 16) hlo = fmaf(f, (float)(1u << ((c + z) & 15)), hlo);
                                        else hhi = fmaf(f, (float)(1u << ((c + z) & 15)), hhi);
                                    }
                                }
                            } else {
#pragma unroll
                                for (int c = 0; c < 32; c += 4) {
                                    const float4 p4 = lds_f4(sc_s + 4 * c);
                                    const float pc[4] = {p4.x, p4.y, p4.z, p4.w};
#pragma unroll
                                    for (int z = 0; z < 4; ++z) {
                                        const float f = fset_ge(fmaxf(__uint_as_float(v[c + z]) + k1, 1.0f) * pc[z], tau[mr]);
                                        if (c + z < 16) hlo = fmaf(f, (float)(1u << ((c + z) & 15)), hlo);
                                        else hhi = fmaf(f, (float)(1u << ((c + z) & 15)), hhi);
                                    }
                                }
                            }
                        } else if (use_col) {
#pragma unroll
                            for (int c = 0; c < 32; c += 4) {
                                const float4 p4 = lds_f4(sc_s + 4 * c);
                                const float pc[4] = {p4.x, p4.y, p4.z, p4.w};
#pragma unroll
                                for (int z = 0; z < 4; ++z) {
                                    const float f = fset_ge((__uint_as_float(v[c + z]) + er) + pc[z], tau[mr]);
                                    if (c + z < 16) hlo = fmaf(f, (float)(1u << ((c + z) & 15)), hlo);
                                    else hhi = fmaf(f, (float)(1u << ((c + z) & 15)), hhi);
                                }
                            }
                        } else {
                            const float thr = tau[mr] - er - fabsf(tau[mr]) * 4e-6f;      // s + er >= tau, rounding-safe
#pragma unroll
                            for (int c = 0; c < 32; ++c) {
                                const float f = fset_ge(__uint_as_float(v[c]), thr);
                                if (c < 16) hlo = fmaf(f, (float)(1u << (c & 15)), hlo);
                                else hhi = fmaf(f, (float)(1u << (c & 15)), hhi);
                            }
                        }
                        uint32_t hits = (uint32_t)hlo | ((uint32_t)hhi << 16);
                        while (hits) {
                            const int c = __ffs(hits) - 1;
                            hits &= hits - 1;
                            const int64_t j = jb + c;
                            if (j < a.N) {
                                if (n_local[mr] < a.seg_cap) my_cand[mr][n_local[mr]] = (int32_t)j;
                                ++n_local[mr];                           // > seg_cap marks the overflow
                            }
                        }
                    }
                }
                if (PASS == 0) {   // this thread's chunk maxima of the tile: one vector store, row-major [row][n_c]
                    float* dst = a.cmax + row[mr] * a.n_c + (t / a.se) * 8 + h * EPI_CH;
                    if (EPI_CH == 4) *reinterpret_cast<float4*>(dst) = make_float4(bests[0], bests[1], bests[EPI_CH - 2], bests[EPI_CH - 1]);
                    else if (EPI_CH == 2) *reinterpret_cast<float2*>(dst) = make_float2(bests[0], bests[1]);
                    else dst[0] = bests[0];
                }
            }
        }
        if (PASS == 1) {
#pragma unroll
            for (int mr = 0; mr < MR; ++mr) a.cnt[row[mr] * a.n_seg + seg] = n_local[mr];
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 1) { __syncwarp(); fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ------------------------------------------------------------------------------------------------------------
// tau selection: one warp per row; keys = chunk maxima with the column packed in the low 5 mantissa bits
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t f2key(float f) { uint32_t b = __float_as_uint(f); return (b & 0x80000000u) ? ~b : (b | 0x80000000u); }
__device__ __forceinline__ float key2f(uint32_t k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }

__device__ __forceinline__ bool csr_row_contains(const int32_t* __restrict__ items, int64_t lo, int64_t hi, int32_t c) {
    const int64_t end = hi;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (__ldg(items + mid) < c) lo = mid + 1; else hi = mid;
    }
    return lo < end && __ldg(items + lo) == c;
}

constexpr int TAU_MAX_KEYS = 2048;

// K-th largest of n 32-bit keys in shared memory (one warp): 4-pass MSB radix select, 256-bin histogram per pass.
// Returns the key value; *n_greater = number of keys strictly greater.  hist: 256 ints of per-warp shared memory.
__device__ __forceinline__ uint32_t warp_kth_largest(const uint32_t* keys, int n, int K, int* hist, int lane, int* n_greater) {
    uint32_t prefix = 0, mask = 0;
    int need = K, above = 0;
    for (int pass = 3; pass >= 0; --pass) {
        const int shift = pass * 8;
#pragma unroll
        for (int b = 0; b < 8; ++b) hist[lane * 8 + b] = 0;
        __syncwarp();
        for (int c = lane; c < n; c += 32) {
            const uint32_t k = keys[c];
            if ((k & mask) == prefix) atomicAdd(&hist[(k >> shift) & 255u], 1);
        }
        __syncwarp();
        // lane L owns bins 255-8L .. 248-8L (descending); find the bin where the count from the top reaches `need`
        int cntb[8], lsum = 0;
#pragma unroll
        for (int b = 0; b < 8; ++b) { cntb[b] = hist[255 - (lane * 8 + b)]; lsum += cntb[b]; }
        int incl = lsum;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += t;
        }
        const int excl = incl - lsum;
        const unsigned bal = __ballot_sync(0xffffffffu, incl >= need);
        const int owner = __ffs(bal) - 1;            // bal != 0 because need <= matching keys
        int bin = 0, before = 0;
        if (lane == owner) {
            int run = excl;
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                if (run + cntb[b] >= need) { bin = 255 - (lane * 8 + b); before = run; break; }
                run += cntb[b];
            }
        }
        bin = __shfl_sync(0xffffffffu, bin, owner);
        before = __shfl_sync(0xffffffffu, before, owner);
        above += before;
        need -= before;
        prefix |= (uint32_t)bin << shift;
        mask |= 255u << shift;
        __syncwarp();
    }
    *n_greater = above;
    return prefix;
}

// one warp per row, n_c keys + 256 histogram bins per warp in dynamic shared memory
__global__ void __launch_bounds__(128) tc_tau_select_kernel(const float* __restrict__ cmax, int n_c, int64_t M, int64_t M_pad,
                                                            int se, const int32_t* __restrict__ users,
                                                            const int64_t* __restrict__ mask_indptr,
                                                            const int32_t* __restrict__ mask_items, int K,
                                                            float* __restrict__ tau) {
    extern __shared__ uint32_t tau_keys[];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * 4 + w;
    if (row >= M_pad) return;
    if (row >= M) { if (lane == 0) tau[row] = INFINITY; return; }
    uint32_t* kk = tau_keys + (size_t)w * (n_c + 256);
    int* hist = reinterpret_cast<int*>(kk + n_c);
    for (int c = lane; c < n_c; c += 32) {
        const float v = cmax[row * n_c + c];
        kk[c] = v > -INFINITY ? f2key(v) : 0u;    // 0 sorts below every real value
    }
    __syncwarp();
    // A sampled chunk that holds ANY train item of this user is dropped: its maximum may belong to a masked item.
    // sampled chunk c = (sampled tile c/8, chunk c%8) covers items [(c/8)*se*256 + (c%8)*32, +32)
    if (mask_indptr) {
        const int u = users[row];
        const int64_t lo = mask_indptr[u], hi = mask_indptr[u + 1];
        for (int64_t z = lo + lane; z < hi; z += 32) {
            const int32_t it = __ldg(mask_items + z);
            const int tile = it / tc::TN;
            if (tile % se == 0) {
                const int c = (tile / se) * 8 + ((it % tc::TN) >> 5);
                if (c < n_c) kk[c] = 0u;
            }
        }
        __syncwarp();
    }
    int n_clean = 0;
    for (int c = lane; c < n_c; c += 32) n_clean += kk[c] != 0u;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) n_clean += __shfl_xor_sync(0xffffffffu, n_clean, off);
    if (n_clean < K) {
        if (lane == 0) tau[row] = INFINITY;     // fewer than K clean maxima: no candidates -> no certificate -> exact kernel
        return;
    }
    int above;
    const uint32_t kth = warp_kth_largest(kk, n_c, K, hist, lane, &above);
    if (lane == 0) {
        const float t = key2f(kth);
        tau[row] = t - fabsf(t) * 1e-5f - 1e-30f;      // the bound arithmetic of the sweep rounds: step down
    }
}

// ------------------------------------------------------------------------------------------------------------
// exact rescoring + top-K + certificate: one warp per row
// ------------------------------------------------------------------------------------------------------------
struct RescoreArgs {
    const float* U; const float* I; int64_t M; int64_t N;
    const int32_t* users;
    int mode; const float* pop; const float* col_bias;
    const int64_t* mask_indptr; const int32_t* mask_items;
    const int32_t* cand; const int32_t* cnt; int n_seg, seg_cap;
    int tiles_per_split;
    const float* tau;
    int K;
    int rc;             // per-row capacity of the compacted candidate list in shared memory
    int32_t* ids_out; float* scores_out;
    int32_t* flag;      // [M] 1 = not certified
};

constexpr int SORT_MAX = 256;

//  1. the row's candidate segments -> one compact id list in shared memory
//  2. train items out: every masked item lives in exactly one segment (ascending ids) -> binary search there
//  3. exact fp32 score (the sequential-k spec) + transform of every remaining candidate
//  4. K-th largest exact score by radix select, survivors (>= it) bitonic-sorted by (score desc, id asc)
//  5. certificate: no overflow and >= K unmasked candidates with exact score >= tau
template <int D>
__global__ void __launch_bounds__(128) tc_rescore_kernel(RescoreArgs a) {
    extern __shared__ __align__(8) unsigned char rs_smem[];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int RC = a.rc;
    // per warp: sortbuf[256] u64 | cid[RC] | ckey[RC] | hist[256] | soff[n_seg + 1]
    const size_t per_warp = (size_t)SORT_MAX * 8 + (size_t)RC * 8 + 1024 + (size_t)((a.n_seg + 2) / 2 * 2) * 4;
    unsigned char* base = rs_smem + (size_t)w * per_warp;
    unsigned long long* sortbuf = reinterpret_cast<unsigned long long*>(base);
    int* cid = reinterpret_cast<int*>(base + (size_t)SORT_MAX * 8);
    uint32_t* ckey = reinterpret_cast<uint32_t*>(base + (size_t)SORT_MAX * 8 + (size_t)RC * 4);
    int* hist = reinterpret_cast<int*>(base + (size_t)SORT_MAX * 8 + (size_t)RC * 8);
    int* soff = hist + 256;
    const int64_t row = (int64_t)blockIdx.x * 4 + w;
    if (row >= a.M) return;
    const int K = a.K;
    const int u = a.users[row];
    const float* ur = a.U + (int64_t)u * D;

    // 1. compact
    bool overflow = false;
    int tot = 0;
    for (int sg0 = 0; sg0 < a.n_seg; sg0 += 32) {
        const int sg = sg0 + lane;
        int n = sg < a.n_seg ? a.cnt[row * a.n_seg + sg] : 0;
        if (n > a.seg_cap) { overflow = true; n = a.seg_cap; }
        int incl = n;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += t;
        }
        if (sg < a.n_seg) soff[sg] = tot + incl - n;
        tot += __shfl_sync(0xffffffffu, incl, 31);
    }
    overflow = __any_sync(0xffffffffu, overflow) || tot > RC;
    if (lane == 0) soff[a.n_seg] = tot;
    __syncwarp();
    if (overflow) { if (lane == 0) a.flag[row] = 1; return; }
    for (int sg = 0; sg < a.n_seg; ++sg) {
        const int o = soff[sg], nn = soff[sg + 1] - o;
        const int32_t* cl = a.cand + (row * a.n_seg + sg) * a.seg_cap;
        for (int c = lane; c < nn; c += 32) cid[o + c] = cl[c];
    }
    __syncwarp();

    // 2. masked items: id -> ~id (negative)
    if (a.mask_indptr) {
        const int64_t lo = a.mask_indptr[u], hi = a.mask_indptr[u + 1];
        for (int64_t z = lo + lane; z < hi; z += 32) {
            const int32_t it = __ldg(a.mask_items + z);
            const int sg = (it / tc::TN / a.tiles_per_split) * tc::EPI_G + (it % tc::TN) / tc::EPI_COLS;
            const int send = soff[sg + 1];
            int l = soff[sg], r = send;
            while (l < r) {
                const int mid = (l + r) >> 1;
                if (cid[mid] < it) l = mid + 1; else r = mid;
            }
            if (l < send && cid[l] == it) cid[l] = ~it;
        }
        __syncwarp();
    }

    // 3. exact scores
    const float tau = a.tau[row];
    int n_cert = 0;
    for (int c = lane; c < tot; c += 32) {
        const int j = cid[c];
        uint32_t key = 0;
        if (j >= 0) {
            const float* ir = a.I + (int64_t)j * D;
            float4 iv[D / 4];
#pragma unroll
            for (int k = 0; k < D / 4; ++k) iv[k] = ldg_f4(ir + 4 * k);     // the whole row in flight at once
            float acc = 0.0f;
#pragma unroll
            for (int k = 0; k < D / 4; ++k) {
                const float4 uv = ldg_f4(ur + 4 * k);
                acc = fadd(acc, fmul(uv.x, iv[k].x)); acc = fadd(acc, fmul(uv.y, iv[k].y));
                acc = fadd(acc, fmul(uv.z, iv[k].z)); acc = fadd(acc, fmul(uv.w, iv[k].w));
            }
            float y;
            if (a.mode == 1) y = fmul(elu_p1(acc), __ldg(a.pop + j));
            else y = a.col_bias ? fadd(acc, __ldg(a.col_bias + j)) : acc;
            key = f2key(y);
            if (key == 0u) key = 1u;          // (only -NaN patterns map to 0) keep 0 for "not a candidate"
            n_cert += y >= tau;
        }
        ckey[c] = key;
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) n_cert += __shfl_xor_sync(0xffffffffu, n_cert, off);
    __syncwarp();
    if (n_cert < K) { if (lane == 0) a.flag[row] = 1; return; }

    // 4. select + sort
    int above;
    const uint32_t kth = warp_kth_largest(ckey, tot, K, hist, lane, &above);
    int ns = 0;
    for (int c0 = 0; c0 < tot; c0 += 32) {
        const int c = c0 + lane;
        const bool keep = c < tot && ckey[c] >= kth;
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (keep) {
            const int pos = ns + __popc(bal & ((1u << lane) - 1u));
            if (pos < SORT_MAX) sortbuf[pos] = ((unsigned long long)ckey[c] << 32) | (uint32_t)(0x7fffffff - cid[c]);
        }
        ns += __popc(bal);
    }
    if (ns > SORT_MAX) { if (lane == 0) a.flag[row] = 1; return; }     // a huge tie at the K-th score: exact kernel
    int n2 = 64;
    while (n2 < ns) n2 <<= 1;
    for (int c = ns + lane; c < n2; c += 32) sortbuf[c] = 0ull;
    __syncwarp();
    for (int k = 2; k <= n2; k <<= 1) {
        for (int jj = k >> 1; jj > 0; jj >>= 1) {
            for (int idx = lane; idx < n2; idx += 32) {
                const int ixj = idx ^ jj;
                if (ixj > idx) {
                    const unsigned long long x = sortbuf[idx], y = sortbuf[ixj];
                    const bool desc = (idx & k) == 0;        // descending overall
                    if (desc ? x < y : x > y) { sortbuf[idx] = y; sortbuf[ixj] = x; }
                }
            }
            __syncwarp();
        }
    }
    if (lane == 0) a.flag[row] = 0;
    for (int e = lane; e < K; e += 32) {
        const unsigned long long kv = sortbuf[e];
        a.ids_out[row * K + e] = 0x7fffffff - (int32_t)(uint32_t)(kv & 0xffffffffu);
        if (a.scores_out) a.scores_out[row * K + e] = key2f((uint32_t)(kv >> 32));
    }
}

// compact the rows that need the exact kernel: list[0..n) = row ids, *n_out = n
__global__ void tc_compact_flags_kernel(const int32_t* __restrict__ flag, int64_t M, const int32_t* __restrict__ users,
                                        int32_t* __restrict__ rows_out, int32_t* __restrict__ users_out, int32_t* n_out) {
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < M; r += (int64_t)gridDim.x * blockDim.x)
        if (flag[r]) {
            const int p = atomicAdd(n_out, 1);
            rows_out[p] = (int32_t)r;
            users_out[p] = users[r];
        }
}

// ------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// bf16 [rows, d] row-major, box = 64 columns (128 B) x box_rows, 128B swizzle
static int make_map(CUtensorMap* m, const void* base, int64_t rows, int d, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return 1;
    cuuint64_t dims[2] = {(cuuint64_t)d, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)d * 2};
    cuuint32_t box[2] = {(cuuint32_t)tc::KB, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : 2;
}

static inline size_t al256(size_t x) { return (x + 255) / 256 * 256; }

bool tc_supported(const EvalArgs& a) {
    return (a.d == 64 || a.d == 128) && a.N >= 4096 && a.K >= 1 && a.K <= 128 && a.N < (1LL << 31) - 512;
}

size_t tc_scratch_bytes(const EvalArgs& a, TcPlan* p) {
    using namespace tc;
    p->M_pad = (a.M + TM * MR - 1) / (TM * MR) * (TM * MR);
    p->N_pad = (a.N + TN - 1) / TN * TN;
    p->n_tiles = (int)(p->N_pad / TN);
    const int64_t n_chunks = (int64_t)p->n_tiles * 8;
    // pass A samples every se-th tile; its cost ~ 1/se, the candidate count ~ se: keep as many chunk maxima per row
    // as the selection kernel holds in shared memory
    int se = 1;
    while ((n_chunks + se - 1) / se > TAU_MAX_KEYS) ++se;
    p->se = se;
    p->n_c = (p->n_tiles + se - 1) / se * 8;
    const int m_tiles = (int)(p->M_pad / (TM * MR));
    int splits = (148 * 2 + m_tiles - 1) / m_tiles;
    if (splits < 1) splits = 1;
    if (splits > p->n_tiles) splits = p->n_tiles;
    p->tiles_per_split = (p->n_tiles + splits - 1) / splits;
    p->tiles_per_split = (p->tiles_per_split + se - 1) / se * se;    // splits start on sampled tiles
    p->splits = (p->n_tiles + p->tiles_per_split - 1) / p->tiles_per_split;
    // candidate lists: one segment per (item split, column group); twice the expected total as head-room
    p->n_seg = p->splits * EPI_G;
    const int cap_total = a.N <= 262144 ? 1024 : 4096;
    p->seg_cap = ((2 * cap_total + p->n_seg - 1) / p->n_seg + 31) / 32 * 32;
    if (p->seg_cap < 64) p->seg_cap = 64;
    // compacted list a rescoring warp holds in shared memory: ~K * (1 + se) * 1.6 entries expected
    p->rc = p->se <= 2 ? 512 : (p->se <= 6 ? 1024 : 2048);
    size_t o = 0;
    p->o_Ib = o; o += al256((size_t)p->N_pad * a.d * 2);
    p->o_Ub = o; o += al256((size_t)p->M_pad * a.d * 2);
    p->o_inorm = o; o += al256((size_t)p->N_pad * 4);
    p->o_unorm = o; o += al256((size_t)p->M_pad * 4);
    p->o_tnorm = o; o += al256((size_t)p->n_tiles * 4);
    p->o_tcolmax = o; o += al256((size_t)p->n_tiles * 4);
    p->o_cmax = o; o += al256((size_t)p->n_c * p->M_pad * 4);
    p->o_tau = o; o += al256((size_t)p->M_pad * 4);
    p->o_cnt = o; o += al256((size_t)p->M_pad * p->n_seg * 4);
    p->o_cand = o; o += al256((size_t)p->M_pad * p->n_seg * p->seg_cap * 4);
    p->o_flag = o; o += al256((size_t)p->M_pad * 4);
    p->o_frows = o; o += al256((size_t)p->M_pad * 4);
    p->o_fusers = o; o += al256((size_t)p->M_pad * 4);
    p->o_nflag = o; o += 256;
    return o;
}

template <int MODE, int PASS>
static int launch_sweep(const CUtensorMap& tmA, const CUtensorMap& tmB, const SweepArgs& s, const TcPlan& p, int m_tiles,
                        cudaStream_t st) {
    using namespace tc;
    const int kblocks = s.d / KB;
    const size_t a_bytes = (size_t)TM * 128 * kblocks * MR, b_bytes = (size_t)TN * 128 * kblocks;
    int n_stages = (int)((204 * 1024 - a_bytes) / b_bytes);
    if (n_stages > 4) n_stages = 4;
    if (n_stages < 1) return 1;
    const size_t smem = 1024 + a_bytes + (size_t)n_stages * b_bytes + 256 + 2048 + 64;
    if (cudaFuncSetAttribute(tc_sweep_kernel<MODE, PASS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return 2;
    dim3 grid(m_tiles, p.splits);
    tc_sweep_kernel<MODE, PASS><<<grid, NT, smem, st>>>(tmA, tmB, s, n_stages);
    return 0;
}

// Runs the whole filter pipeline for one block of users (M <= what tc_scratch_bytes was sized for).
// `scratch` = device buffer of tc_scratch_bytes(); results -> a.ids_out / a.scores_out.
int launch_recommend_tc(const EvalArgs& a, void* scratch, const TcPlan& p, cudaStream_t st) {
    using namespace tc;
    if (!tc_supported(a)) return 1;
    char* b = (char*)scratch;
    __nv_bfloat16* Ib = (__nv_bfloat16*)(b + p.o_Ib);
    __nv_bfloat16* Ub = (__nv_bfloat16*)(b + p.o_Ub);
    float* inorm = (float*)(b + p.o_inorm); float* unorm = (float*)(b + p.o_unorm); float* tnorm = (float*)(b + p.o_tnorm);
    float* cmax = (float*)(b + p.o_cmax); float* tau = (float*)(b + p.o_tau);
    int32_t* cnt = (int32_t*)(b + p.o_cnt); int32_t* cand = (int32_t*)(b + p.o_cand);
    int32_t* flag = (int32_t*)(b + p.o_flag); int32_t* frows = (int32_t*)(b + p.o_frows);
    int32_t* fusers = (int32_t*)(b + p.o_fusers); int32_t* nflag = (int32_t*)(b + p.o_nflag);
    const int m_tiles = (int)(p.M_pad / (TM * MR));

    // prep
    tc_convert_rows_kernel<<<(unsigned)((p.N_pad * 32 + 255) / 256), 256, 0, st>>>(a.I, nullptr, a.N, p.N_pad, a.d, Ib, inorm);
    tc_convert_rows_kernel<<<(unsigned)((p.M_pad * 32 + 255) / 256), 256, 0, st>>>(a.U, a.users, a.M, p.M_pad, a.d, Ub, unorm);
    tc_tile_norm_kernel<<<(unsigned)(((int64_t)p.n_tiles * 32 + 255) / 256), 256, 0, st>>>(inorm, p.n_tiles, tnorm, p.N_pad);
    float* tcolmax = (float*)(b + p.o_tcolmax);
    if (a.mode == 1)
        tc_tile_norm_kernel<<<(unsigned)(((int64_t)p.n_tiles * 32 + 255) / 256), 256, 0, st>>>(a.pop, p.n_tiles, tcolmax, a.N);
    cudaMemsetAsync(nflag, 0, 4, st);

    CUtensorMap tmA, tmB;
    if (make_map(&tmA, Ub, p.M_pad, a.d, TM) || make_map(&tmB, Ib, p.N_pad, a.d, TN)) return 3;

    SweepArgs s;
    memset(&s, 0, sizeof(s));
    s.M = a.M; s.N = a.N; s.M_pad = p.M_pad; s.d = a.d; s.n_tiles = p.n_tiles; s.tiles_per_split = p.tiles_per_split;
    s.se = p.se;
    s.c_err = 1.02f / 256.0f + (float)a.d / 2097152.0f;
    s.unorm = unorm; s.tile_inorm = tnorm;
    s.col = a.mode == 1 ? a.pop : a.col_bias;
    s.tile_colmax = tcolmax;
    s.cmax = cmax; s.n_c = p.n_c; s.tau = tau; s.cand = cand; s.cnt = cnt; s.n_seg = p.n_seg; s.seg_cap = p.seg_cap;

    int rc = a.mode == 1 ? launch_sweep<1, 0>(tmA, tmB, s, p, m_tiles, st) : launch_sweep<0, 0>(tmA, tmB, s, p, m_tiles, st);
    if (rc) return 10 + rc;
    tc_tau_select_kernel<<<(unsigned)((p.M_pad + 3) / 4), 128, (size_t)4 * (p.n_c + 256) * 4, st>>>(cmax, p.n_c, a.M, p.M_pad, p.se, a.users, a.mask_indptr,
                                                                         a.mask_items, a.K, tau);
    rc = a.mode == 1 ? launch_sweep<1, 1>(tmA, tmB, s, p, m_tiles, st) : launch_sweep<0, 1>(tmA, tmB, s, p, m_tiles, st);
    if (rc) return 20 + rc;

    RescoreArgs r;
    memset(&r, 0, sizeof(r));
    r.U = a.U; r.I = a.I; r.M = a.M; r.N = a.N; r.users = a.users; r.mode = a.mode; r.pop = a.pop;
    r.col_bias = a.col_bias; r.mask_indptr = a.mask_indptr; r.mask_items = a.mask_items;
    r.cand = cand; r.cnt = cnt; r.n_seg = p.n_seg; r.seg_cap = p.seg_cap; r.tiles_per_split = p.tiles_per_split;
    r.tau = tau; r.K = a.K; r.rc = p.rc;
    r.ids_out = a.ids_out; r.scores_out = a.scores_out; r.flag = flag;
    const size_t rs_smem = (size_t)4 * ((size_t)SORT_MAX * 8 + (size_t)p.rc * 8 + 1024 + (size_t)((p.n_seg + 2) / 2 * 2) * 4);
    if (a.d == 64) {
        cudaFuncSetAttribute(tc_rescore_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rs_smem);
        tc_rescore_kernel<64><<<(unsigned)((a.M + 3) / 4), 128, rs_smem, st>>>(r);
    } else {
        cudaFuncSetAttribute(tc_rescore_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rs_smem);
        tc_rescore_kernel<128><<<(unsigned)((a.M + 3) / 4), 128, rs_smem, st>>>(r);
    }
    tc_compact_flags_kernel<<<148, 256, 0, st>>>(flag, a.M, a.users, frows, fusers, nflag);

    // rows without a certificate: the exact kernel, sized on the device (CTAs beyond ceil(n/64) exit at once)
    EvalArgs f = a;
    f.users = fusers; f.M_dev = nflag; f.out_rows = frows; f.dense_out = nullptr;
    if (launch_recommend_exact(f, st)) return 30;
    return 0;
}

}  // namespace pda
