// The fused BPR step as a bulk-copy pipeline (d = 128: one 512 B row = one float4 per lane of a warp).
//
// Same arithmetic, same operation order and same results as bpr_step_kernel<32,1,MODE,UMODE> (pda_train.cu) --
// reference: MF/model_api.py:51-53 (gather), :102-121 (PD loss), :123-134 (BPRMF loss), :83 (Adam) -- but the rows
// no longer travel through per-lane LDG.128 into registers that then wait out the HBM latency:
//
//   * every warp owns a ring of D stages in shared memory; one stage = the rows of ONE triple (u, [m, v,] p, n).
//     The lane that holds a triple's ids issues cp.async.bulk (UBLKCP, 512 B per row) for the triple D positions
//     ahead of the one being computed and arms the stage's mbarrier with the byte count; the warp waits on the
//     mbarrier of the stage it consumes.  D x (3 or 5) x 512 B are in flight per warp while it computes -- the memory
//     latency hides behind the (MUFU / FMA-pipe bound) exact Adam replay instead of adding to it.
//   * ids / pops of 32 triples are loaded coalesced (lane l <- triple l of the chunk) one chunk ahead, together with
//     applied[user]; per triple they reach the other lanes by shuffle.
//   * a CTA owns a contiguous range of the batch and its warps draw 32-triple chunks from a shared-memory counter:
//     the replay length of a user row is geometric (mean U/B), so static warp ranges would leave a tail.
//   * L2 policies (`hints`: evict_first on the user rows, touched once per step; evict_last on the Zipf-hot item rows and
//     accumulator) are implemented and measured slower (1.42 vs 1.34 ms): off by default.
//   * the gradient rows of the n_hot most popular positive items (StepArgs::hot_slot) are summed per CTA in shared memory
//     (CAS-loop fp32 atomics) and reach the accumulator once per CTA: 2^20 triples put 25 % of their positive-item rows on
//     28 addresses, and same-address red.global.add serialises in L2 (tools/microbench/gather_bw: 1.14 -> 0.92 ms for the
//     memory skeleton when those reductions disappear).
//   * the exact replay runs with one warp-voted range guard per block of 16 steps, on the negated-v state (pda_common.cuh:
//     lazy_replay4_warp), the row's own Adam step with one guard per float4; the two dot products finish on the two
//     half-warps and elu runs once for both; log(sigmoid + 1e-10) once per 32 triples (lane l <- triple l).
#include <stdlib.h>

#include "pda_kernels.h"

namespace pda {
namespace sp {

constexpr uint32_t FULL = 0xffffffffu;
constexpr int ROW_B = 512;   // d = 128 fp32

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
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
// one lane of a converged warp (elect.sync): the branch stays warp-uniform for the compiler
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t"
        ".reg .b32 rx;\n\t"
        ".reg .pred px;\n\t"
        "elect.sync rx|px, 0xffffffff;\n\t"
        "@px mov.s32 %0, 1;\n\t"
        "}" : "+r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t policy_evict_normal() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
// one row, global -> shared, completion counted in bytes on the stage's mbarrier
__device__ __forceinline__ void bulk_row(uint32_t dst, const void* src, uint32_t bar, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
                 "l"(src), "r"((uint32_t)ROW_B), "r"(bar), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void st_f4_hint(float* p, const float4& v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void red_add_f4_hint(float* p, const float4& v, uint64_t pol) {
    asm volatile("red.global.add.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w),
                 "l"(pol)
                 : "memory");
}
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

struct Chunk {          // lane l holds triple l of the chunk
    int32_t iu, ip, in, done;
    int32_t hs;         // shared-memory slot of a popular positive item, 255 = none
    float pp, pn;
    int n;              // triples in the chunk (0 = no chunk)
};

}  // namespace sp

// MODE 0: BPRMF, 1: PD / PDG.  FUSE: the user table is kept lazily and the warp applies the user row's Adam update
// (UMODE 2 of bpr_step_kernel); otherwise the user gradient is stored into GU (UMODE 1).  Users are distinct.
// V: 0 = the round-2a instruction stream (per-lane replay guards, elu evaluated for value and derivative separately),
//    1 = warp-voted replay guard on the negated-v state + one exp per score.  MINB: resident CTAs per SM the register
//    budget is cut for.
template <int MODE, bool FUSE, int D, int NW, int MINB, int V>
__global__ void __launch_bounds__(NW * 32, MINB) bpr_step_pipe_kernel(StepArgs a, int64_t seg, int hints) {
    using namespace sp;
    constexpr bool POP = MODE == 1;
    constexpr int NR = FUSE ? 5 : 3;
    constexpr int STAGE_B = NR * ROW_B;
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t rows0 = smem_u32(smem) + (uint32_t)(warp * D * STAGE_B);
    const uint32_t bars0 = smem_u32(smem) + (uint32_t)(NW * D * STAGE_B) + (uint32_t)(warp * D * 8);
    int* counter = reinterpret_cast<int*>(smem + NW * D * STAGE_B + NW * D * 8);
    float* hot_acc = reinterpret_cast<float*>(smem + NW * D * STAGE_B + NW * D * 8 + 16);   // [n_hot][128]
    for (int i = threadIdx.x; i < a.n_hot * 128; i += NW * 32) hot_acc[i] = 0.0f;

    const int64_t t_begin = blockIdx.x * seg;
    const int64_t t_end = t_begin + seg < a.B ? t_begin + seg : a.B;
    const int n_chunks = t_end > t_begin ? (int)((t_end - t_begin + 31) >> 5) : 0;
    if (threadIdx.x == 0) *counter = 0;
    if (lane == 0)
        for (int s = 0; s < D; ++s) mbar_init(bars0 + 8u * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    // hints: bit 0 = evict_first on the user rows, bit 1 = evict_last on the item rows / item-gradient accumulator
    const uint64_t pol_user = (hints & 1) ? policy_evict_first() : policy_evict_normal();
    const uint64_t pol_item = (hints & 2) ? policy_evict_last() : policy_evict_normal();
    float lr_t = 0.f;
    bool lr_ok = false;
    if (FUSE) {
        lr_t = fdiv(fmul(a.lr, fsqrt(fsub(1.0f, a.pw[1]))), fsub(1.0f, a.pw[0]));
        lr_ok = lr_in_replay_range(a.lr);
    }

    auto grab = [&]() -> int {
        int c = 0;
        if (lane == 0) c = atomicAdd(counter, 1);
        c = __shfl_sync(FULL, c, 0);
        return c < n_chunks ? c : -1;
    };
    auto load_chunk = [&](int c) -> Chunk {
        Chunk k;
        k.iu = k.ip = k.in = 0; k.done = 0; k.hs = 255; k.pp = k.pn = 1.0f; k.n = 0;
        if (c < 0) return k;
        const int64_t t0 = t_begin + (int64_t)c * 32;
        k.n = t_end - t0 < 32 ? (int)(t_end - t0) : 32;
        if (lane < k.n) {
            const int64_t t = t0 + lane;
            k.iu = __ldg(a.users + t); k.ip = __ldg(a.pos + t); k.in = __ldg(a.neg + t);
            if (POP) { k.pp = __ldg(a.pos_pop + t); k.pn = __ldg(a.neg_pop + t); }
            if (FUSE) k.done = a.appliedU[k.iu];
            if (a.n_hot) k.hs = __ldg(a.hot_slot + k.ip);
        }
        return k;
    };
    // stage `s` <- the rows of stream position `pos` (relative to the head of `cur`).  Warp-uniform control flow: the ids
    // come from the lane that holds them by shuffle, ONE elected lane issues the copies (operands in uniform registers:
    // no per-lane waterfall around UBLKCP)
    auto issue = [&](const Chunk& cur, const Chunk& nxt, int pos, int s) {
        const bool from_next = pos >= cur.n;
        const int p = from_next ? pos - cur.n : pos;
        if (p >= (from_next ? nxt.n : cur.n)) return;
        int64_t iu, ip, in;
        if (from_next) { iu = __shfl_sync(FULL, nxt.iu, p); ip = __shfl_sync(FULL, nxt.ip, p); in = __shfl_sync(FULL, nxt.in, p); }
        else { iu = __shfl_sync(FULL, cur.iu, p); ip = __shfl_sync(FULL, cur.ip, p); in = __shfl_sync(FULL, cur.in, p); }
        if (elect_one()) {
            const uint32_t bar = bars0 + 8u * s, dst = rows0 + (uint32_t)(s * STAGE_B);
            mbar_expect_tx(bar, STAGE_B);
            bulk_row(dst, a.U + iu * 128, bar, pol_user);
            if (FUSE) {
                bulk_row(dst + ROW_B, a.MU + iu * 128, bar, pol_user);
                bulk_row(dst + 2 * ROW_B, a.VU + iu * 128, bar, pol_user);
            }
            bulk_row(dst + (NR - 2) * ROW_B, a.I + ip * 128, bar, pol_item);
            bulk_row(dst + (NR - 1) * ROW_B, a.I + in * 128, bar, pol_item);
        }
    };

    double mf_acc = 0.0, sq_acc = 0.0;
    unsigned long long n_replayed = 0;

    Chunk cur = load_chunk(grab());
    Chunk nxt = load_chunk(cur.n == 32 ? grab() : -1);
#pragma unroll
    for (int i = 0; i < D; ++i) issue(cur, nxt, i, i);
    int s = 0;
    uint32_t par = 0;
    while (cur.n > 0) {
        float sige_mine = 1.0f;   // sigmoid + 1e-10 of the chunk's triple `lane`
        for (int j = 0; j < cur.n; ++j) {
            mbar_wait(bars0 + 8u * s, par);
            const uint32_t st = rows0 + (uint32_t)(s * STAGE_B) + (uint32_t)(lane * 16);
            float4 u = lds_f4(st);
            float4 mu = make_float4(0.f, 0.f, 0.f, 0.f), vu = mu;
            if (FUSE) { mu = lds_f4(st + ROW_B); vu = lds_f4(st + 2 * ROW_B); }
            const float4 p = lds_f4(st + (NR - 2) * ROW_B), n = lds_f4(st + (NR - 1) * ROW_B);
            __syncwarp();                      // every lane has its slice: the stage may be refilled
            issue(cur, nxt, j + D, s);
            if (++s == D) { s = 0; par ^= 1u; }

            const int64_t iu = __shfl_sync(FULL, cur.iu, j), ip = __shfl_sync(FULL, cur.ip, j), in = __shfl_sync(FULL, cur.in, j);
            float pp = 1.0f, pn = 1.0f;
            if (POP) { pp = __shfl_sync(FULL, cur.pp, j); pn = __shfl_sync(FULL, cur.pn, j); }
            if (FUSE) {
                // catch up: zero-gradient steps applied[u] .. step_no-1 (a never-touched row has m = v = 0: identity)
                const int64_t done = __shfl_sync(FULL, cur.done, j);
                if (V == 1) {
                    if (done < a.step_no && lazy_replay4_warp(u, mu, vu, a.lr_hist, done, a.step_no, lr_ok) && lane == 0)
                        n_replayed += (unsigned long long)(a.step_no - done);
                } else if (done < a.step_no) {
                    const bool mine = mu.x != 0.f || mu.y != 0.f || mu.z != 0.f || mu.w != 0.f || vu.x != 0.f || vu.y != 0.f ||
                                      vu.z != 0.f || vu.w != 0.f;
                    if (__any_sync(FULL, mine)) {
                        if (lane == 0) n_replayed += (unsigned long long)(a.step_no - done);
                        lazy_replay4_blocked(u, mu, vu, a.lr_hist, done, a.step_no, lr_ok);
                    }
                }
            }
            float sp_ = 0.0f, sn_ = 0.0f;
            sp_ = fadd(sp_, fmul(u.x, p.x)); sp_ = fadd(sp_, fmul(u.y, p.y)); sp_ = fadd(sp_, fmul(u.z, p.z)); sp_ = fadd(sp_, fmul(u.w, p.w));
            sn_ = fadd(sn_, fmul(u.x, n.x)); sn_ = fadd(sn_, fmul(u.y, n.y)); sn_ = fadd(sn_, fmul(u.z, n.z)); sn_ = fadd(sn_, fmul(u.w, n.w));
            float sq = 0.0f;
            sq += u.x * u.x + u.y * u.y + u.z * u.z + u.w * u.w;
            sq += p.x * p.x + p.y * p.y + p.z * p.z + p.w * p.w;
            sq += n.x * n.x + n.y * n.y + n.z * n.z + n.w * n.w;
            // scalar chain (model_api.py:107-114 / 124-126)
            float x, dp, dn;
            if (V == 1) {
                // the first butterfly step routes the two sums to the two half-warps (lanes 0-15 finish the positive
                // item's dot, lanes 16-31 the negative item's: the same additions on the same operands as the two full
                // butterflies), and elu runs ONCE, on both scores side by side
                const bool lo = lane < 16;
                float mine = fadd(lo ? sp_ : sn_, __shfl_xor_sync(FULL, lo ? sn_ : sp_, 16));
#pragma unroll
                for (int off = 8; off >= 1; off >>= 1) mine = fadd(mine, __shfl_xor_sync(FULL, mine, off));
                if (POP) {
                    const float e = elu_p1(mine), popm = lo ? pp : pn;     // elu_p1_grad(s) == elu_p1(s) for s < 0
                    const float xs = fmul(e, popm), ds = fmul(mine < 0.0f ? e : 1.0f, popm);
                    x = fsub(__shfl_sync(FULL, xs, 0), __shfl_sync(FULL, xs, 16));
                    dp = __shfl_sync(FULL, ds, 0); dn = __shfl_sync(FULL, ds, 16);
                } else {
                    x = fsub(__shfl_sync(FULL, mine, 0), __shfl_sync(FULL, mine, 16)); dp = 1.0f; dn = 1.0f;
                }
            } else {
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) {
                    sp_ = fadd(sp_, __shfl_xor_sync(FULL, sp_, off));
                    sn_ = fadd(sn_, __shfl_xor_sync(FULL, sn_, off));
                }
                if (POP) {
                    x = fsub(fmul(elu_p1(sp_), pp), fmul(elu_p1(sn_), pn));
                    dp = fmul(elu_p1_grad(sp_), pp);
                    dn = fmul(elu_p1_grad(sn_), pn);
                } else {
                    x = fsub(sp_, sn_); dp = 1.0f; dn = 1.0f;
                }
            }
            const float sig = fdiv(1.0f, fadd(1.0f, spec_expf(-x)));
            const float sige = fadd(sig, 1e-10f);
            const float g = fdiv(fmul(sig, fsub(1.0f, sig)), sige);
            const float cp = fmul(fmul(-g, dp), a.invB);
            const float cn = fmul(fmul(g, dn), a.invB);
            if (lane == j) sige_mine = sige;
            sq_acc += (double)sq;

            float4 du, dpv, dnv;
            du.x = fadd(fadd(fmul(cp, p.x), fmul(cn, n.x)), fmul(a.lb, u.x));
            du.y = fadd(fadd(fmul(cp, p.y), fmul(cn, n.y)), fmul(a.lb, u.y));
            du.z = fadd(fadd(fmul(cp, p.z), fmul(cn, n.z)), fmul(a.lb, u.z));
            du.w = fadd(fadd(fmul(cp, p.w), fmul(cn, n.w)), fmul(a.lb, u.w));
            dpv.x = fadd(fmul(cp, u.x), fmul(a.lb, p.x)); dpv.y = fadd(fmul(cp, u.y), fmul(a.lb, p.y));
            dpv.z = fadd(fmul(cp, u.z), fmul(a.lb, p.z)); dpv.w = fadd(fmul(cp, u.w), fmul(a.lb, p.w));
            dnv.x = fadd(fmul(cn, u.x), fmul(a.lb, n.x)); dnv.y = fadd(fmul(cn, u.y), fmul(a.lb, n.y));
            dnv.z = fadd(fmul(cn, u.z), fmul(a.lb, n.z)); dnv.w = fadd(fmul(cn, u.w), fmul(a.lb, n.w));
            float* gp = a.GI + ip * 128 + lane * 4;
            float* gn = a.GI + in * 128 + lane * 4;
            if (FUSE) {
                lazy_grad_step4(u, mu, vu, du, lr_t);
                float* wr = a.Uw + iu * 128 + lane * 4;
                float* mr = a.MU + iu * 128 + lane * 4;
                float* vr = a.VU + iu * 128 + lane * 4;
                st_f4_hint(wr, u, pol_user); st_f4_hint(mr, mu, pol_user); st_f4_hint(vr, vu, pol_user);
                if (lane == 0) { a.appliedU[iu] = (int32_t)(a.step_no + 1); a.stampU[iu] = 1; }
            } else {
                float* gu = a.GU + iu * 128 + lane * 4;
                st_f4_hint(gu, du, pol_user);
            }
            const int hs = a.n_hot ? __shfl_sync(FULL, cur.hs, j) : 255;
            if (hs < a.n_hot) {
                float* h = hot_acc + hs * 128 + lane * 4;
                atomicAdd(h, dpv.x); atomicAdd(h + 1, dpv.y); atomicAdd(h + 2, dpv.z); atomicAdd(h + 3, dpv.w);
            } else {
                red_add_f4_hint(gp, dpv, pol_item);
            }
            red_add_f4_hint(gn, dnv, pol_item);
        }
        if (lane < cur.n) mf_acc += (double)logf(sige_mine);
        cur = nxt;
        nxt = load_chunk(cur.n == 32 ? grab() : -1);
    }

    if (a.n_hot) {   // every CTA adds its sums of the popular rows once
        __syncthreads();
        for (int r = warp; r < a.n_hot; r += NW) {
            const float4 v = *reinterpret_cast<const float4*>(hot_acc + r * 128 + lane * 4);
            if (__any_sync(FULL, v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f))
                red_add_f4_hint(a.GI + (int64_t)__ldg(a.hot_ids + r) * 128 + lane * 4, v, pol_item);
        }
    }
    if (FUSE) {
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) n_replayed += __shfl_xor_sync(FULL, n_replayed, off);
        if (lane == 0 && n_replayed) atomicAdd(a.stats + 1, n_replayed);
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        mf_acc += __shfl_xor_sync(FULL, mf_acc, off);
        sq_acc += __shfl_xor_sync(FULL, sq_acc, off);
    }
    if (lane == 0) {
        atomicAdd(a.loss_acc + 0, mf_acc);
        atomicAdd(a.loss_acc + 1, sq_acc);
    }
}

template <int MODE, bool FUSE, int D, int NW, int MINB, int V>
static int launch_pipe_inst(const StepArgs& a, int hints, cudaStream_t st) {
    constexpr int NR = FUSE ? 5 : 3;
    // the popular-item rows take what the ring leaves of the MINB-CTAs-per-SM budget (227 KB - 1 KB reserved per CTA)
    const size_t ring = (size_t)NW * D * NR * sp::ROW_B + (size_t)NW * D * 8 + 16;
    const size_t per_cta = (size_t)(227 * 1024) / MINB - 1024;
    const int hot_cap = per_cta > ring ? (int)((per_cta - ring) / sp::ROW_B) : 0;
    int n_hot = a.n_hot < PDA_MAX_HOT_ITEMS ? a.n_hot : PDA_MAX_HOT_ITEMS;
    if (n_hot > hot_cap) n_hot = hot_cap;          // slots beyond the shared-memory rows keep the global reduction
    const size_t smem = ring + (size_t)n_hot * sp::ROW_B;
    auto kern = bpr_step_pipe_kernel<MODE, FUSE, D, NW, MINB, V>;
    static int ctas_per_sm = 0, n_sm = 0, occ_for_hot = -1;
    if (!ctas_per_sm || occ_for_hot != n_hot) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(ring + (size_t)hot_cap * sp::ROW_B)) != cudaSuccess)
            return 1;
        int dev = 0, occ = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NW * 32, smem) != cudaSuccess || occ < 1) return 1;
        ctas_per_sm = occ;
        occ_for_hot = n_hot;
    }
    // one resident wave; every CTA owns a contiguous, 32-aligned range of the batch
    int64_t grid = (int64_t)n_sm * ctas_per_sm;
    const int64_t chunks = (a.B + 31) / 32;
    if (grid > chunks) grid = chunks;
    if (grid < 1) grid = 1;
    const int64_t seg = ((chunks + grid - 1) / grid) * 32;
    grid = (a.B + seg - 1) / seg;
    StepArgs b = a;
    b.n_hot = n_hot;
    kern<<<(int)grid, NW * 32, smem, st>>>(b, seg, hints);
    return 0;
}

template <int D, int NW, int MINB, int V>
static int launch_pipe_dn(const StepArgs& a, int hints, cudaStream_t st) {
    const bool fuse = a.fuse_user_adam != 0;
    if (a.pop_mode == 1)
        return fuse ? launch_pipe_inst<1, true, D, NW, MINB, V>(a, hints, st) : launch_pipe_inst<1, false, D, NW, MINB, V>(a, hints, st);
    return fuse ? launch_pipe_inst<0, true, D, NW, MINB, V>(a, hints, st) : launch_pipe_inst<0, false, D, NW, MINB, V>(a, hints, st);
}

template <int V>
static int launch_pipe_v(const StepArgs& a, int hints, int D, int NW, cudaStream_t st) {
    if (NW == 4) {
        if (D == 2) return launch_pipe_dn<2, 4, 6, V>(a, hints, st);
        return launch_pipe_dn<3, 4, 6, V>(a, hints, st);
    }
    if (D == 2) return launch_pipe_dn<2, 8, 3, V>(a, hints, st);
    if (D == 4) return launch_pipe_dn<4, 8, 3, V>(a, hints, st);
    return launch_pipe_dn<3, 8, 3, V>(a, hints, st);
}

// returns 0 when the pipelined kernel took the launch, non-zero when the shape is not its (the caller falls back)
int launch_bpr_step_pipe(const StepArgs& a, cudaStream_t st) {
    if (a.d != 128 || !a.uniq_users || a.pop_mode == 2 || a.Gslots) return 1;
    const char* e = getenv("PDA_STEP_PIPE");
    if (e && atoi(e) == 0) return 1;
    e = getenv("PDA_STEP_PIPE_HINTS");
    const int hints = e ? atoi(e) : 0;   // measured: evict_first on the user rows costs 6 % (1.42 vs 1.34 ms), evict_last on the item rows 3 %
    e = getenv("PDA_STEP_PIPE_D");
    const int D = e ? atoi(e) : 3;
    e = getenv("PDA_STEP_PIPE_NW");
    const int NW = e ? atoi(e) : 8;
    // 4 CTAs per SM (28 / 32 warps on a 72 / 64-register budget, ring depth 3 / 2) measured slower: 1.36 / 1.38 vs 1.31 ms
    e = getenv("PDA_STEP_PIPE_V");
    const int V = e ? atoi(e) : 1;
    return V == 0 ? launch_pipe_v<0>(a, hints, D, NW, st) : launch_pipe_v<1>(a, hints, D, NW, st);
}

}  // namespace pda
