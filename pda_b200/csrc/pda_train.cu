// Training-side kernels of the PDA hot path: Xavier init, device sampler, the fused BPR step
// and the TF1-semantics Adam sweep.  Reference call sites (paths under the upstream repo):
//   init      MF/model_api.py:86-99
//   sampler   MF/train_new_api.py:260-288 (BPRMF), :366-412 (PD), :415-456 (BPR(t)-pop)
//   step      MF/model_api.py:51-53 (gather), :102-121 (PD loss), :123-134 (BPRMF loss)
//   optimizer MF/model_api.py:83,471 -> tf.train.AdamOptimizer on IndexedSlices (dense sweep)
#include <stdlib.h>

#include "pda_kernels.h"

namespace pda {

// ------------------------------------------------------------------------------------------
// a1. Xavier-uniform init: element e = 4q + w takes word w of Philox(ctr=(q_lo,q_hi,table,0)).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) xavier_init_kernel(float* __restrict__ W, int64_t n, float a, uint32_t seed,
                                                          uint32_t table_id) {
    int64_t nq = (n + 3) / 4;
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < nq; q += (int64_t)gridDim.x * blockDim.x) {
        u32x4 r = philox4x32((uint32_t)q, (uint32_t)((uint64_t)q >> 32), table_id, 0u, seed, TAG_INIT);
        float x[4];
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            float u = fmul((float)(r.w[w] >> 8), 5.9604644775390625e-8f);
            x[w] = fmul(fsub(fmul(u, 2.0f), 1.0f), a);
        }
        int64_t e = 4 * q;
        if (e + 3 < n) {
            *reinterpret_cast<float4*>(W + e) = make_float4(x[0], x[1], x[2], x[3]);
        } else {
            for (int w = 0; w < 4 && e + w < n; ++w) W[e + w] = x[w];
        }
    }
}

void launch_xavier_init(float* W, int64_t rows, int cols, uint32_t seed, uint32_t table_id, cudaStream_t st) {
    int64_t n = rows * cols;
    float a = (float)sqrt(6.0 / (double)(rows + cols));
    int64_t nq = (n + 3) / 4;
    int blocks = (int)((nq + 255) / 256 < 148 * 16 ? (nq + 255) / 256 : 148 * 16);
    if (blocks < 1) blocks = 1;
    xavier_init_kernel<<<blocks, 256, 0, st>>>(W, n, a, seed, table_id);
}

// ------------------------------------------------------------------------------------------
// a8. device sampler.  One thread per batch slot; word stream of slot i:
// Philox(ctr=(i, call, step, epoch), key=(seed, TAG_SAMPLE)), call = 0,1,...; word 0 -> pos
// (or the time slot of a user with an empty list), words 1.. -> negative candidates, each
// rejected by binary search in the user's sorted train row.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ bool row_contains(const int32_t* __restrict__ items, int64_t lo, int64_t hi, int32_t c) {
    int64_t end = hi;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (__ldg(items + mid) < c) lo = mid + 1; else hi = mid;
    }
    return lo < end && __ldg(items + lo) == c;
}

__global__ void __launch_bounds__(128) sample_kernel(SamplerArgs a) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < a.B; i += (int64_t)gridDim.x * blockDim.x) {
        int32_t u;
        if (a.B <= a.n_act) {
            uint32_t y = feistel_once((uint32_t)i, a.half_bits, a.keys);
            while (y >= (uint32_t)a.n_act) y = feistel_once(y, a.half_bits, a.keys);
            u = __ldg(a.active_users + y);
        } else {
            u32x4 r = philox4x32((uint32_t)i, 0u, a.step, a.epoch, a.seed, TAG_USER);
            u = __ldg(a.active_users + mulhi32(r.w[0], (uint32_t)a.n_act));
        }
        u32x4 w = philox4x32((uint32_t)i, 0u, a.step, a.epoch, a.seed, TAG_SAMPLE);
        int64_t lo = __ldg(a.indptr + u), hi = __ldg(a.indptr + u + 1);
        uint32_t deg = (uint32_t)(hi - lo);
        int32_t pos, t;
        if (deg > 0) {
            int64_t at = lo + mulhi32(w.w[0], deg);
            pos = __ldg(a.items + at);
            t = a.times ? (int32_t)__ldg(a.times + at) : 0;
        } else {
            pos = 0;
            t = a.n_times > 0 ? __ldg(a.unique_times + mulhi32(w.w[0], (uint32_t)a.n_times)) : 0;
        }
        uint32_t call = 0;
        int32_t neg;
        for (uint32_t k = 1;; ++k) {
            if ((k & 3u) == 0) { ++call; w = philox4x32((uint32_t)i, call, a.step, a.epoch, a.seed, TAG_SAMPLE); }
            int32_t c = (int32_t)mulhi32(w.w[k & 3u], (uint32_t)a.n_items);
            if (!row_contains(a.items, lo, hi, c)) { neg = c; break; }
        }
        a.users_out[i] = u; a.pos_out[i] = pos; a.neg_out[i] = neg;
        if (a.time_out) a.time_out[i] = t;
        if (a.pop_train) {
            int32_t tt = a.T_pop == 1 ? 0 : t;
            a.pos_pop_out[i] = __ldg(a.pop_train + (int64_t)pos * a.T_pop + tt);
            a.neg_pop_out[i] = __ldg(a.pop_train + (int64_t)neg * a.T_pop + tt);
        }
    }
}

void sampler_keys(uint32_t seed, uint32_t epoch, uint32_t step, uint32_t* keys) {
    u32x4 ka = philox4x32(0u, 0u, step, epoch, seed, TAG_PERM), kb = philox4x32(1u, 0u, step, epoch, seed, TAG_PERM);
    keys[0] = ka.w[0]; keys[1] = ka.w[1]; keys[2] = ka.w[2]; keys[3] = ka.w[3];
    keys[4] = kb.w[0]; keys[5] = kb.w[1];
}

void launch_sampler(SamplerArgs a, cudaStream_t st) {
    sampler_keys(a.seed, a.epoch, a.step, a.keys);
    a.half_bits = feistel_half_bits((uint32_t)a.n_act);
    int64_t blocks = (a.B + 127) / 128;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
    sample_kernel<<<(int)blocks, 128, 0, st>>>(a);
}

// ------------------------------------------------------------------------------------------
// a2-a4 + gradient scatter: the fused BPR step.
//
// A group of G lanes owns one triple: each lane holds C float4 chunks of u, p, n (chunk c of
// lane l covers elements 4(l + G c) ..+3), loaded with 128-bit coalesced LDGs (a d=128 row is
// one 512 B warp request; at d=64 a warp carries two triples).  The two dot products are a
// per-lane sequential sum followed by an xor-butterfly over the G lanes (the order the oracle's
// dot_tree restates); every lane then evaluates the scalar chain redundantly (no divergence,
// no broadcast), forms its slice of the three gradient rows from the registers that still
// hold u, p, n and reduces them into the table-shaped gradient accumulators GU / GI with
// 128-bit red.global.add (users are distinct within a batch when B <= #active users, so the
// user row takes a plain store then).  Loss terms: per-lane partials -> warp shuffle ->
// one fp64 atomic pair per warp.
// ------------------------------------------------------------------------------------------
// UMODE: how the user-row gradient leaves the kernel
//   0  red.global.add into GU (users may repeat inside the batch)
//   1  plain store into GU (users distinct: rd.sample)
//   2  users distinct AND the user table is kept lazily: the warp that owns the triple owns the user row, so it
//      replays the row's skipped zero-gradient Adam steps in registers BEFORE the dot products (pda_adam_lazy.cu),
//      and applies this step's Adam update to the row right away -- no GU traffic, no separate catch-up / apply
//      pass over the user table.  Same fp32 operations in the same order as the separate kernels: bit-identical.
template <int G, int C, int MODE, int UMODE>
__global__ void __launch_bounds__(256) bpr_step_kernel(StepArgs a) {
    constexpr bool POP = MODE == 1, TEMP = MODE == 2;
    constexpr bool UNIQ = UMODE >= 1, FUSE = UMODE == 2;
    constexpr int GPW = 32 / G;  // triples per warp per iteration
    const int lane = threadIdx.x & 31;
    const int gl = lane % G;       // lane within the group
    const int gw = lane / G;       // group within the warp
    const int64_t warp_global = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int q = a.d >> 2;  // float4 chunks per row

    double mf_acc = 0.0, sq_acc = 0.0;
    unsigned long long n_replayed = 0;
    float lr_t = 0.f;
    if (FUSE) lr_t = fdiv(fmul(a.lr, fsqrt(fsub(1.0f, a.pw[1]))), fsub(1.0f, a.pw[0]));
    const bool lr_ok = FUSE && lr_in_replay_range(a.lr);

    for (int64_t base = warp_global * GPW; base < a.B; base += n_warps * GPW) {
        const int64_t i = base + gw;
        const bool valid = i < a.B;
        int32_t iu = 0, ip = 0, in = 0;
        float pp = 1.0f, pn = 1.0f;
        int32_t tt = 0;
        float ubf = 1.0f, pib = 0.0f, nib = 0.0f;   // BPR(t)-pop bias terms (model_api.py:342-358)
        if (valid) {
            iu = __ldg(a.users + i); ip = __ldg(a.pos + i); in = __ldg(a.neg + i);
            if (POP) { pp = __ldg(a.pos_pop + i); pn = __ldg(a.neg_pop + i); }
            if (TEMP) {
                tt = __ldg(a.temp + i);
                const int Tc = a.temp_num + 1;
                // gather_nd(user_temp_bias_all[B,1], (row, t)) is out of bounds for t > 0 -> 0 (TF-GPU)
                ubf = fadd(tt == 0 ? __ldg(a.ub + iu) : 0.0f, 1.0f);
                pib = fadd(__ldg(a.ib + (int64_t)ip * Tc + a.temp_num), __ldg(a.ib + (int64_t)ip * Tc + tt));
                nib = fadd(__ldg(a.ib + (int64_t)in * Tc + a.temp_num), __ldg(a.ib + (int64_t)in * Tc + tt));
            }
        }
        const float* ur = a.U + (int64_t)iu * a.d;
        const float* pr = a.I + (int64_t)ip * a.d;
        const float* nr = a.I + (int64_t)in * a.d;
        float4 u[C], p[C], n[C];
        float4 mu[FUSE ? C : 1], vu[FUSE ? C : 1];
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const int ch = gl + G * c;
            if (valid && ch < q) {
                if (FUSE) {   // the row is rewritten below: plain (coherent) loads
                    u[c] = *reinterpret_cast<const float4*>(ur + 4 * ch);
                    mu[c] = *reinterpret_cast<const float4*>(a.MU + (int64_t)iu * a.d + 4 * ch);
                    vu[c] = *reinterpret_cast<const float4*>(a.VU + (int64_t)iu * a.d + 4 * ch);
                } else {
                    u[c] = ldg_f4(ur + 4 * ch);
                }
                p[c] = ldg_f4(pr + 4 * ch); n[c] = ldg_f4(nr + 4 * ch);
            } else {
                u[c] = p[c] = n[c] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (FUSE) mu[c] = vu[c] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        if (FUSE && valid) {
            // catch up: zero-gradient steps applied[u] .. step_no-1 (a never-touched row has m = v = 0: identity)
            const int64_t done = a.appliedU[iu];
            if (done < a.step_no && a.stampU[iu] != 0) {
                if (gl == 0) n_replayed += (unsigned long long)(a.step_no - done);
#pragma unroll
                for (int c = 0; c < C; ++c)
                    if (gl + G * c < q) lazy_replay4_blocked(u[c], mu[c], vu[c], a.lr_hist, done, a.step_no, lr_ok);
            }
        }
        float sp = 0.0f, sn = 0.0f, sq = 0.0f;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            sp = fadd(sp, fmul(u[c].x, p[c].x)); sp = fadd(sp, fmul(u[c].y, p[c].y));
            sp = fadd(sp, fmul(u[c].z, p[c].z)); sp = fadd(sp, fmul(u[c].w, p[c].w));
            sn = fadd(sn, fmul(u[c].x, n[c].x)); sn = fadd(sn, fmul(u[c].y, n[c].y));
            sn = fadd(sn, fmul(u[c].z, n[c].z)); sn = fadd(sn, fmul(u[c].w, n[c].w));
            sq += u[c].x * u[c].x + u[c].y * u[c].y + u[c].z * u[c].z + u[c].w * u[c].w;
            sq += p[c].x * p[c].x + p[c].y * p[c].y + p[c].z * p[c].z + p[c].w * p[c].w;
            sq += n[c].x * n[c].x + n[c].y * n[c].y + n[c].z * n[c].z + n[c].w * n[c].w;
        }
#pragma unroll
        for (int off = G / 2; off >= 1; off >>= 1) {
            sp = fadd(sp, __shfl_xor_sync(0xffffffffu, sp, off));
            sn = fadd(sn, __shfl_xor_sync(0xffffffffu, sn, off));
        }
        // scalar chain (model_api.py:107-114 / 124-126)
        float x, dp, dn;
        if (POP) {
            x = fsub(fmul(elu_p1(sp), pp), fmul(elu_p1(sn), pn));
            dp = fmul(elu_p1_grad(sp), pp);
            dn = fmul(elu_p1_grad(sn), pn);
        } else if (TEMP) {
            x = fsub(fadd(fmul(ubf, pib), sp), fadd(fmul(ubf, nib), sn)); dp = 1.0f; dn = 1.0f;
        } else {
            x = fsub(sp, sn); dp = 1.0f; dn = 1.0f;
        }
        const float sig = fdiv(1.0f, fadd(1.0f, spec_expf(-x)));
        const float sige = fadd(sig, 1e-10f);
        const float g = fdiv(fmul(sig, fsub(1.0f, sig)), sige);
        const float cp = fmul(fmul(-g, dp), a.invB);
        const float cn = fmul(fmul(g, dn), a.invB);
        if (valid) {
            sq_acc += (double)sq;
            if (gl == 0) mf_acc += (double)logf(sige);
            if (TEMP && gl == 0) {   // bias gradients: d/d ub[u,0] (t == 0 only), d/d ib[.,T] and d/d ib[.,t]
                const int Tc = a.temp_num + 1;
                if (tt == 0) atomicAdd(a.Gub + iu, fmul(cp, fsub(pib, nib)));
                const float gpb = fmul(cp, ubf), gnb = fmul(cn, ubf);
                atomicAdd(a.Gib + (int64_t)ip * Tc + a.temp_num, gpb); atomicAdd(a.Gib + (int64_t)ip * Tc + tt, gpb);
                atomicAdd(a.Gib + (int64_t)in * Tc + a.temp_num, gnb); atomicAdd(a.Gib + (int64_t)in * Tc + tt, gnb);
            }
            float* gu = a.GU + (int64_t)iu * a.d;
            float* gp = a.GI + (int64_t)ip * a.d;
            float* gn = a.GI + (int64_t)in * a.d;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const int ch = gl + G * c;
                if (ch < q) {
                    float4 du, dpv, dnv;
                    du.x = fadd(fadd(fmul(cp, p[c].x), fmul(cn, n[c].x)), fmul(a.lb, u[c].x));
                    du.y = fadd(fadd(fmul(cp, p[c].y), fmul(cn, n[c].y)), fmul(a.lb, u[c].y));
                    du.z = fadd(fadd(fmul(cp, p[c].z), fmul(cn, n[c].z)), fmul(a.lb, u[c].z));
                    du.w = fadd(fadd(fmul(cp, p[c].w), fmul(cn, n[c].w)), fmul(a.lb, u[c].w));
                    dpv.x = fadd(fmul(cp, u[c].x), fmul(a.lb, p[c].x));
                    dpv.y = fadd(fmul(cp, u[c].y), fmul(a.lb, p[c].y));
                    dpv.z = fadd(fmul(cp, u[c].z), fmul(a.lb, p[c].z));
                    dpv.w = fadd(fmul(cp, u[c].w), fmul(a.lb, p[c].w));
                    dnv.x = fadd(fmul(cn, u[c].x), fmul(a.lb, n[c].x));
                    dnv.y = fadd(fmul(cn, u[c].y), fmul(a.lb, n[c].y));
                    dnv.z = fadd(fmul(cn, u[c].z), fmul(a.lb, n[c].z));
                    dnv.w = fadd(fmul(cn, u[c].w), fmul(a.lb, n[c].w));
                    if (FUSE) {
                        float4 w = u[c];
                        lazy_grad_step(w.x, mu[c].x, vu[c].x, du.x, lr_t); lazy_grad_step(w.y, mu[c].y, vu[c].y, du.y, lr_t);
                        lazy_grad_step(w.z, mu[c].z, vu[c].z, du.z, lr_t); lazy_grad_step(w.w, mu[c].w, vu[c].w, du.w, lr_t);
                        *reinterpret_cast<float4*>(a.Uw + (int64_t)iu * a.d + 4 * ch) = w;
                        *reinterpret_cast<float4*>(a.MU + (int64_t)iu * a.d + 4 * ch) = mu[c];
                        *reinterpret_cast<float4*>(a.VU + (int64_t)iu * a.d + 4 * ch) = vu[c];
                    } else if (UNIQ) *reinterpret_cast<float4*>(gu + 4 * ch) = du;
                    else if (a.Gslots) *reinterpret_cast<float4*>(a.Gslots + (2 * a.B + i) * a.d + 4 * ch) = du;
                    else red_add_f4(gu + 4 * ch, du);
                    if (a.Gslots) {      // deterministic mode: rows kept per slot, summed in order by pda_segsum.cu
                        *reinterpret_cast<float4*>(a.Gslots + i * a.d + 4 * ch) = dpv;
                        *reinterpret_cast<float4*>(a.Gslots + (a.B + i) * a.d + 4 * ch) = dnv;
                    } else {
                        red_add_f4(gp + 4 * ch, dpv);
                        red_add_f4(gn + 4 * ch, dnv);
                    }
                }
            }
            if (FUSE && gl == 0) { a.appliedU[iu] = (int32_t)(a.step_no + 1); a.stampU[iu] = 1; }
        }
    }
    if (FUSE) {
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) n_replayed += __shfl_xor_sync(0xffffffffu, n_replayed, off);
        if (lane == 0 && n_replayed) atomicAdd(a.stats + 1, n_replayed);
    }
    // loss partials: warp reduce, one fp64 atomic pair per warp
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        mf_acc += __shfl_xor_sync(0xffffffffu, mf_acc, off);
        sq_acc += __shfl_xor_sync(0xffffffffu, sq_acc, off);
    }
    if (lane == 0) {
        atomicAdd(a.loss_acc + 0, mf_acc);
        atomicAdd(a.loss_acc + 1, sq_acc);
    }
}

template <int G, int C>
static void launch_step_gc(const StepArgs& a, int grid, cudaStream_t st) {
    if (a.pop_mode == 1) {
        if (a.uniq_users && a.fuse_user_adam) bpr_step_kernel<G, C, 1, 2><<<grid, 256, 0, st>>>(a);
        else if (a.uniq_users) bpr_step_kernel<G, C, 1, 1><<<grid, 256, 0, st>>>(a);
        else bpr_step_kernel<G, C, 1, 0><<<grid, 256, 0, st>>>(a);
    } else if (a.pop_mode == 2) {
        if (a.uniq_users) bpr_step_kernel<G, C, 2, 1><<<grid, 256, 0, st>>>(a);
        else bpr_step_kernel<G, C, 2, 0><<<grid, 256, 0, st>>>(a);
    } else {
        if (a.uniq_users && a.fuse_user_adam) bpr_step_kernel<G, C, 0, 2><<<grid, 256, 0, st>>>(a);
        else if (a.uniq_users) bpr_step_kernel<G, C, 0, 1><<<grid, 256, 0, st>>>(a);
        else bpr_step_kernel<G, C, 0, 0><<<grid, 256, 0, st>>>(a);
    }
}

// occurrences of every item in the train CSR (one-time, at pda_set_train_csr*: picks the popular items of StepArgs::hot_slot)
__global__ void __launch_bounds__(256) item_count_kernel(const int32_t* __restrict__ items, int64_t nnz, int32_t* __restrict__ cnt) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nnz; i += (int64_t)gridDim.x * blockDim.x)
        atomicAdd(cnt + items[i], 1);
}
void launch_item_count(const int32_t* items, int64_t nnz, int32_t* cnt, cudaStream_t st) {
    if (nnz > 0) item_count_kernel<<<148 * 8, 256, 0, st>>>(items, nnz, cnt);
}

int launch_bpr_step(const StepArgs& a, cudaStream_t st) {
    if (a.d % 4 != 0 || a.d < 4 || a.d > 512) return 1;
    if (launch_bpr_step_pipe(a, st) == 0) return 0;      // d = 128, distinct users: the bulk-copy pipeline (pda_step_pipe.cu)
    const int q = a.d / 4;
    int G = 1;
    while (G < q && G < 32) G *= 2;
    const int C = (q + G - 1) / G;
    // grid: enough warps to cover the batch once, capped at 8 CTAs x 148 SMs (grid-stride beyond that)
    static int per_sm = 0;
    if (!per_sm) { const char* e = getenv("PDA_STEP_GRID_PER_SM"); per_sm = e ? atoi(e) : 8; if (per_sm < 1) per_sm = 8; }
    int64_t warps_needed = (a.B + (32 / G) - 1) / (32 / G);
    int64_t blocks = (warps_needed + 7) / 8;
    if (blocks > 148 * per_sm) blocks = 148 * per_sm;
    if (blocks < 1) blocks = 1;
    const int grid = (int)blocks;
    switch (G) {
        case 1: launch_step_gc<1, 1>(a, grid, st); break;
        case 2: launch_step_gc<2, 1>(a, grid, st); break;
        case 4: launch_step_gc<4, 1>(a, grid, st); break;
        case 8: launch_step_gc<8, 1>(a, grid, st); break;
        case 16: launch_step_gc<16, 1>(a, grid, st); break;
        default:
            if (C == 1) launch_step_gc<32, 1>(a, grid, st);
            else if (C == 2) launch_step_gc<32, 2>(a, grid, st);
            else if (C == 3) launch_step_gc<32, 3>(a, grid, st);
            else launch_step_gc<32, 4>(a, grid, st);
    }
    return 0;
}

// ------------------------------------------------------------------------------------------
// a6. TF1 Adam on IndexedSlices = dense sweep (AdamOptimizer._apply_sparse_shared):
//   m <- b1 m + (1-b1) G ; v <- b2 v + (1-b2) G*G ; W <- W - lr_t m / (sqrt(v) + eps)
// over EVERY row (G is zero on untouched rows).  One launch sweeps both tables; G is consumed
// and zeroed for the next step.  lr_t is formed on the device from the fp32 beta-power
// variables so a captured step needs no host value.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float adam_elem(float& w, float& m, float& v, float g, float lr_t) {
    const float b1 = 0.9f, b2 = 0.999f, eps = 1e-8f;
    const float omb1 = fsub(1.0f, b1), omb2 = fsub(1.0f, b2);
    m = fadd(fmul(m, b1), fmul(g, omb1));
    v = fadd(fmul(v, b2), fmul(fmul(g, g), omb2));
    w = fsub(w, fdiv(fmul(lr_t, m), fadd(fsqrt(v), eps)));
    return w;
}

__global__ void __launch_bounds__(256) adam_dense_kernel(AdamArgs a) {
    const float b1p = a.pw[0], b2p = a.pw[1];
    const float lr_t = fdiv(fmul(a.lr, fsqrt(fsub(1.0f, b2p))), fsub(1.0f, b1p));
    const int64_t e1 = a.n4[0], e2 = e1 + a.n4[1], e3 = e2 + a.n4[2], n4 = e3 + a.n4[3];
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n4; e += (int64_t)gridDim.x * blockDim.x) {
        const int t = e < e1 ? 0 : e < e2 ? 1 : e < e3 ? 2 : 3;
        const int64_t k = e - (t == 0 ? 0 : t == 1 ? e1 : t == 2 ? e2 : e3);
        float4* W = reinterpret_cast<float4*>(a.W[t]) + k;
        float4* M = reinterpret_cast<float4*>(a.m[t]) + k;
        float4* V = reinterpret_cast<float4*>(a.v[t]) + k;
        float4* Gp = reinterpret_cast<float4*>(a.G[t]) + k;
        float4 w = *W, m = *M, v = *V, g = *Gp;
        adam_elem(w.x, m.x, v.x, g.x, lr_t);
        adam_elem(w.y, m.y, v.y, g.y, lr_t);
        adam_elem(w.z, m.z, v.z, g.z, lr_t);
        adam_elem(w.w, m.w, v.w, g.w, lr_t);
        *W = w; *M = m; *V = v;
        if (!a.keep_g) *Gp = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

void launch_adam_dense(const AdamArgs& a, cudaStream_t st) {
    int64_t n4 = a.n4[0] + a.n4[1] + a.n4[2] + a.n4[3];
    int64_t blocks = (n4 + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    adam_dense_kernel<<<(int)blocks, 256, 0, st>>>(a);
}

// Host batches are validated on the device in one pass over the ids: flags[1] is set when any id lies outside its
// table (the reference's embedding_lookup raises for those -- MF/model_api.py:51-53), and, when `seen` is given (an
// n_users scratch array of tags), flags[0] is set when a user repeats: a repeated user finds this call's tag already
// there.  The reference's batches hold distinct users (rd.sample, train_new_api.py:384-385), a host caller may pass anything.
__global__ void batch_check_kernel(const int32_t* __restrict__ users, const int32_t* __restrict__ pos, const int32_t* __restrict__ neg,
                                   int64_t B, int32_t n_users, int32_t n_items, int32_t* __restrict__ seen, int32_t tag,
                                   int32_t* flags) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < B; i += (int64_t)gridDim.x * blockDim.x) {
        const int32_t u = users[i], p = pos[i], n = neg[i];
        if ((uint32_t)u >= (uint32_t)n_users || (uint32_t)p >= (uint32_t)n_items || (uint32_t)n >= (uint32_t)n_items) { flags[1] = 1; continue; }
        if (seen && atomicExch(seen + u, tag) == tag) flags[0] = 1;
    }
}

void launch_batch_check(const int32_t* users, const int32_t* pos, const int32_t* neg, int64_t B, int32_t n_users,
                        int32_t n_items, int32_t* seen, int32_t tag, int32_t* flags, cudaStream_t st) {
    int64_t blocks = (B + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    batch_check_kernel<<<(int)blocks, 256, 0, st>>>(users, pos, neg, B, n_users, n_items, seen, tag, flags);
}

__global__ void fill_i32_kernel(int32_t* p, int64_t n, int32_t value) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = value;
}

void launch_fill_i32(int32_t* p, int64_t n, int32_t value, cudaStream_t st) {
    int64_t blocks = (n + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    fill_i32_kernel<<<(int)blocks, 256, 0, st>>>(p, n, value);
}

// stage labels arrive as fp32 in the pos_pop slot (tf.cast(temp, tf.int32), train_new_api.py:565); clamped to
// [0, max_value] so a bad label can never index outside item_temp_init_bias
__global__ void f32_to_i32_kernel(const float* __restrict__ src, int32_t* __restrict__ dst, int64_t n, int32_t max_value) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int32_t v = (int32_t)src[i];
        dst[i] = v < 0 ? 0 : v > max_value ? max_value : v;
    }
}

void launch_f32_to_i32(const float* src, int32_t* dst, int64_t n, int32_t max_value, cudaStream_t st) {
    int64_t blocks = (n + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    f32_to_i32_kernel<<<(int)blocks, 256, 0, st>>>(src, dst, n, max_value);
}

// BPR(t)-pop inference bias (model_api.py:373-387): (1 + ub[first user of the batch]) * (ib[:,T-1] + ib[:,T])
__global__ void temp_item_bias_kernel(const float* __restrict__ ub, const float* __restrict__ ib, int64_t n_items,
                                      int temp_num, int32_t first_user, float* __restrict__ out) {
    const float f = fadd(ub[first_user], 1.0f);
    const int Tc = temp_num + 1;
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n_items; j += (int64_t)gridDim.x * blockDim.x)
        out[j] = fmul(f, fadd(ib[j * Tc + temp_num - 1], ib[j * Tc + temp_num]));
}

void launch_temp_item_bias(const float* ub, const float* ib, int64_t n_items, int temp_num, int32_t first_user,
                           float* out, cudaStream_t st) {
    int64_t blocks = (n_items + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    temp_item_bias_kernel<<<(int)blocks, 256, 0, st>>>(ub, ib, n_items, temp_num, first_user, out);
}

// loss_sum[0..2] accumulate the fp32 step losses in double (the epoch means of train_new_api.py:1095-1097),
// loss_sum[3] counts the steps.
__global__ void finish_step_kernel(double* loss_acc, float* loss3, double* loss_sum, float* pw, double B, double regs,
                                   double batch_size, int advance_powers, float lr, float* lr_slot) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        // lr_t of the step that just ran (same expression as the Adam kernels), kept for the lazy replay
        if (advance_powers && lr_slot) *lr_slot = fdiv(fmul(lr, fsqrt(fsub(1.0f, pw[1]))), fsub(1.0f, pw[0]));
        double mf = -loss_acc[0] / B;
        double reg = regs * 0.5 * loss_acc[1] / batch_size;
        loss3[0] = (float)(mf + reg); loss3[1] = (float)mf; loss3[2] = (float)reg;
        loss_sum[0] += (double)loss3[0]; loss_sum[1] += (double)loss3[1]; loss_sum[2] += (double)loss3[2];
        loss_sum[3] += 1.0;
        loss_acc[0] = 0.0; loss_acc[1] = 0.0;
        if (advance_powers) { pw[0] = fmul(pw[0], 0.9f); pw[1] = fmul(pw[1], 0.999f); }
    }
}

// loss3 = {loss, mf, reg}; beta powers advance (AdamOptimizer._finish); accumulators reset.
void launch_finish_step(double* loss_acc, float* loss3, double* loss_sum, float* pw, int64_t B, float regs,
                        int batch_size, int advance_powers, float lr, float* lr_slot, cudaStream_t st) {
    finish_step_kernel<<<1, 32, 0, st>>>(loss_acc, loss3, loss_sum, pw, (double)B, (double)regs, (double)batch_size,
                                         advance_powers, lr, lr_slot);
}

}  // namespace pda
