// Exact lazy evaluation of TF1's dense Adam sweep (reference: MF/model_api.py:83,471 ->
// tf.train.AdamOptimizer._apply_sparse_shared of TF 1.14).
//
// TF1's Adam on IndexedSlices is NOT lazy: at every step it decays m and v and moves the variable for EVERY row of
// the table, sampled or not.  For a row whose gradient is zero at step s that update is
//     m <- m*b1 ;  v <- v*b2 ;  w <- w - (lr_s * m) / (sqrt(v) + eps)
// -- a function of the row's own (w, m, v) and the scalar lr_s only.  So instead of sweeping the whole table every
// step (34 GB of traffic per step on the 10M x 1M x 128 synthetic set), a row is brought up to date when it is next
// needed, by REPLAYING the skipped steps in registers with exactly the fp32 operations the dense sweep would have
// executed, in the same order, with the same per-step lr_s (kept in lr_hist[]).  The result is bit-identical to the
// dense sweep (tests/test_gpu_train.py::test_lazy_adam_*); only the memory traffic differs.
//
//   applied[row]  number of Adam steps already applied to the row (0 .. step_no)
//   stamp[row]    claim word: 2*(step+1)+phase of the last kernel phase that owned the row (0 = never touched:
//                 m = v = 0, every skipped step is the identity, nothing to replay)
//
// Per step t:  catch-up  rows of the batch: replay steps applied[row] .. t-1           (before the forward pass)
//              fused BPR step kernel (unchanged)                                        -> G
//              apply     rows of the batch: step t with the summed gradient G, G <- 0, applied[row] = t+1
// Duplicate rows inside a batch are claimed once with atomicExch on stamp[row].
// Flush (before eval / table read-out / dense phases): every row replays up to step_no.
#include "pda_kernels.h"

namespace pda {

// PHASE 0: catch-up to step t.  PHASE 1: apply step t with gradient.  One group of G lanes per batch entry
// (entry e < B: user row, B <= e < 2B: pos item row, else neg item row), C float4 chunks per lane.
template <int G, int C, int PHASE>
__global__ void __launch_bounds__(256, 6) adam_lazy_rows_kernel(LazyArgs a) {
    constexpr int GPW = 32 / G;
    const int lane = threadIdx.x & 31, gl = lane % G, gw = lane / G;
    const int64_t warp_global = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int q = a.d >> 2;
    const int64_t t = a.step_no;
    const int32_t claim = (int32_t)(2 * (t + 1) + PHASE);
    const bool lr_ok = lr_in_replay_range(a.lr);
    float lr_t = 0.f;
    if (PHASE == 1) {
        const float b1p = a.pw[0], b2p = a.pw[1];
        lr_t = fdiv(fmul(a.lr, fsqrt(fsub(1.0f, b2p))), fsub(1.0f, b1p));
    }
    const int64_t n_entries = 3 * a.B;
    const int64_t e_begin = a.lazy[0] ? 0 : a.B;      // user entries are skipped when the step kernel owns the user rows
    unsigned long long stat = 0;   // per-lane tally, reduced once per warp at the end (one same-address atomic per warp)
    for (int64_t base = e_begin + warp_global * GPW; base < n_entries; base += n_warps * GPW) {
        const int64_t e = base + gw;
        const bool valid = e < n_entries;
        int tbl = 0;
        int64_t row = 0;
        if (valid) {
            tbl = e < a.B ? 0 : 1;
            row = e < a.B ? __ldg(a.users + e) : (e < 2 * a.B ? __ldg(a.pos + (e - a.B)) : __ldg(a.neg + (e - 2 * a.B)));
        }
        const bool lazy_tbl = valid && a.lazy[tbl];
        int32_t old = claim;
        if (lazy_tbl && gl == 0) old = atomicExch(a.stamp[tbl] + row, claim);     // first claimant of the row owns it
        old = __shfl_sync(0xffffffffu, old, gw * G);
        const bool mine = lazy_tbl && old != claim;
        if (!mine) continue;
        int32_t* ap = a.applied[tbl] + row;
        // read by the group's first lane only and broadcast: that lane rewrites applied[row] below, and lanes of a
        // group are not guaranteed to stay in lockstep through the replay
        int32_t done32 = 0;
        if (gl == 0) done32 = *ap;
        const uint32_t gmask = G == 32 ? 0xffffffffu : ((1u << G) - 1u) << (gw * G);   // `mine` is uniform within a group
        const int64_t done = __shfl_sync(gmask, done32, gw * G);
        float* Wr = a.W[tbl] + row * a.d;
        float* Mr = a.m[tbl] + row * a.d;
        float* Vr = a.v[tbl] + row * a.d;
        if (PHASE == 0) {
            if (gl == 0 && done < t && old != 0) stat += (unsigned long long)(t - done);
            if (done < t && old != 0) {   // old == 0: never touched, m = v = 0, nothing to replay
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const int ch = gl + G * c;
                    if (ch < q) {
                        float4 w = *reinterpret_cast<float4*>(Wr + 4 * ch), m = *reinterpret_cast<float4*>(Mr + 4 * ch),
                               v = *reinterpret_cast<float4*>(Vr + 4 * ch);
                        lazy_replay4_blocked(w, m, v, a.lr_hist, done, t, lr_ok);
                        *reinterpret_cast<float4*>(Wr + 4 * ch) = w;
                        *reinterpret_cast<float4*>(Mr + 4 * ch) = m;
                        *reinterpret_cast<float4*>(Vr + 4 * ch) = v;
                    }
                }
            }
            if (gl == 0) *ap = (int32_t)t;
        } else {
            float* Gr = a.G[tbl] + row * a.d;
            if (gl == 0) stat += 1ull;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const int ch = gl + G * c;
                if (ch < q) {
                    float4 w = *reinterpret_cast<float4*>(Wr + 4 * ch), m = *reinterpret_cast<float4*>(Mr + 4 * ch),
                           v = *reinterpret_cast<float4*>(Vr + 4 * ch), g = *reinterpret_cast<float4*>(Gr + 4 * ch);
                    lazy_grad_step(w.x, m.x, v.x, g.x, lr_t);
                    lazy_grad_step(w.y, m.y, v.y, g.y, lr_t);
                    lazy_grad_step(w.z, m.z, v.z, g.z, lr_t);
                    lazy_grad_step(w.w, m.w, v.w, g.w, lr_t);
                    *reinterpret_cast<float4*>(Wr + 4 * ch) = w;
                    *reinterpret_cast<float4*>(Mr + 4 * ch) = m;
                    *reinterpret_cast<float4*>(Vr + 4 * ch) = v;
                    *reinterpret_cast<float4*>(Gr + 4 * ch) = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            if (gl == 0) *ap = (int32_t)(t + 1);
        }
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) stat += __shfl_xor_sync(0xffffffffu, stat, off);
    if (lane == 0 && stat) atomicAdd(a.stats + (PHASE == 0 ? 1 : 0), stat);
}

template <int G, int C>
static void launch_rows_gc(const LazyArgs& a, int phase, int grid, cudaStream_t st) {
    if (phase == 0) adam_lazy_rows_kernel<G, C, 0><<<grid, 256, 0, st>>>(a);
    else adam_lazy_rows_kernel<G, C, 1><<<grid, 256, 0, st>>>(a);
}

int launch_adam_lazy_rows(const LazyArgs& a, int phase, cudaStream_t st) {
    if (a.d % 4 != 0 || a.d < 4 || a.d > 512) return 1;
    const int q = a.d / 4;
    int G = 1;
    while (G < q && G < 32) G *= 2;
    const int C = (q + G - 1) / G;
    const int64_t warps_needed = (3 * a.B + (32 / G) - 1) / (32 / G);
    int64_t blocks = (warps_needed + 7) / 8;
    if (blocks > 148 * 6) blocks = 148 * 6;      // one resident wave under the launch bounds
    if (blocks < 1) blocks = 1;
    const int grid = (int)blocks;
    switch (G) {
        case 1: launch_rows_gc<1, 1>(a, phase, grid, st); break;
        case 2: launch_rows_gc<2, 1>(a, phase, grid, st); break;
        case 4: launch_rows_gc<4, 1>(a, phase, grid, st); break;
        case 8: launch_rows_gc<8, 1>(a, phase, grid, st); break;
        case 16: launch_rows_gc<16, 1>(a, phase, grid, st); break;
        default:
            if (C == 1) launch_rows_gc<32, 1>(a, phase, grid, st);
            else if (C == 2) launch_rows_gc<32, 2>(a, phase, grid, st);
            else if (C == 3) launch_rows_gc<32, 3>(a, phase, grid, st);
            else launch_rows_gc<32, 4>(a, phase, grid, st);
    }
    return 0;
}

// flush: every row of table `tbl` replays up to step_no.  One warp per row (grid-stride), lane-strided float4 chunks.
__global__ void __launch_bounds__(256) adam_lazy_flush_kernel(LazyArgs a, int tbl, int64_t n_rows) {
    const int lane = threadIdx.x & 31;
    const int64_t warp_global = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int q = a.d >> 2;
    const int64_t t = a.step_no;
    const bool lr_ok = lr_in_replay_range(a.lr);
    for (int64_t row = warp_global; row < n_rows; row += n_warps) {
        int32_t* ap = a.applied[tbl] + row;
        const int64_t done = *ap;
        if (done >= t) continue;
        if (a.stamp[tbl][row] != 0) {
            float* Wr = a.W[tbl] + row * a.d;
            float* Mr = a.m[tbl] + row * a.d;
            float* Vr = a.v[tbl] + row * a.d;
            for (int ch = lane; ch < q; ch += 32) {
                float4 w = *reinterpret_cast<float4*>(Wr + 4 * ch), m = *reinterpret_cast<float4*>(Mr + 4 * ch),
                       v = *reinterpret_cast<float4*>(Vr + 4 * ch);
                lazy_replay4_blocked(w, m, v, a.lr_hist, done, t, lr_ok);
                *reinterpret_cast<float4*>(Wr + 4 * ch) = w;
                *reinterpret_cast<float4*>(Mr + 4 * ch) = m;
                *reinterpret_cast<float4*>(Vr + 4 * ch) = v;
            }
        }
        __syncwarp();
        if (lane == 0) *ap = (int32_t)t;
    }
}

void launch_adam_lazy_flush(const LazyArgs& a, int tbl, int64_t n_rows, cudaStream_t st) {
    int64_t blocks = (n_rows + 7) / 8;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    adam_lazy_flush_kernel<<<(int)blocks, 256, 0, st>>>(a, tbl, n_rows);
}

}  // namespace pda
