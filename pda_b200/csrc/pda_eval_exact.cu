// Exact all-items scoring + popularity adjust + train-item mask + top-K on CUDA cores.
//
// Reference call sites: MF/train_new_api.py:594-612 (Create_Recommendation), :614-640
// (do_recommendation), :642-669 (testing), MF/model_api.py:62 (batch_ratings = U_b I^T),
// :113 (condition_ratings = (elu(batch_ratings)+1) * pop).
//
// This kernel is the exact-score spec of the library: s[r][j] is the sequential-k fp32 sum
// acc = acc + u[k]*v[k] (two roundings per term), the order oracle/csrc/pda_oracle.c:orc_recommend
// restates, so ids and scores are bit-identical to the oracle.  The tensor-core path
// (pda_eval_tc.cu) only *filters* candidates; every reported score comes from this arithmetic.
//
// Layout: one CTA = 64 users x all items, swept in 128-item tiles.  256 threads hold a 4x8
// register block each; operands are staged k-transposed in shared memory in chunks of 32 k.
// The finished 64x128 tile goes to shared memory, is transformed / masked there, and each warp
// keeps the running top-K lists of its 8 rows (sorted, in shared memory) with ballot-driven
// warp-cooperative insertion.  The score matrix never reaches HBM (unless dense_out is asked
// for, which is the reference's testing()/predict() contract).
#include "pda_kernels.h"

namespace pda {

constexpr int E_TM = 64, E_TN = 128, E_KC = 32, E_NT = 256;
constexpr int E_SLD = E_TN + 1;

__device__ __forceinline__ bool better(float ya, int ia, float yb, int ib) { return ya > yb || (ya == yb && ia < ib); }

// Insert (y, j) into the sorted list val/id[0..K) (best first), dropping the last entry.
__device__ __forceinline__ void warp_insert(float* val, int* id, int K, int Kp, float y, int j, int lane) {
    int pos = 0;
    for (int t = 0; t < Kp; t += 32) {
        const int e = lane + t;
        const bool b = e < K && better(val[e], id[e], y, j);
        pos += __popc(__ballot_sync(0xffffffffu, b));
    }
    float ov[4]; int oi[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const int e = lane + 32 * t;
        if (32 * t < Kp && e > pos && e < K) { ov[t] = val[e - 1]; oi[t] = id[e - 1]; }
    }
    __syncwarp();
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const int e = lane + 32 * t;
        if (32 * t < Kp) {
            if (e > pos && e < K) { val[e] = ov[t]; id[e] = oi[t]; }
            else if (e == pos) { val[e] = y; id[e] = j; }
        }
    }
    __syncwarp();
}

template <int MODE>
__global__ void __launch_bounds__(E_NT, 2) recommend_exact_kernel(EvalArgs a, int Kp) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* Us = reinterpret_cast<float*>(smem_raw);                  // [E_KC][E_TM]
    float* Is = Us + E_KC * E_TM;                                    // [E_KC][E_TN]
    float* S = Is + E_KC * E_TN;                                     // [E_TM][E_SLD]
    float* tval = S + E_TM * E_SLD;                                  // [E_TM][Kp]
    int* tid_ = reinterpret_cast<int*>(tval + E_TM * Kp);            // [E_TM][Kp]
    unsigned* mbits = reinterpret_cast<unsigned*>(tid_ + E_TM * Kp); // [E_TM][4]
    long long* mcur = reinterpret_cast<long long*>(mbits + E_TM * 4);// [E_TM]
    int* urow = reinterpret_cast<int*>(mcur + E_TM);                 // [E_TM] global user id (-1 = padding)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ty = tid >> 4, tx = tid & 15;
    const int64_t m0 = (int64_t)blockIdx.x * E_TM;
    const int K = a.K, d = a.d;
    const int64_t Meff = a.M_dev ? (int64_t)*a.M_dev : a.M;    // device-sized launch (tensor path fallback rows)
    if (m0 >= Meff) return;
    // few flagged rows: one CTA per (64-row block, item-range split), merged afterwards; many: one CTA per block
    const bool split = a.n_split > 1;
    if (split ? Meff > a.split_max_rows : (a.M_dev && Meff <= a.skip_rows_le)) return;
    int64_t j_begin = 0, j_end = a.N;
    if (split) {
        const int64_t span = ((a.N + E_TN - 1) / E_TN + a.n_split - 1) / a.n_split * E_TN;
        j_begin = (int64_t)blockIdx.y * span;
        j_end = j_begin + span < a.N ? j_begin + span : a.N;
    }

    for (int r = tid; r < E_TM; r += E_NT) {
        const int64_t m = m0 + r;
        const int u = m < Meff ? a.users[m] : -1;
        urow[r] = u;
        mcur[r] = (u >= 0 && a.mask_indptr) ? a.mask_indptr[u] : 0;
    }
    for (int e = tid; e < E_TM * Kp; e += E_NT) { tval[e] = -INFINITY; tid_[e] = 0x7fffffff; }
    __syncthreads();

    for (int64_t j0 = j_begin; j0 < j_end; j0 += E_TN) {
        float acc[4][8];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[r][c] = 0.0f;

        for (int kc = 0; kc < d; kc += E_KC) {
            // stage U chunk (64 rows x 32 k) and I chunk (128 rows x 32 k), k-transposed
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                const int idx = tid + E_NT * t, row = idx & (E_TM - 1), k4 = idx >> 6;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                const int u = urow[row];
                if (u >= 0 && kc + 4 * k4 < d) v = ldg_f4(a.U + (int64_t)u * d + kc + 4 * k4);
                Us[(4 * k4 + 0) * E_TM + row] = v.x; Us[(4 * k4 + 1) * E_TM + row] = v.y;
                Us[(4 * k4 + 2) * E_TM + row] = v.z; Us[(4 * k4 + 3) * E_TM + row] = v.w;
            }
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int idx = tid + E_NT * t, row = idx & (E_TN - 1), k4 = idx >> 7;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                const int64_t j = j0 + row;
                if (j < j_end && kc + 4 * k4 < d) v = ldg_f4(a.I + j * d + kc + 4 * k4);
                Is[(4 * k4 + 0) * E_TN + row] = v.x; Is[(4 * k4 + 1) * E_TN + row] = v.y;
                Is[(4 * k4 + 2) * E_TN + row] = v.z; Is[(4 * k4 + 3) * E_TN + row] = v.w;
            }
            __syncthreads();
            const int kmax = d - kc < E_KC ? d - kc : E_KC;
#pragma unroll 4
            for (int k = 0; k < kmax; ++k) {
                const float4 av = *reinterpret_cast<const float4*>(Us + k * E_TM + ty * 4);
                const float4 b0 = *reinterpret_cast<const float4*>(Is + k * E_TN + tx * 4);
                const float4 b1 = *reinterpret_cast<const float4*>(Is + k * E_TN + 64 + tx * 4);
                const float ar[4] = {av.x, av.y, av.z, av.w};
                const float bc[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int c = 0; c < 8; ++c) acc[r][c] = fadd(acc[r][c], fmul(ar[r], bc[c]));
            }
            __syncthreads();
        }
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 8; ++c) S[(ty * 4 + r) * E_SLD + (c < 4 ? tx * 4 + c : 64 + tx * 4 + c - 4)] = acc[r][c];
        __syncthreads();

        // per-warp: rows warp*8 .. +7
        for (int rr = 0; rr < 8; ++rr) {
            const int r = warp * 8 + rr;
            const int u = urow[r];
            if (u < 0) continue;   // warp-uniform
            if (lane < 4) mbits[r * 4 + lane] = 0u;
            __syncwarp();
            if (a.mask_indptr) {
                long long cur = mcur[r];
                const long long end = a.mask_indptr[u + 1];
                while (true) {
                    const long long qi = cur + lane;
                    const int it = qi < end ? __ldg(a.mask_items + qi) : 0x7fffffff;
                    const bool inr = (int64_t)it < j0 + E_TN;
                    if (inr && it >= j0) atomicOr(&mbits[r * 4 + ((it - (int)j0) >> 5)], 1u << ((it - (int)j0) & 31));
                    const int cnt = __popc(__ballot_sync(0xffffffffu, inr));
                    cur += cnt;
                    if (cnt < 32) break;
                }
                __syncwarp();
                if (lane == 0) mcur[r] = cur;
            }
            float* val = tval + r * Kp;
            int* idl = tid_ + r * Kp;
            const int64_t m = m0 + r;
#pragma unroll
            for (int c = 0; c < E_TN / 32; ++c) {
                const int64_t j = j0 + c * 32 + lane;
                const bool valid = j < j_end;
                float y = -INFINITY;
                if (valid) {
                    const float s = S[r * E_SLD + c * 32 + lane];
                    if (MODE == 1) y = fmul(elu_p1(s), __ldg(a.pop + j));
                    else y = a.col_bias ? fadd(s, __ldg(a.col_bias + j)) : s;
                    if (a.dense_out) a.dense_out[m * a.N + j] = y;
                    if ((mbits[r * 4 + c] >> lane) & 1u) y = -INFINITY;
                }
                float tv = val[K - 1]; int ti = idl[K - 1];
                unsigned bal = __ballot_sync(0xffffffffu, valid && better(y, (int)j, tv, ti));
                while (bal) {
                    const int src = __ffs(bal) - 1;
                    bal &= bal - 1;
                    const float yy = __shfl_sync(0xffffffffu, y, src);
                    const int jj = (int)j0 + c * 32 + src;
                    tv = val[K - 1]; ti = idl[K - 1];
                    if (better(yy, jj, tv, ti)) warp_insert(val, idl, K, Kp, yy, jj, lane);
                }
            }
        }
        __syncthreads();
    }
    // write results
    for (int e = tid; e < E_TM * K; e += E_NT) {
        const int r = e / K, k = e - r * K;
        const int64_t m = m0 + r;
        if (m < Meff) {
            const int idv = tid_[r * Kp + k];
            if (split) {
                a.part_ids[(m * a.n_split + blockIdx.y) * K + k] = idv;
                a.part_val[(m * a.n_split + blockIdx.y) * K + k] = tval[r * Kp + k];
                continue;
            }
            const int64_t mo = a.out_rows ? (int64_t)a.out_rows[m] : m;
            if (a.ids_out) a.ids_out[mo * K + k] = idv == 0x7fffffff ? -1 : idv;
            if (a.scores_out) a.scores_out[mo * K + k] = tval[r * Kp + k];
        }
    }
}

// final top-K of a row from its n_split partial lists (each sorted, (score desc, id asc); unfilled slots = (-inf, 0x7fffffff)):
// one CTA per row, bitonic sort of n_split * K 64-bit keys in shared memory
__global__ void __launch_bounds__(256) recommend_merge_kernel(EvalArgs a) {
    extern __shared__ unsigned long long mkeys[];
    const int64_t Meff = a.M_dev ? (int64_t)*a.M_dev : a.M;
    const int64_t m = blockIdx.x;
    if (m >= Meff || Meff > a.split_max_rows) return;
    const int K = a.K, n = a.n_split * K;
    int n2 = 256;
    while (n2 < n) n2 <<= 1;
    for (int e = threadIdx.x; e < n2; e += 256) {
        unsigned long long kv = 0ull;
        if (e < n) {
            const float y = a.part_val[m * n + e];
            const uint32_t b = __float_as_uint(y);
            const uint32_t key = (b & 0x80000000u) ? ~b : (b | 0x80000000u);             // order-preserving; -inf -> 0x007fffff > 0
            kv = ((unsigned long long)key << 32) | (uint32_t)(0x7fffffff - a.part_ids[m * n + e]);
        }
        mkeys[e] = kv;
    }
    __syncthreads();
    for (int k = 2; k <= n2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int idx = threadIdx.x; idx < n2; idx += 256) {
                const int ixj = idx ^ j;
                if (ixj > idx) {
                    const unsigned long long x = mkeys[idx], y = mkeys[ixj];
                    const bool desc = (idx & k) == 0;
                    if (desc ? x < y : x > y) { mkeys[idx] = y; mkeys[ixj] = x; }
                }
            }
            __syncthreads();
        }
    const int64_t mo = a.out_rows ? (int64_t)a.out_rows[m] : m;
    for (int e = threadIdx.x; e < K; e += 256) {
        const unsigned long long kv = mkeys[e];
        const int idv = 0x7fffffff - (int32_t)(uint32_t)(kv & 0xffffffffu);
        const uint32_t key = (uint32_t)(kv >> 32);
        const float y = __uint_as_float((key & 0x80000000u) ? (key & 0x7fffffffu) : ~key);
        if (a.ids_out) a.ids_out[mo * K + e] = idv == 0x7fffffff ? -1 : idv;
        if (a.scores_out) a.scores_out[mo * K + e] = y;
    }
}

int launch_recommend_exact(const EvalArgs& a, cudaStream_t st) {
    if (a.d % 4 != 0 || a.K < 1 || a.K > 128 || a.M < 1) return 1;
    const int Kp = (a.K + 31) / 32 * 32;
    const size_t smem = sizeof(float) * (E_KC * E_TM + E_KC * E_TN + E_TM * E_SLD) + (size_t)E_TM * Kp * 8 +
                        E_TM * 4 * 4 + E_TM * 8 + E_TM * 4;
    int64_t rows = a.M;
    if (a.n_split > 1 && rows > a.split_max_rows) rows = a.split_max_rows;       // beyond that the plain form runs
    const dim3 grid((unsigned)((rows + E_TM - 1) / E_TM), (unsigned)(a.n_split > 1 ? a.n_split : 1));
    if (a.mode == 1) {
        cudaFuncSetAttribute(recommend_exact_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        recommend_exact_kernel<1><<<grid, E_NT, smem, st>>>(a, Kp);
    } else {
        cudaFuncSetAttribute(recommend_exact_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        recommend_exact_kernel<0><<<grid, E_NT, smem, st>>>(a, Kp);
    }
    if (a.n_split > 1) {
        int n2 = 256;
        while (n2 < a.n_split * a.K) n2 <<= 1;
        if (cudaFuncSetAttribute(recommend_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, n2 * 8) != cudaSuccess) return 2;
        recommend_merge_kernel<<<(unsigned)rows, 256, (size_t)n2 * 8, st>>>(a);
    }
    return 0;
}

// ------------------------------------------------------------------------------------------
// a11. metrics from top-K ids (MF/used_metric.py:39-80): one warp per eval user; sums over
// users of precision / recall / ndcg / hit at each K (the caller divides by the user count).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) metrics_kernel(const int32_t* __restrict__ ids, int64_t M, int Kkeep,
                                                      const int32_t* __restrict__ eval_users,
                                                      const int64_t* __restrict__ truth_indptr,
                                                      const int32_t* __restrict__ truth_items,
                                                      const int32_t* __restrict__ Ks, int nK, double* out) {
    const int lane = threadIdx.x & 31;
    const int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (r >= M) return;
    const int u = eval_users[r];
    const int64_t lo = truth_indptr[u], hi = truth_indptr[u + 1];
    const int64_t npos = hi - lo;
    for (int qk = 0; qk < nK; ++qk) {
        const int K = Ks[qk] < Kkeep ? Ks[qk] : Kkeep;
        double hits = 0.0, dcg = 0.0, idcg = 0.0;
        for (int k = lane; k < K; k += 32) {
            const int id = ids[r * Kkeep + k];
            bool hit = false;
            for (int64_t z = lo; z < hi; ++z) if (truth_items[z] == id) { hit = true; break; }
            const double tp = 1.0 / log2((double)(k + 2));
            if (hit) { hits += 1.0; dcg += tp; }
            if (k < npos) idcg += tp;
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
            hits += __shfl_xor_sync(0xffffffffu, hits, off);
            dcg += __shfl_xor_sync(0xffffffffu, dcg, off);
            idcg += __shfl_xor_sync(0xffffffffu, idcg, off);
        }
        if (lane == 0) {
            atomicAdd(out + 0 * nK + qk, hits / K);
            atomicAdd(out + 1 * nK + qk, npos ? hits / (double)npos : 0.0);
            atomicAdd(out + 2 * nK + qk, idcg > 0 ? dcg / idcg : 0.0);
            atomicAdd(out + 3 * nK + qk, hits > 1.0 ? 1.0 : hits);
        }
    }
}

void launch_metrics(const int32_t* ids, int64_t M, int Kkeep, const int32_t* eval_users, const int64_t* truth_indptr,
                    const int32_t* truth_items, const int32_t* Ks, int nK, double* out, cudaStream_t st) {
    const int64_t threads = M * 32;
    metrics_kernel<<<(int)((threads + 127) / 128), 128, 0, st>>>(ids, M, Kkeep, eval_users, truth_indptr, truth_items,
                                                                  Ks, nK, out);
}

}  // namespace pda
