// NeuRec-style native evaluator entry points over the GPU (SURVEY 8b "Native evaluator FFI", row b4):
//   cpp_evaluate_matrix  evaluator/backend/cpp/include/evaluate.h:53  (per user: top-k of a dense rating row, then
//                        Precision / Recall / MAP / NDCG / MRR cumulative over 1..k, metric.h:17-117)
//   arg_top_k_2d         util/cython/include/arg_topk.h:29            (row-wise arg-top-k of a dense matrix)
// Same buffers and meaning as the reference's functions (C-contiguous float32 host matrices, results written into
// caller-owned host arrays), exported with C linkage; the std::vector<std::unordered_set<int>> of test items becomes a CSR.
// Order among equal scores: lower index first (the reference's std::partial_sort_copy leaves ties unspecified).
#include <math.h>

#include "../../include/pda_b200.h"
#include "pda_kernels.h"

namespace pda {

__device__ __forceinline__ uint32_t nr_key(float f) { uint32_t b = __float_as_uint(f); return (b & 0x80000000u) ? ~b : (b | 0x80000000u); }

// one CTA per row: 4-pass radix select of the k-th largest key, then the winners sorted by (score desc, index asc)
__global__ void __launch_bounds__(256) dense_topk_kernel(const float* __restrict__ scores, int cols, int rows, int k, int32_t* __restrict__ out) {
    __shared__ int hist[256];
    __shared__ uint32_t s_prefix, s_mask;
    __shared__ int s_need, s_n;
    __shared__ unsigned long long win[256];
    const int row = blockIdx.x, tid = threadIdx.x;
    const float* r = scores + (size_t)row * cols;
    if (tid == 0) { s_prefix = 0; s_mask = 0; s_need = k; s_n = 0; }
    __syncthreads();
    for (int pass = 3; pass >= 0; --pass) {
        hist[tid] = 0;
        __syncthreads();
        const uint32_t prefix = s_prefix, mask = s_mask;
        for (int j = tid; j < cols; j += 256) {
            const uint32_t key = nr_key(r[j]);
            if ((key & mask) == prefix) atomicAdd(&hist[(key >> (pass * 8)) & 255u], 1);
        }
        __syncthreads();
        if (tid == 0) {
            int need = s_need, run = 0, bin = 255;
            for (; bin > 0; --bin) { if (run + hist[bin] >= need) break; run += hist[bin]; }
            s_need = need - run;
            s_prefix = prefix | ((uint32_t)bin << (pass * 8));
            s_mask = mask | (255u << (pass * 8));
        }
        __syncthreads();
    }
    const uint32_t kth = s_prefix;
    // strictly greater keys all win; equal keys win in ascending index order until k are taken (need_eq of them)
    const int need_eq = s_need;
    for (int j = tid; j < cols; j += 256) {
        const uint32_t key = nr_key(r[j]);
        if (key > kth) { const int p = atomicAdd(&s_n, 1); if (p < 256) win[p] = ((unsigned long long)key << 32) | (uint32_t)(0x7fffffff - j); }
    }
    __syncthreads();
    const int n_gt = s_n;
    __syncthreads();
    if (tid == 0) {      // the equal ones, serially in index order (ties at the k-th score are rare and short)
        int taken = 0;
        for (int j = 0; j < cols && taken < need_eq; ++j)
            if (nr_key(r[j]) == kth) { if (n_gt + taken < 256) win[n_gt + taken] = ((unsigned long long)kth << 32) | (uint32_t)(0x7fffffff - j); ++taken; }
        s_n = n_gt + taken;
    }
    __syncthreads();
    const int n = s_n < 256 ? s_n : 256;
    for (int i = n + tid; i < 256; i += 256) win[i] = 0ull;
    if (tid >= n) win[tid] = 0ull;
    __syncthreads();
    for (int kk = 2; kk <= 256; kk <<= 1)
        for (int jj = kk >> 1; jj > 0; jj >>= 1) {
            const int ixj = tid ^ jj;
            if (ixj > tid) {
                const unsigned long long x = win[tid], y = win[ixj];
                const bool desc = (tid & kk) == 0;
                if (desc ? x < y : x > y) { win[tid] = y; win[ixj] = x; }
            }
            __syncthreads();
        }
    if (tid < k) out[(size_t)row * k + tid] = tid < n ? 0x7fffffff - (int32_t)(uint32_t)(win[tid] & 0xffffffffu) : -1;
}

// metric ids of metric.h:109-114: 1 precision, 2 recall, 3 ap (MAP), 4 ndcg, 5 mrr; one thread per (user, metric)
__global__ void neurec_metrics_kernel(const int32_t* __restrict__ topk, int n_users, int k, const int64_t* __restrict__ tptr,
                                      const int32_t* __restrict__ titems, const int32_t* __restrict__ metric, int n_metric,
                                      float* __restrict__ res) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= (int64_t)n_users * n_metric) return;
    const int u = (int)(t / n_metric), mi = (int)(t % n_metric);
    const int32_t* rank = topk + (size_t)u * k;
    const int64_t lo = tptr[u], hi = tptr[u + 1];
    const double truth_len = (double)(hi - lo);
    float* out = res + ((size_t)u * n_metric + mi) * k;
    auto hit = [&](int32_t id) { for (int64_t z = lo; z < hi; ++z) if (titems[z] == id) return true; return false; };
    const int which = metric[mi];
    int hits = 0;
    float sum_pre = 0.f, dcg = 0.f, idcg = 0.f;
    bool found = false;
    float rr = 0.f;
    for (int i = 0; i < k; ++i) {
        const bool h = hit(rank[i]);
        if (h) hits += 1;
        if (which == 1) out[i] = (float)(1.0 * hits / (i + 1));
        else if (which == 2) out[i] = (float)(1.0 * hits / truth_len);
        else if (which == 3) {
            if (h) sum_pre += (float)(1.0 * hits / (i + 1));
            const float tl = (float)truth_len, denom = tl < (float)(i + 1) ? tl : (float)(i + 1);
            out[i] = hits == 0 ? 0.0f : sum_pre / denom;
        } else if (which == 4) {
            if (h) dcg = (float)((double)dcg + 1.0 / log2((double)(i + 2)));
            if ((double)i < truth_len) idcg = (float)((double)idcg + 1.0 / log2((double)(i + 2)));
            out[i] = dcg / idcg;
        } else {
            if (!found && h) { found = true; rr = (float)(1.0 / (i + 1)); }
            out[i] = found ? rr : 0.0f;
        }
    }
}

}  // namespace pda

using namespace pda;

static int neurec_fail(const char* what) { (void)what; return PDA_ERR_CUDA; }

extern "C" int pda_arg_top_k_2d_host(const float* scores, int32_t cols, int32_t rows, int32_t top_k, int32_t* results) {
    if (!scores || !results || cols < 1 || rows < 1 || top_k < 1 || top_k > 128 || top_k > cols) return PDA_ERR_ARG;
    float* d_s = nullptr; int32_t* d_o = nullptr;
    if (cudaMalloc((void**)&d_s, (size_t)rows * cols * 4) != cudaSuccess) return neurec_fail("malloc");
    if (cudaMalloc((void**)&d_o, (size_t)rows * top_k * 4) != cudaSuccess) { cudaFree(d_s); return neurec_fail("malloc"); }
    cudaMemcpy(d_s, scores, (size_t)rows * cols * 4, cudaMemcpyHostToDevice);
    dense_topk_kernel<<<rows, 256>>>(d_s, cols, rows, top_k, d_o);
    const cudaError_t e = cudaMemcpy(results, d_o, (size_t)rows * top_k * 4, cudaMemcpyDeviceToHost);
    cudaFree(d_s); cudaFree(d_o);
    return e == cudaSuccess ? PDA_OK : neurec_fail("kernel");
}

extern "C" int pda_evaluate_matrix_host(const float* rating_matrix, int32_t rating_len, int32_t n_users, const int64_t* truth_indptr,
                                        const int32_t* truth_items, const int32_t* metric, int32_t n_metric, int32_t top_k,
                                        float* results) {
    if (!rating_matrix || !truth_indptr || !metric || !results || rating_len < 1 || n_users < 1 || n_metric < 1 || top_k < 1 ||
        top_k > 128 || top_k > rating_len)
        return PDA_ERR_ARG;
    for (int i = 0; i < n_metric; ++i) if (metric[i] < 1 || metric[i] > 5) return PDA_ERR_ARG;
    const int64_t nnz = truth_indptr[n_users];
    float *d_s = nullptr, *d_r = nullptr; int32_t *d_o = nullptr, *d_ti = nullptr, *d_m = nullptr; int64_t* d_tp = nullptr;
    cudaError_t e = cudaMalloc((void**)&d_s, (size_t)n_users * rating_len * 4);
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_o, (size_t)n_users * top_k * 4);
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_tp, ((size_t)n_users + 1) * 8);
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_ti, (size_t)(nnz > 0 ? nnz : 1) * 4);
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_m, (size_t)n_metric * 4);
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_r, (size_t)n_users * n_metric * top_k * 4);
    if (e == cudaSuccess) {
        cudaMemcpy(d_s, rating_matrix, (size_t)n_users * rating_len * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(d_tp, truth_indptr, ((size_t)n_users + 1) * 8, cudaMemcpyHostToDevice);
        if (nnz) cudaMemcpy(d_ti, truth_items, (size_t)nnz * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(d_m, metric, (size_t)n_metric * 4, cudaMemcpyHostToDevice);
        dense_topk_kernel<<<n_users, 256>>>(d_s, rating_len, n_users, top_k, d_o);
        const int64_t nt = (int64_t)n_users * n_metric;
        neurec_metrics_kernel<<<(unsigned)((nt + 127) / 128), 128>>>(d_o, n_users, top_k, d_tp, d_ti, d_m, n_metric, d_r);
        e = cudaMemcpy(results, d_r, (size_t)n_users * n_metric * top_k * 4, cudaMemcpyDeviceToHost);
    }
    cudaFree(d_s); cudaFree(d_o); cudaFree(d_tp); cudaFree(d_ti); cudaFree(d_m); cudaFree(d_r);
    return e == cudaSuccess ? PDA_OK : neurec_fail("cuda");
}
