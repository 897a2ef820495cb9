// extern "C" shim around the REFERENCE's own native evaluator headers, compiled in place from
// /root/reference (never copied): evaluator/backend/cpp/include/{evaluate.h,metric.h},
// util/cython/include/{arg_topk.h,thread_pool.h}.  Output: oracle/_ref/libref_eval.so.
// TEST INFRASTRUCTURE ONLY: used to pin the oracle's top-k and metric restatements.
#include <vector>
#include <unordered_set>
#include "evaluate.h"   // cpp_evaluate_matrix  (evaluate.h:53)
#include "arg_topk.h"   // arg_top_k_2d         (arg_topk.h:29)

extern "C" {
void ref_evaluate_matrix(float* ratings, int rating_len, int n_users, const long long* truth_indptr,
                         const int* truth_items, const int* metric, int n_metric, int top_k,
                         int thread_num, float* results) {
    std::vector<std::unordered_set<int>> test_items(n_users);
    for (int u = 0; u < n_users; ++u)
        for (long long q = truth_indptr[u]; q < truth_indptr[u + 1]; ++q) test_items[u].insert(truth_items[q]);
    std::vector<int> m(metric, metric + n_metric);
    cpp_evaluate_matrix(ratings, rating_len, test_items, m, top_k, thread_num, results);
}
void ref_arg_top_k_2d(float* ratings, int rating_len, int rows_num, int top_k, int thread_num, int* results) {
    arg_top_k_2d(ratings, rating_len, rows_num, top_k, thread_num, results);
}
}
