// Internal launch interface between the C-ABI layer (pda_capi.cu) and the kernel files.
#pragma once
#include "pda_common.cuh"

namespace pda {

struct SamplerArgs {
    uint32_t seed, epoch, step;
    int64_t B;
    const int32_t* active_users; int64_t n_act;
    const int64_t* indptr; const int32_t* items; const uint8_t* times;
    int32_t n_items;
    const int32_t* unique_times; int32_t n_times;
    const float* pop_train; int32_t T_pop;
    int32_t *users_out, *pos_out, *neg_out, *time_out;
    float *pos_pop_out, *neg_pop_out;
    uint32_t keys[6]; int half_bits;   // filled by launch_sampler
};

struct StepArgs {
    const float* U; const float* I;
    float* GU; float* GI;
    const int32_t *users, *pos, *neg;
    const float *pos_pop, *neg_pop;
    int64_t B; int d;
    float lb;     // regs / batch_size
    float invB;   // 1 / B
    double* loss_acc;
    int pop_mode;     // 0: create_bpr_loss, 1: create_bpr_loss_with_pop_global, 2: BPRMFTempPop (model_api.py:336-371)
    const int32_t* temp; int temp_num;          // mode 2: stage of each triple, number of train stages T
    const float* ub; const float* ib;           // mode 2: user_temp_bias [n_users], item_temp_init_bias [n_items, T+1]
    float* Gub; float* Gib;
    int uniq_users;   // users distinct within the batch -> plain stores for the user rows
    // fused lazy Adam on the user row (UMODE 2): see bpr_step_kernel
    int fuse_user_adam; float* Uw; float* MU; float* VU; int32_t* appliedU; int32_t* stampU;
    const float* lr_hist; const float* pw; float lr; int64_t step_no; unsigned long long* stats;
    // deterministic mode (pda_segsum.cu): per-triple gradient rows are STORED, slot i = pos row of triple i, B + i = neg
    // row, 2B + i = user row (only when users may repeat); the ordered segment sum fills GU / GI afterwards
    float* Gslots;
    // popular items (pda_set_hot_items; by default the n_hot most frequent items of the train CSR): hot_slot[item] = slot
    // 0..n_hot-1 or 255.  The pipelined kernel sums their positive-item gradient rows per CTA in shared memory and adds
    // each row to GI once per CTA -- the same-row red.global.add traffic of a Zipf head serialises in L2
    const uint8_t* hot_slot; const int32_t* hot_ids; int n_hot;
};
constexpr int PDA_MAX_HOT_ITEMS = 28;   // 14 KB: what a 3-stage ring of 8 warps leaves of a third of an SM's shared memory
// default list length.  Measured on the bench workload (value / e2e, M triples/s): 0 rows 482 / 494, 8 rows 494 / 506,
// 16 rows 491 / 474, 28 rows 490 / 474 -- the first rows remove most of the same-address serialisation, and beyond 8 the
// step kernel's CTAs fill the SM's shared memory, so the batch-check / copy-side kernels of the host path stop overlapping it
constexpr int PDA_DEFAULT_HOT_ITEMS = 8;

struct AdamArgs {
    float* W[4]; float* m[4]; float* v[4]; float* G[4];   // user table, item table, [user bias, item bias]
    int64_t n4[4];       // float4 count of each array (0 = unused)
    const float* pw;     // {beta1_power, beta2_power}
    float lr;
    int keep_g;          // 1: leave G as it is (an exchange buffer the caller owns) instead of zeroing it
};

// exact lazy replay of the dense Adam sweep (pda_adam_lazy.cu)
struct LazyArgs {
    float* W[2]; float* m[2]; float* v[2]; float* G[2];
    int32_t* applied[2]; int32_t* stamp[2];
    int lazy[2];                                // per table: 1 = lazy, 0 = left to the dense sweep
    const int32_t *users, *pos, *neg; int64_t B; int d;
    int64_t step_no;                            // steps applied so far = index of the current step
    const float* lr_hist;                       // lr_hist[s] = lr_t of step s (pointer already offset by the ring base)
    const float* pw; float lr;
    unsigned long long* stats;                  // [0] rows updated with a gradient, [1] zero-gradient row-steps replayed
};

struct EvalArgs {
    const float* U; const float* I; int64_t N; int d;
    const int32_t* users; int64_t M;
    int mode;                 // 0: y = s (+ col_bias), 1: y = (elu(s)+1) * pop
    const float* pop; const float* col_bias;
    const int64_t* mask_indptr; const int32_t* mask_items;   // CSR over global user ids, rows sorted
    int K;
    int32_t* ids_out; float* scores_out;    // [M,K]
    float* dense_out;                       // optional [M,N] transformed scores (unmasked)
    const int32_t* M_dev;                   // optional device-side row count (<= M): CTAs beyond it exit at once
    const int32_t* out_rows;                // optional [M] output row of each input row (scatter of fallback rows)
    // fallback of FEW rows: the item range is split over blockIdx.y, partial top-K lists go to part_val / part_ids
    // [row][split][K] and recommend_merge_kernel picks the final K.  Device-sized launches choose between the two forms:
    // the split form runs only when *M_dev <= split_max_rows, the plain one only when *M_dev > skip_rows_le.
    int n_split; int split_max_rows; int skip_rows_le;
    float* part_val; int32_t* part_ids;
};

// plan + scratch layout of the tensor-core filter (pda_eval_tc.cu)
struct TcPlan {
    int64_t M_pad, N_pad;
    int n_tiles, mr, ts, ordered, n_sel, se, cw, n_c, n_valid, splits, tiles_per_split, n_seg, seg_cap, rc;
    size_t o_Ib, o_Ub, o_Ix, o_Ux, o_inorm, o_unorm, o_tnorm, o_tcolmax, o_targ, o_tcol2, o_torder, o_tpos, o_hotv, o_hoti, o_cmax, o_tau, o_cnt, o_cand, o_flag, o_frows, o_fusers, o_nflag, o_clist, o_ckeys, o_ccount, o_work, o_nwork, o_partv, o_parti;
};

void launch_xavier_init(float* W, int64_t rows, int cols, uint32_t seed, uint32_t table_id, cudaStream_t st);
void sampler_keys(uint32_t seed, uint32_t epoch, uint32_t step, uint32_t* keys);
void launch_sampler(SamplerArgs a, cudaStream_t st);
int launch_bpr_step(const StepArgs& a, cudaStream_t st);
void launch_item_count(const int32_t* items, int64_t nnz, int32_t* cnt, cudaStream_t st);
int launch_bpr_step_pipe(const StepArgs& a, cudaStream_t st);   // pda_step_pipe.cu; non-zero = shape not covered
void launch_adam_dense(const AdamArgs& a, cudaStream_t st);
void launch_finish_step(double* loss_acc, float* loss3, double* loss_sum, float* pw, int64_t B, float regs,
                        int batch_size, int advance_powers, float lr, float* lr_slot, cudaStream_t st);
// cross-rank barrier state of the fused exchange kernels (flags in symmetric memory; local == nullptr: none)
struct DpSync { uint32_t* local; uint32_t* peer[8]; uint32_t epoch; int world, rank; };
void launch_dp_exchange_adam(const float* mcG, float* mcW, const float* Gl, float* W, float* M, float* V, int64_t n4,
                             const float* pw, float lr, const DpSync* sy, cudaStream_t st);   // pda_exchange.cu
size_t segsum_temp_bytes(int64_t n);
int launch_segment_sum(const int32_t* a, const int32_t* b, int64_t B, const float* rows, int d, float* G, int32_t* work, void* temp,
                       size_t temp_bytes, int key_bits, cudaStream_t st);   // pda_segsum.cu
void launch_dp_exchange_adam_p2p(const float* const* G, float* const* W, int world, int self, int64_t off, float* M, float* V,
                                 int64_t n4, const float* pw, float lr, const DpSync* sy, cudaStream_t st);
int launch_adam_lazy_rows(const LazyArgs& a, int phase, cudaStream_t st);
void launch_adam_lazy_flush(const LazyArgs& a, int tbl, int64_t n_rows, cudaStream_t st);
void launch_batch_check(const int32_t* users, const int32_t* pos, const int32_t* neg, int64_t B, int32_t n_users,
                        int32_t n_items, int32_t* seen, int32_t tag, int32_t* flags, cudaStream_t st);
void launch_fill_i32(int32_t* p, int64_t n, int32_t value, cudaStream_t st);
void launch_f32_to_i32(const float* src, int32_t* dst, int64_t n, int32_t max_value, cudaStream_t st);
void launch_temp_item_bias(const float* ub, const float* ib, int64_t n_items, int temp_num, int32_t first_user,
                           float* out, cudaStream_t st);
int launch_recommend_exact(const EvalArgs& a, cudaStream_t st);
constexpr int FALLBACK_SPLIT_ROWS = 1024, FALLBACK_SPLITS = 64;   // few-row fallback of the tensor path (pda_eval_exact.cu)
bool tc_supported(const EvalArgs& a);
size_t tc_scratch_bytes(const EvalArgs& a, TcPlan* plan);
// ev: optional 4 events recorded around the sampled sweep (ev[0], ev[1]) and the full sweep (ev[2], ev[3])
int launch_recommend_tc(const EvalArgs& a, void* scratch, const TcPlan& plan, bool prep_items, cudaStream_t st,
                        cudaEvent_t* ev = nullptr);
int launch_tc_debug_dense(const EvalArgs& a, void* scratch, const TcPlan& plan, float* dense, cudaStream_t st);
void launch_metrics(const int32_t* ids, int64_t M, int Kkeep, const int32_t* eval_users, const int64_t* truth_indptr,
                    const int32_t* truth_items, const int32_t* Ks, int nK, double* out, cudaStream_t st);

}  // namespace pda
