/* pda_b200 -- C ABI of the B200-native PDA hot path (libpda_b200.so, sm_100a only).
 *
 * What this replaces in the reference (zyang1580/PDA; paths relative to its root):
 * the TF1 session objects behind MF/train_new_api.py -- the model graph of MF/model_api.py
 * (BPRMF :419-471,695-706; ConditionalBPRMF :19-134), the sampler generators
 * (MF/train_new_api.py:260-456), the inference graph (Create_Recommendation :594-612,
 * do_recommendation :614-640, testing :642-669, predict :683-696), the metric code
 * (MF/used_metric.py:39-80) and, for NeuRec-style callers, the native evaluator entry points
 * cpp_evaluate_matrix (evaluator/backend/cpp/include/evaluate.h:53) and arg_top_k_2d
 * (util/cython/include/arg_topk.h:29).  INTEGRATION.md shows the ctypes binding a maintainer
 * of the reference would add.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on error; pda_last_error() gives the text
 *     of the calling thread's last error.  No exceptions cross the ABI, no ownership transfer:
 *     buffers passed in are only read/written during the call.
 *   - a pda_model owns all device state of one GPU (tables, Adam slots, train CSR, popularity
 *     tables, batch buffers).  One model per GPU/process; calls on one model are not
 *     thread-safe (the reference's sess.run loop is single-threaded as well).
 *   - *_host entry points take HOST pointers, copy in/out inside the call and return after the
 *     result is on the host (the sess.run contract).  *_device entry points take DEVICE
 *     pointers, enqueue on `stream` (a cudaStream_t passed as void*; NULL = default stream)
 *     and return without synchronising.
 *   - there is no CPU fallback: without a CUDA device pda_create fails with PDA_ERR_CUDA.
 */
#ifndef PDA_B200_H
#define PDA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PDA_OK 0
#define PDA_ERR_ARG 1
#define PDA_ERR_CUDA 2
#define PDA_ERR_STATE 3

/* --train of MF/parse.py:11 */
#define PDA_TRAIN_NORMAL 0      /* BPRMF           : create_bpr_loss                (model_api.py:695-706) */
#define PDA_TRAIN_S_CONDITION 1 /* PD / PDA / PDG  : create_bpr_loss_with_pop_global (model_api.py:102-121) */
#define PDA_TRAIN_TEMP_POP 2    /* BPR(t)-pop      : BPRMFTempPop (model_api.py:300-401); TF-GPU gather_nd semantics */

/* rec_type of do_recommendation (train_new_api.py:626-633) */
#define PDA_REC_MAIN_BRANCH 0   /* y = u.i                      (PD, BPRMF)   */
#define PDA_REC_WITH_POP 1      /* y = (elu(u.i)+1) * pop[i]    (PDA 'condition', BPRMF-A 'main_with_pop') */

/* table selectors for pda_get_table / pda_set_table / pda_table_ptr */
#define PDA_TABLE_USER 0
#define PDA_TABLE_ITEM 1
#define PDA_TABLE_USER_M 2
#define PDA_TABLE_USER_V 3
#define PDA_TABLE_ITEM_M 4
#define PDA_TABLE_ITEM_V 5
/* BPR(t)-pop only: user_temp_bias [n_users, 1], item_temp_init_bias [n_items, temp_num + 1] and their Adam slots */
#define PDA_TABLE_USER_BIAS 6
#define PDA_TABLE_ITEM_BIAS 7
#define PDA_TABLE_USER_BIAS_M 8
#define PDA_TABLE_USER_BIAS_V 9
#define PDA_TABLE_ITEM_BIAS_M 10
#define PDA_TABLE_ITEM_BIAS_V 11

/* eval back end */
#define PDA_EVAL_AUTO 0
#define PDA_EVAL_EXACT 1        /* CUDA-core exact scorer over all items */
#define PDA_EVAL_TENSOR 2       /* tcgen05 bf16 filter + exact rescoring of the certified candidates */

typedef struct pda_model pda_model;

typedef struct pda_config {
    int32_t device;      /* CUDA ordinal */
    int64_t n_users;     /* data_config['n_users'] */
    int64_t n_items;     /* data_config['n_items'] */
    int32_t embed_size;  /* --embed_size (multiple of 4, <= 512) */
    int32_t train_mode;  /* PDA_TRAIN_* */
    int32_t batch_size;  /* --batch_size: divisor of the L2 term (model_api.py:118,128) and batch capacity */
    float lr;            /* --lr */
    float regs;          /* --regs */
    int64_t max_batch;   /* capacity of the internal batch buffers (0 -> batch_size) */
    int32_t temp_num;    /* PDA_TRAIN_TEMP_POP: number of train stages T = data_config['temp_num'] (else ignored) */
} pda_config;

const char* pda_last_error(void);
int pda_version(void);
int pda_device_count(void);

/* model life cycle -- replaces ConditionalBPRMF/BPRMF.__init__ + tf.Session + initializer */
int pda_create(const pda_config* cfg, pda_model** out);
void pda_destroy(pda_model* m);
/* Xavier-uniform tables from the Philox stream (model_api.py:86-99); zero Adam slots; beta powers reset */
int pda_init_tables(pda_model* m, uint32_t seed);
int pda_set_table(pda_model* m, int which, const float* host_src);
int pda_get_table(pda_model* m, int which, float* host_dst);
void* pda_table_ptr(pda_model* m, int which);                 /* device pointer, row-major [rows, d] fp32 */
int pda_get_adam_powers(pda_model* m, float* b1p_b2p_host);   /* beta1_power, beta2_power */
int pda_set_adam_powers(pda_model* m, const float* b1p_b2p_host);
int pda_synchronize(pda_model* m);

/* How the TF1 Adam sweep is evaluated.  TF1's Adam on IndexedSlices updates EVERY row at EVERY step (model_api.py:83 ->
 * AdamOptimizer._apply_sparse_shared).  PDA_ADAM_DENSE does exactly that sweep (default while tables, slots
 * and accumulators fit the L2: <= 256 MB).  PDA_ADAM_LAZY (default for larger BPRMF / PD / PDG models) produces bit-identical tables without the sweep: a row replays the zero-gradient steps it skipped, in
 * registers, with the same fp32 operations and the same per-step lr_t, when it is next sampled or when the tables
 * are read (eval, pda_get_table, ...).  PDA_ADAM_LAZY_USERS: lazy user table, dense item table (data-parallel runs whose
 * item gradient is all-reduced).  BPR(t)-pop always runs dense. */
#define PDA_ADAM_DENSE 0
#define PDA_ADAM_LAZY 1
#define PDA_ADAM_LAZY_USERS 2
int pda_set_adam_mode(pda_model* m, int mode);
/* Duplicate rows inside a batch (items): TF1 sums them before the optimizer (_deduplicate_indexed_slices behind
 * MF/model_api.py:83).  Default: fp32 red.global.add in L2 (order not fixed: 1e-5 per step against the oracle).  on != 0:
 * the per-triple gradient rows are stored and summed in occurrence order (pos slots, then neg slots) -- the whole
 * trajectory is bit-identical to the CPU oracle, at about twice the step cost.  $PDA_DETERMINISTIC=1 sets it at create. */
int pda_set_deterministic(pda_model* m, int on);
/* Popular items of the step kernel: the gradient rows of these items (as POSITIVE item of a triple) are summed per thread
 * block in shared memory and reach the gradient accumulator once per block -- a Zipf head puts a quarter of a batch on a
 * few dozen rows, and same-row reductions serialise in L2.  A performance hint only (the sums are the same for any list;
 * d = 128 pipeline only).  pda_set_train_csr* installs the most frequent items of the train CSR (8 by default, at most 28:
 * $PDA_STEP_HOT, 0 = none); this call replaces the list (n = 0 clears it).  ids: host array, distinct. */
int pda_set_hot_items(pda_model* m, const int32_t* ids, int32_t n);
/* out[0] = rows updated with a gradient by the lazy apply kernel, out[1] = zero-gradient row-steps replayed, both since
 * the last call with reset != 0 (synchronises the device) */
int pda_adam_stats(pda_model* m, int64_t* out2, int reset);

/* per-kernel device timing with CUDA events recorded on the launching stream around each kernel:
 * kinds PDA_PROF_*; pda_profile_read synchronises, returns the summed milliseconds and launch count
 * per kind since the last read (arrays of PDA_PROF_KINDS) and resets the counters. */
#define PDA_PROF_SAMPLER 0
#define PDA_PROF_STEP 1
#define PDA_PROF_ADAM 2
#define PDA_PROF_EVAL 3
#define PDA_PROF_EVAL_TC 4
#define PDA_PROF_ADAM_CATCHUP 5   /* lazy Adam: replay of skipped steps for the rows of the batch (+ flushes) */
#define PDA_PROF_EVAL_SWEEP_A 6   /* tensor eval: the sampled tcgen05 sweep (lower bounds -> tau) */
#define PDA_PROF_EVAL_SWEEP_B 7   /* tensor eval: the full tcgen05 sweep (candidates) -- the dominant eval kernel */
#define PDA_PROF_KINDS 8
int pda_profile_enable(pda_model* m, int on);
int pda_profile_read(pda_model* m, double* ms_sum, int32_t* count);

/* pinned host memory for callers that want true async H2D/D2H (cudaHostAlloc / cudaFreeHost) */
void* pda_host_alloc(int64_t bytes);
void pda_host_free(void* p);

/* training data -- replaces Data/Data2.train_user_list (+ item times) and add_expo_popularity.
 * CSR over user ids: indptr[n_users+1], items sorted ascending within each row, times (stage of
 * each interaction, may be NULL), unique_times = data.unique_times (may be NULL).
 * Also the train-item mask of the recommender (train_new_api.py:730-733). */
int pda_set_train_csr(pda_model* m, const int64_t* indptr, const int32_t* items, const uint8_t* times, int64_t nnz,
                      const int32_t* unique_times, int32_t n_times);
/* same, from DEVICE arrays already sorted by the caller (no validation; used for large synthetic sets);
 * active_d = ascending ids of users with at least one interaction; unique_times is a HOST pointer */
int pda_set_train_csr_device(pda_model* m, const int64_t* indptr_d, const int32_t* items_d, const uint8_t* times_d,
                             int64_t nnz, const int32_t* active_d, int64_t n_act, const int32_t* unique_times,
                             int32_t n_times);
/* P = pop[:, :-1] ** gamma as fp32 [n_items, T_pop] (train_new_api.py:988-990); T_pop == 1 -> PDG */
int pda_set_train_pop(pda_model* m, const float* pop, int32_t T_pop);

/* sampler -- replaces generator_n_batch / generator_n_batch_with_pop (train_new_api.py:260-412).
 * Fills the model's internal batch buffers on the device. */
int pda_sample_batch(pda_model* m, uint32_t seed, uint32_t epoch, uint32_t step, int64_t B, void* stream);
/* copy the internal batch to the host (any pointer may be NULL) */
int pda_get_batch(pda_model* m, int64_t B, int32_t* users, int32_t* pos, int32_t* neg, int32_t* time, float* pos_pop,
                  float* neg_pop);

/* one optimisation step -- replaces sess.run([opt, loss, mf_loss, reg_loss]) (train_new_api.py:1080-1090).
 * loss3_out = {loss, mf_loss, reg_loss}.  pos_pop/neg_pop are ignored for PDA_TRAIN_NORMAL.  For
 * PDA_TRAIN_TEMP_POP the two fp32 slots carry what the reference's iterator carries in them
 * (train_new_api.py:544-545,563-565): pos_pop = the stage `temp` of each triple as fp32 (cast to int32 on the
 * device), neg_pop = `raw` (= arange(B); may be NULL, never read). */
int pda_train_step_host(pda_model* m, const int32_t* users, const int32_t* pos, const int32_t* neg,
                        const float* pos_pop, const float* neg_pop, int64_t B, float* loss3_out);
/* n_batches consecutive train steps from PINNED host arrays [n_batches, B] (the generator-fed epoch loop of
 * train_new_api.py:1078-1098, several sess.run calls at once): batch k+1's host->device copies run on a copy stream
 * while step k computes.  loss3_out: fp32 [n_batches, 3] = {loss, mf_loss, reg_loss} of every step.  Same results as
 * n_batches calls of pda_train_step_host. */
int pda_train_steps_host(pda_model* m, const int32_t* users, const int32_t* pos, const int32_t* neg, const float* pos_pop,
                         const float* neg_pop, int32_t n_batches, int64_t B, float* loss3_out);
/* device pointers; users == NULL -> use the internal batch written by pda_sample_batch */
int pda_train_step_device(pda_model* m, const int32_t* users, const int32_t* pos, const int32_t* neg,
                          const float* pos_pop, const float* neg_pop, int64_t B, void* stream);
/* n_steps x (sample + step) enqueued back to back; the loss of every step is kept on the device */
int pda_train_steps_sampled(pda_model* m, uint32_t seed, uint32_t epoch, uint32_t step0, int32_t n_steps, int64_t B,
                            void* stream);
/* The step in two halves, for data-parallel callers that reduce gradients between them:
 * pda_forward_backward_device = the fused kernel only (gradients stay in the accumulators returned by
 * pda_grad_ptr, loss partial sums in pda_loss_acc_ptr: double[2]); pda_adam_apply = Adam sweep + loss /
 * beta-power bookkeeping.  pda_set_global_batch(Bg > 0) makes the loss mean and the gradient scale use Bg
 * (the batch summed over all ranks) instead of the local B.  pda_stage_batch_host copies a host batch into
 * the internal batch buffers (then pass users == NULL). */
int pda_set_global_batch(pda_model* m, int64_t global_batch);
void* pda_grad_ptr(pda_model* m, int which);
void* pda_loss_acc_ptr(pda_model* m);
int pda_forward_backward_device(pda_model* m, const int32_t* users, const int32_t* pos, const int32_t* neg,
                                const float* pos_pop, const float* neg_pop, int64_t B, void* stream);
int pda_adam_apply(pda_model* m, void* stream);
/* the same in two halves: part 1 = tables kept lazily (rank-local gradients; may run while the exchange of the other
 * gradient is in flight), part 2 = dense sweep of the remaining variables + loss / beta-power bookkeeping; 3 = both */
int pda_adam_apply_part(pda_model* m, int part, void* stream);
/* finer still, for callers that pipeline the exchange of a dense gradient with its consumption: the dense Adam sweep of
 * rows [row_lo, row_hi) of one (densely kept) table; after the last range call pda_adam_apply_part(m, 8, stream) --
 * loss / beta-power bookkeeping only */
int pda_adam_dense_rows(pda_model* m, int which, int64_t row_lo, int64_t row_hi, void* stream);
/* the same sweep with the gradient of those rows taken from a caller-owned DEVICE buffer grad[row_hi - row_lo, d] (the
 * output of an out-of-place reduce-scatter); the buffer is left as it is, the model's own accumulator is not touched */
int pda_adam_dense_rows_ext(pda_model* m, int which, int64_t row_lo, int64_t row_hi, const float* grad, void* stream);
/* Data-parallel exchange over NVLink multicast.  pda_adopt_item_buffers moves the item table and its gradient accumulator
 * into caller-owned device memory ([n_items, d] fp32 each; symmetric memory bound to a multicast object), keeping the
 * contents; the buffers must outlive the model.  pda_dp_exchange_adam is reduce-scatter + TF1 Adam sweep of rows
 * [row_lo, row_hi) + all-gather in ONE kernel: mcG / mcW are the MULTICAST addresses of the adopted accumulator / table
 * (multimem.ld_reduce sums the accumulators of all ranks in the switch, multimem.st writes the updated rows to every
 * replica).  The caller brackets it with two cross-rank barriers on `stream` and zeroes the accumulator afterwards. */
int pda_adopt_item_buffers(pda_model* m, float* W_ext, float* G_ext);
/* cross-rank barriers inside the exchange kernels: flags_local / peer_flags[r] = 64 zero-initialised uint32 per rank in
 * symmetric memory (peer_flags[r] = rank r's flags as mapped in this process, self included).  Once set, the kernel itself
 * waits for every rank at its start and completes only when every rank's writes have landed: no barrier launches around
 * it.  flags_local == NULL switches the in-kernel barriers off again. */
int pda_dp_set_barrier(pda_model* m, uint32_t* flags_local, uint32_t* const* peer_flags, int32_t world, int32_t rank);
/* switch the accumulator to another caller-owned ZERO-filled buffer (double buffering; after pda_adopt_item_buffers) */
int pda_set_item_grad_buffer(pda_model* m, float* G_ext);
int pda_dp_exchange_adam(pda_model* m, const float* mcG, float* mcW, int64_t row_lo, int64_t row_hi, void* stream);
/* the same kernel over UNICAST peer pointers (2 ranks: same wire volume, nothing for the switch to reduce): peer_G[r] /
 * peer_W[r] = rank r's accumulator / table as mapped in this process, r = 0..world-1 (<= 8), self included; the partial
 * gradients are added in rank order */
int pda_dp_exchange_adam_p2p(pda_model* m, const float* const* peer_G, float* const* peer_W, int32_t world, int32_t self,
                             int64_t row_lo, int64_t row_hi, void* stream);
int pda_stage_batch_host(pda_model* m, const int32_t* users, const int32_t* pos, const int32_t* neg,
                         const float* pos_pop, const float* neg_pop, int64_t B, void* stream);
/* pda_stage_batch_host without blocking: the copies and the device-side id check (ids inside their tables, users
 * distinct) are enqueued on `copy_stream`; pda_staged_batch_wait blocks the HOST until they are done, returns
 * PDA_ERR_ARG for an id outside its table and makes `stream` wait for the copies.  Lets a data-parallel caller move
 * batch k+1 while step k's gradient exchange is in flight. */
int pda_stage_batch_host_async(pda_model* m, const int32_t* users, const int32_t* pos, const int32_t* neg,
                               const float* pos_pop, const float* neg_pop, int64_t B, void* copy_stream);
int pda_staged_batch_wait(pda_model* m, void* stream);
/* {loss, mf_loss, reg_loss} of the last enqueued step -> PINNED host memory, enqueued on `stream`, no synchronisation */
int pda_read_loss_async(pda_model* m, float* pinned_dst3, void* stream);
/* loss3 of the last enqueued step (synchronises `stream`) */
int pda_read_loss(pda_model* m, float* loss3_out, void* stream);
/* running sums of the per-step fp32 {loss, mf_loss, reg_loss} in double + the step count since the last
 * reset -- the epoch means of train_new_api.py:1095-1097 without a host sync per step (synchronises `stream`) */
int pda_read_loss_sums(pda_model* m, double* out4, int reset, void* stream);
/* forward/backward only (no optimizer): gradients of the two tables as dense [rows, d] host arrays; test hook */
int pda_gradients_host(pda_model* m, const int32_t* users, const int32_t* pos, const int32_t* neg,
                       const float* pos_pop, const float* neg_pop, int64_t B, float* gU_out, float* gI_out,
                       float* loss3_out);

/* BPR(t)-pop test hook: like pda_gradients_host plus the bias gradients gub [n_users], gib [n_items, temp_num+1] */
int pda_gradients_temp_host(pda_model* m, const int32_t* users, const int32_t* pos, const int32_t* neg, const float* temp,
                            int64_t B, float* gU_out, float* gI_out, float* gub_out, float* gib_out, float* loss3_out);
/* BPR(t)-pop inference bias of one eval batch (model_api.py:373-387): out[j] = (1 + user_temp_bias[first_user]) *
 * (item_bias[j, T-1] + item_bias[j, T]), fp32 [n_items] on the host; pass it as col_bias of pda_recommend_* */
int pda_temp_item_bias_host(pda_model* m, int32_t first_user, float* out);

/* recommendation -- replaces do_recommendation (train_new_api.py:614-640).  users: M global user ids;
 * all n_items are scored; pop: fp32 [n_items] (PDA_REC_WITH_POP) or NULL; col_bias: optional fp32
 * [n_items] added in PDA_REC_MAIN_BRANCH (BPR(t)-pop item bias, model_api.py:387); use_mask != 0
 * removes each user's train items (the CSR of pda_set_train_csr).  ids_out int32 [M,K] sorted by
 * (score desc, id asc) -- tf.nn.top_k's order; scores_out fp32 [M,K] or NULL.  K <= 128. */
int pda_recommend_host(pda_model* m, const int32_t* users, int64_t M, int rec_type, const float* pop,
                       const float* col_bias, int use_mask, int K, int backend, int32_t* ids_out, float* scores_out);
int pda_recommend_device(pda_model* m, const int32_t* users, int64_t M, int rec_type, const float* pop,
                         const float* col_bias, int use_mask, int K, int backend, int32_t* ids_out, float* scores_out,
                         void* stream);
/* diagnostics of the last tensor-core eval block: out[8] = {rows, rows recomputed by the exact kernel (no certificate),
 * candidates rescored, max candidates of a row, rows whose candidate list overflowed, pass-A tile stride, sampled
 * chunks per row, item-range splits}.  Synchronises the device. */
int pda_tc_last_stats(pda_model* m, int64_t* out);
/* diagnostics, host arithmetic only (no device needed): the launch plan / scratch layout the tensor-core eval would use for
 * M users x N items: out[24] = {M_pad, N_pad, n_tiles, user tiles per CTA, ts, ordered, n_sel, stride, chunk width, n_c,
 * n_valid, splits, tiles_per_split, n_seg, seg_cap, rc, total bytes, o_Ib, o_Ub, o_cmax, o_cand, o_clist, o_work, o_nwork} */
int pda_tc_plan_host(int64_t M, int64_t N, int32_t d, int32_t K, int64_t* out);
/* diagnostics: the raw tensor-core accumulators the filter compares -- out fp32 [M, n_items]:
 * v[r][j] ~ (u_r . i_j + 1) * pop_j  (PDA_REC_WITH_POP),  u_r . i_j + col_bias_j,  or  u_r . i_j  (bf16 operands, fp32
 * accumulation in TMEM); err_coef[2] = {cAB, cB} of the bound |v - exact| <= cAB |u| |c_j i_j| + cB |x_j|.
 * Lets a test check operand layouts and the bound against fp64.  M <= 32768. */
int pda_tc_debug_dense_host(pda_model* m, const int32_t* users, int64_t M, int rec_type, const float* pop,
                            const float* col_bias, float* out, float* err_coef);
/* dense scores -- replaces testing()/predict() (train_new_api.py:642-696): out fp32 [M, n_items], no mask */
int pda_scores_host(pda_model* m, const int32_t* users, int64_t M, int rec_type, const float* pop, float* out);

/* metrics -- replaces test_one_batch/get_performance (train_new_api.py:741-758, used_metric.py:69-80).
 * ids int32 [M,Kkeep] (host), truth CSR over global user ids (host).  out double [4, nK]:
 * precision, recall, ndcg, hit_ratio SUMMED over the M users (the caller divides). */
int pda_metrics_host(pda_model* m, const int32_t* ids, int64_t M, int Kkeep, const int32_t* eval_users,
                     const int64_t* truth_indptr, const int32_t* truth_items, int64_t n_truth_rows, const int32_t* Ks,
                     int nK, double* out);

/* NeuRec-style native evaluator entry points over the GPU (the reference's optional native FFI; both stateless, device 0):
 *   pda_arg_top_k_2d_host    = arg_top_k_2d (util/cython/include/arg_topk.h:29): results int32 [rows, top_k], the indices of
 *                              the top_k largest scores of every row of the C-contiguous float32 matrix scores[rows, cols],
 *                              best first; equal scores: lower index first.  top_k <= 128.
 *   pda_evaluate_matrix_host = cpp_evaluate_matrix (evaluator/backend/cpp/include/evaluate.h:53): per user the top_k of its
 *                              rating row, then for every requested metric (1 Precision, 2 Recall, 3 MAP, 4 NDCG, 5 MRR;
 *                              metric.h:109-114) its value at cut-offs 1..top_k -> results float32 [n_users, n_metric * top_k].
 *                              The reference's vector<unordered_set<int>> of test items is passed as a CSR over users. */
int pda_arg_top_k_2d_host(const float* scores, int32_t cols, int32_t rows, int32_t top_k, int32_t* results);
int pda_evaluate_matrix_host(const float* rating_matrix, int32_t rating_len, int32_t n_users, const int64_t* truth_indptr,
                             const int32_t* truth_items, const int32_t* metric, int32_t n_metric, int32_t top_k, float* results);

/* diagnostics behind the bit-exactness claims of the exact Adam replay: runs a check kernel on the current device and
 * returns out5 = {operands checked, mismatching results, first mismatch: 3 raw words}.
 *   kind 3: every fp32 bit pattern in [lo_or_seed, hi] through the straight-line sqrt refinement vs __fsqrt_rn
 *   kind 0: per_thread x 2 x 303104 random in-range (a, b) pairs through the straight-line quotient vs __fdiv_rn
 *   kind 1 / 2: per_thread x 4 x 303104 random in-range elements through one packed zero-gradient / gradient Adam step
 *               vs the generic separately rounded form (pda_common.cuh)
 *   kind 4: the same for three consecutive zero-gradient steps in the negated-v form the pipelined step kernel replays with */
int pda_debug_numerics(int kind, uint32_t lo_or_seed, uint32_t hi, uint64_t per_thread, uint64_t* out5);

#ifdef __cplusplus
}
#endif
#endif /* PDA_B200_H */
