// C-ABI layer of libpda_b200.so: model handle, device state, host<->device staging.
// See include/pda_b200.h for the contract and the reference interfaces each entry point replaces.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>

#include "../../include/pda_b200.h"
#include "pda_kernels.h"

#define PDA_PROF_SLOTS 1024
#define PDA_LR_CAP (1 << 22)   /* lr_t history ring of the lazy Adam replay (steps between two forced flushes) */

using namespace pda;

static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CK(call)                                                                                     \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess)                                                                       \
            return fail(PDA_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

struct pda_model {
    pda_config cfg;
    int64_t nU, nI;
    int d;
    float *W[4], *Mo[4], *Vo[4], *G[4];   // 0 = user table, 1 = item table, 2 = user_temp_bias, 3 = item_temp_init_bias
    int64_t rows[4]; int cols[4]; int64_t n4[4];   // logical shape and padded float4 count of each array
    int n_arr;                                      // 2, or 4 for BPR(t)-pop
    // exact lazy replay of the dense Adam sweep (pda_adam_lazy.cu)
    int adam_lazy[2]; int32_t* applied[2]; int32_t* stamp[2];
    float* lr_hist; int64_t step_no, lr_base; unsigned long long* lazy_stats;
    const int32_t *cur_users, *cur_pos, *cur_neg; int64_t cur_B;   // batch of the step in flight
    int cur_fused, fuse_user_adam;
    int32_t* seen; int32_t seen_tag;   // scratch of the distinct-users check of host batches
    int32_t* chk_flags;                // device {users repeat, id out of range} x 2 slots (+ the asynchronous staging slot)
    int32_t* chk_pinned;               // pinned mirror of chk_flags
    cudaEvent_t ev_staged; int staged_pending;   // pda_stage_batch_host_async / pda_staged_batch_wait
    int32_t max_time;                  // largest stage label of the host-validated train CSR (-1: unknown)
    int item_ext;                      // W[1] / G[1] live in caller-owned (symmetric) memory: not freed here
    // deterministic duplicate-row accumulation (pda_segsum.cu): slot buffer [3 B, d], sort work arrays, cub temp storage
    int deterministic;
    DpSync dp_sync;                    // in-kernel barriers of the fused exchange (pda_dp_set_barrier)
    void* gslots; size_t gslots_bytes; void* seg_work; size_t seg_work_bytes; void* seg_temp; size_t seg_temp_bytes;
    float* pw;          // {beta1_power, beta2_power}
    double* loss_acc;   // {sum log(sigmoid+1e-10), sum of squares}
    float* loss3;       // device {loss, mf, reg}
    double* loss_sum;   // device {sum loss, sum mf, sum reg, steps} since the last pda_read_loss_sums(reset)
    float* loss3_pinned;
    // train CSR / mask
    int64_t* indptr; int32_t* items; uint8_t* times; int64_t nnz;
    int32_t* active; int64_t n_act;
    int32_t* unique_times; int32_t n_times;
    float* pop_train; int32_t T_pop;
    uint8_t* hot_slot; int32_t* hot_ids; int n_hot;   // popular items of the step kernel (pda_set_hot_items)
    // batch buffers
    int64_t cap, last_B, global_batch;
    int batch_uniq;   // internal batch holds distinct users (device sampler with B <= #active users)
    int32_t *b_users, *b_pos, *b_neg, *b_time;
    float *b_pp, *b_np;
    // pda_train_steps_host: second batch slot, copy stream, per-slot events, pinned loss ring
    int32_t *b2_users, *b2_pos, *b2_neg; float *b2_pp, *b2_np;
    cudaStream_t copy_st; cudaEvent_t ev_copied[2], ev_stepped[2], ev_start; int pipe_ready;
    float* loss_ring; size_t loss_ring_bytes;
    void* stage_pinned; size_t stage_bytes;
    // eval scratch
    void* ev_buf; size_t ev_bytes;
    void* tc_buf; size_t tc_bytes;   // scratch of the tensor-core filter
    TcPlan tc_last_plan; int64_t tc_last_M;
    float* grad_bias_out[2];   // pda_gradients_temp_host: host destinations of the bias gradients
    void* ev_pinned; size_t ev_pinned_bytes;
    // optional per-kernel CUDA-event timing (pda_profile_*): [kernel kind][slot][begin/end]
    int prof_on; int prof_n[PDA_PROF_KINDS];
    cudaEvent_t (*prof_ev)[PDA_PROF_SLOTS][2];
};

struct ProfScope {   // records an event pair around one kernel launch on the launching stream
    pda_model* m; int kind; cudaStream_t st; int slot;
    ProfScope(pda_model* m_, int kind_, cudaStream_t st_) : m(m_), kind(kind_), st(st_), slot(-1) {
        if (m->prof_on && m->prof_ev && m->prof_n[kind] < PDA_PROF_SLOTS) {
            slot = m->prof_n[kind]++;
            cudaEventRecord(m->prof_ev[kind][slot][0], st);
        }
    }
    ~ProfScope() { if (slot >= 0) cudaEventRecord(m->prof_ev[kind][slot][1], st); }
};

template <typename T>
static cudaError_t dmalloc(T** p, size_t n) { return cudaMalloc((void**)p, n * sizeof(T) > 0 ? n * sizeof(T) : 16); }

static cudaError_t ensure_dev(void** buf, size_t* have, size_t need) {
    if (*have >= need) return cudaSuccess;
    if (*buf) cudaFree(*buf);
    *buf = nullptr; *have = 0;
    cudaError_t e = cudaMalloc(buf, need);
    if (e == cudaSuccess) *have = need;
    return e;
}
static cudaError_t ensure_pinned(void** buf, size_t* have, size_t need) {
    if (*have >= need) return cudaSuccess;
    if (*buf) cudaFreeHost(*buf);
    *buf = nullptr; *have = 0;
    cudaError_t e = cudaHostAlloc(buf, need, cudaHostAllocDefault);
    if (e == cudaSuccess) *have = need;
    return e;
}

extern "C" {

const char* pda_last_error(void) { return g_err; }
int pda_version(void) { return 100; }
int pda_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

void* pda_host_alloc(int64_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, (size_t)bytes, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void pda_host_free(void* p) { if (p) cudaFreeHost(p); }

int pda_create(const pda_config* cfg, pda_model** out) {
    if (!cfg || !out) return fail(PDA_ERR_ARG, "pda_create: null argument");
    if (cfg->embed_size % 4 != 0 || cfg->embed_size < 4 || cfg->embed_size > 512)
        return fail(PDA_ERR_ARG, "embed_size must be a multiple of 4 in [4, 512], got %d", cfg->embed_size);
    if (cfg->n_users < 1 || cfg->n_items < 1 || cfg->batch_size < 1)
        return fail(PDA_ERR_ARG, "n_users, n_items and batch_size must be positive");
    if (cfg->n_users > 0x7fffffffLL || cfg->n_items > 0x7fffffffLL)
        return fail(PDA_ERR_ARG, "ids are int32: n_users/n_items must be < 2^31");
    if (cfg->train_mode != PDA_TRAIN_NORMAL && cfg->train_mode != PDA_TRAIN_S_CONDITION &&
        cfg->train_mode != PDA_TRAIN_TEMP_POP)
        return fail(PDA_ERR_ARG, "unknown train_mode %d", cfg->train_mode);
    if (cfg->train_mode == PDA_TRAIN_TEMP_POP && (cfg->temp_num < 1 || cfg->temp_num > 255))
        return fail(PDA_ERR_ARG, "train_mode temp_pop needs temp_num in [1, 255], got %d", cfg->temp_num);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(PDA_ERR_CUDA, "no CUDA device: pda_b200 has no CPU fallback (%s)", cudaGetErrorString(e));
    }
    CK(cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major != 10)
        return fail(PDA_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", cfg->device, prop.major,
                    prop.minor);
    pda_model* m = new (std::nothrow) pda_model();
    if (!m) return fail(PDA_ERR_STATE, "out of host memory");
    memset(m, 0, sizeof(*m));
    m->cfg = *cfg;
    m->nU = cfg->n_users; m->nI = cfg->n_items; m->d = cfg->embed_size;
    m->cap = cfg->max_batch > 0 ? cfg->max_batch : cfg->batch_size;
    m->n_arr = cfg->train_mode == PDA_TRAIN_TEMP_POP ? 4 : 2;
    m->rows[0] = m->nU; m->cols[0] = m->d; m->rows[1] = m->nI; m->cols[1] = m->d;
    m->rows[2] = m->nU; m->cols[2] = 1; m->rows[3] = m->nI; m->cols[3] = cfg->temp_num + 1;
    for (int t = 0; t < m->n_arr; ++t) {
        m->n4[t] = (m->rows[t] * m->cols[t] + 3) / 4;   // padded to whole float4s (the pad stays zero)
        size_t n = (size_t)m->n4[t] * 4;
        CK(dmalloc(&m->W[t], n)); CK(dmalloc(&m->Mo[t], n)); CK(dmalloc(&m->Vo[t], n)); CK(dmalloc(&m->G[t], n));
        CK(cudaMemset(m->W[t], 0, n * 4)); CK(cudaMemset(m->Mo[t], 0, n * 4));
        CK(cudaMemset(m->Vo[t], 0, n * 4)); CK(cudaMemset(m->G[t], 0, n * 4));
    }
    for (int t = 0; t < 2; ++t) {
        CK(dmalloc(&m->applied[t], (size_t)m->rows[t])); CK(dmalloc(&m->stamp[t], (size_t)m->rows[t]));
        CK(cudaMemset(m->applied[t], 0, (size_t)m->rows[t] * 4)); CK(cudaMemset(m->stamp[t], 0, (size_t)m->rows[t] * 4));
    }
    {   // default per table: the lazy replay pays off when (a) tables + slots + accumulators spill the 126 MB L2 (the
        // dense sweep is then HBM traffic; small tables stay L2-resident: Douban 0.1 s vs 0.3 s per epoch) and (b) a step
        // touches a small part of the table (rows > 4 x row references per step); a table that is mostly touched
        // every step (1M items vs 2 x 2^20 item references) is swept densely at less traffic than catch-up + apply.
        const double bytes = (double)(m->nU + m->nI) * m->d * 16.0;
        const bool big = cfg->train_mode != PDA_TRAIN_TEMP_POP && bytes > 256.0 * 1024 * 1024;
        m->adam_lazy[0] = big && m->nU > 4 * m->cap;
        m->adam_lazy[1] = big && m->nI > 8 * m->cap;
        const char* e = getenv("PDA_FUSE_USER_ADAM");
        m->fuse_user_adam = e ? atoi(e) : 1;
        e = getenv("PDA_DETERMINISTIC");
        m->deterministic = e ? atoi(e) != 0 : 0;
    }
    CK(dmalloc(&m->lr_hist, (size_t)PDA_LR_CAP));
    CK(dmalloc(&m->lazy_stats, 2)); CK(cudaMemset(m->lazy_stats, 0, 16));
    CK(dmalloc(&m->pw, 2)); CK(dmalloc(&m->loss_acc, 2)); CK(dmalloc(&m->loss3, 4)); CK(dmalloc(&m->loss_sum, 4));
    const float pw0[2] = {0.9f, 0.999f};
    CK(cudaMemcpy(m->pw, pw0, 8, cudaMemcpyHostToDevice));
    CK(cudaMemset(m->loss_acc, 0, 16)); CK(cudaMemset(m->loss3, 0, 16)); CK(cudaMemset(m->loss_sum, 0, 32));
    CK(cudaHostAlloc((void**)&m->loss3_pinned, 64, cudaHostAllocDefault));
    CK(dmalloc(&m->chk_flags, 8)); CK(cudaMemset(m->chk_flags, 0, 32));   // 3 slots x {repeat, bad id} (+ pad)
    CK(cudaHostAlloc((void**)&m->chk_pinned, 32, cudaHostAllocDefault));
    CK(cudaEventCreateWithFlags(&m->ev_staged, cudaEventDisableTiming));
    m->max_time = -1;
    CK(dmalloc(&m->b_users, (size_t)m->cap)); CK(dmalloc(&m->b_pos, (size_t)m->cap)); CK(dmalloc(&m->b_neg, (size_t)m->cap));
    CK(dmalloc(&m->b_time, (size_t)m->cap)); CK(dmalloc(&m->b_pp, (size_t)m->cap)); CK(dmalloc(&m->b_np, (size_t)m->cap));
    CK(cudaDeviceSynchronize());
    *out = m;
    return PDA_OK;
}

void pda_destroy(pda_model* m) {
    if (!m) return;
    cudaSetDevice(m->cfg.device);
    cudaDeviceSynchronize();
    if (m->item_ext) { m->W[1] = nullptr; m->G[1] = nullptr; }
    for (int t = 0; t < 4; ++t) { cudaFree(m->W[t]); cudaFree(m->Mo[t]); cudaFree(m->Vo[t]); cudaFree(m->G[t]); }
    for (int t = 0; t < 2; ++t) { cudaFree(m->applied[t]); cudaFree(m->stamp[t]); }
    cudaFree(m->lr_hist); cudaFree(m->lazy_stats); cudaFree(m->seen); cudaFree(m->chk_flags);
    cudaFreeHost(m->chk_pinned); cudaEventDestroy(m->ev_staged);
    cudaFree(m->pw); cudaFree(m->loss_acc); cudaFree(m->loss3); cudaFree(m->loss_sum); cudaFreeHost(m->loss3_pinned);
    cudaFree(m->indptr); cudaFree(m->items); cudaFree(m->times); cudaFree(m->active); cudaFree(m->unique_times);
    cudaFree(m->pop_train); cudaFree(m->hot_slot); cudaFree(m->hot_ids);
    cudaFree(m->b_users); cudaFree(m->b_pos); cudaFree(m->b_neg); cudaFree(m->b_time); cudaFree(m->b_pp); cudaFree(m->b_np);
    if (m->pipe_ready) {
        cudaFree(m->b2_users); cudaFree(m->b2_pos); cudaFree(m->b2_neg); cudaFree(m->b2_pp); cudaFree(m->b2_np);
        cudaStreamDestroy(m->copy_st);
        for (int i = 0; i < 2; ++i) { cudaEventDestroy(m->ev_copied[i]); cudaEventDestroy(m->ev_stepped[i]); }
        cudaEventDestroy(m->ev_start);
    }
    if (m->loss_ring) cudaFreeHost(m->loss_ring);
    if (m->stage_pinned) cudaFreeHost(m->stage_pinned);
    if (m->gslots) cudaFree(m->gslots);
    if (m->seg_work) cudaFree(m->seg_work);
    if (m->seg_temp) cudaFree(m->seg_temp);
    if (m->ev_buf) cudaFree(m->ev_buf);
    if (m->tc_buf) cudaFree(m->tc_buf);
    if (m->ev_pinned) cudaFreeHost(m->ev_pinned);
    if (m->prof_ev) {
        for (int k = 0; k < PDA_PROF_KINDS; ++k)
            for (int i = 0; i < PDA_PROF_SLOTS; ++i) { cudaEventDestroy(m->prof_ev[k][i][0]); cudaEventDestroy(m->prof_ev[k][i][1]); }
        free(m->prof_ev);
    }
    delete m;
}

int pda_profile_enable(pda_model* m, int on) {
    if (!m) return fail(PDA_ERR_ARG, "null model");
    CK(cudaSetDevice(m->cfg.device));
    if (on && !m->prof_ev) {
        m->prof_ev = (cudaEvent_t(*)[PDA_PROF_SLOTS][2])calloc(PDA_PROF_KINDS, sizeof(cudaEvent_t) * PDA_PROF_SLOTS * 2);
        if (!m->prof_ev) return fail(PDA_ERR_STATE, "out of host memory");
        for (int k = 0; k < PDA_PROF_KINDS; ++k)
            for (int i = 0; i < PDA_PROF_SLOTS; ++i) { CK(cudaEventCreate(&m->prof_ev[k][i][0])); CK(cudaEventCreate(&m->prof_ev[k][i][1])); }
    }
    m->prof_on = on;
    for (int k = 0; k < PDA_PROF_KINDS; ++k) m->prof_n[k] = 0;
    return PDA_OK;
}

int pda_profile_read(pda_model* m, double* ms_sum, int32_t* count) {
    if (!m || !ms_sum || !count) return fail(PDA_ERR_ARG, "null argument");
    CK(cudaSetDevice(m->cfg.device));
    CK(cudaDeviceSynchronize());
    for (int k = 0; k < PDA_PROF_KINDS; ++k) {
        ms_sum[k] = 0.0; count[k] = m->prof_ev ? m->prof_n[k] : 0;
        for (int i = 0; i < count[k]; ++i) {
            float ms = 0.f;
            CK(cudaEventElapsedTime(&ms, m->prof_ev[k][i][0], m->prof_ev[k][i][1]));
            ms_sum[k] += ms;
        }
        m->prof_n[k] = 0;
    }
    return PDA_OK;
}

int pda_synchronize(pda_model* m) {
    if (!m) return fail(PDA_ERR_ARG, "null model");
    CK(cudaSetDevice(m->cfg.device));
    CK(cudaDeviceSynchronize());
    return PDA_OK;
}

int pda_init_tables(pda_model* m, uint32_t seed) {
    if (!m) return fail(PDA_ERR_ARG, "null model");
    CK(cudaSetDevice(m->cfg.device));
    for (int t = 0; t < m->n_arr; ++t) launch_xavier_init(m->W[t], m->rows[t], m->cols[t], seed, (uint32_t)t, 0);
    CK(cudaGetLastError());
    for (int t = 0; t < m->n_arr; ++t) {
        size_t n = (size_t)m->n4[t] * 16;
        CK(cudaMemsetAsync(m->Mo[t], 0, n)); CK(cudaMemsetAsync(m->Vo[t], 0, n)); CK(cudaMemsetAsync(m->G[t], 0, n));
    }
    const float pw0[2] = {0.9f, 0.999f};
    CK(cudaMemcpy(m->pw, pw0, 8, cudaMemcpyHostToDevice));
    CK(cudaMemset(m->loss_acc, 0, 16));
    CK(cudaMemset(m->loss_sum, 0, 32));
    for (int t = 0; t < 2; ++t) {
        CK(cudaMemset(m->applied[t], 0, (size_t)m->rows[t] * 4)); CK(cudaMemset(m->stamp[t], 0, (size_t)m->rows[t] * 4));
    }
    m->step_no = 0; m->lr_base = 0;
    CK(cudaDeviceSynchronize());
    return PDA_OK;
}

// ---- lazy Adam plumbing ----
static void lazy_args(pda_model* m, LazyArgs* a) {
    memset(a, 0, sizeof(*a));
    for (int t = 0; t < 2; ++t) {
        a->W[t] = m->W[t]; a->m[t] = m->Mo[t]; a->v[t] = m->Vo[t]; a->G[t] = m->G[t];
        a->applied[t] = m->applied[t]; a->stamp[t] = m->stamp[t]; a->lazy[t] = m->adam_lazy[t];
    }
    a->users = m->cur_users; a->pos = m->cur_pos; a->neg = m->cur_neg; a->B = m->cur_B; a->d = m->d;
    a->step_no = m->step_no; a->lr_hist = m->lr_hist - m->lr_base; a->pw = m->pw; a->lr = m->cfg.lr;
    a->stats = m->lazy_stats;
}

// every lazily maintained row replays up to step_no: after this the tables hold what the dense sweep would hold
static void flush_lazy(pda_model* m, cudaStream_t st) {
    if (!m->adam_lazy[0] && !m->adam_lazy[1]) return;
    LazyArgs a;
    lazy_args(m, &a);
    for (int t = 0; t < 2; ++t)
        if (m->adam_lazy[t]) { ProfScope ps(m, PDA_PROF_ADAM_CATCHUP, st); launch_adam_lazy_flush(a, t, m->rows[t], st); }
}

// *n = number of fp32 elements of the selected array (rows x cols)
static float* table_of(pda_model* m, int which, int64_t* n) {
    int t, kind;   // kind 0 = variable, 1 = Adam m, 2 = Adam v
    switch (which) {
        case PDA_TABLE_USER: t = 0; kind = 0; break;
        case PDA_TABLE_ITEM: t = 1; kind = 0; break;
        case PDA_TABLE_USER_M: t = 0; kind = 1; break;
        case PDA_TABLE_USER_V: t = 0; kind = 2; break;
        case PDA_TABLE_ITEM_M: t = 1; kind = 1; break;
        case PDA_TABLE_ITEM_V: t = 1; kind = 2; break;
        case PDA_TABLE_USER_BIAS: t = 2; kind = 0; break;
        case PDA_TABLE_ITEM_BIAS: t = 3; kind = 0; break;
        case PDA_TABLE_USER_BIAS_M: t = 2; kind = 1; break;
        case PDA_TABLE_USER_BIAS_V: t = 2; kind = 2; break;
        case PDA_TABLE_ITEM_BIAS_M: t = 3; kind = 1; break;
        case PDA_TABLE_ITEM_BIAS_V: t = 3; kind = 2; break;
        default: return nullptr;
    }
    if (t >= m->n_arr) return nullptr;
    *n = m->rows[t] * m->cols[t];
    return kind == 0 ? m->W[t] : kind == 1 ? m->Mo[t] : m->Vo[t];
}

int pda_set_table(pda_model* m, int which, const float* src) {
    if (!m || !src) return fail(PDA_ERR_ARG, "null argument");
    int64_t rows; float* p = table_of(m, which, &rows);
    if (!p) return fail(PDA_ERR_ARG, "unknown table %d", which);
    CK(cudaSetDevice(m->cfg.device));
    flush_lazy(m, 0);
    CK(cudaMemcpy(p, src, (size_t)rows * 4, cudaMemcpyHostToDevice));
    // a row whose slots were just overwritten may hold non-zero m / v: it is no longer "never touched"
    for (int t = 0; t < 2; ++t) launch_fill_i32(m->stamp[t], m->rows[t], 1, 0);
    return PDA_OK;
}
int pda_get_table(pda_model* m, int which, float* dst) {
    if (!m || !dst) return fail(PDA_ERR_ARG, "null argument");
    int64_t rows; float* p = table_of(m, which, &rows);
    if (!p) return fail(PDA_ERR_ARG, "unknown table %d", which);
    CK(cudaSetDevice(m->cfg.device));
    flush_lazy(m, 0);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(dst, p, (size_t)rows * 4, cudaMemcpyDeviceToHost));
    return PDA_OK;
}
void* pda_table_ptr(pda_model* m, int which) {
    if (!m) return nullptr;
    int64_t rows;
    cudaSetDevice(m->cfg.device);
    flush_lazy(m, 0);          // the caller reads the table itself: make it current first
    cudaDeviceSynchronize();
    return table_of(m, which, &rows);
}

int pda_adam_stats(pda_model* m, int64_t* out2, int reset) {
    if (!m || !out2) return fail(PDA_ERR_ARG, "null argument");
    CK(cudaSetDevice(m->cfg.device));
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(out2, m->lazy_stats, 16, cudaMemcpyDeviceToHost));
    if (reset) CK(cudaMemset(m->lazy_stats, 0, 16));
    return PDA_OK;
}

// Deterministic accumulation of duplicate rows (items that repeat inside a batch): on = per-triple gradient rows are stored
// and summed in occurrence order (the oracle's dedup_sum order) instead of reduced with fp32 atomics -- trajectories become
// bit-identical to the CPU oracle at ~2x the step cost.  BPRMF / PD / PDG (the bias tables of BPR(t)-pop keep atomics).
int pda_set_deterministic(pda_model* m, int on) {
    if (!m) return fail(PDA_ERR_ARG, "null model");
    m->deterministic = on ? 1 : 0;
    return PDA_OK;
}

// The popular items whose positive-item gradient rows the pipelined step kernel sums per CTA in shared memory before they
// reach the accumulator (same-row red.global.add serialises in L2).  A performance hint only: any list gives the same sums.
int pda_set_hot_items(pda_model* m, const int32_t* ids, int32_t n) {
    if (!m || n < 0 || (n > 0 && !ids)) return fail(PDA_ERR_ARG, "bad argument");
    if (n > PDA_MAX_HOT_ITEMS) n = PDA_MAX_HOT_ITEMS;
    for (int32_t i = 0; i < n; ++i) {
        if (ids[i] < 0 || ids[i] >= m->nI) return fail(PDA_ERR_ARG, "hot item id %d outside [0, %lld)", ids[i], (long long)m->nI);
        for (int32_t j = 0; j < i; ++j)
            if (ids[j] == ids[i]) return fail(PDA_ERR_ARG, "hot item id %d listed twice", ids[i]);
    }
    CK(cudaSetDevice(m->cfg.device));
    CK(cudaDeviceSynchronize());
    m->n_hot = 0;
    if (n == 0) return PDA_OK;
    if (!m->hot_slot) {
        CK(cudaMalloc((void**)&m->hot_slot, (size_t)m->nI));
        CK(cudaMalloc((void**)&m->hot_ids, sizeof(int32_t) * PDA_MAX_HOT_ITEMS));
    }
    uint8_t* h = (uint8_t*)malloc((size_t)m->nI);
    if (!h) return fail(PDA_ERR_STATE, "out of host memory");
    memset(h, 255, (size_t)m->nI);
    for (int32_t i = 0; i < n; ++i) h[ids[i]] = (uint8_t)i;
    cudaError_t e = cudaMemcpy(m->hot_slot, h, (size_t)m->nI, cudaMemcpyHostToDevice);
    free(h);
    if (e != cudaSuccess) return fail(PDA_ERR_CUDA, "cudaMemcpy failed: %s", cudaGetErrorString(e));
    CK(cudaMemcpy(m->hot_ids, ids, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice));
    m->n_hot = n;
    return PDA_OK;
}

// default hot list: the most frequent items of the train CSR ($PDA_STEP_HOT = how many, default 8, 0 = none; at most PDA_MAX_HOT_ITEMS)
static int hot_items_from_csr(pda_model* m) {
    const char* e = getenv("PDA_STEP_HOT");
    int want = e ? atoi(e) : PDA_DEFAULT_HOT_ITEMS;
    if (want > PDA_MAX_HOT_ITEMS) want = PDA_MAX_HOT_ITEMS;
    if (want > m->nI) want = (int)m->nI;
    if (want <= 0 || m->nnz <= 0 || m->d != 128) return pda_set_hot_items(m, nullptr, 0);   // only the d = 128 pipeline uses it
    int32_t* cnt_d = nullptr;
    CK(cudaMalloc((void**)&cnt_d, sizeof(int32_t) * (size_t)m->nI));
    CK(cudaMemset(cnt_d, 0, sizeof(int32_t) * (size_t)m->nI));
    launch_item_count(m->items, m->nnz, cnt_d, 0);
    int32_t* cnt = (int32_t*)malloc(sizeof(int32_t) * (size_t)m->nI);
    if (!cnt) { cudaFree(cnt_d); return fail(PDA_ERR_STATE, "out of host memory"); }
    cudaError_t ce = cudaMemcpy(cnt, cnt_d, sizeof(int32_t) * (size_t)m->nI, cudaMemcpyDeviceToHost);
    cudaFree(cnt_d);
    if (ce != cudaSuccess) { free(cnt); return fail(PDA_ERR_CUDA, "item count failed: %s", cudaGetErrorString(ce)); }
    int32_t ids[PDA_MAX_HOT_ITEMS];
    int n = 0;
    for (int k = 0; k < want; ++k) {        // `want` selection passes over the counts: <= 28 x n_items, once per CSR
        int64_t best = -1;
        for (int64_t i = 0; i < m->nI; ++i)
            if (cnt[i] > 0 && (best < 0 || cnt[i] > cnt[best])) best = i;
        // an item is worth a shared-memory row only if it repeats inside a batch often: >= ~1 triple in 4096 of the CSR
        if (best < 0 || (int64_t)cnt[best] * 4096 < m->nnz) break;
        ids[n++] = (int32_t)best;
        cnt[best] = 0;
    }
    free(cnt);
    return pda_set_hot_items(m, ids, n);
}

int pda_set_adam_mode(pda_model* m, int mode) {
    if (!m) return fail(PDA_ERR_ARG, "null model");
    if (mode != PDA_ADAM_DENSE && mode != PDA_ADAM_LAZY && mode != PDA_ADAM_LAZY_USERS)
        return fail(PDA_ERR_ARG, "unknown adam mode %d", mode);
    if (mode != PDA_ADAM_DENSE && m->cfg.train_mode == PDA_TRAIN_TEMP_POP)
        return fail(PDA_ERR_ARG, "the lazy Adam replay covers the two embedding tables only; BPR(t)-pop runs the dense sweep");
    CK(cudaSetDevice(m->cfg.device));
    flush_lazy(m, 0);
    const int want[2] = {mode != PDA_ADAM_DENSE, mode == PDA_ADAM_LAZY};
    for (int t = 0; t < 2; ++t) {
        if (want[t] && !m->adam_lazy[t]) {   // dense -> lazy: every row is current and may carry non-zero slots
            launch_fill_i32(m->applied[t], m->rows[t], (int32_t)m->step_no, 0);
            launch_fill_i32(m->stamp[t], m->rows[t], 1, 0);
        }
        m->adam_lazy[t] = want[t];
    }
    CK(cudaDeviceSynchronize());
    return PDA_OK;
}
int pda_get_adam_powers(pda_model* m, float* out) {
    if (!m || !out) return fail(PDA_ERR_ARG, "null argument");
    CK(cudaSetDevice(m->cfg.device));
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(out, m->pw, 8, cudaMemcpyDeviceToHost));
    return PDA_OK;
}
int pda_set_adam_powers(pda_model* m, const float* in) {
    if (!m || !in) return fail(PDA_ERR_ARG, "null argument");
    CK(cudaSetDevice(m->cfg.device));
    CK(cudaMemcpy(m->pw, in, 8, cudaMemcpyHostToDevice));
    return PDA_OK;
}

int pda_set_train_csr(pda_model* m, const int64_t* indptr, const int32_t* items, const uint8_t* times, int64_t nnz,
                      const int32_t* unique_times, int32_t n_times) {
    if (!m || !indptr || (!items && nnz > 0)) return fail(PDA_ERR_ARG, "null argument");
    if (indptr[0] != 0 || indptr[m->nU] != nnz) return fail(PDA_ERR_ARG, "indptr must span [0, nnz] over n_users rows");
    CK(cudaSetDevice(m->cfg.device));
    // rows must be sorted: the sampler's rejection test and the eval mask walk rely on it
    int64_t n_act = 0;
    for (int64_t u = 0; u < m->nU; ++u) {
        if (indptr[u + 1] < indptr[u]) return fail(PDA_ERR_ARG, "indptr not monotone at row %lld", (long long)u);
        if (indptr[u + 1] > indptr[u]) ++n_act;
        for (int64_t q = indptr[u] + 1; q < indptr[u + 1]; ++q)
            if (items[q] < items[q - 1]) return fail(PDA_ERR_ARG, "items of user %lld are not sorted", (long long)u);
    }
    for (int64_t q = 0; q < nnz; ++q)
        if (items[q] < 0 || items[q] >= m->nI) return fail(PDA_ERR_ARG, "item id out of range at %lld", (long long)q);
    int32_t max_time = -1;
    if (times) for (int64_t q = 0; q < nnz; ++q) if ((int32_t)times[q] > max_time) max_time = times[q];
    if (unique_times) for (int32_t q = 0; q < n_times; ++q) if (unique_times[q] > max_time) max_time = unique_times[q];
    if (m->pop_train && m->T_pop > 1 && max_time >= m->T_pop)
        return fail(PDA_ERR_ARG, "stage label %d has no column in the popularity table (T_pop = %d)", max_time, m->T_pop);
    if (m->cfg.train_mode == PDA_TRAIN_TEMP_POP && max_time >= m->cfg.temp_num)
        return fail(PDA_ERR_ARG, "stage label %d outside the temp_num = %d train stages", max_time, m->cfg.temp_num);
    cudaFree(m->indptr); cudaFree(m->items); cudaFree(m->times); cudaFree(m->active); cudaFree(m->unique_times);
    m->indptr = nullptr; m->items = nullptr; m->times = nullptr; m->active = nullptr; m->unique_times = nullptr;
    CK(dmalloc(&m->indptr, (size_t)m->nU + 1));
    CK(dmalloc(&m->items, (size_t)nnz));
    CK(cudaMemcpy(m->indptr, indptr, ((size_t)m->nU + 1) * 8, cudaMemcpyHostToDevice));
    if (nnz) CK(cudaMemcpy(m->items, items, (size_t)nnz * 4, cudaMemcpyHostToDevice));
    if (times) {
        CK(dmalloc(&m->times, (size_t)nnz));
        if (nnz) CK(cudaMemcpy(m->times, times, (size_t)nnz, cudaMemcpyHostToDevice));
    }
    // users with training data, ascending (the reference samples from train_user_list.keys())
    int32_t* act = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n_act > 0 ? n_act : 1));
    if (!act) return fail(PDA_ERR_STATE, "out of host memory");
    int64_t k = 0;
    for (int64_t u = 0; u < m->nU; ++u) if (indptr[u + 1] > indptr[u]) act[k++] = (int32_t)u;
    CK(dmalloc(&m->active, (size_t)n_act));
    if (n_act) CK(cudaMemcpy(m->active, act, (size_t)n_act * 4, cudaMemcpyHostToDevice));
    free(act);
    m->n_act = n_act; m->nnz = nnz; m->max_time = max_time;
    m->n_times = 0;
    if (unique_times && n_times > 0) {
        CK(dmalloc(&m->unique_times, (size_t)n_times));
        CK(cudaMemcpy(m->unique_times, unique_times, (size_t)n_times * 4, cudaMemcpyHostToDevice));
        m->n_times = n_times;
    }
    return hot_items_from_csr(m);
}

int pda_set_train_pop(pda_model* m, const float* pop, int32_t T_pop) {
    if (!m || !pop || T_pop < 1) return fail(PDA_ERR_ARG, "bad argument");
    // the sampler reads pop[item, stage]: every stage label of the train CSR needs a column (the reference's numpy
    // lookup raises IndexError otherwise, train_new_api.py:402-403)
    if (T_pop > 1 && m->max_time >= T_pop)
        return fail(PDA_ERR_ARG, "stage label %d has no column in the popularity table (T_pop = %d)", m->max_time, T_pop);
    CK(cudaSetDevice(m->cfg.device));
    cudaFree(m->pop_train); m->pop_train = nullptr;
    CK(dmalloc(&m->pop_train, (size_t)m->nI * T_pop));
    CK(cudaMemcpy(m->pop_train, pop, (size_t)m->nI * T_pop * 4, cudaMemcpyHostToDevice));
    m->T_pop = T_pop;
    return PDA_OK;
}

int pda_set_train_csr_device(pda_model* m, const int64_t* indptr_d, const int32_t* items_d, const uint8_t* times_d,
                             int64_t nnz, const int32_t* active_d, int64_t n_act, const int32_t* unique_times, int32_t n_times) {
    if (!m || !indptr_d || !items_d || !active_d || n_act < 1) return fail(PDA_ERR_ARG, "bad argument");
    CK(cudaSetDevice(m->cfg.device));
    cudaFree(m->indptr); cudaFree(m->items); cudaFree(m->times); cudaFree(m->active); cudaFree(m->unique_times);
    m->indptr = nullptr; m->items = nullptr; m->times = nullptr; m->active = nullptr; m->unique_times = nullptr;
    CK(dmalloc(&m->indptr, (size_t)m->nU + 1));
    CK(dmalloc(&m->items, (size_t)nnz));
    CK(dmalloc(&m->active, (size_t)n_act));
    CK(cudaMemcpy(m->indptr, indptr_d, ((size_t)m->nU + 1) * 8, cudaMemcpyDeviceToDevice));
    CK(cudaMemcpy(m->items, items_d, (size_t)nnz * 4, cudaMemcpyDeviceToDevice));
    CK(cudaMemcpy(m->active, active_d, (size_t)n_act * 4, cudaMemcpyDeviceToDevice));
    if (times_d) {
        CK(dmalloc(&m->times, (size_t)nnz));
        CK(cudaMemcpy(m->times, times_d, (size_t)nnz, cudaMemcpyDeviceToDevice));
    }
    m->n_act = n_act; m->nnz = nnz; m->n_times = 0; m->max_time = -1;   // device arrays are taken as they are
    if (unique_times && n_times > 0) {
        CK(dmalloc(&m->unique_times, (size_t)n_times));
        CK(cudaMemcpy(m->unique_times, unique_times, (size_t)n_times * 4, cudaMemcpyHostToDevice));
        m->n_times = n_times;
    }
    return hot_items_from_csr(m);
}

static int do_sample(pda_model* m, uint32_t seed, uint32_t epoch, uint32_t step, int64_t B, cudaStream_t st, int slot = 0) {
    if (!m->indptr) return fail(PDA_ERR_STATE, "pda_set_train_csr has not been called");
    if (m->n_act < 1) return fail(PDA_ERR_STATE, "no user has training data");
    if (B < 1 || B > m->cap) return fail(PDA_ERR_ARG, "B=%lld exceeds the batch capacity %lld", (long long)B, (long long)m->cap);
    if (m->cfg.train_mode == PDA_TRAIN_S_CONDITION && !m->pop_train)
        return fail(PDA_ERR_STATE, "train_mode s_condition needs pda_set_train_pop");
    if (m->pop_train && m->T_pop > 1 && !m->times) return fail(PDA_ERR_STATE, "time-dependent popularity needs interaction times");
    if (m->cfg.train_mode == PDA_TRAIN_TEMP_POP && !m->times) return fail(PDA_ERR_STATE, "train_mode temp_pop needs interaction times");
    SamplerArgs a;
    memset(&a, 0, sizeof(a));
    a.seed = seed; a.epoch = epoch; a.step = step; a.B = B;
    a.active_users = m->active; a.n_act = m->n_act;
    a.indptr = m->indptr; a.items = m->items; a.times = m->times; a.n_items = (int32_t)m->nI;
    a.unique_times = m->unique_times; a.n_times = m->n_times;
    a.pop_train = m->cfg.train_mode == PDA_TRAIN_S_CONDITION ? m->pop_train : nullptr; a.T_pop = m->T_pop;
    a.users_out = m->b_users; a.pos_out = m->b_pos; a.neg_out = m->b_neg; a.time_out = m->b_time;
    a.pos_pop_out = m->b_pp; a.neg_pop_out = m->b_np;
    if (slot) {      // second batch slot (pipelined callers; BPR(t)-pop never gets here: it needs b_time)
        a.users_out = m->b2_users; a.pos_out = m->b2_pos; a.neg_out = m->b2_neg; a.pos_pop_out = m->b2_pp; a.neg_pop_out = m->b2_np;
    }
    { ProfScope ps(m, PDA_PROF_SAMPLER, st); launch_sampler(a, st); }
    m->batch_uniq = B <= m->n_act;
    return PDA_OK;
}

int pda_sample_batch(pda_model* m, uint32_t seed, uint32_t epoch, uint32_t step, int64_t B, void* stream) {
    if (!m) return fail(PDA_ERR_ARG, "null model");
    CK(cudaSetDevice(m->cfg.device));
    int rc = do_sample(m, seed, epoch, step, B, (cudaStream_t)stream);
    if (rc) return rc;
    CK(cudaGetLastError());
    return PDA_OK;
}

int pda_get_batch(pda_model* m, int64_t B, int32_t* users, int32_t* pos, int32_t* neg, int32_t* time, float* pp,
                  float* np_) {
    if (!m || B < 1 || B > m->cap) return fail(PDA_ERR_ARG, "bad argument");
    CK(cudaSetDevice(m->cfg.device));
    CK(cudaDeviceSynchronize());
    if (users) CK(cudaMemcpy(users, m->b_users, (size_t)B * 4, cudaMemcpyDeviceToHost));
    if (pos) CK(cudaMemcpy(pos, m->b_pos, (size_t)B * 4, cudaMemcpyDeviceToHost));
    if (neg) CK(cudaMemcpy(neg, m->b_neg, (size_t)B * 4, cudaMemcpyDeviceToHost));
    if (time) CK(cudaMemcpy(time, m->b_time, (size_t)B * 4, cudaMemcpyDeviceToHost));
    if (pp) CK(cudaMemcpy(pp, m->b_pp, (size_t)B * 4, cudaMemcpyDeviceToHost));
    if (np_) CK(cudaMemcpy(np_, m->b_np, (size_t)B * 4, cudaMemcpyDeviceToHost));
    return PDA_OK;
}

// gather -> loss -> gradient scatter: ONE kernel (gradients land in the table-shaped accumulators G)
static int enqueue_fwd_bwd(pda_model* m, const int32_t* users, const int32_t* pos, const int32_t* neg, const float* pp,
                           const float* np_, int64_t B, int uniq, cudaStream_t st, bool will_apply = true) {
    StepArgs s;
    memset(&s, 0, sizeof(s));
    s.U = m->W[0]; s.I = m->W[1]; s.GU = m->G[0]; s.GI = m->G[1];
    s.users = users; s.pos = pos; s.neg = neg; s.pos_pop = pp; s.neg_pop = np_;
    s.B = B; s.d = m->d;
    s.lb = (float)((double)m->cfg.regs / (double)m->cfg.batch_size);
    m->last_B = m->global_batch > 0 ? m->global_batch : B;
    s.invB = 1.0f / (float)m->last_B;
    s.loss_acc = m->loss_acc;
    m->cur_users = users; m->cur_pos = pos; m->cur_neg = neg; m->cur_B = B;
    // distinct users + lazily kept user table + an Adam step that is certain to follow: the step kernel itself
    // catches the user rows up and applies their update (bpr_step_kernel UMODE 2)
    m->cur_fused = m->adam_lazy[0] && uniq && will_apply && m->fuse_user_adam && !m->deterministic;
    if (m->deterministic) {
        // per-triple gradient rows go to a slot buffer; pda_segsum.cu sums the slots of each row in occurrence order
        const size_t n2 = (size_t)2 * B;
        CK(ensure_dev(&m->gslots, &m->gslots_bytes, (size_t)3 * B * m->d * 4));
        CK(ensure_dev(&m->seg_work, &m->seg_work_bytes, n2 * 4 * 4));
        CK(ensure_dev(&m->seg_temp, &m->seg_temp_bytes, segsum_temp_bytes((int64_t)n2) + 256));
        s.Gslots = (float*)m->gslots;
    }
    if ((m->adam_lazy[0] && !m->cur_fused) || m->adam_lazy[1]) {   // rows of this batch replay the steps they skipped
        LazyArgs la;
        lazy_args(m, &la);
        if (m->cur_fused) la.lazy[0] = 0;
        ProfScope ps(m, PDA_PROF_ADAM_CATCHUP, st);
        if (launch_adam_lazy_rows(la, 0, st)) return fail(PDA_ERR_ARG, "unsupported embed_size %d", m->d);
    }
    if (m->cur_fused) {
        s.fuse_user_adam = 1; s.Uw = m->W[0]; s.MU = m->Mo[0]; s.VU = m->Vo[0]; s.appliedU = m->applied[0];
        s.stampU = m->stamp[0]; s.lr_hist = m->lr_hist - m->lr_base; s.pw = m->pw; s.lr = m->cfg.lr; s.step_no = m->step_no;
        s.stats = m->lazy_stats;
    }
    s.pop_mode = m->cfg.train_mode == PDA_TRAIN_S_CONDITION ? 1 : m->cfg.train_mode == PDA_TRAIN_TEMP_POP ? 2 : 0;
    s.uniq_users = uniq;
    s.hot_slot = m->hot_slot; s.hot_ids = m->hot_ids; s.n_hot = m->n_hot;
    if (s.pop_mode == 2) {   // BPR(t)-pop: the stage of each triple rides in the internal batch (b_time)
        if (pp) {   // explicit batch: the reference passes `temp` as fp32 through the pos_pop slot (train_new_api.py:544,565)
            if (B > m->cap) return fail(PDA_ERR_ARG, "B exceeds the batch capacity");
            launch_f32_to_i32(pp, m->b_time, B, m->cfg.temp_num - 1, st);
        }
        s.temp = m->b_time; s.temp_num = m->cfg.temp_num;
        s.ub = m->W[2]; s.ib = m->W[3]; s.Gub = m->G[2]; s.Gib = m->G[3];
    }
    {
        ProfScope ps(m, PDA_PROF_STEP, st);
        if (launch_bpr_step(s, st)) return fail(PDA_ERR_ARG, "unsupported embed_size %d", m->d);
    }
    if (m->deterministic) {
        auto bits_of = [](int64_t n) { int b = 1; while (((int64_t)1 << b) < n) ++b; return b; };
        ProfScope ps(m, PDA_PROF_ADAM_CATCHUP, st);
        if (launch_segment_sum(pos, neg, B, (const float*)m->gslots, m->d, m->G[1], (int32_t*)m->seg_work, m->seg_temp, m->seg_temp_bytes,
                               bits_of(m->nI), st))
            return fail(PDA_ERR_CUDA, "segment sum (items) failed");
        if (!uniq && launch_segment_sum(users, nullptr, B, (const float*)m->gslots + (size_t)2 * B * m->d, m->d, m->G[0], (int32_t*)m->seg_work,
                                        m->seg_temp, m->seg_temp_bytes, bits_of(m->nU), st))
            return fail(PDA_ERR_CUDA, "segment sum (users) failed");
    }
    return PDA_OK;
}

// TF1 Adam sweep over both tables (one kernel) + loss / beta-power bookkeeping
// parts: 1 = the rank-local half (lazily maintained tables: their gradient never leaves the GPU),
//        2 = the exchanged half (dense sweep of the remaining variables) + loss / beta-power bookkeeping, 3 = both,
//        8 = bookkeeping only (the dense sweep was issued in row ranges through pda_adam_dense_rows).
// Data-parallel callers run part 1 while the item-gradient all-reduce is in flight, then part 2.
static int enqueue_adam(pda_model* m, bool apply_adam, cudaStream_t st, int parts = 3) {
    float* lr_slot = nullptr;
    if (apply_adam && (parts & 1)) {
        const bool any_lazy = (m->adam_lazy[0] && !m->cur_fused) || m->adam_lazy[1];
        if (any_lazy) {
            if (!m->cur_users) return fail(PDA_ERR_STATE, "pda_adam_apply without a preceding forward/backward");
            LazyArgs la;
            lazy_args(m, &la);
            if (m->cur_fused) la.lazy[0] = 0;
            ProfScope ps(m, PDA_PROF_ADAM, st);
            if (launch_adam_lazy_rows(la, 1, st)) return fail(PDA_ERR_ARG, "unsupported embed_size %d", m->d);
        }
    }
    if (!(parts & (2 | 8))) return PDA_OK;
    if (apply_adam && (parts & 2)) {
        AdamArgs a;
        memset(&a, 0, sizeof(a));
        bool any_dense = false;
        for (int t = 0; t < m->n_arr; ++t) {
            if (t < 2 && m->adam_lazy[t]) continue;     // n4 = 0: left to the lazy kernels
            a.W[t] = m->W[t]; a.m[t] = m->Mo[t]; a.v[t] = m->Vo[t]; a.G[t] = m->G[t]; a.n4[t] = m->n4[t];
            any_dense = true;
        }
        a.pw = m->pw; a.lr = m->cfg.lr;
        if (any_dense) { ProfScope ps(m, PDA_PROF_ADAM, st); launch_adam_dense(a, st); }
    }
    if (apply_adam) lr_slot = m->lr_hist + (m->step_no - m->lr_base);
    launch_finish_step(m->loss_acc, m->loss3, m->loss_sum, m->pw, m->last_B, m->cfg.regs, m->cfg.batch_size, apply_adam ? 1 : 0,
                       m->cfg.lr, lr_slot, st);
    if (apply_adam) {
        ++m->step_no;
        if (m->step_no - m->lr_base >= PDA_LR_CAP) {   // history ring full: bring every row up to date, start a new ring
            flush_lazy(m, st);
            m->lr_base = m->step_no;
        }
    }
    return PDA_OK;
}

static int enqueue_step(pda_model* m, const int32_t* users, const int32_t* pos, const int32_t* neg, const float* pp,
                        const float* np_, int64_t B, int uniq, bool apply_adam, cudaStream_t st) {
    int rc = enqueue_fwd_bwd(m, users, pos, neg, pp, np_, B, uniq, st, apply_adam);
    if (rc) return rc;
    return enqueue_adam(m, apply_adam, st);
}

int pda_set_global_batch(pda_model* m, int64_t global_batch) {
    if (!m || global_batch < 0) return fail(PDA_ERR_ARG, "bad argument");
    m->global_batch = global_batch;
    return PDA_OK;
}

void* pda_grad_ptr(pda_model* m, int which) {
    if (!m) return nullptr;
    if (which == PDA_TABLE_USER) return m->G[0];
    if (which == PDA_TABLE_ITEM) return m->G[1];
    return nullptr;
}

void* pda_loss_acc_ptr(pda_model* m) { return m ? m->loss_acc : nullptr; }

int pda_forward_backward_device(pda_model* m, const int32_t* users, const int32_t* pos, const int32_t* neg,
                                const float* pp, const float* np_, int64_t B, void* stream) {
    if (!m) return fail(PDA_ERR_ARG, "null model");
    CK(cudaSetDevice(m->cfg.device));
    int uniq = 0;
    if (!users) {
        if (B < 1 || B > m->cap) return fail(PDA_ERR_ARG, "B exceeds the batch capacity");
        users = m->b_users; pos = m->b_pos; neg = m->b_neg; pp = m->b_pp; np_ = m->b_np;
        if (m->cfg.train_mode == PDA_TRAIN_TEMP_POP) pp = np_ = nullptr;   // stages already sit in b_time
        uniq = m->batch_uniq;
    } else if (m->cfg.train_mode == PDA_TRAIN_TEMP_POP && !pp) return fail(PDA_ERR_ARG, "temp_pop needs the stage array in the pos_pop slot");
    if (!pos || !neg) return fail(PDA_ERR_ARG, "null index pointer");
    if (m->cfg.train_mode == PDA_TRAIN_S_CONDITION && (!pp || !np_)) return fail(PDA_ERR_ARG, "s_condition needs pos_pop/neg_pop");
    int rc = enqueue_fwd_bwd(m, users, pos, neg, pp, np_, B, uniq, (cudaStream_t)stream);
    if (rc) return rc;
    CK(cudaGetLastError());
    return PDA_OK;
}

int pda_adam_dense_rows(pda_model* m, int which, int64_t row_lo, int64_t row_hi, void* stream) {
    if (!m || (which != PDA_TABLE_USER && which != PDA_TABLE_ITEM)) return fail(PDA_ERR_ARG, "bad argument");
    const int t = which == PDA_TABLE_USER ? 0 : 1;
    if (row_lo < 0 || row_hi > m->rows[t] || row_lo > row_hi) return fail(PDA_ERR_ARG, "row range outside the table");
    if (m->adam_lazy[t]) return fail(PDA_ERR_STATE, "the table is kept lazily: there is no dense sweep to run on it");
    if (row_lo == row_hi) return PDA_OK;
    CK(cudaSetDevice(m->cfg.device));
    AdamArgs a;
    memset(&a, 0, sizeof(a));
    const size_t off = (size_t)row_lo * m->d;
    a.W[0] = m->W[t] + off; a.m[0] = m->Mo[t] + off; a.v[0] = m->Vo[t] + off; a.G[0] = m->G[t] + off;
    a.n4[0] = (row_hi - row_lo) * m->d / 4;
    a.pw = m->pw; a.lr = m->cfg.lr;
    { ProfScope ps(m, PDA_PROF_ADAM, (cudaStream_t)stream); launch_adam_dense(a, (cudaStream_t)stream); }
    CK(cudaGetLastError());
    return PDA_OK;
}

int pda_adam_dense_rows_ext(pda_model* m, int which, int64_t row_lo, int64_t row_hi, const float* grad, void* stream) {
    if (!m || !grad || (which != PDA_TABLE_USER && which != PDA_TABLE_ITEM)) return fail(PDA_ERR_ARG, "bad argument");
    const int t = which == PDA_TABLE_USER ? 0 : 1;
    if (row_lo < 0 || row_hi > m->rows[t] || row_lo > row_hi) return fail(PDA_ERR_ARG, "row range outside the table");
    if (m->adam_lazy[t]) return fail(PDA_ERR_STATE, "the table is kept lazily: there is no dense sweep to run on it");
    if (row_lo == row_hi) return PDA_OK;
    CK(cudaSetDevice(m->cfg.device));
    AdamArgs a;
    memset(&a, 0, sizeof(a));
    const size_t off = (size_t)row_lo * m->d;
    a.W[0] = m->W[t] + off; a.m[0] = m->Mo[t] + off; a.v[0] = m->Vo[t] + off; a.G[0] = const_cast<float*>(grad);
    a.n4[0] = (row_hi - row_lo) * m->d / 4;
    a.pw = m->pw; a.lr = m->cfg.lr; a.keep_g = 1;
    { ProfScope ps(m, PDA_PROF_ADAM, (cudaStream_t)stream); launch_adam_dense(a, (cudaStream_t)stream); }
    CK(cudaGetLastError());
    return PDA_OK;
}

// Move the item table and its gradient accumulator into caller-owned device memory (symmetric / multicast-mapped
// buffers of a data-parallel group): current contents are copied, the library's own arrays are released, and every
// kernel of the model uses the new storage from now on.  The caller keeps the buffers alive until pda_destroy.
int pda_adopt_item_buffers(pda_model* m, float* W_ext, float* G_ext) {
    if (!m || !W_ext || !G_ext) return fail(PDA_ERR_ARG, "null argument");
    if (((uintptr_t)W_ext | (uintptr_t)G_ext) & 15) return fail(PDA_ERR_ARG, "buffers must be 16-byte aligned");
    CK(cudaSetDevice(m->cfg.device));
    flush_lazy(m, 0);
    CK(cudaDeviceSynchronize());
    const size_t bytes = (size_t)m->n4[1] * 16;
    CK(cudaMemcpy(W_ext, m->W[1], bytes, cudaMemcpyDeviceToDevice));
    CK(cudaMemcpy(G_ext, m->G[1], bytes, cudaMemcpyDeviceToDevice));
    if (!m->item_ext) { cudaFree(m->W[1]); cudaFree(m->G[1]); }
    m->W[1] = W_ext; m->G[1] = G_ext; m->item_ext = 1;
    return PDA_OK;
}

// Cross-rank barriers INSIDE the fused exchange kernels: flags = 64 zero-initialised uint32 per rank in symmetric memory
// (flags_local = this rank's, peer_flags[r] = rank r's as mapped here, self included).  With this set the callers of
// pda_dp_exchange_adam[_p2p] no longer bracket the kernel with barriers of their own: the kernel waits at its start until
// every rank has reached it and completes only when every rank's writes have landed.  flags_local == NULL switches it off.
int pda_dp_set_barrier(pda_model* m, uint32_t* flags_local, uint32_t* const* peer_flags, int32_t world, int32_t rank) {
    if (!m) return fail(PDA_ERR_ARG, "null model");
    memset(&m->dp_sync, 0, sizeof(m->dp_sync));
    if (!flags_local) return PDA_OK;
    if (!peer_flags || world < 1 || world > 8 || rank < 0 || rank >= world) return fail(PDA_ERR_ARG, "bad argument");
    m->dp_sync.local = flags_local; m->dp_sync.world = world; m->dp_sync.rank = rank; m->dp_sync.epoch = 0;
    for (int r = 0; r < world; ++r) m->dp_sync.peer[r] = peer_flags[r];
    return PDA_OK;
}

// Point the item-gradient accumulator at another caller-owned, ZERO-filled [n_items, d] buffer (double buffering: the
// previous one is zeroed off the critical path while the next step accumulates into this one).
int pda_set_item_grad_buffer(pda_model* m, float* G_ext) {
    if (!m || !G_ext) return fail(PDA_ERR_ARG, "null argument");
    if (!m->item_ext) return fail(PDA_ERR_STATE, "pda_adopt_item_buffers has not been called");
    if ((uintptr_t)G_ext & 15) return fail(PDA_ERR_ARG, "buffer must be 16-byte aligned");
    m->G[1] = G_ext;
    return PDA_OK;
}

// reduce-scatter + sliced Adam + all-gather of the item table in one kernel over NVLink multicast (pda_exchange.cu).
// mcG / mcW: the MULTICAST addresses of the (adopted) accumulator / table, rows [row_lo, row_hi) = this rank's slice.
// The caller orders it between two cross-rank barriers on `stream`; the accumulator is left as it is (zero it after
// the second barrier).
int pda_dp_exchange_adam(pda_model* m, const float* mcG, float* mcW, int64_t row_lo, int64_t row_hi, void* stream) {
    if (!m || !mcG || !mcW) return fail(PDA_ERR_ARG, "null argument");
    if (row_lo < 0 || row_hi > m->nI || row_lo > row_hi) return fail(PDA_ERR_ARG, "row range outside the item table");
    if (m->adam_lazy[1]) return fail(PDA_ERR_STATE, "the item table is kept lazily: there is no dense sweep to run on it");
    if (row_lo == row_hi) return PDA_OK;
    CK(cudaSetDevice(m->cfg.device));
    const size_t off = (size_t)row_lo * m->d;
    { ProfScope ps(m, PDA_PROF_ADAM, (cudaStream_t)stream);
      if (m->dp_sync.local) ++m->dp_sync.epoch;
      launch_dp_exchange_adam(mcG + off, mcW + off, m->G[1] + off, m->W[1] + off, m->Mo[1] + off, m->Vo[1] + off, (row_hi - row_lo) * m->d / 4, m->pw,
                              m->cfg.lr, &m->dp_sync, (cudaStream_t)stream); }
    CK(cudaGetLastError());
    return PDA_OK;
}

// the same over unicast peer pointers: peer_G[r] / peer_W[r] = rank r's accumulator / table as mapped in THIS process
// (symmetric memory), r = 0..world-1, self included
int pda_dp_exchange_adam_p2p(pda_model* m, const float* const* peer_G, float* const* peer_W, int32_t world, int32_t self,
                             int64_t row_lo, int64_t row_hi, void* stream) {
    if (!m || !peer_G || !peer_W || world < 1 || world > 8 || self < 0 || self >= world) return fail(PDA_ERR_ARG, "bad argument");
    if (row_lo < 0 || row_hi > m->nI || row_lo > row_hi) return fail(PDA_ERR_ARG, "row range outside the item table");
    if (m->adam_lazy[1]) return fail(PDA_ERR_STATE, "the item table is kept lazily: there is no dense sweep to run on it");
    if (peer_W[self] != m->W[1]) return fail(PDA_ERR_STATE, "peer_W[self] is not the model's (adopted) item table");
    if (row_lo == row_hi) return PDA_OK;
    CK(cudaSetDevice(m->cfg.device));
    const size_t off = (size_t)row_lo * m->d;
    { ProfScope ps(m, PDA_PROF_ADAM, (cudaStream_t)stream);
      if (m->dp_sync.local) ++m->dp_sync.epoch;
      launch_dp_exchange_adam_p2p(peer_G, peer_W, world, self, (int64_t)off, m->Mo[1] + off, m->Vo[1] + off, (row_hi - row_lo) * m->d / 4,
                                  m->pw, m->cfg.lr, &m->dp_sync, (cudaStream_t)stream); }
    CK(cudaGetLastError());
    return PDA_OK;
}

int pda_adam_apply_part(pda_model* m, int part, void* stream) {
    if (!m || (part != 1 && part != 2 && part != 3 && part != 8)) return fail(PDA_ERR_ARG, "bad argument");
    CK(cudaSetDevice(m->cfg.device));
    int rc = enqueue_adam(m, true, (cudaStream_t)stream, part);
    if (rc) return rc;
    CK(cudaGetLastError());
    return PDA_OK;
}

int pda_adam_apply(pda_model* m, void* stream) {
    if (!m) return fail(PDA_ERR_ARG, "null model");
    CK(cudaSetDevice(m->cfg.device));
    int rc = enqueue_adam(m, true, (cudaStream_t)stream);
    if (rc) return rc;
    CK(cudaGetLastError());
    return PDA_OK;
}

int pda_stage_batch_host(pda_model* m, const int32_t* users, const int32_t* pos, const int32_t* neg, const float* pp,
                         const float* np_, int64_t B, void* stream);

int pda_train_step_device(pda_model* m, const int32_t* users, const int32_t* pos, const int32_t* neg,
                          const float* pp, const float* np_, int64_t B, void* stream) {
    if (!m) return fail(PDA_ERR_ARG, "null model");
    CK(cudaSetDevice(m->cfg.device));
    int uniq = 0;
    if (!users) {   // internal batch from the device sampler: users are distinct iff B <= #active users
        if (B < 1 || B > m->cap) return fail(PDA_ERR_ARG, "B exceeds the batch capacity");
        users = m->b_users; pos = m->b_pos; neg = m->b_neg; pp = m->b_pp; np_ = m->b_np;
        if (m->cfg.train_mode == PDA_TRAIN_TEMP_POP) pp = np_ = nullptr;   // stages already sit in b_time
        uniq = m->batch_uniq;
    } else if (m->cfg.train_mode == PDA_TRAIN_TEMP_POP && !pp) return fail(PDA_ERR_ARG, "temp_pop needs the stage array in the pos_pop slot");
    if (!pos || !neg) return fail(PDA_ERR_ARG, "null index pointer");
    if (m->cfg.train_mode == PDA_TRAIN_S_CONDITION && (!pp || !np_)) return fail(PDA_ERR_ARG, "s_condition needs pos_pop/neg_pop");
    int rc = enqueue_step(m, users, pos, neg, pp, np_, B, uniq, true, (cudaStream_t)stream);
    if (rc) return rc;
    CK(cudaGetLastError());
    return PDA_OK;
}

// enqueue the id check of a device batch on `st`: flags land in chk_pinned[2*slot] (users repeat) and [2*slot+1] (bad id)
static int check_batch(pda_model* m, const int32_t* users, const int32_t* pos, const int32_t* neg, int64_t B, bool want_uniq,
                       int slot, cudaStream_t st) {
    if (want_uniq && !m->seen) { CK(dmalloc(&m->seen, (size_t)m->nU)); CK(cudaMemset(m->seen, 0, (size_t)m->nU * 4)); }
    int32_t* fl = m->chk_flags + 2 * slot;
    CK(cudaMemsetAsync(fl, 0, 8, st));
    launch_batch_check(users, pos, neg, B, (int32_t)m->nU, (int32_t)m->nI, want_uniq ? m->seen : nullptr, ++m->seen_tag, fl, st);
    CK(cudaMemcpyAsync(m->chk_pinned + 2 * slot, fl, 8, cudaMemcpyDeviceToHost, st));
    return PDA_OK;
}

static int stage_batch(pda_model* m, const int32_t* users, const int32_t* pos, const int32_t* neg, const float* pp,
                       const float* np_, int64_t B, cudaStream_t st) {
    if (B < 1 || B > m->cap) return fail(PDA_ERR_ARG, "B=%lld exceeds the batch capacity %lld", (long long)B, (long long)m->cap);
    const bool temp = m->cfg.train_mode == PDA_TRAIN_TEMP_POP;
    const bool pop = m->cfg.train_mode == PDA_TRAIN_S_CONDITION || temp;   // two fp32 side arrays travel with the batch
    if (temp && pp && !np_) np_ = pp;                                       // `raw` (= arange(B)) is never read
    if (!users || !pos || !neg || (pop && (!pp || !np_))) return fail(PDA_ERR_ARG, "null batch pointer");
    // one pinned staging block, one DMA per array
    m->batch_uniq = 0;
    const size_t nb = (size_t)B * 4;
    // caller buffers already pinned (pda_host_alloc / cudaHostAlloc): DMA straight from them
    cudaPointerAttributes at;
    bool pinned = cudaPointerGetAttributes(&at, users) == cudaSuccess && at.type == cudaMemoryTypeHost;
    if (!pinned) cudaGetLastError();
    const char *su = (const char*)users, *sp = (const char*)pos, *sn = (const char*)neg, *spp = (const char*)pp,
               *snp = (const char*)np_;
    if (!pinned) {   // pageable memory: one pinned staging block, then one DMA per array
        CK(ensure_pinned(&m->stage_pinned, &m->stage_bytes, nb * 5));
        char* s = (char*)m->stage_pinned;
        memcpy(s, users, nb); memcpy(s + nb, pos, nb); memcpy(s + 2 * nb, neg, nb);
        su = s; sp = s + nb; sn = s + 2 * nb;
        if (pop) { memcpy(s + 3 * nb, pp, nb); memcpy(s + 4 * nb, np_, nb); spp = s + 3 * nb; snp = s + 4 * nb; }
    }
    CK(cudaMemcpyAsync(m->b_users, su, nb, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(m->b_pos, sp, nb, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(m->b_neg, sn, nb, cudaMemcpyHostToDevice, st));
    if (pop) {
        CK(cudaMemcpyAsync(m->b_pp, spp, nb, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(m->b_np, snp, nb, cudaMemcpyHostToDevice, st));
    }
    return PDA_OK;
}

int pda_stage_batch_host(pda_model* m, const int32_t* users, const int32_t* pos, const int32_t* neg, const float* pp,
                         const float* np_, int64_t B, void* stream) {
    if (!m) return fail(PDA_ERR_ARG, "null model");
    CK(cudaSetDevice(m->cfg.device));
    return stage_batch(m, users, pos, neg, pp, np_, B, (cudaStream_t)stream);
}

// pda_stage_batch_host for callers that overlap the copies with device work (data-parallel host-batch loops): copies +
// the id / distinct-users check are enqueued on `copy_stream` and the call returns; pda_staged_batch_wait blocks the
// HOST until they are done (nothing else), validates, and makes `stream` wait for them.
int pda_stage_batch_host_async(pda_model* m, const int32_t* users, const int32_t* pos, const int32_t* neg, const float* pp,
                               const float* np_, int64_t B, void* copy_stream) {
    if (!m) return fail(PDA_ERR_ARG, "null model");
    CK(cudaSetDevice(m->cfg.device));
    cudaStream_t cs = (cudaStream_t)copy_stream;
    int rc = stage_batch(m, users, pos, neg, pp, np_, B, cs);
    if (rc) return rc;
    rc = check_batch(m, m->b_users, m->b_pos, m->b_neg, B, m->adam_lazy[0] && m->fuse_user_adam, 2, cs);
    if (rc) return rc;
    CK(cudaEventRecord(m->ev_staged, cs));
    m->staged_pending = 1;
    return PDA_OK;
}

int pda_staged_batch_wait(pda_model* m, void* stream) {
    if (!m) return fail(PDA_ERR_ARG, "null model");
    if (!m->staged_pending) return fail(PDA_ERR_STATE, "no batch was staged with pda_stage_batch_host_async");
    CK(cudaSetDevice(m->cfg.device));
    CK(cudaEventSynchronize(m->ev_staged));
    CK(cudaStreamWaitEvent((cudaStream_t)stream, m->ev_staged, 0));
    m->staged_pending = 0;
    if (m->chk_pinned[5]) return fail(PDA_ERR_ARG, "batch holds a user / item id outside [0, n_users) / [0, n_items)");
    m->batch_uniq = (m->adam_lazy[0] && m->fuse_user_adam && m->chk_pinned[4] == 0) ? 1 : 0;
    return PDA_OK;
}

// {loss, mf_loss, reg_loss} of the last enqueued step -> pinned host memory, enqueued on `stream` (no synchronisation:
// the caller reads it after its own sync)
int pda_read_loss_async(pda_model* m, float* pinned_dst3, void* stream) {
    if (!m || !pinned_dst3) return fail(PDA_ERR_ARG, "null argument");
    CK(cudaSetDevice(m->cfg.device));
    CK(cudaMemcpyAsync(pinned_dst3, m->loss3, 12, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return PDA_OK;
}

int pda_train_step_host(pda_model* m, const int32_t* users, const int32_t* pos, const int32_t* neg,
                        const float* pp, const float* np_, int64_t B, float* loss3_out) {
    if (!m) return fail(PDA_ERR_ARG, "null model");
    CK(cudaSetDevice(m->cfg.device));
    int rc = stage_batch(m, users, pos, neg, pp, np_, B, 0);
    if (rc) return rc;
    // One pass over the ids on the device (~50 us at B = 2^20): ids outside their table are an error (the reference's
    // embedding_lookup raises), and the fused user-row path applies only when the users are distinct -- the reference's
    // batches are (rd.sample, train_new_api.py:384-385) but a host caller may pass anything.
    int uniq = 0;
    {
        const bool want_uniq = m->adam_lazy[0] && m->fuse_user_adam;
        int rc2 = check_batch(m, m->b_users, m->b_pos, m->b_neg, B, want_uniq, 0, 0);
        if (rc2) return rc2;
        CK(cudaStreamSynchronize(0));
        if (m->chk_pinned[1]) return fail(PDA_ERR_ARG, "batch holds a user / item id outside [0, n_users) / [0, n_items)");
        uniq = want_uniq && m->chk_pinned[0] == 0;
    }
    rc = enqueue_step(m, m->b_users, m->b_pos, m->b_neg, m->b_pp, m->b_np, B, uniq, true, 0);
    if (rc) return rc;
    CK(cudaMemcpyAsync(m->loss3_pinned, m->loss3, 12, cudaMemcpyDeviceToHost, 0));
    CK(cudaStreamSynchronize(0));
    CK(cudaGetLastError());
    if (loss3_out) memcpy(loss3_out, m->loss3_pinned, 12);
    return PDA_OK;
}

// second batch slot + side stream + events of the pipelined paths (pda_train_steps_host, pda_train_steps_sampled)
static int ensure_pipe(pda_model* m) {
    if (m->pipe_ready) return PDA_OK;
    CK(dmalloc(&m->b2_users, (size_t)m->cap)); CK(dmalloc(&m->b2_pos, (size_t)m->cap)); CK(dmalloc(&m->b2_neg, (size_t)m->cap));
    CK(dmalloc(&m->b2_pp, (size_t)m->cap)); CK(dmalloc(&m->b2_np, (size_t)m->cap));
    CK(cudaStreamCreateWithFlags(&m->copy_st, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
        CK(cudaEventCreateWithFlags(&m->ev_copied[i], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&m->ev_stepped[i], cudaEventDisableTiming));
    }
    CK(cudaEventCreateWithFlags(&m->ev_start, cudaEventDisableTiming));
    m->pipe_ready = 1;
    return PDA_OK;
}

// n host batches, pipelined: the H2D copies (+ the distinct-users check) of batch k+1 run on a copy stream while step
// k computes; two device batch slots, events in both directions, one host wait per batch on the COPY side only.
int pda_train_steps_host(pda_model* m, const int32_t* users, const int32_t* pos, const int32_t* neg, const float* pp,
                         const float* np_, int32_t n_batches, int64_t B, float* loss3_out) {
    if (!m || !users || !pos || !neg || n_batches < 1) return fail(PDA_ERR_ARG, "bad argument");
    if (B < 1 || B > m->cap) return fail(PDA_ERR_ARG, "B=%lld exceeds the batch capacity %lld", (long long)B, (long long)m->cap);
    if (m->cfg.train_mode == PDA_TRAIN_TEMP_POP) return fail(PDA_ERR_STATE, "pda_train_steps_host covers BPRMF / PD / PDG batches");
    const bool pop = m->cfg.train_mode == PDA_TRAIN_S_CONDITION;
    if (pop && (!pp || !np_)) return fail(PDA_ERR_ARG, "s_condition needs pos_pop/neg_pop");
    CK(cudaSetDevice(m->cfg.device));
    cudaPointerAttributes at;
    if (!(cudaPointerGetAttributes(&at, users) == cudaSuccess && at.type == cudaMemoryTypeHost)) {
        cudaGetLastError();
        return fail(PDA_ERR_ARG, "pda_train_steps_host needs pinned host batches (pda_host_alloc): pageable memory cannot overlap with compute");
    }
    { int rc = ensure_pipe(m); if (rc) return rc; }
    CK(ensure_pinned((void**)&m->loss_ring, &m->loss_ring_bytes, (size_t)n_batches * 16 + 64));
    const bool check = m->adam_lazy[0] && m->fuse_user_adam;
    int32_t* bu[2] = {m->b_users, m->b2_users}; int32_t* bp[2] = {m->b_pos, m->b2_pos}; int32_t* bn[2] = {m->b_neg, m->b2_neg};
    float* bpp[2] = {m->b_pp, m->b2_pp}; float* bnp[2] = {m->b_np, m->b2_np};
    const size_t nb = (size_t)B * 4;
    CK(cudaStreamSynchronize(0));      // everything enqueued before this call is done: both slots are free
    auto issue_copy = [&](int k) -> cudaError_t {
        const int sl = k & 1;
        cudaError_t e;
        if (k >= 2 && (e = cudaStreamWaitEvent(m->copy_st, m->ev_stepped[sl], 0)) != cudaSuccess) return e;   // step k-2 has read the slot
        if ((e = cudaMemcpyAsync(bu[sl], users + (size_t)k * B, nb, cudaMemcpyHostToDevice, m->copy_st)) != cudaSuccess) return e;
        if ((e = cudaMemcpyAsync(bp[sl], pos + (size_t)k * B, nb, cudaMemcpyHostToDevice, m->copy_st)) != cudaSuccess) return e;
        if ((e = cudaMemcpyAsync(bn[sl], neg + (size_t)k * B, nb, cudaMemcpyHostToDevice, m->copy_st)) != cudaSuccess) return e;
        if (pop) {
            if ((e = cudaMemcpyAsync(bpp[sl], pp + (size_t)k * B, nb, cudaMemcpyHostToDevice, m->copy_st)) != cudaSuccess) return e;
            if ((e = cudaMemcpyAsync(bnp[sl], np_ + (size_t)k * B, nb, cudaMemcpyHostToDevice, m->copy_st)) != cudaSuccess) return e;
        }
        // id range + distinct users, on the copy stream (see pda_train_step_host)
        if (check_batch(m, bu[sl], bp[sl], bn[sl], B, check, sl, m->copy_st)) return cudaErrorUnknown;
        return cudaEventRecord(m->ev_copied[sl], m->copy_st);
    };
    CK(issue_copy(0));
    for (int k = 0; k < n_batches; ++k) {
        const int sl = k & 1;
        if (k + 1 < n_batches) CK(issue_copy(k + 1));
        CK(cudaEventSynchronize(m->ev_copied[sl]));            // host: batch k is on the device (step k-1 may still be running)
        if (m->chk_pinned[2 * sl + 1]) {
            cudaDeviceSynchronize();
            return fail(PDA_ERR_ARG, "batch %d holds a user / item id outside [0, n_users) / [0, n_items)", k);
        }
        const int uniq = check ? m->chk_pinned[2 * sl] == 0 : 0;
        CK(cudaStreamWaitEvent(0, m->ev_copied[sl], 0));
        m->batch_uniq = 0;
        int rc = enqueue_step(m, bu[sl], bp[sl], bn[sl], pop ? bpp[sl] : nullptr, pop ? bnp[sl] : nullptr, B, uniq, true, 0);
        if (rc) return rc;
        CK(cudaMemcpyAsync(m->loss_ring + (size_t)k * 4, m->loss3, 12, cudaMemcpyDeviceToHost, 0));
        CK(cudaEventRecord(m->ev_stepped[sl], 0));
    }
    CK(cudaStreamSynchronize(0));
    CK(cudaGetLastError());
    if (loss3_out)
        for (int k = 0; k < n_batches; ++k) memcpy(loss3_out + (size_t)k * 3, m->loss_ring + (size_t)k * 4, 12);
    return PDA_OK;
}

int pda_train_steps_sampled(pda_model* m, uint32_t seed, uint32_t epoch, uint32_t step0, int32_t n_steps, int64_t B,
                            void* stream) {
    if (!m) return fail(PDA_ERR_ARG, "null model");
    CK(cudaSetDevice(m->cfg.device));
    cudaStream_t st = (cudaStream_t)stream;
    const bool tmode = m->cfg.train_mode == PDA_TRAIN_TEMP_POP;
    static const bool overlap = !(getenv("PDA_OVERLAP_SAMPLER") && atoi(getenv("PDA_OVERLAP_SAMPLER")) == 0);
    if (n_steps >= 2 && !tmode && overlap) {
        // The sampler of step k+1 depends on nothing step k computes: it runs on a side stream into the other batch slot
        // under step k's Adam sweep (latency-bound CSR walks under an HBM-bound sweep) -- it starts when the fused step
        // kernel of step k is done, so that kernel keeps the GPU to itself.  Slot of step k = (n_steps-1-k) & 1, so the
        // last batch ends in the primary buffers like in the sequential path.
        { int rc = ensure_pipe(m); if (rc) return rc; }
        int32_t* bu[2] = {m->b_users, m->b2_users}; int32_t* bp[2] = {m->b_pos, m->b2_pos}; int32_t* bn[2] = {m->b_neg, m->b2_neg};
        float* bpp[2] = {m->b_pp, m->b2_pp}; float* bnp[2] = {m->b_np, m->b2_np};
        CK(cudaEventRecord(m->ev_start, st));                       // both slots are free once the caller's earlier work is done
        CK(cudaStreamWaitEvent(m->copy_st, m->ev_start, 0));
        auto slot_of = [&](int k) { return (n_steps - 1 - k) & 1; };
        int rc = do_sample(m, seed, epoch, step0, B, m->copy_st, slot_of(0));
        if (rc) return rc;
        CK(cudaEventRecord(m->ev_copied[slot_of(0)], m->copy_st));
        for (int32_t k = 0; k < n_steps; ++k) {
            const int sl = slot_of(k);
            CK(cudaStreamWaitEvent(st, m->ev_copied[sl], 0));
            rc = enqueue_fwd_bwd(m, bu[sl], bp[sl], bn[sl], bpp[sl], bnp[sl], B, m->batch_uniq, st, true);
            if (rc) return rc;
            if (k + 1 < n_steps) {
                const int sn = slot_of(k + 1);
                // the step kernel of k is done => so is all of step k-1, the last reader of slot sn
                CK(cudaEventRecord(m->ev_stepped[sl], st));
                CK(cudaStreamWaitEvent(m->copy_st, m->ev_stepped[sl], 0));
                rc = do_sample(m, seed, epoch, step0 + (uint32_t)(k + 1), B, m->copy_st, sn);
                if (rc) return rc;
                CK(cudaEventRecord(m->ev_copied[sn], m->copy_st));
            }
            rc = enqueue_adam(m, true, st);
            if (rc) return rc;
        }
        CK(cudaGetLastError());
        return PDA_OK;
    }
    for (int32_t k = 0; k < n_steps; ++k) {
        int rc = do_sample(m, seed, epoch, step0 + (uint32_t)k, B, st);
        if (rc) return rc;
        rc = enqueue_step(m, m->b_users, m->b_pos, m->b_neg, tmode ? nullptr : m->b_pp, tmode ? nullptr : m->b_np, B,
                          m->batch_uniq, true, st);
        if (rc) return rc;
    }
    CK(cudaGetLastError());
    return PDA_OK;
}

int pda_read_loss(pda_model* m, float* loss3_out, void* stream) {
    if (!m || !loss3_out) return fail(PDA_ERR_ARG, "null argument");
    CK(cudaSetDevice(m->cfg.device));
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaMemcpyAsync(m->loss3_pinned, m->loss3, 12, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    memcpy(loss3_out, m->loss3_pinned, 12);
    return PDA_OK;
}

int pda_read_loss_sums(pda_model* m, double* out4, int reset, void* stream) {
    if (!m || !out4) return fail(PDA_ERR_ARG, "null argument");
    CK(cudaSetDevice(m->cfg.device));
    cudaStream_t st = (cudaStream_t)stream;
    double* pin = (double*)((char*)m->loss3_pinned + 16);
    CK(cudaMemcpyAsync(pin, m->loss_sum, 32, cudaMemcpyDeviceToHost, st));
    if (reset) CK(cudaMemsetAsync(m->loss_sum, 0, 32, st));
    CK(cudaStreamSynchronize(st));
    memcpy(out4, pin, 32);
    return PDA_OK;
}

int pda_gradients_host(pda_model* m, const int32_t* users, const int32_t* pos, const int32_t* neg, const float* pp,
                       const float* np_, int64_t B, float* gU_out, float* gI_out, float* loss3_out) {
    if (!m) return fail(PDA_ERR_ARG, "null model");
    CK(cudaSetDevice(m->cfg.device));
    int rc = stage_batch(m, users, pos, neg, pp, np_, B, 0);
    if (rc) return rc;
    rc = check_batch(m, m->b_users, m->b_pos, m->b_neg, B, false, 0, 0);
    if (rc) return rc;
    CK(cudaStreamSynchronize(0));
    if (m->chk_pinned[1]) return fail(PDA_ERR_ARG, "batch holds a user / item id outside [0, n_users) / [0, n_items)");
    rc = enqueue_step(m, m->b_users, m->b_pos, m->b_neg, m->b_pp, m->b_np, B, 0, false, 0);
    if (rc) return rc;
    CK(cudaStreamSynchronize(0));
    CK(cudaGetLastError());
    if (gU_out) CK(cudaMemcpy(gU_out, m->G[0], (size_t)m->nU * m->d * 4, cudaMemcpyDeviceToHost));
    if (gI_out) CK(cudaMemcpy(gI_out, m->G[1], (size_t)m->nI * m->d * 4, cudaMemcpyDeviceToHost));
    if (loss3_out) CK(cudaMemcpy(loss3_out, m->loss3, 12, cudaMemcpyDeviceToHost));
    if (m->n_arr == 4 && m->grad_bias_out[0]) CK(cudaMemcpy(m->grad_bias_out[0], m->G[2], (size_t)m->nU * 4, cudaMemcpyDeviceToHost));
    if (m->n_arr == 4 && m->grad_bias_out[1])
        CK(cudaMemcpy(m->grad_bias_out[1], m->G[3], (size_t)m->nI * (m->cfg.temp_num + 1) * 4, cudaMemcpyDeviceToHost));
    for (int t = 0; t < m->n_arr; ++t) CK(cudaMemset(m->G[t], 0, (size_t)m->n4[t] * 16));
    return PDA_OK;
}

int pda_gradients_temp_host(pda_model* m, const int32_t* users, const int32_t* pos, const int32_t* neg, const float* temp,
                            int64_t B, float* gU_out, float* gI_out, float* gub_out, float* gib_out, float* loss3_out) {
    if (!m) return fail(PDA_ERR_ARG, "null model");
    if (m->cfg.train_mode != PDA_TRAIN_TEMP_POP) return fail(PDA_ERR_STATE, "model is not in train_mode temp_pop");
    m->grad_bias_out[0] = gub_out; m->grad_bias_out[1] = gib_out;
    int rc = pda_gradients_host(m, users, pos, neg, temp, temp, B, gU_out, gI_out, loss3_out);
    m->grad_bias_out[0] = m->grad_bias_out[1] = nullptr;
    return rc;
}

int pda_temp_item_bias_host(pda_model* m, int32_t first_user, float* out) {
    if (!m || !out) return fail(PDA_ERR_ARG, "null argument");
    if (m->cfg.train_mode != PDA_TRAIN_TEMP_POP) return fail(PDA_ERR_STATE, "model is not in train_mode temp_pop");
    if (first_user < 0 || first_user >= m->nU) return fail(PDA_ERR_ARG, "user id out of range");
    CK(cudaSetDevice(m->cfg.device));
    CK(ensure_dev(&m->ev_buf, &m->ev_bytes, (size_t)m->nI * 4));
    launch_temp_item_bias(m->W[2], m->W[3], m->nI, m->cfg.temp_num, first_user, (float*)m->ev_buf, 0);
    CK(cudaMemcpyAsync(out, m->ev_buf, (size_t)m->nI * 4, cudaMemcpyDeviceToHost, 0));
    CK(cudaStreamSynchronize(0));
    CK(cudaGetLastError());
    return PDA_OK;
}

// ---- recommendation ----
// PDA_EVAL_AUTO: the tcgen05 filter when the shape supports it (d in {64,128}, >= 4096 items; few rows are spread
// over item-range splits), else the exact CUDA-core kernel.  Both give identical ids and scores.
static int do_recommend(pda_model* m, const EvalArgs& a, int backend, cudaStream_t st) {
    flush_lazy(m, st);     // scoring reads both tables: rows that skipped Adam steps catch up first
    if (a.K < 1 || a.K > 128 || a.M < 1) return fail(PDA_ERR_ARG, "bad eval arguments (K in [1,128], M >= 1)");
    const bool tc_ok = tc_supported(a);
    if (backend == PDA_EVAL_TENSOR && !tc_ok)
        return fail(PDA_ERR_ARG, "tensor-core eval needs embed_size in {64,128}, n_items >= 4096 (got d=%d, n_items=%lld)", a.d,
                    (long long)a.N);
    const bool use_tc = backend == PDA_EVAL_TENSOR || (backend == PDA_EVAL_AUTO && tc_ok);
    if (!use_tc) {
        ProfScope ps(m, PDA_PROF_EVAL, st);
        if (launch_recommend_exact(a, st)) return fail(PDA_ERR_ARG, "bad eval arguments (K in [1,128], M >= 1)");
        return PDA_OK;
    }
    // user blocks bound the scratch (chunk maxima + candidate lists are per row)
    const int64_t MB = 32768;
    EvalArgs blk = a;
    blk.M = a.M < MB ? a.M : MB;
    TcPlan plan;
    const size_t need = tc_scratch_bytes(blk, &plan);
    CK(ensure_dev(&m->tc_buf, &m->tc_bytes, need));
    for (int64_t m0 = 0; m0 < a.M; m0 += MB) {
        blk = a;
        blk.users = a.users + m0;
        blk.M = a.M - m0 < MB ? a.M - m0 : MB;
        blk.ids_out = a.ids_out + m0 * a.K;
        blk.scores_out = a.scores_out ? a.scores_out + m0 * a.K : nullptr;
        tc_scratch_bytes(blk, &plan);
        m->tc_last_plan = plan; m->tc_last_M = blk.M;
        ProfScope ps(m, PDA_PROF_EVAL_TC, st);
        cudaEvent_t ev[4];
        bool sub = false;
        if (m->prof_on && m->prof_ev && m->prof_n[PDA_PROF_EVAL_SWEEP_A] < PDA_PROF_SLOTS && m->prof_n[PDA_PROF_EVAL_SWEEP_B] < PDA_PROF_SLOTS) {
            const int sa = m->prof_n[PDA_PROF_EVAL_SWEEP_A]++, sb = m->prof_n[PDA_PROF_EVAL_SWEEP_B]++;
            ev[0] = m->prof_ev[PDA_PROF_EVAL_SWEEP_A][sa][0]; ev[1] = m->prof_ev[PDA_PROF_EVAL_SWEEP_A][sa][1];
            ev[2] = m->prof_ev[PDA_PROF_EVAL_SWEEP_B][sb][0]; ev[3] = m->prof_ev[PDA_PROF_EVAL_SWEEP_B][sb][1];
            sub = true;
        }
        const int rc = launch_recommend_tc(blk, m->tc_buf, plan, m0 == 0, st, sub ? ev : nullptr);
        if (rc) return fail(PDA_ERR_CUDA, "tensor-core eval launch failed (stage %d): %s", rc, cudaGetErrorString(cudaGetLastError()));
    }
    return PDA_OK;
}

int pda_recommend_device(pda_model* m, const int32_t* users, int64_t M, int rec_type, const float* pop,
                         const float* col_bias, int use_mask, int K, int backend, int32_t* ids_out, float* scores_out,
                         void* stream) {
    if (!m || !users || !ids_out) return fail(PDA_ERR_ARG, "null argument");
    if (rec_type == PDA_REC_WITH_POP && !pop) return fail(PDA_ERR_ARG, "rec_type with_pop needs pop");
    if (use_mask && !m->indptr) return fail(PDA_ERR_STATE, "mask requested but pda_set_train_csr was not called");
    CK(cudaSetDevice(m->cfg.device));
    EvalArgs a;
    memset(&a, 0, sizeof(a));
    a.U = m->W[0]; a.I = m->W[1]; a.N = m->nI; a.d = m->d; a.users = users; a.M = M;
    a.mode = rec_type == PDA_REC_WITH_POP ? 1 : 0;
    a.pop = pop; a.col_bias = col_bias;
    a.mask_indptr = use_mask ? m->indptr : nullptr; a.mask_items = use_mask ? m->items : nullptr;
    a.K = K; a.ids_out = ids_out; a.scores_out = scores_out;
    int rc = do_recommend(m, a, backend, (cudaStream_t)stream);
    if (rc) return rc;
    CK(cudaGetLastError());
    return PDA_OK;
}

int pda_recommend_host(pda_model* m, const int32_t* users, int64_t M, int rec_type, const float* pop,
                       const float* col_bias, int use_mask, int K, int backend, int32_t* ids_out, float* scores_out) {
    if (!m || !users || !ids_out || M < 1 || K < 1) return fail(PDA_ERR_ARG, "bad argument");
    for (int64_t r = 0; r < M; ++r)
        if (users[r] < 0 || users[r] >= m->nU) return fail(PDA_ERR_ARG, "user id %d outside [0, n_users) at position %lld", users[r], (long long)r);
    CK(cudaSetDevice(m->cfg.device));
    const size_t nu = ((size_t)M * 4 + 255) / 256 * 256, nv = ((size_t)m->nI * 4 + 255) / 256 * 256;
    const size_t nk = ((size_t)M * K * 4 + 255) / 256 * 256;
    CK(ensure_dev(&m->ev_buf, &m->ev_bytes, nu + 2 * nv + 2 * nk));
    char* b = (char*)m->ev_buf;
    int32_t* d_users = (int32_t*)b; float* d_pop = (float*)(b + nu); float* d_bias = (float*)(b + nu + nv);
    int32_t* d_ids = (int32_t*)(b + nu + 2 * nv); float* d_sc = (float*)(b + nu + 2 * nv + nk);
    CK(cudaMemcpyAsync(d_users, users, (size_t)M * 4, cudaMemcpyHostToDevice, 0));
    if (pop) CK(cudaMemcpyAsync(d_pop, pop, (size_t)m->nI * 4, cudaMemcpyHostToDevice, 0));
    if (col_bias) CK(cudaMemcpyAsync(d_bias, col_bias, (size_t)m->nI * 4, cudaMemcpyHostToDevice, 0));
    int rc = pda_recommend_device(m, d_users, M, rec_type, pop ? d_pop : nullptr, col_bias ? d_bias : nullptr, use_mask, K,
                                  backend, d_ids, scores_out ? d_sc : nullptr, 0);
    if (rc) return rc;
    CK(cudaMemcpyAsync(ids_out, d_ids, (size_t)M * K * 4, cudaMemcpyDeviceToHost, 0));
    if (scores_out) CK(cudaMemcpyAsync(scores_out, d_sc, (size_t)M * K * 4, cudaMemcpyDeviceToHost, 0));
    CK(cudaStreamSynchronize(0));
    CK(cudaGetLastError());
    return PDA_OK;
}

int pda_tc_last_stats(pda_model* m, int64_t* out) {
    if (!m || !out) return fail(PDA_ERR_ARG, "null argument");
    if (!m->tc_buf || m->tc_last_M < 1) return fail(PDA_ERR_STATE, "the tensor-core eval path has not run yet");
    CK(cudaSetDevice(m->cfg.device));
    CK(cudaDeviceSynchronize());
    const TcPlan& p = m->tc_last_plan;
    const int64_t M = m->tc_last_M;
    int32_t nflag = 0;
    CK(cudaMemcpy(&nflag, (char*)m->tc_buf + p.o_nflag, 4, cudaMemcpyDeviceToHost));
    int32_t* cnt = (int32_t*)malloc((size_t)M * p.n_seg * 4);
    if (!cnt) return fail(PDA_ERR_STATE, "out of host memory");
    CK(cudaMemcpy(cnt, (char*)m->tc_buf + p.o_cnt, (size_t)M * p.n_seg * 4, cudaMemcpyDeviceToHost));
    int64_t tot = 0, mx = 0, over = 0;
    for (int64_t r = 0; r < M; ++r) {
        int64_t row_tot = 0; bool ov = false;
        for (int sg = 0; sg < p.n_seg; ++sg) {
            const int32_t c = cnt[r * p.n_seg + sg];
            row_tot += c < p.seg_cap ? c : p.seg_cap;
            if (c > p.seg_cap) ov = true;
        }
        tot += row_tot;
        if (row_tot > mx) mx = row_tot;
        if (ov) ++over;
    }
    free(cnt);
    out[0] = M; out[1] = nflag; out[2] = tot; out[3] = mx; out[4] = over; out[5] = p.se; out[6] = p.n_c; out[7] = p.splits;
    return PDA_OK;
}

// pure host arithmetic (no CUDA call): the scratch layout / launch plan of the tensor-core eval for a shape
int pda_tc_plan_host(int64_t M, int64_t N, int32_t d, int32_t K, int64_t* out) {
    if (!out || M < 1) return fail(PDA_ERR_ARG, "bad argument");
    EvalArgs a;
    memset(&a, 0, sizeof(a));
    a.M = M; a.N = N; a.d = d; a.K = K;
    if (!tc_supported(a)) return fail(PDA_ERR_ARG, "tensor-core eval needs embed_size in {64,128}, n_items >= 4096, K <= 128");
    TcPlan p;
    const size_t total = tc_scratch_bytes(a, &p);
    const int64_t v[24] = {p.M_pad, p.N_pad, p.n_tiles, p.mr, p.ts, p.ordered, p.n_sel, p.se, p.cw, p.n_c, p.n_valid, p.splits,
                           p.tiles_per_split, p.n_seg, p.seg_cap, p.rc, (int64_t)total, (int64_t)p.o_Ib, (int64_t)p.o_Ub,
                           (int64_t)p.o_cmax, (int64_t)p.o_cand, (int64_t)p.o_clist, (int64_t)p.o_work, (int64_t)p.o_nwork};
    memcpy(out, v, sizeof(v));
    return PDA_OK;
}

int pda_tc_debug_dense_host(pda_model* m, const int32_t* users, int64_t M, int rec_type, const float* pop,
                            const float* col_bias, float* out, float* err_coef) {
    if (!m || !users || !out || M < 1 || M > 32768) return fail(PDA_ERR_ARG, "bad argument");
    if (rec_type == PDA_REC_WITH_POP && !pop) return fail(PDA_ERR_ARG, "rec_type with_pop needs pop");
    CK(cudaSetDevice(m->cfg.device));
    EvalArgs a;
    memset(&a, 0, sizeof(a));
    a.U = m->W[0]; a.I = m->W[1]; a.N = m->nI; a.d = m->d; a.M = M; a.K = 1;
    a.mode = rec_type == PDA_REC_WITH_POP ? 1 : 0;
    if (!tc_supported(a)) return fail(PDA_ERR_ARG, "tensor-core eval needs embed_size in {64,128}, n_items >= 4096");
    TcPlan plan;
    const size_t need = tc_scratch_bytes(a, &plan);
    CK(ensure_dev(&m->tc_buf, &m->tc_bytes, need));
    const size_t nu = ((size_t)M * 4 + 255) / 256 * 256, nv = ((size_t)m->nI * 4 + 255) / 256 * 256;
    const size_t nd = (size_t)M * plan.N_pad * 4;
    CK(ensure_dev(&m->ev_buf, &m->ev_bytes, nu + 2 * nv + nd));
    char* b = (char*)m->ev_buf;
    int32_t* d_users = (int32_t*)b; float* d_pop = (float*)(b + nu); float* d_bias = (float*)(b + nu + nv);
    float* d_dense = (float*)(b + nu + 2 * nv);
    CK(cudaMemcpyAsync(d_users, users, (size_t)M * 4, cudaMemcpyHostToDevice, 0));
    if (pop) CK(cudaMemcpyAsync(d_pop, pop, (size_t)m->nI * 4, cudaMemcpyHostToDevice, 0));
    if (col_bias) CK(cudaMemcpyAsync(d_bias, col_bias, (size_t)m->nI * 4, cudaMemcpyHostToDevice, 0));
    a.users = d_users; a.pop = pop ? d_pop : nullptr; a.col_bias = col_bias ? d_bias : nullptr;
    flush_lazy(m, 0);
    const int rc = launch_tc_debug_dense(a, m->tc_buf, plan, d_dense, 0);
    if (rc) return fail(PDA_ERR_CUDA, "tensor-core sweep launch failed (stage %d): %s", rc, cudaGetErrorString(cudaGetLastError()));
    CK(cudaMemcpy2DAsync(out, (size_t)m->nI * 4, d_dense, (size_t)plan.N_pad * 4, (size_t)m->nI * 4, (size_t)M,
                         cudaMemcpyDeviceToHost, 0));
    CK(cudaStreamSynchronize(0));
    CK(cudaGetLastError());
    if (err_coef) {
        const float cA = 1.02f / 256.0f + (float)m->d / 2097152.0f, cB = (float)(m->d / 16 + 5) / 524288.0f;
        err_coef[0] = cA + cB; err_coef[1] = cB;
    }
    return PDA_OK;
}

int pda_scores_host(pda_model* m, const int32_t* users, int64_t M, int rec_type, const float* pop, float* out) {
    if (!m || !users || !out || M < 1) return fail(PDA_ERR_ARG, "bad argument");
    if (rec_type == PDA_REC_WITH_POP && !pop) return fail(PDA_ERR_ARG, "rec_type with_pop needs pop");
    CK(cudaSetDevice(m->cfg.device));
    const size_t nu = ((size_t)M * 4 + 255) / 256 * 256, nv = ((size_t)m->nI * 4 + 255) / 256 * 256;
    const size_t nd = (size_t)M * m->nI * 4;
    CK(ensure_dev(&m->ev_buf, &m->ev_bytes, nu + nv + 256 + nd));
    char* b = (char*)m->ev_buf;
    int32_t* d_users = (int32_t*)b; float* d_pop = (float*)(b + nu); int32_t* d_ids = (int32_t*)(b + nu + nv);
    float* d_dense = (float*)(b + nu + nv + 256);
    CK(cudaMemcpyAsync(d_users, users, (size_t)M * 4, cudaMemcpyHostToDevice, 0));
    if (pop) CK(cudaMemcpyAsync(d_pop, pop, (size_t)m->nI * 4, cudaMemcpyHostToDevice, 0));
    EvalArgs a;
    memset(&a, 0, sizeof(a));
    a.U = m->W[0]; a.I = m->W[1]; a.N = m->nI; a.d = m->d; a.users = d_users; a.M = M;
    a.mode = rec_type == PDA_REC_WITH_POP ? 1 : 0; a.pop = pop ? d_pop : nullptr;
    a.K = 1; a.ids_out = nullptr; a.scores_out = nullptr; a.dense_out = d_dense;
    flush_lazy(m, 0);
    (void)d_ids;
    if (launch_recommend_exact(a, 0)) return fail(PDA_ERR_ARG, "bad eval arguments");
    CK(cudaMemcpyAsync(out, d_dense, nd, cudaMemcpyDeviceToHost, 0));
    CK(cudaStreamSynchronize(0));
    CK(cudaGetLastError());
    return PDA_OK;
}

int pda_metrics_host(pda_model* m, const int32_t* ids, int64_t M, int Kkeep, const int32_t* eval_users,
                     const int64_t* truth_indptr, const int32_t* truth_items, int64_t n_truth_rows, const int32_t* Ks,
                     int nK, double* out) {
    if (!m || !ids || !eval_users || !truth_indptr || !Ks || !out || M < 1 || nK < 1 || nK > 16)
        return fail(PDA_ERR_ARG, "bad argument");
    for (int64_t r = 0; r < M; ++r)
        if (eval_users[r] < 0 || eval_users[r] >= n_truth_rows)
            return fail(PDA_ERR_ARG, "eval user id %d outside the truth CSR at position %lld", eval_users[r], (long long)r);
    CK(cudaSetDevice(m->cfg.device));
    const int64_t nnz = truth_indptr[n_truth_rows];
    auto al = [](size_t x) { return (x + 255) / 256 * 256; };
    const size_t o_ids = 0, o_users = o_ids + al((size_t)M * Kkeep * 4), o_ptr = o_users + al((size_t)M * 4);
    const size_t o_items = o_ptr + al(((size_t)n_truth_rows + 1) * 8), o_ks = o_items + al((size_t)(nnz > 0 ? nnz : 1) * 4);
    const size_t o_out = o_ks + al((size_t)nK * 4), total = o_out + al((size_t)4 * nK * 8);
    CK(ensure_dev(&m->ev_buf, &m->ev_bytes, total));
    char* b = (char*)m->ev_buf;
    CK(cudaMemcpyAsync(b + o_ids, ids, (size_t)M * Kkeep * 4, cudaMemcpyHostToDevice, 0));
    CK(cudaMemcpyAsync(b + o_users, eval_users, (size_t)M * 4, cudaMemcpyHostToDevice, 0));
    CK(cudaMemcpyAsync(b + o_ptr, truth_indptr, ((size_t)n_truth_rows + 1) * 8, cudaMemcpyHostToDevice, 0));
    if (nnz) CK(cudaMemcpyAsync(b + o_items, truth_items, (size_t)nnz * 4, cudaMemcpyHostToDevice, 0));
    CK(cudaMemcpyAsync(b + o_ks, Ks, (size_t)nK * 4, cudaMemcpyHostToDevice, 0));
    CK(cudaMemsetAsync(b + o_out, 0, (size_t)4 * nK * 8, 0));
    launch_metrics((const int32_t*)(b + o_ids), M, Kkeep, (const int32_t*)(b + o_users), (const int64_t*)(b + o_ptr),
                   (const int32_t*)(b + o_items), (const int32_t*)(b + o_ks), nK, (double*)(b + o_out), 0);
    CK(cudaMemcpyAsync(out, b + o_out, (size_t)4 * nK * 8, cudaMemcpyDeviceToHost, 0));
    CK(cudaStreamSynchronize(0));
    CK(cudaGetLastError());
    return PDA_OK;
}

}  // extern "C"
