/* CPU oracle (plain C + OpenMP) for the PDA hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * Same spec as oracle/pda_oracle.py (which documents the parity pinning); this file
 * exists so that full-size parity checks and the bench's cpu_baseline finish in seconds.
 * Build: oracle/Makefile  (gcc -O3 -mavx2 -ffp-contract=off -fopenmp; contraction must be
 * off: every fp32 op below is a separately rounded IEEE op, as in the CUDA kernels).
 *
 * Reference call sites restated (paths under /root/reference):
 *   sampler     MF/train_new_api.py:260-288, 366-412, 415-456
 *   step math   MF/model_api.py:51-53, 102-134 (PD / BPRMF), 336-371 (BPR(t)-pop)
 *   optimizer   MF/model_api.py:83,371,471 -> TF1.14 AdamOptimizer._apply_sparse_shared
 *   scoring     MF/train_new_api.py:594-612, MF/model_api.py:62,113
 *   metrics     MF/used_metric.py:39-80
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define TAG_INIT 0x1717AB01u
#define TAG_SAMPLE 0x5A4D9E02u
#define TAG_PERM 0x0FE15703u
#define TAG_USER 0x7C3B2A04u

typedef struct { uint32_t w[4]; } u32x4;

static inline u32x4 philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    u32x4 o = {{c0, c1, c2, c3}};
    return o;
}

void orc_philox(const uint32_t* ctr, const uint32_t* key, uint32_t* out) {
    u32x4 o = philox(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1]);
    memcpy(out, o.w, 16);
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* bench.py's reference arm: use every host core even when the launcher exported OMP_NUM_THREADS=1 (torchrun does) */
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

static inline uint32_t mulhi(uint32_t r, uint32_t n) { return (uint32_t)(((uint64_t)r * n) >> 32); }

/* ---- a1: Xavier init (MF/model_api.py:86-99) ---- */
void orc_xavier_init(float* W, int64_t rows, int cols, uint32_t seed, uint32_t table_id) {
    int64_t n = rows * cols, nq = (n + 3) / 4;
    float a = (float)sqrt(6.0 / (double)(rows + cols));
#pragma omp parallel for schedule(static)
    for (int64_t q = 0; q < nq; ++q) {
        u32x4 r = philox((uint32_t)q, (uint32_t)((uint64_t)q >> 32), table_id, 0, seed, TAG_INIT);
        for (int w = 0; w < 4; ++w) {
            int64_t e = 4 * q + w;
            if (e < n) {
                float u = (float)(r.w[w] >> 8) * 5.9604644775390625e-8f;
                W[e] = (u * 2.0f - 1.0f) * a;
            }
        }
    }
}

/* ---- a8: sampler ---- */
static inline uint32_t feistel_once(uint32_t x, int half, const uint32_t* keys) {
    uint32_t mask = (1u << half) - 1u, L = (x >> half) & mask, R = x & mask;
    for (int r = 0; r < 6; ++r) {
        uint32_t f = R * 0x9E3779B1u + keys[r];
        f ^= f >> 15; f *= 0x85EBCA6Bu; f ^= f >> 13; f *= 0xC2B2AE35u; f ^= f >> 16;
        uint32_t nR = L ^ (f & mask);
        L = R; R = nR;
    }
    return (L << half) | R;
}

static inline int half_bits(uint32_t n) {
    int bits = 0;
    uint32_t v = n - 1;
    while (v) { ++bits; v >>= 1; }
    if (bits < 2) bits = 2;
    return (bits + 1) / 2;
}

static inline int row_contains(const int32_t* items, int64_t lo, int64_t hi, int32_t c) {
    int64_t end = hi;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (items[mid] < c) lo = mid + 1; else hi = mid;
    }
    return lo < end && items[lo] == c;
}

/* pop_train: [n_items, T_pop] fp32 or NULL; T_pop==1 -> global popularity (PDG). */
void orc_sample_batch(uint32_t seed, uint32_t epoch, uint32_t step, int64_t B,
                      const int32_t* active_users, int64_t n_act, const int64_t* indptr,
                      const int32_t* items, const uint8_t* times, int32_t n_items,
                      const int32_t* unique_times, int32_t n_times, const float* pop_train,
                      int32_t T_pop, int32_t* users_out, int32_t* pos_out, int32_t* neg_out,
                      int32_t* time_out, float* pos_pop_out, float* neg_pop_out) {
    uint32_t keys[6];
    u32x4 ka = philox(0, 0, step, epoch, seed, TAG_PERM), kb = philox(1, 0, step, epoch, seed, TAG_PERM);
    keys[0] = ka.w[0]; keys[1] = ka.w[1]; keys[2] = ka.w[2]; keys[3] = ka.w[3];
    keys[4] = kb.w[0]; keys[5] = kb.w[1];
    int half = half_bits((uint32_t)n_act);
#pragma omp parallel for schedule(dynamic, 1024)
    for (int64_t i = 0; i < B; ++i) {
        int32_t u;
        if (B <= n_act) {
            uint32_t y = feistel_once((uint32_t)i, half, keys);
            while (y >= (uint32_t)n_act) y = feistel_once(y, half, keys);
            u = active_users[y];
        } else {
            u32x4 r = philox((uint32_t)i, 0, step, epoch, seed, TAG_USER);
            u = active_users[mulhi(r.w[0], (uint32_t)n_act)];
        }
        u32x4 w = philox((uint32_t)i, 0, step, epoch, seed, TAG_SAMPLE);
        int64_t lo = indptr[u], hi = indptr[u + 1];
        uint32_t deg = (uint32_t)(hi - lo);
        int32_t pos, t;
        if (deg > 0) {
            int64_t at = lo + mulhi(w.w[0], deg);
            pos = items[at]; t = times[at];
        } else {
            pos = 0; t = unique_times[mulhi(w.w[0], (uint32_t)n_times)];
        }
        uint32_t call = 0;
        int32_t neg;
        for (uint32_t a = 1;; ++a) {
            if ((a & 3u) == 0) { ++call; w = philox((uint32_t)i, call, step, epoch, seed, TAG_SAMPLE); }
            int32_t c = (int32_t)mulhi(w.w[a & 3u], (uint32_t)n_items);
            if (!row_contains(items, lo, hi, c)) { neg = c; break; }
        }
        users_out[i] = u; pos_out[i] = pos; neg_out[i] = neg;
        if (time_out) time_out[i] = t;
        if (pop_train) {
            int32_t tt = T_pop == 1 ? 0 : t;
            pos_pop_out[i] = pop_train[(int64_t)pos * T_pop + tt];
            neg_pop_out[i] = pop_train[(int64_t)neg * T_pop + tt];
        }
    }
}

/* ---- a2-a4: forward/backward ---- */
static void lanes_for_dim(int d, int* G, int* C) {
    int q = d / 4, g = 1;
    while (g < q && g < 32) g *= 2;
    *G = g; *C = (q + g - 1) / g;
}

static float dot_tree(const float* a, const float* b, int d, int G, int C) {
    float p[32];
    for (int l = 0; l < G; ++l) {
        float acc = 0.0f;
        for (int c = 0; c < C; ++c) {
            int base = 4 * (l + G * c);
            for (int k = 0; k < 4; ++k)
                if (base + k < d) { float pr = a[base + k] * b[base + k]; acc = acc + pr; }
        }
        p[l] = acc;
    }
    for (int off = G / 2; off >= 1; off >>= 1) {
        float q[32];
        for (int l = 0; l < G; ++l) q[l] = p[l] + p[l ^ off];
        memcpy(p, q, sizeof(float) * G);
    }
    return p[0];
}

/* exp() of the numerical spec: Cody-Waite reduction + degree-5 polynomial
 * (Cephes expf constants), every step ONE rounded fp32 mul or add (this file is compiled with
 * -ffp-contract=off), so the CUDA kernels (pda_common.cuh:spec_expf) and the numpy oracle
 * (pda_oracle.py:spec_expf) produce the same bits.  TF1's own exp is Eigen's pexp, equally a
 * polynomial approximation; tests/test_oracle_math.py bounds spec_expf vs libm exp by 2 ulp. */
static inline float bits_f(int32_t b) { float f; memcpy(&f, &b, 4); return f; }
float orc_spec_expf(float x) {
    if (x > 88.72283f) return INFINITY;
    if (x < -103.9f) return 0.0f;
    float t = x * 1.44269504088896341f;
    float n = nearbyintf(t);
    float a = n * 0.693359375f;
    float r = x - a;
    float b = n * 2.12194440e-4f;
    r = r + b;
    float p = 1.9875691500e-4f;
    p = p * r; p = p + 1.3981999507e-3f;
    p = p * r; p = p + 8.3334519073e-3f;
    p = p * r; p = p + 4.1665795894e-2f;
    p = p * r; p = p + 1.6666665459e-1f;
    p = p * r; p = p + 5.0000001201e-1f;
    float r2 = r * r;
    float y = p * r2; y = y + r; y = y + 1.0f;
    int ni = (int)n, n1 = ni >> 1, n2 = ni - n1;
    y = y * bits_f((n1 + 127) << 23);
    return y * bits_f((n2 + 127) << 23);
}
void orc_spec_expf_vec(const float* x, float* y, int64_t n) { for (int64_t i = 0; i < n; ++i) y[i] = orc_spec_expf(x[i]); }
#define spec_expf orc_spec_expf

static inline float elu_p1(float s) { return (s < 0.0f ? spec_expf(s) - 1.0f : s) + 1.0f; }
static inline float elu_p1_grad(float s) { return s < 0.0f ? (spec_expf(s) - 1.0f) + 1.0f : 1.0f; }

/* mode 0 = normal (BPRMF), 1 = pop (PD / PDG).  Writes per-triple row gradients
 * gU,gP,gN [B,d] and returns loss3 = {loss, mf, reg}. */
void orc_bpr_forward_backward(const float* U, const float* I, int d, const int32_t* users,
                              const int32_t* pos, const int32_t* neg, const float* pos_pop,
                              const float* neg_pop, int64_t B, float regs, int batch_size, int mode,
                              float* gU, float* gP, float* gN, float* loss3) {
    int G, C;
    lanes_for_dim(d, &G, &C);
    double mf_sum = 0.0, sq_sum = 0.0;
    float invB = 1.0f / (float)B, lb = (float)((double)regs / batch_size);
#pragma omp parallel for schedule(static) reduction(+ : mf_sum, sq_sum)
    for (int64_t i = 0; i < B; ++i) {
        const float *u = U + (int64_t)users[i] * d, *p = I + (int64_t)pos[i] * d, *n = I + (int64_t)neg[i] * d;
        float sp = dot_tree(u, p, d, G, C), sn = dot_tree(u, n, d, G, C);
        float x, dp, dn;
        if (mode == 0) { x = sp - sn; dp = 1.0f; dn = 1.0f; }
        else {
            x = elu_p1(sp) * pos_pop[i] - elu_p1(sn) * neg_pop[i];
            dp = elu_p1_grad(sp) * pos_pop[i]; dn = elu_p1_grad(sn) * neg_pop[i];
        }
        float sig = 1.0f / (1.0f + spec_expf(-x));
        mf_sum += (double)logf(sig + 1e-10f);
        float g = sig * (1.0f - sig) / (sig + 1e-10f);
        float cp = -g * dp * invB, cn = g * dn * invB;
        double sq = 0.0;
        for (int k = 0; k < d; ++k) {
            sq += (double)u[k] * u[k] + (double)p[k] * p[k] + (double)n[k] * n[k];
            if (gU) {
                gU[i * d + k] = cp * p[k] + cn * n[k] + lb * u[k];
                gP[i * d + k] = cp * u[k] + lb * p[k];
                gN[i * d + k] = cn * u[k] + lb * n[k];
            }
        }
        sq_sum += sq;
    }
    double mf = -mf_sum / (double)B, reg = (double)regs * 0.5 * sq_sum / batch_size;
    loss3[0] = (float)(mf + reg); loss3[1] = (float)mf; loss3[2] = (float)reg;
}

/* dedup-sum in occurrence order, parallel over row owners (row % T == tid keeps the order). */
void orc_scatter_add_rows(float* G, int d, const int32_t* idx, const float* rows, int64_t n) {
#pragma omp parallel
    {
#ifdef _OPENMP
        int tid = omp_get_thread_num(), nt = omp_get_num_threads();
#else
        int tid = 0, nt = 1;
#endif
        for (int64_t i = 0; i < n; ++i) {
            int32_t r = idx[i];
            if (r % nt != tid) continue;
            float* g = G + (int64_t)r * d;
            const float* s = rows + i * d;
            for (int k = 0; k < d; ++k) g[k] = g[k] + s[k];
        }
    }
}

/* the same sum in REVERSED occurrence order: a second, equally valid fp32 summation order (TF's unsorted_segment_sum does not
 * promise one) -- used only to measure the oracle's own noise band (SURVEY 7 "trajectory divergence" (iii)) */
void orc_scatter_add_rows_rev(float* G, int d, const int32_t* idx, const float* rows, int64_t n) {
#pragma omp parallel
    {
#ifdef _OPENMP
        int tid = omp_get_thread_num(), nt = omp_get_num_threads();
#else
        int tid = 0, nt = 1;
#endif
        for (int64_t i = n - 1; i >= 0; --i) {
            int32_t r = idx[i];
            if (r % nt != tid) continue;
            float* g = G + (int64_t)r * d;
            const float* s = rows + i * d;
            for (int k = 0; k < d; ++k) g[k] = g[k] + s[k];
        }
    }
}

/* ---- a6: TF1 Adam, dense sweep; G is consumed and zeroed ---- */
void orc_adam_dense(float* W, float* m, float* v, float* G, int64_t n, float lr_t, int zero_g) {
    const float b1 = 0.9f, b2 = 0.999f, eps = 1e-8f;
    const float omb1 = 1.0f - b1, omb2 = 1.0f - b2;
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < n; ++e) {
        float g = G[e];
        float mm = m[e] * b1; mm = mm + g * omb1;
        float vv = v[e] * b2; float g2 = g * g; vv = vv + g2 * omb2;
        float num = lr_t * mm, den = sqrtf(vv) + eps;
        W[e] = W[e] - num / den;
        m[e] = mm; v[e] = vv;
        if (zero_g) G[e] = 0.0f;
    }
}

float orc_lr_t(float lr, float b1p, float b2p) { float s = sqrtf(1.0f - b2p); float a = lr * s; return a / (1.0f - b1p); }

/* One full reference-semantics step on CPU: gather -> loss -> grads -> dedup -> dense Adam on
 * both tables.  scratch: gU,gP,gN [B,d]; GU [n_users,d], GI [n_items,d] zero on entry/exit.
 * pw = {beta1_power, beta2_power} (updated). */
void orc_train_step(float* U, float* mU, float* vU, float* GU, int64_t n_users, float* I, float* mI,
                    float* vI, float* GI, int64_t n_items, int d, const int32_t* users,
                    const int32_t* pos, const int32_t* neg, const float* pos_pop, const float* neg_pop,
                    int64_t B, float lr, float regs, int batch_size, int mode, float* gU, float* gP,
                    float* gN, float* pw, float* loss3) {
    orc_bpr_forward_backward(U, I, d, users, pos, neg, pos_pop, neg_pop, B, regs, batch_size, mode, gU,
                             gP, gN, loss3);
    orc_scatter_add_rows(GU, d, users, gU, B);
    orc_scatter_add_rows(GI, d, pos, gP, B);
    orc_scatter_add_rows(GI, d, neg, gN, B);
    float lr_t = orc_lr_t(lr, pw[0], pw[1]);
    orc_adam_dense(U, mU, vU, GU, n_users * d, lr_t, 1);
    orc_adam_dense(I, mI, vI, GI, n_items * d, lr_t, 1);
    pw[0] = pw[0] * 0.9f; pw[1] = pw[1] * 0.999f;
}

/* the same step with the duplicate rows summed in the reversed order (neg before pos, last occurrence first) */
void orc_train_step_rev(float* U, float* mU, float* vU, float* GU, int64_t n_users, float* I, float* mI,
                        float* vI, float* GI, int64_t n_items, int d, const int32_t* users,
                        const int32_t* pos, const int32_t* neg, const float* pos_pop, const float* neg_pop,
                        int64_t B, float lr, float regs, int batch_size, int mode, float* gU, float* gP,
                        float* gN, float* pw, float* loss3) {
    orc_bpr_forward_backward(U, I, d, users, pos, neg, pos_pop, neg_pop, B, regs, batch_size, mode, gU,
                             gP, gN, loss3);
    orc_scatter_add_rows_rev(GU, d, users, gU, B);
    orc_scatter_add_rows_rev(GI, d, neg, gN, B);
    orc_scatter_add_rows_rev(GI, d, pos, gP, B);
    float lr_t = orc_lr_t(lr, pw[0], pw[1]);
    orc_adam_dense(U, mU, vU, GU, n_users * d, lr_t, 1);
    orc_adam_dense(I, mI, vI, GI, n_items * d, lr_t, 1);
    pw[0] = pw[0] * 0.9f; pw[1] = pw[1] * 0.999f;
}

/* ---- a9: scoring + transform + mask + top-K ---- */
typedef struct { float y; int32_t id; } cand_t;
static inline int better(float ya, int32_t ia, float yb, int32_t ib) { return ya > yb || (ya == yb && ia < ib); }

/* mode 0: y = s (+ col_bias[j] if col_bias);  mode 1: y = (elu(s)+1) * pop[j].
 * mask CSR over global user ids (sorted rows not required).  IT = item table transposed [d, N].
 * ids_out [M,K] int32, scores_out [M,K] fp32 or NULL.  If fewer than K unmasked items exist the
 * tail is filled with masked items in ascending id (tf.nn.top_k on -inf ties). */
void orc_recommend(const float* U, const float* IT, int64_t N, int d, const int32_t* users, int64_t M,
                   int mode, const float* pop, const float* col_bias, const int64_t* mask_indptr,
                   const int32_t* mask_items, int K, int32_t* ids_out, float* scores_out) {
#pragma omp parallel
    {
        float* acc = (float*)malloc(sizeof(float) * N);
        cand_t* heap = (cand_t*)malloc(sizeof(cand_t) * (K + 1));
#pragma omp for schedule(dynamic, 4)
        for (int64_t r = 0; r < M; ++r) {
            const float* u = U + (int64_t)users[r] * d;
            for (int64_t j = 0; j < N; ++j) acc[j] = 0.0f;
            for (int k = 0; k < d; ++k) {
                float uk = u[k];
                const float* row = IT + (int64_t)k * N;
                for (int64_t j = 0; j < N; ++j) { float pr = uk * row[j]; acc[j] = acc[j] + pr; }
            }
            if (mode == 0) { if (col_bias) for (int64_t j = 0; j < N; ++j) acc[j] = acc[j] + col_bias[j]; }
            else for (int64_t j = 0; j < N; ++j) acc[j] = elu_p1(acc[j]) * pop[j];
            if (mask_indptr)
                for (int64_t q = mask_indptr[users[r]]; q < mask_indptr[users[r] + 1]; ++q)
                    acc[mask_items[q]] = -INFINITY;
            /* bounded insertion: keep the K best under (y desc, id asc); heap[0..cnt) sorted best-first */
            int cnt = 0;
            for (int64_t j = 0; j < N; ++j) {
                float y = acc[j];
                if (cnt == K && !better(y, (int32_t)j, heap[K - 1].y, heap[K - 1].id)) continue;
                int p = cnt < K ? cnt : K - 1;
                while (p > 0 && better(y, (int32_t)j, heap[p - 1].y, heap[p - 1].id)) { heap[p] = heap[p - 1]; --p; }
                heap[p].y = y; heap[p].id = (int32_t)j;
                if (cnt < K) ++cnt;
            }
            for (int k = 0; k < K; ++k) {
                ids_out[r * K + k] = k < cnt ? heap[k].id : -1;
                if (scores_out) scores_out[r * K + k] = k < cnt ? heap[k].y : -INFINITY;
            }
        }
        free(acc); free(heap);
    }
}

/* exact scores for explicit (row, item) pairs: s = sequential-k accumulate */
void orc_exact_scores_pairs(const float* U, const float* I, int d, const int32_t* users,
                            const int32_t* items, int64_t n, float* out) {
#pragma omp parallel for schedule(static)
    for (int64_t q = 0; q < n; ++q) {
        const float *u = U + (int64_t)users[q] * d, *v = I + (int64_t)items[q] * d;
        float acc = 0.0f;
        for (int k = 0; k < d; ++k) { float pr = u[k] * v[k]; acc = acc + pr; }
        out[q] = acc;
    }
}

/* ---- a11: metrics (MF/used_metric.py:69-80), summed over users (caller divides) ---- */
void orc_metrics_sum(const int32_t* ids, int64_t M, int Kkeep, const int32_t* eval_users,
                     const int64_t* truth_indptr, const int32_t* truth_items, const int32_t* Ks, int nK,
                     double* out /* [4, nK]: precision, recall, ndcg, hit */) {
    for (int i = 0; i < 4 * nK; ++i) out[i] = 0.0;
    for (int64_t r = 0; r < M; ++r) {
        int64_t lo = truth_indptr[eval_users[r]], hi = truth_indptr[eval_users[r] + 1];
        int64_t npos = hi - lo;
        for (int q = 0; q < nK; ++q) {
            int K = Ks[q] < Kkeep ? Ks[q] : Kkeep;
            double hits = 0.0, dcg = 0.0, idcg = 0.0;
            for (int k = 0; k < K; ++k) {
                int32_t id = ids[r * Kkeep + k];
                int hit = 0;
                for (int64_t z = lo; z < hi; ++z) if (truth_items[z] == id) { hit = 1; break; }
                double tp = 1.0 / log2((double)(k + 2));
                if (hit) { hits += 1.0; dcg += tp; }
                if (k < npos) idcg += tp;
            }
            out[0 * nK + q] += hits / K;
            out[1 * nK + q] += npos ? hits / (double)npos : 0.0;
            out[2 * nK + q] += idcg > 0 ? dcg / idcg : 0.0;
            out[3 * nK + q] += hits > 1.0 ? 1.0 : hits;
        }
    }
}
