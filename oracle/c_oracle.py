"""ctypes front-end of oracle/csrc/pda_oracle.c and oracle/_ref/libref_eval.so.

TEST INFRASTRUCTURE (see oracle/pda_oracle.py header): importable only from tests/,
__graft_entry__.smoke() and bench.py's CPU legs.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF = None

c_f = C.POINTER(C.c_float)
c_i32 = C.POINTER(C.c_int32)
c_i64 = C.POINTER(C.c_int64)
c_u8 = C.POINTER(C.c_uint8)
c_u32 = C.POINTER(C.c_uint32)
c_d = C.POINTER(C.c_double)


def build(force: bool = False) -> None:
    """Compile the C restatement (and oracle/_ref when /root/reference is present)."""
    so = os.path.join(_HERE, "_build", "libpda_oracle.so")
    src = os.path.join(_HERE, "csrc", "pda_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "_build/libpda_oracle.so"], stdout=subprocess.DEVNULL)
    if os.path.isdir("/root/reference/evaluator/backend/cpp/include"):
        ref = os.path.join(_HERE, "_ref", "libref_eval.so")
        if force or not os.path.exists(ref):
            subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)


def _p(a, t):
    if a is None:
        return None
    return a.ctypes.data_as(t)


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "_build", "libpda_oracle.so")
        if not os.path.exists(so):
            build()
        _LIB = C.CDLL(so)
        _LIB.orc_lr_t.restype = C.c_float
        _LIB.orc_lr_t.argtypes = [C.c_float, C.c_float, C.c_float]
        _LIB.orc_num_threads.restype = C.c_int
    return _LIB


def ref_lib():
    """The reference's own C++ evaluator (None when oracle/_ref was never built)."""
    global _REF
    if _REF is None:
        so = os.path.join(_HERE, "_ref", "libref_eval.so")
        if not os.path.exists(so):
            return None
        _REF = C.CDLL(so)
    return _REF


def spec_expf(x):
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.empty_like(x)
    lib().orc_spec_expf_vec(_p(x, c_f), _p(y, c_f), C.c_int64(x.size))
    return y


def num_threads() -> int:
    return int(lib().orc_num_threads())


def set_num_threads(n) -> int:
    """OpenMP team size of every later call (bench.py: all host cores, whatever OMP_NUM_THREADS the launcher exported)."""
    lib().orc_set_num_threads(C.c_int(int(n)))
    return num_threads()


def xavier_init(rows, cols, seed, table_id):
    W = np.empty((rows, cols), dtype=np.float32)
    lib().orc_xavier_init(_p(W, c_f), C.c_int64(rows), C.c_int(cols), C.c_uint32(seed), C.c_uint32(table_id))
    return W


def sample_batch(seed, epoch, step, B, active_users, indptr, items, times, n_items, unique_times,
                 pop_train=None):
    active_users = np.ascontiguousarray(active_users, dtype=np.int32)
    indptr = np.ascontiguousarray(indptr, dtype=np.int64)
    items = np.ascontiguousarray(items, dtype=np.int32)
    times = np.ascontiguousarray(times, dtype=np.uint8)
    ut = np.ascontiguousarray(unique_times, dtype=np.int32)
    out = {k: np.empty(B, dtype=np.int32) for k in ("users", "pos", "neg", "time")}
    T_pop = 0
    pp = npop = None
    if pop_train is not None:
        pop_train = np.ascontiguousarray(pop_train, dtype=np.float32)
        T_pop = 1 if pop_train.ndim == 1 else pop_train.shape[1]
        pp = np.empty(B, dtype=np.float32)
        npop = np.empty(B, dtype=np.float32)
        out["pos_pop"], out["neg_pop"] = pp, npop
    lib().orc_sample_batch(C.c_uint32(seed), C.c_uint32(epoch), C.c_uint32(step), C.c_int64(B),
                           _p(active_users, c_i32), C.c_int64(len(active_users)), _p(indptr, c_i64),
                           _p(items, c_i32), _p(times, c_u8), C.c_int32(n_items), _p(ut, c_i32),
                           C.c_int32(len(ut)), _p(pop_train, c_f), C.c_int32(T_pop),
                           _p(out["users"], c_i32), _p(out["pos"], c_i32), _p(out["neg"], c_i32),
                           _p(out["time"], c_i32), _p(pp, c_f), _p(npop, c_f))
    return out


class CModel:
    """Reference-semantics trainer on CPU (dense TF1 Adam), state in numpy arrays."""

    def __init__(self, U, I, lr, regs, batch_size, mode, copy=True, reversed_sum=False):
        # reversed_sum: duplicate rows summed in the reversed occurrence order -- a second valid fp32 order, for noise bands
        self._step = "orc_train_step_rev" if reversed_sum else "orc_train_step"
        self.U = np.array(U, dtype=np.float32, order="C") if copy else np.ascontiguousarray(U, dtype=np.float32)
        self.I = np.array(I, dtype=np.float32, order="C") if copy else np.ascontiguousarray(I, dtype=np.float32)
        self.d = self.U.shape[1]
        self.mU, self.vU, self.GU = (np.zeros_like(self.U) for _ in range(3))
        self.mI, self.vI, self.GI = (np.zeros_like(self.I) for _ in range(3))
        self.pw = np.array([0.9, 0.999], dtype=np.float32)
        self.lr, self.regs, self.batch_size = lr, regs, batch_size
        self.mode = 0 if mode == "normal" else 1
        self._scratch = None

    def train_step(self, users, pos, neg, pos_pop=None, neg_pop=None):
        B = len(users)
        if self._scratch is None or self._scratch[0].shape[0] != B:
            self._scratch = [np.empty((B, self.d), dtype=np.float32) for _ in range(3)]
        gU, gP, gN = self._scratch
        users = np.ascontiguousarray(users, dtype=np.int32)
        pos = np.ascontiguousarray(pos, dtype=np.int32)
        neg = np.ascontiguousarray(neg, dtype=np.int32)
        if pos_pop is not None:
            pos_pop = np.ascontiguousarray(pos_pop, dtype=np.float32)
            neg_pop = np.ascontiguousarray(neg_pop, dtype=np.float32)
        loss3 = np.zeros(3, dtype=np.float32)
        getattr(lib(), self._step)(_p(self.U, c_f), _p(self.mU, c_f), _p(self.vU, c_f), _p(self.GU, c_f),
                             C.c_int64(self.U.shape[0]), _p(self.I, c_f), _p(self.mI, c_f),
                             _p(self.vI, c_f), _p(self.GI, c_f), C.c_int64(self.I.shape[0]),
                             C.c_int(self.d), _p(users, c_i32), _p(pos, c_i32), _p(neg, c_i32),
                             _p(pos_pop, c_f), _p(neg_pop, c_f), C.c_int64(B), C.c_float(self.lr),
                             C.c_float(self.regs), C.c_int(self.batch_size), C.c_int(self.mode),
                             _p(gU, c_f), _p(gP, c_f), _p(gN, c_f), _p(self.pw, c_f), _p(loss3, c_f))
        return loss3


def forward_backward(U, I, users, pos, neg, regs, batch_size, mode, pos_pop=None, neg_pop=None):
    U = np.ascontiguousarray(U, dtype=np.float32)
    I = np.ascontiguousarray(I, dtype=np.float32)
    B, d = len(users), U.shape[1]
    gU, gP, gN = (np.empty((B, d), dtype=np.float32) for _ in range(3))
    loss3 = np.zeros(3, dtype=np.float32)
    users = np.ascontiguousarray(users, dtype=np.int32)
    pos = np.ascontiguousarray(pos, dtype=np.int32)
    neg = np.ascontiguousarray(neg, dtype=np.int32)
    if pos_pop is not None:
        pos_pop = np.ascontiguousarray(pos_pop, dtype=np.float32)
        neg_pop = np.ascontiguousarray(neg_pop, dtype=np.float32)
    lib().orc_bpr_forward_backward(_p(U, c_f), _p(I, c_f), C.c_int(d), _p(users, c_i32), _p(pos, c_i32),
                                   _p(neg, c_i32), _p(pos_pop, c_f), _p(neg_pop, c_f), C.c_int64(B),
                                   C.c_float(regs), C.c_int(batch_size), C.c_int(0 if mode == "normal" else 1),
                                   _p(gU, c_f), _p(gP, c_f), _p(gN, c_f), _p(loss3, c_f))
    return loss3, gU, gP, gN


def adam_dense(W, m, v, G, lr_t, zero_g=True):
    lib().orc_adam_dense(_p(W, c_f), _p(m, c_f), _p(v, c_f), _p(G, c_f), C.c_int64(W.size),
                         C.c_float(lr_t), C.c_int(1 if zero_g else 0))


def lr_t(lr, b1p, b2p) -> float:
    return float(lib().orc_lr_t(C.c_float(lr), C.c_float(b1p), C.c_float(b2p)))


def recommend(U, I, users, rec_type, K, mask_indptr=None, mask_items=None, pop=None, col_bias=None,
              IT=None):
    """ids [M,K] int32, scores [M,K] fp32 under the exact-score spec."""
    U = np.ascontiguousarray(U, dtype=np.float32)
    if IT is None:
        IT = np.ascontiguousarray(np.asarray(I, dtype=np.float32).T)
    d, N = IT.shape
    users = np.ascontiguousarray(users, dtype=np.int32)
    M = len(users)
    ids = np.empty((M, K), dtype=np.int32)
    sc = np.empty((M, K), dtype=np.float32)
    mode = 0 if rec_type == "main_branch" else 1
    if pop is not None:
        pop = np.ascontiguousarray(pop, dtype=np.float32)
    if col_bias is not None:
        col_bias = np.ascontiguousarray(col_bias, dtype=np.float32)
    if mask_indptr is not None:
        mask_indptr = np.ascontiguousarray(mask_indptr, dtype=np.int64)
        mask_items = np.ascontiguousarray(mask_items, dtype=np.int32)
    lib().orc_recommend(_p(U, c_f), _p(IT, c_f), C.c_int64(N), C.c_int(d), _p(users, c_i32), C.c_int64(M),
                        C.c_int(mode), _p(pop, c_f), _p(col_bias, c_f), _p(mask_indptr, c_i64),
                        _p(mask_items, c_i32), C.c_int(K), _p(ids, c_i32), _p(sc, c_f))
    return ids, sc


def exact_scores_pairs(U, I, users, items):
    U = np.ascontiguousarray(U, dtype=np.float32)
    I = np.ascontiguousarray(I, dtype=np.float32)
    users = np.ascontiguousarray(users, dtype=np.int32)
    items = np.ascontiguousarray(items, dtype=np.int32)
    out = np.empty(len(users), dtype=np.float32)
    lib().orc_exact_scores_pairs(_p(U, c_f), _p(I, c_f), C.c_int(U.shape[1]), _p(users, c_i32),
                                 _p(items, c_i32), C.c_int64(len(users)), _p(out, c_f))
    return out


def metrics_sum(ids, eval_users, truth_indptr, truth_items, Ks):
    ids = np.ascontiguousarray(ids, dtype=np.int32)
    eval_users = np.ascontiguousarray(eval_users, dtype=np.int32)
    truth_indptr = np.ascontiguousarray(truth_indptr, dtype=np.int64)
    truth_items = np.ascontiguousarray(truth_items, dtype=np.int32)
    Ks = np.ascontiguousarray(Ks, dtype=np.int32)
    out = np.zeros((4, len(Ks)), dtype=np.float64)
    lib().orc_metrics_sum(_p(ids, c_i32), C.c_int64(ids.shape[0]), C.c_int(ids.shape[1]),
                          _p(eval_users, c_i32), _p(truth_indptr, c_i64), _p(truth_items, c_i32),
                          _p(Ks, c_i32), C.c_int(len(Ks)), _p(out, c_d))
    return dict(precision=out[0], recall=out[1], ndcg=out[2], hit_ratio=out[3])


# ---- the reference's own native code (oracle/_ref) ----
def ref_evaluate_matrix(ratings, truth_indptr, truth_items, metric, top_k, threads=4):
    r = ref_lib()
    ratings = np.ascontiguousarray(ratings, dtype=np.float32)
    n_users, rating_len = ratings.shape
    truth_indptr = np.ascontiguousarray(truth_indptr, dtype=np.int64)
    truth_items = np.ascontiguousarray(truth_items, dtype=np.int32)
    metric = np.ascontiguousarray(metric, dtype=np.int32)
    res = np.zeros((n_users, len(metric) * top_k), dtype=np.float32)
    r.ref_evaluate_matrix(_p(ratings, c_f), C.c_int(rating_len), C.c_int(n_users), _p(truth_indptr, c_i64),
                          _p(truth_items, c_i32), _p(metric, c_i32), C.c_int(len(metric)), C.c_int(top_k),
                          C.c_int(threads), _p(res, c_f))
    return res


def ref_arg_top_k_2d(ratings, top_k, threads=4):
    r = ref_lib()
    ratings = np.ascontiguousarray(ratings, dtype=np.float32)
    rows, cols = ratings.shape
    out = np.zeros((rows, top_k), dtype=np.int32)
    r.ref_arg_top_k_2d(_p(ratings, c_f), C.c_int(cols), C.c_int(rows), C.c_int(top_k), C.c_int(threads),
                       _p(out, c_i32))
    return out
