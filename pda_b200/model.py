"""Host-side mirror of the reference's model / recommender objects over the C ABI.

Reference interfaces mirrored (paths under the upstream repo):
  MF/model_api.py:19-134 ConditionalBPRMF, :419-471 BPRMF          -> PDAModel(train=...)
  MF/train_new_api.py:538-696 DatasetApi_Model                      -> do_recommendation / testing / predict
  MF/train_new_api.py:1080-1090 sess.run([opt, loss, mf, reg])      -> train_step
  MF/train_new_api.py:260-412 generator_n_batch*                    -> sample_batch / train_sampled
Everything numeric happens in libpda_b200.so; numpy arrays only carry data across the boundary.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import PdaConfig, PdaError, check, ptr

TRAIN_MODES = {"normal": 0, "s_condition": 1, "condition": 1, "temp_pop": 2}
REC_TYPES = {"main_branch": 0, "main_with_pop": 1, "condition": 1}
BACKENDS = {"auto": 0, "exact": 1, "tensor": 2}
TOPK_MAX = 50  # Create_Recommendation(topk_max=50), train_new_api.py:594


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class PDAModel:
    """One model on one GPU: embedding tables + TF1-Adam state + train CSR, owned by the library."""

    def __init__(self, n_users, n_items, embed_size=64, train="s_condition", batch_size=2048, lr=1e-3, regs=1e-5,
                 device=0, max_batch=0, seed=2021, init=True, temp_num=0):
        if train not in TRAIN_MODES:
            raise NotImplementedError("not implement this model: " + str(train))
        self.lib = _lib.load()
        self.n_users, self.n_items, self.emb_dim = int(n_users), int(n_items), int(embed_size)
        self.train, self.batch_size, self.lr, self.regs = train, int(batch_size), float(lr), float(regs)
        self.device = int(device)
        self.temp_num = int(temp_num)
        cfg = PdaConfig(self.device, self.n_users, self.n_items, self.emb_dim, TRAIN_MODES[train], self.batch_size,
                        self.lr, self.regs, int(max_batch), self.temp_num)
        h = C.c_void_p()
        check(self.lib.pda_create(C.byref(cfg), C.byref(h)))
        self._h = h
        self.max_batch = int(max_batch) if max_batch else self.batch_size
        self.testing_model_type = "o"
        self.testing_popularity = None
        self._has_csr = False
        if init:
            self.init_tables(seed)

    # ---- life cycle ----
    def close(self):
        if getattr(self, "_h", None):
            self.lib.pda_destroy(self._h)
            self._h = None
        for p in getattr(self, "_pinned", []):
            self.lib.pda_host_free(p)
        self._pinned = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def init_tables(self, seed=2021):
        check(self.lib.pda_init_tables(self._h, seed))

    ADAM_MODES = {"dense": 0, "lazy": 1, "lazy_users": 2}

    def set_adam_mode(self, mode):
        """'dense': TF1's every-row sweep as such; 'lazy' (default): bit-identical exact replay without the sweep;
        'lazy_users': lazy user table + dense item table (data-parallel item-gradient all-reduce)."""
        check(self.lib.pda_set_adam_mode(self._h, self.ADAM_MODES[mode]))

    def set_deterministic(self, on=True):
        """duplicate item rows of a batch summed in occurrence order instead of with fp32 atomics: trajectories
        bit-identical to the CPU oracle (about twice the step cost)"""
        check(self.lib.pda_set_deterministic(self._h, 1 if on else 0))

    def set_hot_items(self, ids):
        """popular items whose positive-item gradient rows the d = 128 step kernel pre-sums per thread block in shared
        memory (at most 28; a performance hint -- set_train_csr* installs the most frequent items of the CSR)"""
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        check(self.lib.pda_set_hot_items(self._h, ptr(ids) if ids.size else None, int(ids.size)))

    def adam_stats(self, reset=True):
        """(rows updated with a gradient, zero-gradient row-steps replayed) by the lazy Adam kernels since the last reset."""
        out = np.zeros(2, dtype=np.int64)
        check(self.lib.pda_adam_stats(self._h, ptr(out), 1 if reset else 0))
        return int(out[0]), int(out[1])

    def synchronize(self):
        check(self.lib.pda_synchronize(self._h))

    PROF_KINDS = ("sampler", "bpr_step", "adam", "eval_exact", "eval_tensor", "adam_catchup", "eval_sweep_a", "eval_sweep_b")

    def profile(self, on=True):
        check(self.lib.pda_profile_enable(self._h, 1 if on else 0))

    def profile_read(self):
        """{kind: (total_ms, launches)} since the last read (CUDA events on the launching stream)."""
        ms = np.zeros(len(self.PROF_KINDS), dtype=np.float64)
        cnt = np.zeros(len(self.PROF_KINDS), dtype=np.int32)
        check(self.lib.pda_profile_read(self._h, ptr(ms), ptr(cnt)))
        return {k: (float(ms[i]), int(cnt[i])) for i, k in enumerate(self.PROF_KINDS)}

    # ---- state access (checkpoint interop: parameter/user_embedding, parameter/item_embedding) ----
    _TABLES = {"user_embedding": 0, "item_embedding": 1, "user_m": 2, "user_v": 3, "item_m": 4, "item_v": 5,
               "user_temp_bias": 6, "item_temp_bias": 7, "user_temp_bias_m": 8, "user_temp_bias_v": 9,
               "item_temp_bias_m": 10, "item_temp_bias_v": 11}

    def _shape(self, name):
        k = self._TABLES[name]
        if k in (0, 2, 3):
            return (self.n_users, self.emb_dim)
        if k in (1, 4, 5):
            return (self.n_items, self.emb_dim)
        if self.train != "temp_pop":
            raise KeyError(name + " exists only for --train temp_pop")
        return (self.n_users, 1) if k in (6, 8, 9) else (self.n_items, self.temp_num + 1)

    def _rows(self, name):
        return self._shape(name)[0]

    def get_table(self, name):
        out = np.empty(self._shape(name), dtype=np.float32)
        check(self.lib.pda_get_table(self._h, self._TABLES[name], ptr(out)))
        return out

    def set_table(self, name, value):
        v = _f32(value)
        if v.shape != self._shape(name):
            raise ValueError(f"{name}: expected {self._shape(name)}, got {v.shape}")
        check(self.lib.pda_set_table(self._h, self._TABLES[name], ptr(v)))

    def table_ptr(self, name) -> int:
        return int(self.lib.pda_table_ptr(self._h, self._TABLES[name]))

    def get_adam_powers(self):
        out = np.empty(2, dtype=np.float32)
        check(self.lib.pda_get_adam_powers(self._h, ptr(out)))
        return out

    def set_adam_powers(self, b1p, b2p):
        v = np.array([b1p, b2p], dtype=np.float32)
        check(self.lib.pda_set_adam_powers(self._h, ptr(v)))

    def _var_names(self):
        # TF variable names of the reference (model_api.py:91-92, 397-400: the item bias variable is *named* item_temp_bias)
        v = ["user_embedding", "item_embedding"]
        return v + ["user_temp_bias", "item_temp_bias"] if self.train == "temp_pop" else v

    def _slot_names(self):
        s = ["user_m", "user_v", "item_m", "item_v"]
        if self.train == "temp_pop":
            s += ["user_temp_bias_m", "user_temp_bias_v", "item_temp_bias_m", "item_temp_bias_v"]
        return s

    def state_dict(self):
        d = {"parameter/" + k: self.get_table(k) for k in self._var_names()}
        for k in self._slot_names():
            d["adam/" + k] = self.get_table(k)
        d["adam/beta_powers"] = self.get_adam_powers()
        return d

    def load_state_dict(self, d):
        for k in self._var_names():
            self.set_table(k, d["parameter/" + k])
        for k in self._slot_names():
            if "adam/" + k in d:
                self.set_table(k, d["adam/" + k])
        if "adam/beta_powers" in d:
            self.set_adam_powers(*d["adam/beta_powers"])

    # ---- data ----
    def set_train_csr(self, indptr, items, times=None, unique_times=None):
        indptr = np.ascontiguousarray(indptr, dtype=np.int64)
        items = _i32(items)
        t = None if times is None else np.ascontiguousarray(times, dtype=np.uint8)
        ut = None if unique_times is None else _i32(unique_times)
        if len(indptr) != self.n_users + 1:
            raise ValueError("indptr must have n_users + 1 entries")
        check(self.lib.pda_set_train_csr(self._h, ptr(indptr), ptr(items), ptr(t), len(items), ptr(ut),
                                         0 if ut is None else len(ut)))
        self._has_csr = True

    def set_train_csr_device(self, indptr_ptr, items_ptr, times_ptr, nnz, active_ptr, n_act, unique_times=None):
        ut = None if unique_times is None else _i32(unique_times)
        check(self.lib.pda_set_train_csr_device(self._h, ptr(indptr_ptr), ptr(items_ptr), ptr(times_ptr), int(nnz),
                                                ptr(active_ptr), int(n_act), ptr(ut), 0 if ut is None else len(ut)))
        self._has_csr = True

    def set_train_pop(self, pop):
        """pop: fp32 [n_items, T] = (pop[:, :-1]) ** gamma, or [n_items] for global popularity (PDG)."""
        p = _f32(pop)
        T = 1 if p.ndim == 1 else p.shape[1]
        if p.shape[0] != self.n_items:
            raise ValueError("popularity table must have n_items rows")
        check(self.lib.pda_set_train_pop(self._h, ptr(p), T))

    # ---- sampler ----
    def sample_batch(self, seed, epoch, step, B=None, stream=0, fetch=True):
        B = self.batch_size if B is None else int(B)
        check(self.lib.pda_sample_batch(self._h, seed, epoch, step, B, ptr(stream) if stream else None))
        if not fetch:
            return None
        out = {k: np.empty(B, dtype=np.int32) for k in ("users", "pos", "neg", "time")}
        out["pos_pop"] = np.empty(B, dtype=np.float32)
        out["neg_pop"] = np.empty(B, dtype=np.float32)
        check(self.lib.pda_get_batch(self._h, B, ptr(out["users"]), ptr(out["pos"]), ptr(out["neg"]), ptr(out["time"]),
                                     ptr(out["pos_pop"]), ptr(out["neg_pop"])))
        return out

    # ---- training ----
    def train_step(self, users, pos_items, neg_items, pos_pop=None, neg_pop=None):
        """One optimisation step on a host batch; returns (loss, mf_loss, reg_loss) like
        sess.run([opt, loss, mf_loss, reg_loss])[1:].  For --train temp_pop the fourth array is the stage `temp`
        of each triple (the slot the reference's iterator uses for it) and the fifth (`raw`) is optional."""
        u, p, n = _i32(users), _i32(pos_items), _i32(neg_items)
        pp = None if pos_pop is None else _f32(pos_pop)
        npop = None if neg_pop is None else _f32(neg_pop)
        out = np.zeros(3, dtype=np.float32)
        check(self.lib.pda_train_step_host(self._h, ptr(u), ptr(p), ptr(n), ptr(pp), ptr(npop), len(u), ptr(out)))
        return float(out[0]), float(out[1]), float(out[2])

    def train_step_device(self, users_ptr=None, pos_ptr=None, neg_ptr=None, pos_pop_ptr=None, neg_pop_ptr=None, B=None,
                          stream=0):
        B = self.batch_size if B is None else int(B)
        check(self.lib.pda_train_step_device(self._h, ptr(users_ptr), ptr(pos_ptr), ptr(neg_ptr), ptr(pos_pop_ptr),
                                             ptr(neg_pop_ptr), B, ptr(stream) if stream else None))

    def train_sampled(self, seed, epoch, step0, n_steps, B=None, stream=0):
        B = self.batch_size if B is None else int(B)
        check(self.lib.pda_train_steps_sampled(self._h, seed, epoch, step0, n_steps, B, ptr(stream) if stream else None))

    # ---- the step in two halves (data-parallel callers reduce gradients in between) ----
    def set_global_batch(self, Bg):
        check(self.lib.pda_set_global_batch(self._h, int(Bg)))

    def train_steps(self, users, pos_items, neg_items, pos_pop=None, neg_pop=None):
        """n consecutive train steps from PINNED host arrays of shape [n, B] (allocate them with pinned_array): the
        host->device copies of batch k+1 overlap step k.  Returns the [n, 3] losses {loss, mf_loss, reg_loss}; the
        result equals n calls of train_step."""
        u = np.ascontiguousarray(users, dtype=np.int32)
        n, B = u.shape
        p, ng = np.ascontiguousarray(pos_items, dtype=np.int32), np.ascontiguousarray(neg_items, dtype=np.int32)
        pp = None if pos_pop is None else np.ascontiguousarray(pos_pop, dtype=np.float32)
        npop = None if neg_pop is None else np.ascontiguousarray(neg_pop, dtype=np.float32)
        out = np.empty((n, 3), dtype=np.float32)
        check(self.lib.pda_train_steps_host(self._h, ptr(u), ptr(p), ptr(ng), ptr(pp), ptr(npop), n, B, ptr(out)))
        return out

    def pinned_array(self, shape, dtype):
        """numpy view of page-locked host memory (pda_host_alloc); freed with the model"""
        import ctypes as C
        dt = np.dtype(dtype)
        nbytes = int(np.prod(shape)) * dt.itemsize
        p = self.lib.pda_host_alloc(nbytes)
        if not p:
            raise PdaError("pda_host_alloc failed")
        self._pinned = getattr(self, "_pinned", [])
        self._pinned.append(p)
        return np.frombuffer((C.c_byte * nbytes).from_address(p), dtype=dt).reshape(shape)

    def grad_ptr(self, name) -> int:
        return int(self.lib.pda_grad_ptr(self._h, self._TABLES[name]))

    def loss_acc_ptr(self) -> int:
        return int(self.lib.pda_loss_acc_ptr(self._h))

    def forward_backward_device(self, B=None, stream=0):
        B = self.batch_size if B is None else int(B)
        check(self.lib.pda_forward_backward_device(self._h, None, None, None, None, None, B, ptr(stream) if stream else None))

    def adam_apply(self, stream=0, part=3):
        """part 1: lazily kept (rank-local) tables; part 2: dense sweep of the rest + bookkeeping; 3: both."""
        check(self.lib.pda_adam_apply_part(self._h, int(part), ptr(stream) if stream else None))

    def adam_dense_rows(self, name, row_lo, row_hi, stream=0):
        """dense Adam sweep of rows [row_lo, row_hi) of one densely kept table (pipelined gradient exchange)."""
        check(self.lib.pda_adam_dense_rows(self._h, self._TABLES[name], int(row_lo), int(row_hi),
                                           ptr(stream) if stream else None))

    def adam_dense_rows_ext(self, name, row_lo, row_hi, grad_ptr, stream=0):
        """the same sweep with the rows' gradient read from a caller-owned device buffer [row_hi - row_lo, d]
        (out-of-place reduce-scatter output); the model's own accumulator is not touched."""
        check(self.lib.pda_adam_dense_rows_ext(self._h, self._TABLES[name], int(row_lo), int(row_hi), ptr(int(grad_ptr)),
                                               ptr(stream) if stream else None))

    def adopt_item_buffers(self, W_ptr, G_ptr):
        """item table + item-gradient accumulator move into caller-owned device memory (symmetric / multicast-mapped)"""
        check(self.lib.pda_adopt_item_buffers(self._h, ptr(int(W_ptr)), ptr(int(G_ptr))))

    def dp_set_barrier(self, flags_local, peer_flags, rank):
        """in-kernel cross-rank barriers of the fused exchange (flags in symmetric memory); flags_local = 0 switches them off"""
        if not flags_local:
            check(self.lib.pda_dp_set_barrier(self._h, None, None, 0, 0))
            return
        w = len(peer_flags)
        arr = (C.c_void_p * w)(*[int(x) for x in peer_flags])
        check(self.lib.pda_dp_set_barrier(self._h, ptr(int(flags_local)), C.cast(arr, C.c_void_p), w, int(rank)))

    def set_item_grad_buffer(self, G_ptr):
        check(self.lib.pda_set_item_grad_buffer(self._h, ptr(int(G_ptr))))

    def dp_exchange_adam(self, mc_G, mc_W, row_lo, row_hi, stream=0):
        """reduce-scatter + sliced Adam + all-gather of the item table in one NVLink-multicast kernel (pda_exchange.cu)"""
        check(self.lib.pda_dp_exchange_adam(self._h, ptr(int(mc_G)), ptr(int(mc_W)), int(row_lo), int(row_hi),
                                            ptr(stream) if stream else None))

    def dp_exchange_adam_p2p(self, peer_G, peer_W, self_rank, row_lo, row_hi, stream=0):
        """the fused exchange over unicast peer pointers (lists of device addresses, one per rank, self included)"""
        w = len(peer_G)
        g = (C.c_void_p * w)(*[int(x) for x in peer_G])
        ww = (C.c_void_p * w)(*[int(x) for x in peer_W])
        check(self.lib.pda_dp_exchange_adam_p2p(self._h, C.cast(g, C.c_void_p), C.cast(ww, C.c_void_p), w, int(self_rank), int(row_lo),
                                                int(row_hi), ptr(stream) if stream else None))

    def stage_batch_async(self, users, pos_items, neg_items, pos_pop=None, neg_pop=None, copy_stream=0):
        """enqueue the host->device copies (+ the id / distinct-users check) of a PINNED host batch on `copy_stream` and
        return; staged_batch_wait() completes it.  The arrays must stay alive until then."""
        u, p, n = _i32(users), _i32(pos_items), _i32(neg_items)
        pp = None if pos_pop is None else _f32(pos_pop)
        npop = None if neg_pop is None else _f32(neg_pop)
        self._staged_keepalive = (u, p, n, pp, npop)
        check(self.lib.pda_stage_batch_host_async(self._h, ptr(u), ptr(p), ptr(n), ptr(pp), ptr(npop), len(u),
                                                  ptr(copy_stream) if copy_stream else None))
        return len(u)

    def staged_batch_wait(self, stream=0):
        check(self.lib.pda_staged_batch_wait(self._h, ptr(stream) if stream else None))
        self._staged_keepalive = None

    def read_loss_async(self, pinned_dst, stream=0):
        """loss3 of the last enqueued step -> a pinned fp32[3] view (pinned_array), no synchronisation"""
        check(self.lib.pda_read_loss_async(self._h, ptr(pinned_dst), ptr(stream) if stream else None))

    def stage_batch(self, users, pos_items, neg_items, pos_pop=None, neg_pop=None, stream=0):
        u, p, n = _i32(users), _i32(pos_items), _i32(neg_items)
        pp = None if pos_pop is None else _f32(pos_pop)
        npop = None if neg_pop is None else _f32(neg_pop)
        check(self.lib.pda_stage_batch_host(self._h, ptr(u), ptr(p), ptr(n), ptr(pp), ptr(npop), len(u),
                                            ptr(stream) if stream else None))
        return len(u)

    def read_loss(self, stream=0):
        out = np.zeros(3, dtype=np.float32)
        check(self.lib.pda_read_loss(self._h, ptr(out), ptr(stream) if stream else None))
        return float(out[0]), float(out[1]), float(out[2])

    def read_loss_sums(self, reset=True, stream=0):
        """(sum loss, sum mf, sum reg, n_steps) accumulated on the device since the last reset."""
        out = np.zeros(4, dtype=np.float64)
        check(self.lib.pda_read_loss_sums(self._h, ptr(out), 1 if reset else 0, ptr(stream) if stream else None))
        return out

    def gradients(self, users, pos_items, neg_items, pos_pop=None, neg_pop=None):
        u, p, n = _i32(users), _i32(pos_items), _i32(neg_items)
        pp = None if pos_pop is None else _f32(pos_pop)
        npop = None if neg_pop is None else _f32(neg_pop)
        gU = np.empty((self.n_users, self.emb_dim), dtype=np.float32)
        gI = np.empty((self.n_items, self.emb_dim), dtype=np.float32)
        loss3 = np.zeros(3, dtype=np.float32)
        check(self.lib.pda_gradients_host(self._h, ptr(u), ptr(p), ptr(n), ptr(pp), ptr(npop), len(u), ptr(gU), ptr(gI),
                                          ptr(loss3)))
        return gU, gI, loss3

    def gradients_temp(self, users, pos_items, neg_items, temp):
        """BPR(t)-pop test hook: (gU, gI, g_user_temp_bias [n_users], g_item_temp_bias [n_items, T+1], loss3)."""
        u, p, n, t = _i32(users), _i32(pos_items), _i32(neg_items), _f32(temp)
        gU = np.empty((self.n_users, self.emb_dim), dtype=np.float32)
        gI = np.empty((self.n_items, self.emb_dim), dtype=np.float32)
        gub = np.empty(self.n_users, dtype=np.float32)
        gib = np.empty((self.n_items, self.temp_num + 1), dtype=np.float32)
        loss3 = np.zeros(3, dtype=np.float32)
        check(self.lib.pda_gradients_temp_host(self._h, ptr(u), ptr(p), ptr(n), ptr(t), len(u), ptr(gU), ptr(gI), ptr(gub),
                                               ptr(gib), ptr(loss3)))
        return gU, gI, gub, gib, loss3

    def temp_item_bias_for_eval(self, first_user):
        """model_api.py:373-387: (1 + user_temp_bias[first user of the batch]) * (item_bias[:, T-1] + item_bias[:, T])."""
        out = np.empty(self.n_items, dtype=np.float32)
        check(self.lib.pda_temp_item_bias_host(self._h, int(first_user), ptr(out)))
        return out

    # ---- inference (DatasetApi_Model) ----
    def do_recommendation(self, batch_users, items=None, rec_type="main_branch", pos_pop=None, sparse_cliked_matrix=None,
                          K=TOPK_MAX, mask=True, col_bias=None, backend="auto", return_scores=False):
        """train_new_api.py:614-640.  `items` must be None / range(n_items) (the reference always passes all
        items); `sparse_cliked_matrix` is accepted for signature compatibility -- the mask is the train CSR
        already resident on the device (it holds exactly the (row, train item) pairs of :730-733)."""
        if rec_type not in REC_TYPES:
            raise NotImplementedError("we have only implement recommendation method: main main+pop condition")
        if items is not None and len(items) != self.n_items:
            raise NotImplementedError("only all-item recommendation is implemented (as used by the reference)")
        u = _i32(batch_users)
        pop = None
        if REC_TYPES[rec_type] == 1:
            if pos_pop is None:
                raise ValueError("rec_type %s needs pos_pop" % rec_type)
            pop = _f32(np.asarray(pos_pop).reshape(-1))
            if len(pop) != self.n_items:
                raise ValueError("pos_pop must have n_items entries")
        cb = None if col_bias is None else _f32(col_bias)
        ids = np.empty((len(u), K), dtype=np.int32)
        sc = np.empty((len(u), K), dtype=np.float32) if return_scores else None
        check(self.lib.pda_recommend_host(self._h, ptr(u), len(u), REC_TYPES[rec_type], ptr(pop), ptr(cb),
                                          1 if (mask and self._has_csr) else 0, K, BACKENDS[backend], ptr(ids), ptr(sc)))
        return (ids, sc) if return_scores else ids

    def tc_last_stats(self):
        out = np.zeros(8, dtype=np.int64)
        check(self.lib.pda_tc_last_stats(self._h, ptr(out)))
        return dict(zip(("rows", "rows_exact_fallback", "candidates", "max_candidates_row", "rows_overflow", "tile_stride",
                         "sampled_chunks", "item_splits"), (int(x) for x in out)))

    def tc_debug_dense(self, batch_users, rec_type="main_branch", pos_pop=None, col_bias=None):
        """Diagnostics: the raw tensor-core accumulators the filter compares (fp32 [len(users), n_items]) and the
        coefficients (cAB, cB) of its error bound -- see pda_tc_debug_dense_host in include/pda_b200.h."""
        u = _i32(batch_users)
        pop = None if pos_pop is None else _f32(np.asarray(pos_pop).reshape(-1))
        cb = None if col_bias is None else _f32(col_bias)
        out = np.empty((len(u), self.n_items), dtype=np.float32)
        coef = np.zeros(2, dtype=np.float32)
        check(self.lib.pda_tc_debug_dense_host(self._h, ptr(u), len(u), REC_TYPES[rec_type], ptr(pop), ptr(cb), ptr(out),
                                               ptr(coef)))
        return out, (float(coef[0]), float(coef[1]))

    def testing(self, batch_users, items=None, model_type="main_branch", pos_pop=None):
        """train_new_api.py:642-669: dense fp32 [len(batch_users), n_items] ratings (no mask)."""
        if model_type not in ("main_branch", "condition"):
            raise NotImplementedError("error -- not implement this type testing method...")
        u = _i32(batch_users)
        pop = None if model_type == "main_branch" else _f32(np.asarray(pos_pop).reshape(-1))
        out = np.empty((len(u), self.n_items), dtype=np.float32)
        check(self.lib.pda_scores_host(self._h, ptr(u), len(u), REC_TYPES[model_type], ptr(pop), ptr(out)))
        return out

    def set_testing_way(self, model_type, popularity_exp):
        self.testing_model_type = model_type
        self.testing_popularity = popularity_exp

    def predict(self, user_batch, item_batch=None):
        """NeuRec evaluator protocol (train_new_api.py:683-696)."""
        mt = self.testing_model_type
        if mt == "o":
            return self.testing(user_batch, None, "main_branch")
        if mt == "condition":
            return self.testing(user_batch, None, "condition", pos_pop=self.testing_popularity)
        raise NotImplementedError("not implement this type testing methods")

    def metrics_sum(self, ids, eval_users, truth_indptr, truth_items, Ks):
        ids = _i32(ids)
        eu = _i32(eval_users)
        ti = np.ascontiguousarray(truth_indptr, dtype=np.int64)
        tt = _i32(truth_items)
        ks = _i32(Ks)
        out = np.zeros((4, len(ks)), dtype=np.float64)
        check(self.lib.pda_metrics_host(self._h, ptr(ids), ids.shape[0], ids.shape[1], ptr(eu), ptr(ti), ptr(tt),
                                        len(ti) - 1, ptr(ks), len(ks), ptr(out)))
        return dict(precision=out[0], recall=out[1], ndcg=out[2], hit_ratio=out[3])
