"""world_size-2 check (gloo, CPU) of the data-parallel driver logic in pda_b200/parallel.py: user shards + replicated
item table + item-gradient all-reduce must reproduce ONE process stepping on the union batch.  The GPU model is
replaced by a host stand-in with the same split-step interface (forward_backward / adam_apply) built on the oracle."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_covers_everything():
    from pda_b200.parallel import shard_range
    for n, w in [(10, 3), (7, 8), (1000, 4), (5, 1)]:
        parts = [shard_range(n, w, r) for r in range(w)]
        assert parts[0][0] == 0 and parts[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
        sizes = [hi - lo for lo, hi in parts]
        assert max(sizes) - min(sizes) <= 1


class HostStandIn:
    """Same split-step surface as PDAModel (stage_batch / forward_backward_device / adam_apply / read_loss)."""

    def __init__(self, U, I, lr, regs, batch_size):
        import torch
        from oracle import pda_oracle as po
        self.po, self.U, self.I = po, U.copy(), I.copy()
        self.n_items, self.emb_dim = I.shape
        self.lr, self.regs, self.batch_size = lr, regs, batch_size
        self.aU, self.aI, self.pw = po.AdamState(U.shape), po.AdamState(I.shape), po.AdamPowers()
        self.GI = torch.zeros(I.shape, dtype=torch.float32)
        self.acc = torch.zeros(2, dtype=torch.float64)
        self.GU = np.zeros_like(U)
        self.Bg = 0

    def exchange_tensors(self):
        import torch
        return self.GI, self.acc, torch.from_numpy(self.I)       # the item table shares memory with self.I

    def slot_tensor(self, name):
        import torch
        return torch.from_numpy(self.aI.m if name == "item_m" else self.aI.v)

    def set_global_batch(self, Bg):
        self.Bg = Bg

    def stage_batch(self, users, pos, neg, pp, npop, stream=0):
        self.batch = (users, pos, neg, pp, npop)
        return len(users)

    def forward_backward_device(self, B, stream=0):
        po = self.po
        users, pos, neg, pp, npop = self.batch
        r = po.bpr_step_forward_backward(self.U, self.I, users, pos, neg, self.regs, self.batch_size, "s_condition", pp, npop)
        scale = np.float32(len(users) / self.Bg)          # the kernel scales by 1/global_batch instead of 1/B
        lb = np.float32(np.float64(np.float32(self.regs)) / self.batch_size)
        u, p, n = self.U[users], self.I[pos], self.I[neg]
        gU = (r["gU_rows"] - lb * u) * scale + lb * u
        gP = (r["gP_rows"] - lb * p) * scale + lb * p
        gN = (r["gN_rows"] - lb * n) * scale + lb * n
        GU, _ = po.dedup_sum(self.U.shape[0], self.emb_dim, [users], [gU])
        GI, _ = po.dedup_sum(self.n_items, self.emb_dim, [pos, neg], [gP, gN])
        self.GU += GU
        self.GI += __import__("torch").from_numpy(GI)
        self.acc[0] += -float(r["mf_loss"]) * len(users)
        self.acc[1] += float(r["reg_loss"]) * self.batch_size / (0.5 * float(np.float32(self.regs)))

    # the split optimizer of PDAModel: part 1 = rank-local (user) table, part 2 = dense item sweep + bookkeeping,
    # part 8 = bookkeeping only (the item sweep ran in row ranges)
    def adam_apply(self, stream=0, part=3):
        po = self.po
        lr_t = self.pw.lr_t(self.lr)
        if part & 1:
            po.adam_apply_dense(self.U, self.aU, self.GU, lr_t)
            self.GU[:] = 0
        if part & 2:
            self.adam_dense_rows("item_embedding", 0, self.n_items)
        if part & (2 | 8):
            self.pw.finish()
            mf = -float(self.acc[0]) / self.Bg
            reg = float(np.float32(self.regs)) * 0.5 * float(self.acc[1]) / self.batch_size
            self.loss3 = (mf + reg, mf, reg)
            self.acc.zero_()

    def _item_rows(self, lo, hi, G):
        po = self.po
        st = po.AdamState((hi - lo, self.emb_dim))
        st.m, st.v = self.aI.m[lo:hi], self.aI.v[lo:hi]
        W = self.I[lo:hi]
        po.adam_apply_dense(W, st, G, self.pw.lr_t(self.lr))      # W is a view: updated in place
        self.aI.m[lo:hi], self.aI.v[lo:hi] = st.m, st.v

    def adam_dense_rows(self, name, lo, hi, stream=0):
        self._item_rows(lo, hi, self.GI.numpy()[lo:hi].copy())
        self.GI[lo:hi] = 0

    def adam_dense_rows_ext(self, name, lo, hi, grad, stream=0):
        self._item_rows(lo, hi, grad.numpy().copy())

    def read_loss(self, stream=0):
        return self.loss3


def _worker(rank, world, port, q, exchange="scatter"):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from pda_b200.parallel import ShardedTrainer, shard_range
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(0)                         # same data on every rank
        n_users, n_items, d, B = 120, 48, 16, 32
        U = rng.normal(0, 0.3, (n_users, d)).astype(np.float32)
        I = rng.normal(0, 0.3, (n_items, d)).astype(np.float32)
        lo, hi = shard_range(n_users, world, rank)
        model = HostStandIn(U[lo:hi], I, 1e-2, 1e-3, B * world)
        os.environ["PDA_DP_CHUNKS"] = "4"
        tr = ShardedTrainer(model, world, rank, exchange=exchange)
        assert (tr._own is not None) == (exchange == "scatter")
        assert exchange != "scatter" or tr.nch == 4      # 48 item rows: 4 exchange chunks of 12 rows, 6 per rank
        losses = []
        for step in range(4):
            srng = np.random.default_rng(100 + step)
            batches = []
            for r in range(world):
                rlo, rhi = shard_range(n_users, world, r)
                users = srng.permutation(rhi - rlo)[:B].astype(np.int32)      # rank-local user ids
                batches.append((users, srng.integers(0, n_items, B).astype(np.int32),
                                srng.integers(0, n_items, B).astype(np.int32), srng.random(B).astype(np.float32),
                                srng.random(B).astype(np.float32)))
            losses.append(tr.train_step_host(*batches[rank]))
        tr.sync_item_slots()                                   # scatter: every rank ends with the full item Adam slots
        q.put((rank, lo, hi, model.U, model.I, losses, model.aI.m.copy(), model.aI.v.copy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("exchange", ["scatter", "allreduce"])
def test_two_ranks_equal_one_process_on_the_union_batch(exchange):
    import torch.multiprocessing as mp
    from oracle import pda_oracle as po
    from pda_b200.parallel import shard_range
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world = 2
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, exchange)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=240) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    # single process, union batch (global user ids), same init
    rng = np.random.default_rng(0)
    n_users, n_items, d, B = 120, 48, 16, 32
    U = rng.normal(0, 0.3, (n_users, d)).astype(np.float32)
    I = rng.normal(0, 0.3, (n_items, d)).astype(np.float32)
    om = po.OracleModel(n_users, n_items, d, 1e-2, 1e-3, B * world, "s_condition", U=U, I=I)
    ref_losses = []
    for step in range(4):
        srng = np.random.default_rng(100 + step)
        parts = []
        for r in range(world):
            rlo, rhi = shard_range(n_users, world, r)
            users = srng.permutation(rhi - rlo)[:B].astype(np.int32) + rlo
            parts.append((users, srng.integers(0, n_items, B).astype(np.int32), srng.integers(0, n_items, B).astype(np.int32),
                          srng.random(B).astype(np.float32), srng.random(B).astype(np.float32)))
        cat = [np.concatenate([p[k] for p in parts]) for k in range(5)]
        ref_losses.append(om.train_step(*cat))
    I0, I1 = res[0][4], res[1][4]
    assert np.array_equal(I0, I1)                                   # replicas stay identical
    assert np.abs(I0 - om.I).max() <= 2e-5 * np.abs(om.I).max()     # = the union-batch step (fp32 sum order differs)
    for rank, lo, hi, Ur, _, losses, mI, vI in res:
        assert np.abs(Ur - om.U[lo:hi]).max() <= 2e-5 * np.abs(om.U).max()
        for got, want in zip(losses, ref_losses):
            assert np.allclose(got, want, rtol=1e-5)
        # item Adam slots complete on every rank (scatter: after sync_item_slots)
        assert np.abs(mI - om.aI.m).max() <= 2e-5 * np.abs(om.aI.m).max()
        assert np.abs(vI - om.aI.v).max() <= 2e-5 * np.abs(om.aI.v).max()


# ---------------------------------------------------------------------------------------------
# sharded evaluation: per-rank metric sums + one all-reduce = the single-process means
# ---------------------------------------------------------------------------------------------
class EvalStandIn:
    """do_recommendation / metrics_sum of PDAModel on the CPU oracle (rank-local user ids)."""

    def __init__(self, U, I, indptr, items):
        from oracle import pda_oracle as po
        self.po, self.U, self.I, self.indptr, self.items = po, U, I, indptr, items

    def do_recommendation(self, users, items=None, rec_type="main_branch", pos_pop=None, K=50):
        return self.po.recommend(self.U, self.I, np.asarray(users, dtype=np.int32), rec_type, K, self.indptr, self.items, pop=pos_pop)

    def metrics_sum(self, ids, eval_users, truth_indptr, truth_items, Ks):
        return _metrics_sum(self.po, ids, eval_users, truth_indptr, truth_items, Ks)


def _metrics_sum(po, ids, eval_users, truth_indptr, truth_items, Ks):
    res = {k: np.zeros(len(Ks)) for k in ("precision", "recall", "ndcg", "hit_ratio")}
    for r, u in enumerate(eval_users):
        one = po.get_performance(truth_items[truth_indptr[u]:truth_indptr[u + 1]], ids[r], Ks)
        for k in res:
            res[k] += one[k]
    return res


def _eval_problem():
    from oracle import pda_oracle as po
    rng = np.random.default_rng(5)
    n_users, n_items, d = 90, 120, 8
    U = rng.normal(0, 1, (n_users, d)).astype(np.float32)
    I = rng.normal(0, 1, (n_items, d)).astype(np.float32)
    uid = np.repeat(np.arange(n_users), 6); iid = rng.integers(0, n_items, len(uid))
    key = np.unique(uid.astype(np.int64) * n_items + iid)
    tr_ptr, tr_items, _ = po.build_csr(n_users, key // n_items, key % n_items)
    uid2 = np.repeat(np.arange(n_users), 4); iid2 = rng.integers(0, n_items, len(uid2))
    key2 = np.unique(uid2.astype(np.int64) * n_items + iid2)
    te_ptr, te_items, _ = po.build_csr(n_users, key2 // n_items, key2 % n_items)
    pop = rng.random(n_items).astype(np.float32)
    return U, I, tr_ptr, tr_items, te_ptr, te_items, pop


def _csr_slice(indptr, items, lo, hi):
    return (indptr[lo:hi + 1] - indptr[lo]).astype(np.int64), items[indptr[lo]:indptr[hi]]


def _eval_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from pda_b200.parallel import ShardedEvaluator, shard_range
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        U, I, tr_ptr, tr_items, te_ptr, te_items, pop = _eval_problem()
        lo, hi = shard_range(len(U), world, rank)
        mp_, mi_ = _csr_slice(tr_ptr, tr_items, lo, hi)
        tp_, ti_ = _csr_slice(te_ptr, te_items, lo, hi)
        ev = ShardedEvaluator(EvalStandIn(U[lo:hi], I, mp_, mi_), world, rank)
        users = np.arange(hi - lo, dtype=np.int32)[:: 2 if rank == 0 else 1]       # uneven shards
        out = ev.eval(users, tp_, ti_, [5, 20], rec_type="condition", pos_pop=pop, K=20)
        q.put((rank, lo, users + lo, out))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_eval_equals_single_process():
    import torch.multiprocessing as mp
    from oracle import pda_oracle as po
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world = 2
    procs = [ctx.Process(target=_eval_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    U, I, tr_ptr, tr_items, te_ptr, te_items, pop = _eval_problem()
    users = np.concatenate([r[2] for r in res]).astype(np.int32)
    ids = po.recommend(U, I, users, "condition", 20, tr_ptr, tr_items, pop=pop)
    want = _metrics_sum(po, ids, users, te_ptr, te_items, [5, 20])
    for r in res:
        for k in ("precision", "recall", "ndcg", "hit_ratio"):
            assert np.allclose(r[3][k], np.asarray(want[k]) / len(users), rtol=1e-12, atol=1e-15), k
