"""Data-parallel drivers of the train step and of the evaluation over the GPUs of one box.

Partition: USERS are split into contiguous shards, one per rank; every rank samples its B triples from
its own users (rank-local sampling), so user rows, user gradients and the user table's Adam state never
leave the GPU.  The ITEM table (small: n_items x d) is replicated; its gradient is the one real exchange
step.  Default ("scatter", when n_items divides by the world size): NCCL reduce-scatter of the dense
item-gradient accumulator -> every rank runs the Adam sweep on ITS row slice only (1/world of the sweep, and
only that slice of the item Adam slots is live) -> NCCL all-gather of the updated rows, waited for right
before the next step's kernel reads the table.  Fallback ("allreduce"): all-reduce of the gradient in row
chunks pipelined with the full dense sweep on every rank.  Either way the result equals a single-process
step on the union batch of world*B triples (loss mean and L2 divisor use the global batch).

torch.distributed is plumbing: it owns the NCCL communicator and the stream; the kernels are the library's.
With world == 1 this class adds nothing to the single-GPU path.
"""
from __future__ import annotations


class _DevArray:
    """Exposes a raw device pointer owned by libpda_b200 through __cuda_array_interface__."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"data": (int(ptr), False), "shape": tuple(shape), "typestr": typestr,
                                         "version": 2, "strides": None}


def shard_range(n, world, rank):
    """Contiguous shard [lo, hi) of n rows for `rank`; the first n % world ranks get one extra row."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class ShardedTrainer:
    def __init__(self, model, world=1, rank=0, reducer=None, chunks=4, exchange=None):
        """reducer(tensor) -> None sums `tensor` in place over ranks (default: torch.distributed.all_reduce).
        exchange: "scatter" | "allreduce" | None (= $PDA_DP_EXCHANGE, else scatter where it applies)."""
        import os
        self.model, self.world, self.rank = model, int(world), int(rank)
        self._wi = None
        self._pending = None
        self._own = None
        exchange = exchange or os.environ.get("PDA_DP_EXCHANGE", "scatter")
        if int(world) > 1 and getattr(model, "train", "") == "temp_pop":
            raise NotImplementedError("data-parallel training covers BPRMF / PD / PDG (the bias tables of BPR(t)-pop are not exchanged)")
        self._gi = self._acc = None
        self._reduce = reducer
        self._async_reduce = None
        self.chunks = int(chunks)
        if self.world > 1:
            if hasattr(model, "set_adam_mode"):
                # the item gradient is the sum over ranks: which item rows were touched is not known locally,
                # so the replicated item table keeps the dense sweep; the rank-local user table stays lazy
                model.set_adam_mode("lazy_users")
            import torch
            import torch.distributed as dist
            if reducer is None:
                self._reduce = lambda t: dist.all_reduce(t, op=dist.ReduceOp.SUM)
                if hasattr(model, "lib"):     # the GPU model: split optimizer + asynchronous NCCL work handles
                    self._async_reduce = lambda t: dist.all_reduce(t, op=dist.ReduceOp.SUM, async_op=True)
            if hasattr(model, "exchange_tensors"):     # host stand-ins (tests) hand their buffers over directly
                self._gi, self._acc = model.exchange_tensors()
            else:
                dev = torch.device("cuda", model.device)
                self._gi = torch.as_tensor(_DevArray(model.grad_ptr("item_embedding"), (model.n_items, model.emb_dim),
                                                     "<f4"), device=dev)
                self._acc = torch.as_tensor(_DevArray(model.loss_acc_ptr(), (2,), "<f8"), device=dev)
                if exchange == "scatter" and self._async_reduce is not None and model.n_items % self.world == 0:
                    self._wi = torch.as_tensor(_DevArray(model.table_ptr("item_embedding"), (model.n_items, model.emb_dim),
                                                         "<f4"), device=dev)
                    rows = model.n_items // self.world
                    self._own = (self.rank * rows, (self.rank + 1) * rows)

    def _exchange(self):
        self._reduce(self._gi)     # dense item-gradient block, summed over ranks (NVLink / NVSwitch)
        self._reduce(self._acc)    # 2 doubles: loss partial sums -> global mean

    def _exchange_and_apply(self, stream):
        """item-gradient all-reduce overlapped with the rank-local half of the optimizer: the user table's gradient
        never leaves the GPU, so its Adam update runs on the compute stream while NCCL moves the item gradient."""
        m = self.model
        if self._async_reduce is None:
            self._exchange()
            m.adam_apply(stream)
            return
        if self._own is not None:
            import torch.distributed as dist
            lo, hi = self._own
            # in place: this rank's slice of the accumulator receives the sum over ranks of that slice
            w_rs = dist.reduce_scatter_tensor(self._gi[lo:hi], self._gi, op=dist.ReduceOp.SUM, async_op=True)
            wacc = self._async_reduce(self._acc)
            m.adam_apply(stream, part=1)          # rank-local tables (nothing when the step kernel already did it)
            w_rs.wait()                           # stream-level dependency, the host does not block
            m.adam_dense_rows("item_embedding", lo, hi, stream)      # consumes and zeroes rows [lo, hi) of the accumulator
            if lo > 0:
                self._gi[:lo].zero_()             # the other slices hold this rank's partial sums
            if hi < m.n_items:
                self._gi[hi:].zero_()
            self._pending = dist.all_gather_into_tensor(self._wi, self._wi[lo:hi], async_op=True)
            wacc.wait()
            m.adam_apply(stream, part=8)
            return
        # the item gradient travels in row chunks: chunk k's all-reduce overlaps the dense Adam sweep of chunk k-1
        # (and, first of all, the rank-local half of the optimizer)
        n = m.n_items
        nch = self.chunks if n >= 4096 * self.chunks else 1
        bounds = [n * k // nch for k in range(nch + 1)]
        works = [self._async_reduce(self._gi[bounds[k]:bounds[k + 1]]) for k in range(nch)]
        wacc = self._async_reduce(self._acc)
        m.adam_apply(stream, part=1)
        for k in range(nch):
            works[k].wait()        # stream-level dependency, the host does not block
            m.adam_dense_rows("item_embedding", bounds[k], bounds[k + 1], stream)
        wacc.wait()
        m.adam_apply(stream, part=8)

    def train_sampled(self, seed, epoch, step0, n_steps, B, stream=0):
        m = self.model
        if self.world == 1:
            m.train_sampled(seed, epoch, step0, n_steps, B, stream)
            return
        m.set_global_batch(B * self.world)
        for k in range(n_steps):
            m.sample_batch(seed, epoch, step0 + k, B, stream, fetch=False)      # overlaps the all-gather of the last step
            self.finish()
            m.forward_backward_device(B, stream)
            self._exchange_and_apply(stream)
        self.finish()

    def finish(self):
        """the item table is complete on this rank's compute stream (all-gather of the last step done)"""
        if self._pending is not None:
            self._pending.wait()
            self._pending = None

    def train_step_host(self, users, pos, neg, pos_pop=None, neg_pop=None, stream=0):
        m = self.model
        if self.world == 1:
            return m.train_step(users, pos, neg, pos_pop, neg_pop)
        m.set_global_batch(len(users) * self.world)
        B = m.stage_batch(users, pos, neg, pos_pop, neg_pop, stream)
        self.finish()
        m.forward_backward_device(B, stream)
        self._exchange_and_apply(stream)
        self.finish()
        return m.read_loss(stream)


class ShardedEvaluator:
    """Evaluation over user shards: every rank scores ITS eval users against all items (the item table is
    replicated, so the recommender needs no exchange), reduces Recall / Precision / NDCG / Hit sums on its GPU, and one
    all-reduce of the 4 x len(Ks) sums (+ the user count) gives the global means -- train_new_api.py:741-778 averaged
    over all ranks' users.  `users` are rank-local row ids of the model's user table; `truth_*` is the CSR of the
    rank's own eval users."""

    def __init__(self, model, world=1, rank=0, reducer=None):
        self.model, self.world, self.rank = model, int(world), int(rank)
        self._reduce = reducer
        if self.world > 1 and reducer is None:
            import torch.distributed as dist
            self._reduce = lambda t: dist.all_reduce(t, op=dist.ReduceOp.SUM)

    def eval(self, users, truth_indptr, truth_items, Ks, rec_type="main_branch", pos_pop=None, K=50, device=None):
        import numpy as np
        Ks = list(Ks)
        keys = ("precision", "recall", "ndcg", "hit_ratio")
        sums = np.zeros(4 * len(Ks) + 1, dtype=np.float64)
        if len(users):
            ids = self.model.do_recommendation(users, None, rec_type, pos_pop=pos_pop, K=K)
            s = self.model.metrics_sum(ids, users, truth_indptr, truth_items, Ks)
            sums[:-1] = np.concatenate([np.asarray(s[k], dtype=np.float64) for k in keys])
            sums[-1] = len(users)
        if self.world > 1:
            import torch
            t = torch.from_numpy(sums)
            if device is not None:
                t = t.to(device)
            self._reduce(t)
            sums = t.cpu().numpy()
        n = max(sums[-1], 1.0)
        return {k: sums[i * len(Ks):(i + 1) * len(Ks)] / n for i, k in enumerate(keys)}
