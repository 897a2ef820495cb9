"""Data-parallel drivers of the train step and of the evaluation over the GPUs of one box.

Partition: USERS are split into contiguous shards, one per rank; every rank samples its B triples from
its own users (rank-local sampling), so user rows, user gradients and the user table's Adam state never
leave the GPU.  The ITEM table (small: n_items x d) is replicated; its gradient is the one real exchange
step.  Default ("scatter", when n_items divides by the world size), per step:

    sampler k+1 (side stream)  ||  step kernel k -> reduce-scatter (out of place) -> sliced Adam -> all-gather
                                                      \\-> zero the accumulator (side stream, under the all-gather)

  * NCCL reduce-scatter of the dense item-gradient accumulator into a separate [rows/world, d] buffer, so the
    accumulator is free to be zeroed while the all-gather runs;
  * every rank runs the Adam sweep on ITS row slice only (1/world of the sweep; only that slice of the item
    Adam slots is live -- sync_item_slots() all-gathers them for checkpoints);
  * NCCL all-gather of the updated rows, waited for right before the next step kernel reads the table;
  * the sampler of step k+1 (or the host->device copies of host batch k+1) starts when step k's kernel has read
    the batch buffers and runs under the exchange.
Fallback ("allreduce"): all-reduce of the gradient in row chunks pipelined with the full dense sweep on every
rank.  Either way the result equals a single-process step on the union batch of world*B triples (loss mean and
L2 divisor use the global batch).

torch.distributed is plumbing: it owns the NCCL communicator and the streams; the kernels are the library's.
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


class _NullCtx:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class ShardedTrainer:
    def __init__(self, model, world=1, rank=0, reducer=None, chunks=4, exchange=None):
        """reducer(tensor) -> None sums `tensor` in place over ranks (default: torch.distributed.all_reduce).
        exchange: "scatter" | "allreduce" | None (= $PDA_DP_EXCHANGE, else scatter where it applies)."""
        import os
        self.model, self.world, self.rank = model, int(world), int(rank)
        self._wi = None
        self._own = None
        self._gslice = None
        self._cuda = hasattr(model, "lib")      # the GPU model (host stand-ins of the tests run the same logic on CPU)
        self._side = self._zs = self._copy = None
        self._zero_pending = False
        self._pending = []
        self.nch = 1
        exchange = exchange or os.environ.get("PDA_DP_EXCHANGE", "auto")     # auto: nvls where available, else scatter
        self.exchange = exchange
        self._nvls = None
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
                if hasattr(model, "adam_dense_rows_ext"):     # split optimizer available: asynchronous work handles
                    self._async_reduce = lambda t: dist.all_reduce(t, op=dist.ReduceOp.SUM, async_op=True)
            scatter_ok = exchange in ("scatter", "auto", "nvls", "p2p") and self._async_reduce is not None and model.n_items % self.world == 0
            if hasattr(model, "exchange_tensors"):     # host stand-ins (tests) hand their buffers over directly
                ex = model.exchange_tensors()
                self._gi, self._acc = ex[0], ex[1]
                if scatter_ok and len(ex) > 2:
                    self._wi = ex[2]
            else:
                dev = torch.device("cuda", model.device)
                self._gi = torch.as_tensor(_DevArray(model.grad_ptr("item_embedding"), (model.n_items, model.emb_dim),
                                                     "<f4"), device=dev)
                self._acc = torch.as_tensor(_DevArray(model.loss_acc_ptr(), (2,), "<f8"), device=dev)
                if scatter_ok:
                    self._wi = torch.as_tensor(_DevArray(model.table_ptr("item_embedding"), (model.n_items, model.emb_dim),
                                                         "<f4"), device=dev)
                self._side, self._zs, self._copy = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev)
                # auto: symmetric memory + the fused exchange kernel -- over the multicast mapping from 4 ranks on (in-switch
                # reduction: 8 GPUs 2.49 vs 2.93 ms/step with NCCL), over unicast peer pointers below ("p2p")
                if scatter_ok and exchange in ("nvls", "p2p", "auto"):
                    unicast = exchange == "p2p" or (exchange == "auto" and self.world < 4)
                    self._nvls = self._setup_nvls(dev, need_multicast=not unicast)
                    if self._nvls is not None:
                        self._nvls["unicast"] = unicast
                    if self._nvls is None and exchange in ("nvls", "p2p"):
                        raise RuntimeError("exchange=%r requested but symmetric memory / NVLink multicast is not available" % exchange)
            if self._nvls is not None:
                rows = model.n_items // self.world
                self.nch = 1
                self._own = [(self.rank * rows, (self.rank + 1) * rows)]
                self._chunk = [(0, model.n_items)]
                self.exchange = "p2p" if self._nvls["unicast"] else "nvls"
            elif self._wi is not None:
                self.exchange = "scatter"
                # The exchange runs in `nch` row chunks; inside chunk c (rows [c R, (c+1) R), R = n_items / nch) rank r
                # owns rows [c R + r R/world, c R + (r+1) R/world).  Reduce-scatter of chunk c+1 (send-heavy with in-switch
                # reduction) and all-gather of chunk c (receive-heavy with multicast) travel on two communicators and
                # overlap; so does the sliced Adam sweep.
                nch = int(os.environ.get("PDA_DP_CHUNKS", "2"))     # measured at 8 GPUs: 1 -> 2.96, 2 -> 2.92, 4 -> 2.94, 8 -> 3.12 ms/step
                while nch > 1 and model.n_items % (nch * self.world) != 0:
                    nch -= 1
                self.nch = nch
                R = model.n_items // nch
                sub = R // self.world
                self._own = [(c * R + self.rank * sub, c * R + (self.rank + 1) * sub) for c in range(nch)]
                self._chunk = [(c * R, (c + 1) * R) for c in range(nch)]
                self._gslice = [torch.empty((sub, model.emb_dim), dtype=torch.float32, device=self._gi.device) for _ in range(nch)]
                self._pg_ag = None
                if self._cuda and nch > 1 and os.environ.get("PDA_DP_TWO_COMMS", "1") != "0":
                    try:
                        opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
                        self._pg_ag = dist.new_group(backend="nccl", pg_options=opts)
                    except Exception:
                        self._pg_ag = dist.new_group(backend="nccl")
        self._prof = None

    def _setup_nvls(self, dev, need_multicast=True):
        """item table + item-gradient accumulator into symmetric memory with a multicast mapping (torch plumbing); None when
        the platform has no NVLink multicast.  Collective: every rank calls it."""
        import os
        import torch
        import torch.distributed as dist
        m = self.model
        n = m.n_items * m.emb_dim
        ok, st = 1, None
        try:
            import torch.distributed._symmetric_memory as symm
            try:
                symm.enable_symm_mem_for_group(dist.group.WORLD.group_name)
            except Exception:
                pass
            buf = symm.empty(3 * n + 64, dtype=torch.float32, device=dev)      # [W | G0 | G1 | 64 barrier flags]: accumulator double-buffered
            hdl = symm.rendezvous(buf, dist.group.WORLD)
            mc = int(hdl.multicast_ptr)
            if mc == 0 and need_multicast:
                ok = 0
            ptrs = [int(x) for x in hdl.buffer_ptrs]
            st = dict(buf=buf, hdl=hdl, mcW=mc, mcG=[mc + 4 * n, mc + 8 * n], cur=0, peerW=ptrs,
                      peerG=[[x + 4 * n for x in ptrs], [x + 8 * n for x in ptrs]])
        except Exception as e:      # no symmetric memory on this platform / torch build
            import sys
            print("pda_b200: NVLink multicast exchange unavailable (%r); using the NCCL exchange" % (e,), file=sys.stderr)
            ok = 0
        flag = torch.tensor([ok], device=dev, dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            return None
        buf = st["buf"]
        buf[n:].zero_()
        torch.cuda.synchronize()
        st["hdl"].barrier(channel=0)              # every rank's flags are zero before anyone signals
        torch.cuda.synchronize()
        m.adopt_item_buffers(buf.data_ptr(), buf.data_ptr() + 4 * n)
        # cross-rank barriers inside the exchange kernel (flags in the symmetric buffer) instead of two barrier launches
        # measured: 2 GPUs 2.33 vs 2.39 ms/step in favour of the in-kernel barriers; 8 GPUs 2.67 vs 2.49 against them (the
        # exchange kernel then starts together with the sampler of the next step, which crawls beside it: 1.39 vs 0.57 ms)
        st["inkernel"] = os.environ.get("PDA_DP_INKERNEL_BARRIER", "1" if self.world < 4 else "0") != "0"
        if st["inkernel"]:
            m.dp_set_barrier(buf.data_ptr() + 12 * n, [x + 12 * n for x in st["peerW"]], self.rank)
        st["flags"] = buf[3 * n:].view(torch.int32)
        self._wi = buf[:n].view(m.n_items, m.emb_dim)
        st["G"] = [buf[n:2 * n].view(m.n_items, m.emb_dim), buf[2 * n:3 * n].view(m.n_items, m.emb_dim)]
        self._gi = st["G"][0]
        return st

    # ---- stream plumbing: the library enqueues on the raw `stream`; torch (NCCL, memsets) must order against the same one
    def _cs(self, stream):
        if not self._cuda:
            return None
        import torch
        return torch.cuda.ExternalStream(stream) if stream else torch.cuda.default_stream(self._gi.device)

    def _on(self, cs):
        if cs is None:
            return _NullCtx()
        import torch
        return torch.cuda.stream(cs)

    # ---- optional per-phase timeline (CUDA events on the compute stream) ----
    def profile(self, on=True):
        self._prof = {"fwd": [], "rs": [], "adam": [], "ag": []} if on else None

    def _mark(self, name):
        if self._prof is None or not self._cuda:
            return
        import torch
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        self._prof[name].append(e)

    def profile_summary(self):
        """mean ms per step on the compute stream: step kernel end -> last reduce-scatter chunk waited (`rs`: exposed
        reduce-scatter), -> last sliced Adam sweep enqueued and done (`adam`), -> all-gather waited by the next step
        (`ag`: exposed all-gather, the sampler of the next step runs under it)."""
        if self._prof is None or not self._prof["fwd"]:
            return None
        import torch
        torch.cuda.synchronize()
        p = self._prof
        n = min(len(p["fwd"]), len(p["rs"]), len(p["adam"]), len(p["ag"]))
        out = {"steps": n, "exchange": self.exchange, "chunks": self.nch, "two_communicators": getattr(self, "_pg_ag", None) is not None}
        if self.exchange in ("nvls", "p2p"):
            out["in_kernel_barriers"] = bool(self._nvls["inkernel"])
            out["barrier_timeouts"] = int(self._nvls["flags"][17].item())
            out["phases"] = ("rs_exposed = barrier + fused multimem kernel (the other accumulator is zeroed under it), "
                             "adam_after_rs = second barrier, ag_exposed = 0 (nothing left on the critical path)")
        out["rs_exposed_ms"] = sum(p["fwd"][i].elapsed_time(p["rs"][i]) for i in range(n)) / n
        out["adam_after_rs_ms"] = sum(p["rs"][i].elapsed_time(p["adam"][i]) for i in range(n)) / n
        out["ag_exposed_ms"] = sum(p["adam"][i].elapsed_time(p["ag"][i]) for i in range(n)) / n
        out["exchange_total_ms"] = sum(p["fwd"][i].elapsed_time(p["ag"][i]) for i in range(n)) / n
        return out

    def _exchange(self):
        self._reduce(self._gi)     # dense item-gradient block, summed over ranks (NVLink / NVSwitch)
        self._reduce(self._acc)    # 2 doubles: loss partial sums -> global mean

    def _exchange_and_apply(self, stream, cs=None):
        """item-gradient exchange + optimizer.  The user table's gradient never leaves the GPU: its Adam update runs on
        the compute stream (or already ran inside the step kernel) while NCCL moves the item gradient."""
        m = self.model
        if self._async_reduce is None:
            self._exchange()
            m.adam_apply(stream)
            return
        if self._nvls is not None:
            # reduce-scatter + sliced Adam + all-gather in ONE kernel over NVLink multicast (pda_exchange.cu), bracketed by
            # two cross-rank barriers on the compute stream (symmetric-memory signal pads)
            nv = self._nvls
            h, cur = nv["hdl"], nv["cur"]
            wacc = self._async_reduce(self._acc)
            m.adam_apply(stream, part=1)
            if not nv["inkernel"]:
                h.barrier(channel=0)              # every rank's step kernel is done: the accumulators are complete
            if cs is not None:
                # the OTHER accumulator (step k-1's, fully read by every rank since that step's second barrier) is zeroed on a
                # side stream under this NVLink-bound kernel; the next step accumulates into it
                self._zs.wait_stream(cs)
                with self._on(self._zs):
                    nv["G"][1 - cur].zero_()
                self._zero_pending = True
            else:
                nv["G"][1 - cur].zero_()
            lo, hi = self._own[0]
            if nv["unicast"]:
                m.dp_exchange_adam_p2p(nv["peerG"][cur], nv["peerW"], self.rank, lo, hi, stream)
            else:
                m.dp_exchange_adam(nv["mcG"][cur], nv["mcW"], lo, hi, stream)
            self._mark("rs")
            if not nv["inkernel"]:
                h.barrier(channel=1)              # every replica written, every accumulator read
            self._mark("adam")
            nv["cur"] = 1 - cur
            self._gi = nv["G"][1 - cur]
            m.set_item_grad_buffer(self._gi.data_ptr())
            self._mark("ag")
            wacc.wait()
            m.adam_apply(stream, part=8)
            return
        if self._own is not None:
            import torch.distributed as dist
            # out of place: the sum over ranks of this rank's rows of chunk c lands in _gslice[c]; the accumulator is only read
            w_rs = [dist.reduce_scatter_tensor(self._gslice[c], self._gi[lo:hi], op=dist.ReduceOp.SUM, async_op=True)
                    for c, (lo, hi) in enumerate(self._chunk)]
            wacc = self._async_reduce(self._acc)
            m.adam_apply(stream, part=1)          # rank-local tables (nothing when the step kernel already did it)
            for c, (lo, hi) in enumerate(self._own):
                w_rs[c].wait()                    # stream-level dependency, the host does not block
                if c == self.nch - 1:             # the accumulator has been consumed: zero it under the all-gather
                    self._mark("rs")
                    if cs is not None:
                        self._zs.wait_stream(cs)
                        with self._on(self._zs):
                            self._gi.zero_()
                        self._zero_pending = True
                    else:
                        self._gi.zero_()
                g = self._gslice[c] if not self._cuda else self._gslice[c].data_ptr()
                m.adam_dense_rows_ext("item_embedding", lo, hi, g, stream)
                clo, chi = self._chunk[c]
                self._pending.append(dist.all_gather_into_tensor(self._wi[clo:chi], self._wi[lo:hi], group=self._pg_ag, async_op=True))
            self._mark("adam")
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
        cs = self._cs(stream)
        with self._on(cs):
            if cs is None or self._side is None:
                for k in range(n_steps):
                    m.sample_batch(seed, epoch, step0 + k, B, stream, fetch=False)
                    self.finish(cs)
                    m.forward_backward_device(B, stream)
                    self._exchange_and_apply(stream, cs)
                self.finish(cs)
                return
            side = self._side
            side.wait_stream(cs)            # earlier work on the compute stream may still read the batch buffers
            m.sample_batch(seed, epoch, step0, B, side.cuda_stream, fetch=False)
            for k in range(n_steps):
                cs.wait_stream(side)        # batch k is sampled
                self.finish(cs)             # item table complete (all-gather of step k-1), accumulator zeroed
                m.forward_backward_device(B, stream)
                self._mark("fwd")
                if k + 1 < n_steps:         # the sampler of step k+1 runs under step k's exchange
                    side.wait_stream(cs)
                    m.sample_batch(seed, epoch, step0 + k + 1, B, side.cuda_stream, fetch=False)
                self._exchange_and_apply(stream, cs)
            self.finish(cs)

    def finish(self, cs=None):
        """the item table is complete on this rank's compute stream (all-gather of the last step done) and the
        gradient accumulator is zero again"""
        if self._pending:
            for w in self._pending:
                w.wait()
            self._pending = []
            self._mark("ag")
        if self._zero_pending:
            import torch
            (cs if cs is not None else torch.cuda.current_stream()).wait_stream(self._zs)
            self._zero_pending = False

    def train_step_host(self, users, pos, neg, pos_pop=None, neg_pop=None, stream=0):
        m = self.model
        if self.world == 1:
            return m.train_step(users, pos, neg, pos_pop, neg_pop)
        m.set_global_batch(len(users) * self.world)
        cs = self._cs(stream)
        with self._on(cs):
            B = m.stage_batch(users, pos, neg, pos_pop, neg_pop, stream)
            self.finish(cs)
            m.forward_backward_device(B, stream)
            self._exchange_and_apply(stream, cs)
            self.finish(cs)
        return m.read_loss(stream)

    def train_steps_host(self, users, pos, neg, pos_pop=None, neg_pop=None, stream=0):
        """n consecutive steps from PINNED host arrays [n, B] (PDAModel.pinned_array): the host->device copies of
        batch k+1 (and the device-side id / distinct-users check) run on a copy stream under step k's gradient
        exchange; one async D2H of the 3 loss scalars per step.  Returns the [n, 3] losses; same results as n calls
        of train_step_host."""
        import numpy as np
        m = self.model
        n, B = users.shape
        if self.world == 1:
            return m.train_steps(users, pos, neg, pos_pop, neg_pop)
        m.set_global_batch(B * self.world)
        cs = self._cs(stream)
        ring = m.pinned_array((n, 4), np.float32)
        pp = (lambda k: None) if pos_pop is None else (lambda k: pos_pop[k])
        pn = (lambda k: None) if neg_pop is None else (lambda k: neg_pop[k])
        with self._on(cs):
            self._copy.wait_stream(cs)
            m.stage_batch_async(users[0], pos[0], neg[0], pp(0), pn(0), self._copy.cuda_stream)
            for k in range(n):
                m.staged_batch_wait(stream)          # the HOST waits for batch k's copies only
                self.finish(cs)
                m.forward_backward_device(B, stream)
                if k + 1 < n:
                    self._copy.wait_stream(cs)       # the step kernel has read the batch buffers
                    m.stage_batch_async(users[k + 1], pos[k + 1], neg[k + 1], pp(k + 1), pn(k + 1), self._copy.cuda_stream)
                self._exchange_and_apply(stream, cs)
                m.read_loss_async(ring[k], stream)
            self.finish(cs)
        m.synchronize()
        return np.array(ring[:, :3])

    def sync_item_slots(self):
        """scatter mode keeps only this rank's row slice of the item Adam slots (m, v) live; all-gather them so that
        every rank holds the full state (checkpoints, switching to the all-reduce exchange or to one GPU)."""
        if self._own is None:
            return
        import torch
        import torch.distributed as dist
        self.finish()
        m = self.model
        for name in ("item_m", "item_v"):
            if self._cuda:
                t = torch.as_tensor(_DevArray(m.table_ptr(name), (m.n_items, m.emb_dim), "<f4"), device=self._gi.device)
            else:
                t = m.slot_tensor(name)
            for (lo, hi), (clo, chi) in zip(self._own, self._chunk):
                dist.all_gather_into_tensor(t[clo:chi], t[lo:hi].clone())
        if self._cuda:
            torch.cuda.synchronize()

    def state_dict(self):
        """PDAModel.state_dict() of this rank with complete item Adam slots (rank-local user rows)."""
        self.sync_item_slots()
        return self.model.state_dict()


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
