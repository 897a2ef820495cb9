"""2-GPU check (NCCL) of the data-parallel train step: user shards + replicated item table + item-gradient exchange
(reduce-scatter -> sliced Adam -> all-gather, or the chunked all-reduce) must equal ONE process stepping on the union
batch (CPU oracle).
Skipped on boxes with fewer than 2 GPUs (run it with `gpurun --gpus 2`)."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    import pda_b200
    return pda_b200.load().pda_device_count()


def _worker(rank, world, port, q, adam_mode, exchange):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import pda_b200
    from pda_b200.parallel import ShardedTrainer, shard_range
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        rng = np.random.default_rng(0)
        n_users, n_items, d, B = 4000, 900, 64, 256
        U = rng.normal(0, 0.3, (n_users, d)).astype(np.float32)
        I = rng.normal(0, 0.3, (n_items, d)).astype(np.float32)
        lo, hi = shard_range(n_users, world, rank)
        m = pda_b200.PDAModel(hi - lo, n_items, d, train="s_condition", batch_size=B * world, lr=1e-2, regs=1e-3,
                              device=rank, max_batch=B, init=False)
        m.set_table("user_embedding", U[lo:hi])
        m.set_table("item_embedding", I)
        if adam_mode == "lazy":
            m.set_adam_mode("lazy")       # the trainer narrows it to lazy users + dense (all-reduced) items
        try:
            tr = ShardedTrainer(m, world, rank, exchange=exchange)     # "nvls": the fused NVLink-multicast kernel
        except RuntimeError as e:
            if exchange not in ("nvls", "p2p") or "not available" not in str(e):
                raise
            q.put((rank, "skip", str(e)))
            return
        if rank == 0:
            print("exchange used:", tr.exchange, file=sys.stderr)
        assert exchange == "auto" or tr.exchange == exchange, (exchange, tr.exchange)
        stream = torch.cuda.current_stream().cuda_stream
        losses = []
        for step in range(6):
            srng = np.random.default_rng(100 + step)
            mine = None
            for r in range(world):
                rlo, rhi = shard_range(n_users, world, r)
                b = (srng.permutation(rhi - rlo)[:B].astype(np.int32), srng.integers(0, n_items, B).astype(np.int32),
                     srng.integers(0, n_items, B).astype(np.int32), srng.random(B).astype(np.float32),
                     srng.random(B).astype(np.float32))
                if r == rank:
                    mine = b
            losses.append(tr.train_step_host(*mine, stream=stream))
        q.put((rank, lo, hi, m.get_table("user_embedding"), m.get_table("item_embedding"), losses))
        m.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("adam_mode,exchange", [("dense", "scatter"), ("lazy", "scatter"), ("lazy", "allreduce"), ("lazy", "nvls"), ("lazy", "p2p")])
def test_two_gpus_equal_one_process_on_the_union_batch(c_oracle, adam_mode, exchange):
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from pda_b200.parallel import shard_range
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world = 2
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, adam_mode, exchange)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=400) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    if res[0][1] == "skip":
        pytest.skip("NVLink multicast not available here: " + res[0][2])
    rng = np.random.default_rng(0)
    n_users, n_items, d, B = 4000, 900, 64, 256
    U = rng.normal(0, 0.3, (n_users, d)).astype(np.float32)
    I = rng.normal(0, 0.3, (n_items, d)).astype(np.float32)
    ref = c_oracle.CModel(U, I, 1e-2, 1e-3, B * world, "s_condition")
    ref_losses = []
    for step in range(6):
        srng = np.random.default_rng(100 + step)
        parts = []
        for r in range(world):
            rlo, rhi = shard_range(n_users, world, r)
            parts.append((srng.permutation(rhi - rlo)[:B].astype(np.int32) + rlo, srng.integers(0, n_items, B).astype(np.int32),
                          srng.integers(0, n_items, B).astype(np.int32), srng.random(B).astype(np.float32),
                          srng.random(B).astype(np.float32)))
        cat = [np.concatenate([p[k] for p in parts]) for k in range(5)]
        ref_losses.append(ref.train_step(*cat))
    assert np.array_equal(res[0][4], res[1][4])                       # item replicas stay bit-identical
    assert np.abs(res[0][4] - ref.I).max() <= 1e-4 * np.abs(ref.I).max()
    for rank, lo, hi, Ur, _, losses in res:
        assert np.abs(Ur - ref.U[lo:hi]).max() <= 1e-4 * np.abs(ref.U).max()
        for got, want in zip(losses, ref_losses):
            assert np.allclose(got, want, rtol=1e-5)
